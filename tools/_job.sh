set -x
python -m pytest tests/test_gpu_stages.py -x -q -k "exact_tc" 2>&1 | tail -5 > gpurun_out/pytest_pairx.log
python -m pytest tests/test_gpu_e2e.py -x -q -k "golden" 2>&1 | tail -5 >> gpurun_out/pytest_pairx.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --fast-mode none"
VLTK_TCX_CTA2=0 $B > gpurun_out/bench_pairx_0.json 2> gpurun_out/bench_pairx_0.err
$B --profile-csv gpurun_out/ev_pairx_1.csv > gpurun_out/bench_pairx_1.json 2> gpurun_out/bench_pairx_1.err
VLTK_TCX_CTA2=0 $B --profile-csv gpurun_out/ev_pairx_0.csv > gpurun_out/bench_pairx_0b.json 2>> gpurun_out/bench_pairx_0.err
$B > gpurun_out/bench_pairx_1b.json 2>> gpurun_out/bench_pairx_1.err
cat gpurun_out/pytest_pairx.log; for f in gpurun_out/bench_pairx_*.json; do cut -c1-220 $f; done
