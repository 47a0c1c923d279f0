set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r02_final.log
python -m pytest tests/test_gpu_e2e.py -q -s -k fresh 2>&1 | grep -E "seed|certified|passed|failed" > gpurun_out/pytest_r02_fresh_seeds.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_final.log 2>&1
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2>> gpurun_out/bench_r02_final.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --fast-mode none --streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_final_all.csv $CMD > gpurun_out/ncu1.log 2>&1
echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_tcx_kernel<\(int\)256, \(int\)5' -s 36 -c 9 -o gpurun_out/prof_tcx_final $CMD > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
cat gpurun_out/pytest_r02_final.log gpurun_out/smoke_r02_final.log; tail -3 gpurun_out/pytest_r02_fresh_seeds.log; cut -c1-300 gpurun_out/bench_r02_final.json
