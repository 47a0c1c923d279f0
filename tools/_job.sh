set -x
python -m pytest tests/test_gpu_stages.py -x -q -k "exact" 2>&1 | tail -3 > gpurun_out/pytest_pool.log
python -m pytest tests/test_gpu_e2e.py -x -q 2>&1 | tail -3 >> gpurun_out/pytest_pool.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --fast-mode none"
VLTK_FUSE_MEAN=0 $B > gpurun_out/bench_pool_0.json 2> gpurun_out/bench_pool_0.err
$B > gpurun_out/bench_pool_1.json 2> gpurun_out/bench_pool_1.err
VLTK_FUSE_MEAN=0 $B > gpurun_out/bench_pool_0b.json 2>> gpurun_out/bench_pool_0.err
$B > gpurun_out/bench_pool_1b.json 2>> gpurun_out/bench_pool_1.err
cat gpurun_out/pytest_pool.log
