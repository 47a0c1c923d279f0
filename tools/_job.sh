set -x
python -m pytest tests/test_gpu_stages.py -x -q -k "exact or pair" 2>&1 | tail -3 > gpurun_out/pytest_ep.log
python -m pytest tests/test_gpu_e2e.py -x -q -k "golden" 2>&1 | tail -3 >> gpurun_out/pytest_ep.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs"
$B > gpurun_out/bench_ep_1.json 2> gpurun_out/bench_ep_1.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max
PROBE_MODE=exact_tc PROBE_REPS=2 ncu --metrics $M --clock-control none -k regex:conv_tcx --csv --log-file gpurun_out/probe_ep.csv python tools/layer_probe.py conv1 conv2 conv3res c512 > gpurun_out/probe_ep.log 2>&1
cat gpurun_out/pytest_ep.log; cut -c1-220 gpurun_out/bench_ep_1.json
