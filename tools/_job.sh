set -x
python -m pytest tests/test_gpu_stages.py -x -q -k "exact" 2>&1 | tail -5 > gpurun_out/pytest_cc.log
python -m pytest tests/test_gpu_e2e.py -x -q -k "golden or batch8" 2>&1 | tail -5 >> gpurun_out/pytest_cc.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --fast-mode none"
VLTK_FUSE_SC=0 $B > gpurun_out/bench_cc_0.json 2> gpurun_out/bench_cc_0.err
$B > gpurun_out/bench_cc_1.json 2> gpurun_out/bench_cc_1.err
$B --streams 1 --profile-csv gpurun_out/ev_cc_s1.csv > gpurun_out/bench_cc_s1.json 2>> gpurun_out/bench_cc_1.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.max,lts__t_sectors_srcunit_tex_op_read.sum
PROBE_MODE=exact_tc PROBE_REPS=2 ncu --metrics $M --clock-control none -k regex:conv_tcx --csv --log-file gpurun_out/probe_tcx_pairs.csv python tools/layer_probe.py conv1 conv2 conv3res c512 > gpurun_out/probe_tcx.log 2>&1
VLTK_TCX_CTA2=0 PROBE_MODE=exact_tc PROBE_REPS=2 ncu --metrics $M --clock-control none -k regex:conv_tcx --csv --log-file gpurun_out/probe_tcx_single.csv python tools/layer_probe.py conv1 conv2 conv3res c512 >> gpurun_out/probe_tcx.log 2>&1
cat gpurun_out/pytest_cc.log; for f in gpurun_out/bench_cc_*.json; do cut -c1-220 $f; done
