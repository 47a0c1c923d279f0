M=gpu__time_duration.sum,sm__cycles_elapsed.max,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum
for st in 5 4 3; do
VLTK_TCX_PROBE_STAGES=$st PROBE_MODE=exact_tc PROBE_REPS=2 ncu --metrics $M --clock-control none -k regex:conv_tcx --csv --log-file gpurun_out/probe_stages_$st.csv python tools/layer_probe.py conv2 conv1 > gpurun_out/probe_stages.log 2>&1
done
python -m pytest tests/test_gpu_stages.py -x -q -k "hbm_scale" 2>&1 | tail -3
