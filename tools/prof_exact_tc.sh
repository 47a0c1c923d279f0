set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|Error|assert" 
python -m pytest tests/test_gpu_e2e.py -q -s -k bf16_tensor_core 2>&1 | grep -E "^\[|passed|failed|^E " | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --fast-mode none --streams 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 872 -c 440 --csv --log-file gpurun_out/r02_launches_exact_tc.csv $CMD > gpurun_out/ncu1.log 2>&1
echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:conv_tcx -s 527 -c 13 -o gpurun_out/prof_tcx $CMD > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out/ | tail -8
