B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs"
run() { tag=$1; shift; env "$@" $B > gpurun_out/bench_cap_$tag.json 2> gpurun_out/bench_cap_$tag.err; python - gpurun_out/bench_cap_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    f=d['fast_mode']
    print(sys.argv[2], 'value %.1f e2e %.1f frac %.3f clk %s parity %s | bf16 %.1f e2e %.1f frac %.3f parity %s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['clocks']['sm_mhz'],d['parity']['ok'],f['value'],f['e2e']['value'],f['roofline']['frac'],f['parity']['ok']))
except Exception as e: print(sys.argv[2],'FAILED',e)
PY
}
run cap0 VLTK_SPLIT_CAP=0
run cap1 A=1
run cap0b VLTK_SPLIT_CAP=0
run cap1b A=1
python -m pytest tests/test_gpu_e2e.py -x -q 2>&1 | tail -2
