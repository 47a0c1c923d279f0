B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-configs --fast-mode none"
run() { tag=$1; shift; env "$@" $B > gpurun_out/bench_x_$tag.json 2> gpurun_out/bench_x_$tag.err; python - gpurun_out/bench_x_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value %.1f e2e %.1f tc_ms %.2f frac %.3f clk %s parity %s'%(d['value'],d['e2e']['value'],d['roofline']['ms_per_step'],d['roofline']['frac'],d['clocks']['sm_mhz'],d['parity']['ok']))
except Exception as e: print(sys.argv[2],'FAILED',e)
PY
}
run base A=1
run cta2_4096 VLTK_TCX_CTA2=4096
run cta2_1 VLTK_TCX_CTA2=1
run chunk8 VLTK_TCX_CHUNK=8
run smallk128 VLTK_TCX_SMALLK=128
run smallk512 VLTK_TCX_SMALLK=512
run nosplit VLTK_SPLIT_BACKBONE=0
run base2 A=1
