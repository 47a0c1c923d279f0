python tools/determinism_probe.py exact_tc 24 2>&1 | grep -v Warn | tail -12
python tools/determinism_probe.py exact_tc 24 --stress 2>&1 | grep -v Warn | tail -12
VLTK_PDL=0 python tools/determinism_probe.py exact_tc 12 --stress 2>&1 | grep -v Warn | tail -6
