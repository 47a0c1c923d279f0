"""Counts the Blackwell-native instructions per kernel in the built library (evidence that the hot path is tcgen05 / TMEM /
TMA code and not a recompiled mma.sync kernel or a library call):
    python tools/sass_summary.py [vltk_b200/libvltk_frcnn.so] > profiles/rNN_sass_summary.txt
Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA load
(.IM2COL = im2col mode), UTMASTG = TMA store, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, FFMA2 = packed fp32x2."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vltk_b200", "libvltk_frcnn.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("UTMALDG.IM2COL", r"\bUTMALDG\.\dD\.IM2COL"),
       ("UTMALDG", r"\bUTMALDG\.\dD(?!\.IM2COL)"), ("UTMASTG", r"\bUTMASTG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"),
       ("FFMA2", r"\bFFMA2"), ("HMMA/IMMA (mma.sync)", r"\b[HI]MMA\."), ("LDG.256", r"\bLDG\.[A-Z0-9.]*\.256")]
per = collections.OrderedDict()
name = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    if name is None:
        continue
    for key, pat in PAT:
        if re.search(pat, ln):
            per[name][key] += 1
demangled = dict(zip(per, subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()))
tot = collections.Counter()
print(f"# {os.path.basename(lib)}: {len(per)} kernels; `ldd` dependencies: "
      + ", ".join(sorted({l.split()[0] for l in subprocess.run(['ldd', lib], capture_output=True, text=True).stdout.splitlines() if l.strip()})))
print("# arch:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", subprocess.run(['cuobjdump', '-lelf', lib], capture_output=True, text=True).stdout)))))
print("%-110s %s" % ("kernel", "  ".join(k for k, _ in PAT)))
for n, c in per.items():
    tot.update(c)
    if not any(c[k] for k, _ in PAT[:8]):
        continue
    short = demangled.get(n, n).replace("vltk::(anonymous namespace)::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    short = re.sub(r"\((CUtensorMap|float|int|unsigned|const|__half|__nv|vltk).*", "", short)
    print("%-110s %s" % (short[:110], "  ".join(str(c[k]).rjust(len(k)) for k, _ in PAT)))
print("%-110s %s" % ("TOTAL (all kernels)", "  ".join(str(tot[k]).rjust(len(k)) for k, _ in PAT)))
