"""Opcode histogram per kernel of the built library (cuobjdump -sass): the evidence that the pruning kernels are
sm_100a code using the bulk-copy engine (UBLKCP), mbarriers (SYNCS) and the FP64 pipe (DFMA/DMUL/DADD).
    python scripts/sass_histogram.py [lib.so] > profiles/r2_sass_opcode_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vbsky_b200/libvbsky_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {lib}: cubins for {arch}")
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEYS = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "FFMA", "LDS", "STS", "LDG", "STG", "ATOMG", "SHFL", "MUFU", "BAR"]
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':110s} {'total':>7s} " + " ".join(f"{k:>6s}" for k in KEYS))
for (k, h), name in sorted(zip(hist.items(), demangle), key=lambda x: -sum(x[0][1].values())):
    name = re.sub(r"\(.*", "", name)
    print(f"{name[:110]:110s} {sum(h.values()):7d} " + " ".join(f"{h.get(x, 0):6d}" for x in KEYS))
