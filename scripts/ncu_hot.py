"""Top SASS instructions by stall samples from an .ncu-rep: python scripts/ncu_hot.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
totinst = sum(int(r[ix["Instructions Executed"]]) for r in body)
print("total samples", tot, "total warp instr", totinst)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in body) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# opcode histogram by executed count
from collections import Counter
opc = Counter()
for r in body:
    op = r[ix["Source"]].strip().split()[0]
    if op.startswith("@"):
        op = r[ix["Source"]].strip().split()[1]
    opc[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
print("executed by opcode:", [(k, round(v / totinst * 100, 1)) for k, v in opc.most_common(25)])
body.sort(key=lambda r: -int(r[ix["# Samples"]]))
for r in body[:N]:
    st = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]])}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{int(r[ix["# Samples"]]):7d} {int(r[ix["Instructions Executed"]]):9d}  {r[ix["Source"]].strip()[:70]:70s} {top}')
