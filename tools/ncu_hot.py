"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` output (read here, no GPU).
usage: ncu -i rep --page source --csv --kernel-name regex:NAME > f.csv ; python tools/ncu_hot.py f.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if r and r[0].startswith("0x") and len(r) == len(hdr)]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:N]
for i in sorted(idx):
    r = data[i]
    st = {h[6:]: int(r[hdr.index(h)]) for h in stalls if r[hdr.index(h)] not in ("0", "")}
    print(i, r[iS], r[iSrc].strip()[:64], st)
agg = {}
for r in data:
    for h in stalls:
        v = r[hdr.index(h)]
        if v not in ("0", ""):
            agg[h[6:]] = agg.get(h[6:], 0) + int(v)
print("by reason:", dict(sorted(agg.items(), key=lambda kv: -kv[1])))
