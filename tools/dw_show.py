"""Summarise a tools/dw_replay.py output file."""
import json, sys
d = json.load(open(sys.argv[1]))
print("ms %.4f  GB/s %.0f  frac %.4f" % (d["ms"], d["GBs"], d["GBs"] / 6553))
tot = {}
for p in d["per_launch"]:
    k = p["op"] + ("" if p["op"] == "gather" else (":v2" if p["n"] >= 10 else ":v1"))
    tot.setdefault(k, [0, 0]); tot[k][0] += p["us"]; tot[k][1] += p["moved_MB"]
for k, v in tot.items():
    print("  %-10s %7.1f us %8.1f MB  %5.0f GB/s" % (k, v[0], v[1], v[1] / v[0] * 1e3))
if len(sys.argv) > 2:
    for p in d["per_launch"]:
        print(p["op"], p.get("n", p.get("consumers")), p.get("C"), p["us"], p["moved_MB"], p["moved_GBs"])
