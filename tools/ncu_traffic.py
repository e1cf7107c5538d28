"""DRAM traffic of the depth-wise attention launches of one step, from an ncu capture (read here, no GPU).

GPU box (76 = launches of the recorded step; the first 76 are the warm-up pass of --once and are skipped):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:dw_ -s 76 -c 76 -o gpurun_out/r02_dw_traffic python tools/dw_replay.py --trace <trace> --once
here:
    python tools/ncu_traffic.py gpurun_out/r02_dw_traffic.ncu-rep profiles/r02_dw_trace_cifar16_b40.json

Writes profiles/r02_ncu_dw_attn_traffic.json (read by bench.py for roofline.traffic) and prints a per-launch table:
ncu's dram__bytes_read.sum + dram__bytes_write.sum next to the bytes the launch has to move and the algorithmic bytes
of the stacked-tensor formula."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}[unit]


def main():
    rep, trace_path = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    import dw_replay                                  # algorithmic_bytes / moved_bytes only (no torch use here)
    trace = json.load(open(trace_path))
    assert len(data) == len(trace), (len(data), len(trace))
    per, tot = [], {"dram": 0.0, "moved": 0, "alg": 0, "us": 0.0}
    for r, e in zip(data, trace):
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("davi::", "")
        moved, alg = dw_replay.moved_bytes(e), dw_replay.algorithmic_bytes(e)
        size = len(e["problems"]) if e["op"] == "gather" else e["n"]
        per.append({"op": e["op"], "size": size, "kernel": name, "dram_read_MB": round(rd / 1e6, 2),
                    "dram_write_MB": round(wr / 1e6, 2), "moved_MB": round(moved / 1e6, 2),
                    "algorithmic_MB": round(alg / 1e6, 2), "ncu_us": round(us, 2)})
        tot["dram"] += rd + wr; tot["moved"] += moved; tot["alg"] += alg; tot["us"] += us
    res = {"source": os.path.basename(rep) + " (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                     "gpu__time_duration.sum, one cold launch per recorded launch, after a 252 MB L2 flush)",
           "launches": len(per), "traffic_bytes_per_launch_sum": tot["dram"], "moved_bytes": tot["moved"],
           "algorithmic_bytes": tot["alg"], "traffic_over_algorithmic": tot["dram"] / tot["alg"],
           "ncu_time_us_sum": round(tot["us"], 1), "per_launch": per}
    with open(os.path.join(ROOT, "profiles", "r02_ncu_dw_attn_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("launches %d  dram %.1f MB  moved %.1f MB  algorithmic %.1f MB  dram/algorithmic %.3f  ncu time %.1f us" % (
        len(per), tot["dram"] / 1e6, tot["moved"] / 1e6, tot["alg"] / 1e6, tot["dram"] / tot["alg"], tot["us"]))
    for p in per:
        print("%-6s %3d  %-34s rd %8.2f wr %7.2f | moved %8.2f alg %8.2f MB | %7.2f us" % (
            p["op"], p["size"], p["kernel"][:34], p["dram_read_MB"], p["dram_write_MB"], p["moved_MB"], p["algorithmic_MB"],
            p["ncu_us"]))


if __name__ == "__main__":
    main()
