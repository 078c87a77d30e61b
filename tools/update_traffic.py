"""Regenerate profiles/traffic.json (the `roofline.traffic` field of bench.py) from ncu captures:

    ncu --set full --clock-control none -k regex:<kernel> -c 1 -o gpurun_out/<name> -f python <workload>
    python tools/update_traffic.py key=gpurun_out/<name>.ncu-rep [key=...]

key is the name bench.py looks up (k_gap_global_blk, k_pathwise_tr_m5, k_pathwise_tr_m9). Stores per launch
dram__bytes_read.sum + dram__bytes_write.sum, the kernel's duration under the profiler, and where the numbers came from."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for h, u, v in zip(hdr, units, vals):
        m[h] = (u, v)
    return m


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    src = []
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        m = metrics(rep)
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u, v = m[name]
            tot += float(v) * UNIT[u]
        out[key] = int(tot)
        out[key + "_ms_under_ncu"] = float(m["gpu__time_duration.sum"][1]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[m["gpu__time_duration.sum"][0]]
        src.append(f"{key}: {os.path.basename(rep)}")
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out["_source"] = "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum per launch), tools/update_traffic.py at " + head + "; " + "; ".join(src)
    out.pop("_comment", None)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
