"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line:
   python tools/ncu_src_lines.py dump.csv [top]   -> line, instructions executed (share), stall samples (share), source"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        ln = int(r[0])
        inst = int(r[7]) if r[7] not in ("-", "") else 0
        samp = int(r[4]) if r[4] not in ("-", "") else 0
    except ValueError:
        continue
    k = (cur_file, ln)
    a = agg.setdefault(k, [0, 0, r[1]])
    a[0] += inst
    a[1] += samp
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total instructions {ti:.3e}, samples {ts}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:5d}  inst {100 * a[0] / ti:5.1f}%  stall {100 * a[1] / ts:5.1f}%  {a[2].strip()[:110]}")
