"""Per-source-line view of an ncu capture: joins `ncu --page source --print-source sass --csv` with the line table
of the cubin (nvdisasm -g) and prints, per CUDA source line range, executed warp instructions per DP row and the
share of stall samples.

  python tools/ncu_lines.py report.ncu-rep obj.o 'mangled-kernel-substring' ROWS [bucket]
"""
import csv
import re
import subprocess
import sys
import tempfile


def line_table(obj, kern):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
        import glob
        cub = glob.glob(d + "/*.cubin")[0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    tab, cur, inside = {}, 0, False
    for ln in txt.splitlines():
        if ln.startswith("//----") and ".text." in ln:
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*);", ln)
        if m:
            tab[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return tab


def main():
    rep, obj, kern, rows = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    bucket = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    tab = line_table(obj, kern)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rws = list(csv.reader(out.splitlines()))
    hdr = rws[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rws[2:]
    base = int(data[0][ix["Address"]], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0

    agg = {}
    tot_s = tot_i = 0.0
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        (fn, ln), _ = tab.get(off, (("?", 0), ""))
        key = (fn, ln // bucket * bucket)
        e = agg.setdefault(key, {"inst": 0.0, "samp": 0.0, "st": {}})
        e["inst"] += f(r[ix["Instructions Executed"]])
        e["samp"] += f(r[ix["# Samples"]])
        for s in stalls:
            e["st"][s] = e["st"].get(s, 0.0) + f(r[ix[s]])
        tot_s += f(r[ix["# Samples"]])
        tot_i += f(r[ix["Instructions Executed"]])
    print(f"total warp instructions {tot_i:.4g} = {tot_i / rows:.1f} per row; samples {tot_s:.0f}")
    print("file:line   inst/row  inst%  samples%  top stalls")
    for key in sorted(agg):
        e = agg[key]
        if e["inst"] / tot_i < 0.002 and e["samp"] / tot_s < 0.002:
            continue
        top = sorted(e["st"].items(), key=lambda x: -x[1])[:3]
        print(f"{key[0]}:{key[1]:<5d} {e['inst'] / rows:8.1f} {100 * e['inst'] / tot_i:5.1f}% {100 * e['samp'] / tot_s:6.1f}%  " +
              " ".join(f"{s[6:]}={100 * v / max(1.0, e['samp']):.0f}%" for s, v in top))


if __name__ == "__main__":
    main()
