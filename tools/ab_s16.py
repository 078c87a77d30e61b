"""A/B check of the mode-2 kernel: packed 16-bit paths against the 32-bit-only paths (RG_NO_S16) on the same inputs.
Prints the first differing read of every case. Debug helper, not a test (tests compare with the oracle)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "tests", "golden", "example")


def records(al, reads):
    codes, off = al.pack_reads(reads)
    res = al.align_packed(2, codes, off)
    out = []
    for i in range(res.n_reads):
        r = res.reads[i]
        runs = [(res.runs[r.run_off + k].row, res.runs[r.run_off + k].op_count >> 28,
                 res.runs[r.run_off + k].op_count & 0x0fffffff) for k in range(r.n_runs)]
        out.append((r.status, r.score, r.end_row, r.end_col, r.start_row, r.start_col, r.cells, tuple(runs)))
    return out


def fasta_reads(path):
    seqs, cur = [], []
    for ln in open(path):
        if ln.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(ln.strip())
    if cur:
        seqs.append("".join(cur))
    return seqs


def run_case(name, gfa_text, reads, **sc):
    os.environ.pop("RG_NO_S16", None)
    a = Aligner(0)
    a.load_gfa_text(gfa_text)
    a.set_scoring(**sc)
    ra = records(a, reads)
    os.environ["RG_NO_S16"] = "1"
    b = Aligner(0)
    b.load_gfa_text(gfa_text)
    b.set_scoring(**sc)
    rb = records(b, reads)
    os.environ.pop("RG_NO_S16", None)
    os.environ["RG_FORCE_STRIPED"] = "1"
    c = Aligner(0)
    c.load_gfa_text(gfa_text)
    c.set_scoring(**sc)
    rc_ = records(c, reads)
    os.environ.pop("RG_FORCE_STRIPED", None)
    if rc_ != rb:
        print(f"{name}: striped kernel differs from the blocked 32-bit paths on", sum(1 for x, y in zip(rc_, rb) if x != y), "reads")
    bad = [i for i in range(len(reads)) if ra[i] != rb[i] or rc_[i] != rb[i]]
    print(f"{name}: {len(reads)} reads, {len(bad)} differ")
    for i in bad[:2]:
        x, y = ra[i], rb[i]
        print("  read", i, "len", len(reads[i]))
        print("   s16 :", x[:7], "runs", len(x[7]))
        print("   s32 :", y[:7], "runs", len(y[7]))
        for k, (p, q) in enumerate(zip(x[7], y[7])):
            if p != q:
                print("   first run difference at", k, p, q, "| before:", x[7][max(0, k - 3):k])
                break
    return len(bad)


def main():
    ex_gfa = open(os.path.join(EX, "graph.gfa")).read()
    ex_reads = fasta_reads(os.path.join(EX, "reads.fa"))
    n = 0
    n += run_case("example -b 50", ex_gfa, ex_reads, extra_b=50)
    n += run_case("example -b 1000", ex_gfa, ex_reads, extra_b=1000)
    for bp, paths, nr, rl, seed in [(1200, 5, 24, 150, 21), (6000, 8, 16, 700, 22), (800, 4, 20, 31, 23), (20000, 8, 64, 1000, 5),
                                   (3000, 6, 32, 100, 7), (3000, 6, 32, 300, 8)]:
        g = synth.make_graph(bp, paths, seed=seed)
        reads = synth.make_reads(g, nr, rl, err=0.05, seed=seed + 100)
        n += run_case(f"synth {bp}bp {rl}bp reads -b 3000", g.gfa(), reads, extra_b=3000)
        n += run_case(f"synth {bp}bp {rl}bp reads default", g.gfa(), reads)
    # read lengths around the points where the lane width changes (128 / 256 / 512 / 1024 columns incl. '$')
    g = synth.make_graph(4000, 6, seed=9)
    for rl in (126, 127, 128, 254, 255, 256, 510, 511, 512, 990, 1003, 1004, 1012, 1022, 1023):
        reads = synth.make_reads(g, 6, rl, err=0.0, seed=rl)
        reads = [r[:rl] for r in reads]
        n += run_case(f"lane boundary: reads of exactly {rl} bp, full band", g.gfa(), reads, extra_b=3000)
    print("TOTAL DIFFERING", n)
    return n


if __name__ == "__main__":
    main()
