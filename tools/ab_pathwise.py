"""A/B of the two pathwise implementations on the device: the score-transport kernel (pathwise_tr.cu, default) against the
per-path kernel of round 1 (pathwise.cu, RG_PW_V1=1). Same records and run lists are required on every case; prints the
device time of both. Used by tests/test_gpu_pathwise_ab.py and by hand (`python tools/ab_pathwise.py --time`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

EXAMPLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "example")


def records(al, mode, reads):
    codes, off = al.pack_reads(reads)
    res = al.align_packed(mode, codes, off)
    out = []
    for i in range(res.n_reads):
        r = res.reads[i]
        runs = [(res.runs[r.run_off + k].row, res.runs[r.run_off + k].op_count) for k in range(r.n_runs + r.n_runs_rev)]
        out.append((r.status, r.score, r.score_f32, r.displacement, r.end_row, r.end_col, r.start_row, r.best_path,
                    r.rev_best_path, r.fen, r.rsn, r.rec_col, r.rev_end_row, r.n_runs, r.n_runs_rev, tuple(runs)))
    ms, _l, _c = al.kernel_stats()
    return out, ms


def make(v1, gfa_text, sc):
    if v1:
        os.environ["RG_PW_V1"] = "1"
    try:
        al = Aligner(0)
    finally:
        os.environ.pop("RG_PW_V1", None)
    al.load_gfa_text(gfa_text)
    al.set_scoring(**sc)
    return al


def compare(name, gfa_text, reads, modes, sc=None, verbose=True):
    sc = sc or {}
    bad = 0
    a, b = make(False, gfa_text, sc), make(True, gfa_text, sc)
    for mode in modes:
        ra, ta = records(a, mode, reads)
        rb, tb = records(b, mode, reads)
        diff = [i for i in range(len(reads)) if ra[i] != rb[i]]
        if verbose or diff:
            print(f"{name} -m {mode}: {len(reads)} reads, transport {ta:.2f} ms, per-path {tb:.2f} ms, differing reads: {len(diff)}", flush=True)
        for i in diff[:3]:
            x, y = ra[i], rb[i]
            print("   read", i, "tr :", x[:15], x[15][:6])
            print("   read", i, "v1 :", y[:15], y[15][:6])
        bad += len(diff)
    a.close()
    b.close()
    return bad


def example_reads():
    seqs = []
    for ln in open(os.path.join(EXAMPLE, "reads.fa")):
        if not ln.startswith(">"):
            seqs.append(ln.strip())
    return seqs


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true", help="also run the C3 / C4 sized cases")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args(argv)
    bad = 0
    bad += compare("example", open(os.path.join(EXAMPLE, "graph.gfa")).read(), example_reads()[:16 if a.quick else 52], [4, 5, 8, 9])
    cases = [("small", 1200, 5, 24, 150, 0.05, 21, 0), ("mid", 6000, 8, 12, 700, 0.05, 22, 0), ("short", 800, 4, 20, 31, 0.1, 23, 0),
             ("long", 900, 3, 6, 1400, 0.03, 24, 0), ("mos1", 900, 6, 16, 200, 0.02, 31, 2), ("mos2", 1500, 10, 12, 350, 0.03, 32, 1),
             ("p40", 2500, 40, 10, 300, 0.04, 33, 2), ("p70", 2000, 70, 8, 260, 0.03, 34, 1), ("p128", 1500, 128, 6, 200, 0.03, 35, 2)]
    if a.quick:
        cases = cases[:2] + cases[4:5]
    for name, bp, paths, nreads, rlen, err, seed, breaks in cases:
        g = synth.make_graph(bp, paths, seed=seed)
        reads = synth.make_reads(g, nreads, rlen, err=err, seed=seed + 100, mosaic_breaks=breaks)
        bad += compare(name, g.gfa(), reads, [4, 5, 8, 9])
        bad += compare(name + " R=1 r=0.05", g.gfa(), reads, [8, 9], dict(base_rec_cost=1, multi_rec_cost=0.05), verbose=False)
    if a.time:
        g = synth.make_graph(10000, 32, seed=1)
        reads = synth.make_reads(g, 296, 2000, err=0.05, seed=3)
        bad += compare("C3", g.gfa(), reads, [5])
        g = synth.make_graph(5000, 64, seed=1)
        reads = synth.make_reads(g, 148, 1000, err=0.02, seed=3, mosaic_breaks=2)
        bad += compare("C4", g.gfa(), reads, [9], dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0))
    print("differences:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
