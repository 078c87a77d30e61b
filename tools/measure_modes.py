"""Throughput of the pathwise (C3: -m 5) and recombination (C4: -m 9) configurations of BASELINE.json, device time of
the kernels (CUDA events inside the library), inputs resident in HBM. Reported next to the headline bench in DESIGN.md."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c3-reads", type=int, default=592)
ap.add_argument("--c4-reads", type=int, default=148)
a = ap.parse_args()
out = {}
for name, mode, bp, paths, nreads, rlen, err, mosaic, sc in [
        ("C3 -m 5: 32 paths, 10 kbp graph, 2 kbp reads", 5, 10000, 32, a.c3_reads, 2000, 0.05, 0, {}),
        ("C4 -m 9: 64 paths, 5 kbp graph, 1 kbp mosaic reads, R=4 r=0.1 B=1", 9, 5000, 64, a.c4_reads, 1000, 0.02, 2,
         dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0))]:
    g = synth.make_graph(bp, paths, seed=1)
    reads = synth.make_reads(g, nreads, rlen, err=err, seed=3, mosaic_breaks=mosaic)
    al = Aligner(0)
    al.load_gfa_text(g.gfa())
    al.set_scoring(**sc)
    n_rows, _s, P = al.graph_info()
    codes, off = al.pack_reads(reads)
    al.upload(codes, off)
    best = None
    for _ in range(3):
        al.align_staged(mode)
        ms, nl, _c = al.kernel_stats()
        best = ms if best is None else min(best, ms)
    res = al.fetch()
    bad = sum(1 for i in range(res.n_reads) if res.reads[i].status & ~(1 | 16 | 32))
    rows_cols = sum((n_rows - 1) * (len(r) + 1) for r in reads)
    dirs = 2 if mode >= 8 else 1
    out[name] = {"reads": nreads, "kernel_ms": best, "reads_per_s": nreads / (best * 1e-3),
                 "gcups_row_col": dirs * rows_cols / (best * 1e-3) / 1e9,
                 "path_cells_per_s_upper": dirs * rows_cols * P / (best * 1e-3), "rows": n_rows, "paths": P, "bad_status": bad}
    print(name, json.dumps(out[name]), flush=True)

# POA modes with full-matrix semantics (0: AVX2 routine of global_abpoa, 1: local, 3: affine local) on the headline graph
g = synth.make_graph(100000, 8, seed=1)
reads = synth.make_reads(g, 296, 1000, err=0.05, seed=3)
al = Aligner(0)
al.load_gfa_text(g.gfa())
al.set_scoring()
codes, off = al.pack_reads(reads)
al.upload(codes, off)
for mode in (0, 1, 3):
    al.align_staged(mode)
    al.align_staged(mode)
    ms, _nl, _c = al.kernel_stats()
    res = al.fetch()
    cells = sum(res.reads[i].cells for i in range(res.n_reads))
    print("C2 graph -m %d: 296 reads x 1 kbp" % mode, json.dumps({"kernel_ms": ms, "reads_per_s": 296 / (ms * 1e-3), "gcups": cells / (ms * 1e-3) / 1e9}), flush=True)
