"""One short pass of the hot path for ncu (never a bench number)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", type=int, default=2)
ap.add_argument("--reads", type=int, default=1184)
ap.add_argument("--read-len", type=int, default=1000)
ap.add_argument("--graph-bp", type=int, default=100000)
ap.add_argument("--paths", type=int, default=8)
ap.add_argument("--passes", type=int, default=1)
ap.add_argument("--mosaic", type=int, default=0)
ap.add_argument("--err", type=float, default=0.05)
a = ap.parse_args()
g = synth.make_graph(a.graph_bp, a.paths, seed=1)
reads = synth.make_reads(g, a.reads, a.read_len, err=a.err, seed=3, mosaic_breaks=a.mosaic)
al = Aligner(0)
al.load_gfa_text(g.gfa())
al.set_scoring()
codes, off = al.pack_reads(reads)
al.upload(codes, off)
import time
for _ in range(a.passes):
    t0 = time.perf_counter()
    al.align_staged(a.mode)
    wall = 1e3 * (time.perf_counter() - t0)
    print("kernel_ms, launches, cells:", al.kernel_stats(), "wall ms of the call %.1f" % wall)
res = al.fetch()
print("cells", sum(res.reads[i].cells for i in range(res.n_reads)))
import hashlib
h = hashlib.sha256()
for i in range(res.n_reads):
    r = res.reads[i]
    h.update(repr((r.status, r.score, r.end_row, r.end_col, r.start_row, r.start_col, r.n_runs, r.cells)).encode())
print("result checksum", h.hexdigest()[:16], "gaf sha", hashlib.sha256(al.format_gaf_all(a.mode, res, off).encode()).hexdigest()[:16])
import numpy as np
dp = np.array([res.reads[i].fen for i in range(res.n_reads)], dtype=np.float64)
tb = np.array([res.reads[i].rsn for i in range(res.n_reads)], dtype=np.float64)
if a.mode == 2:
    print("kcycles per read: forward mean %.0f, traceback mean %.0f (%.1f%% of total); runs/read %.0f" % (dp.mean(), tb.mean(), 100 * tb.sum() / (dp.sum() + tb.sum()), res.n_runs_total / max(1, res.n_reads)))
if a.mode == 2 and os.environ.get("RG_ROWSTATS"):
    f = lambda name: np.array([float(getattr(res.reads[i], name)) for i in range(res.n_reads)]).sum()
    cnt = [f("best_path"), f("rev_best_path"), f("rec_col"), f("rev_end_row"), f("displacement")]
    cyc = [f("score_f32"), f("n_runs_rev"), f("end_col"), f("start_row"), f("start_col")]
    tot = sum(cnt)
    for nm, c, k in zip(["f16", "g16", "plain32", "band32", "general/i0"], cnt, cyc):
        print("%-10s rows %6.2f%%  cycles %6.2f%%  cycles/row %8.0f" % (nm, 100 * c / tot, 100 * k / sum(cyc), 1024 * k / max(1, c)))
