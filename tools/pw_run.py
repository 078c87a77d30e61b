"""One pathwise run for profiling: python tools/pw_run.py <c3|c4> <mode> <reads> [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

cfg, mode, nreads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if cfg == "c3":
    g = synth.make_graph(10000, 32, seed=1)
    reads = synth.make_reads(g, nreads, 2000, err=0.05, seed=3)
    sc = {}
else:
    g = synth.make_graph(5000, 64, seed=1)
    reads = synth.make_reads(g, nreads, 1000, err=0.02, seed=3, mosaic_breaks=2)
    sc = dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0)
al = Aligner(0)
al.load_gfa_text(g.gfa())
al.set_scoring(**sc)
codes, off = al.pack_reads(reads)
al.upload(codes, off)
for _ in range(reps):
    al.align_staged(mode)
    ms, _l, _c = al.kernel_stats()
    print(f"{cfg} -m {mode}: {nreads} reads {ms:.2f} ms = {nreads / ms * 1e3:.0f} reads/s", flush=True)
