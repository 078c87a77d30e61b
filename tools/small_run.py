"""Small device run for compute-sanitizer: python tools/small_run.py <mode> [n_reads] [graph_bp paths read_len]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

mode = int(sys.argv[1])
nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
bp, paths, rlen = (int(x) for x in sys.argv[3:6]) if len(sys.argv) > 5 else (600, 6, 120)
g = synth.make_graph(bp, paths, seed=7)
reads = synth.make_reads(g, nreads, rlen, err=0.04, seed=8, mosaic_breaks=1 if mode >= 8 else 0)
al = Aligner(0)
al.load_gfa_text(g.gfa())
al.set_scoring()
recs, text = al.align(mode, reads)
print("mode", mode, "reads", len(reads), "status", [r.status for r in recs], "scores", [r.score for r in recs][:8])
