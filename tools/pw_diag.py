"""Phase timings of the score-transport kernel (RG_PW_DIAG=1: kilo-cycles per read in the record's spare fields)."""
import os
import sys
os.environ["RG_PW_DIAG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import Aligner, synth  # noqa: E402

for name, mode, bp, paths, nreads, rlen, err, mosaic, sc in [
        ("C3 -m 5", 5, 10000, 32, int(os.environ.get("C3_READS", 888)), 2000, 0.05, 0, {}),
        ("C3 -m 4", 4, 10000, 32, 444, 2000, 0.05, 0, {}),
        ("C4 -m 9", 9, 5000, 64, int(os.environ.get("C4_READS", 444)), 1000, 0.02, 2, dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0)),
        ("C4 -m 8", 8, 5000, 64, 148, 1000, 0.02, 2, dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0))]:
    g = synth.make_graph(bp, paths, seed=1)
    reads = synth.make_reads(g, nreads, rlen, err=err, seed=3, mosaic_breaks=mosaic)
    al = Aligner(0)
    al.load_gfa_text(g.gfa())
    al.set_scoring(**sc)
    codes, off = al.pack_reads(reads)
    al.upload(codes, off)
    al.align_staged(mode)
    al.align_staged(mode)
    ms, _l, _c = al.kernel_stats()
    res = al.fetch()
    n = res.n_reads
    f = lambda k: sum(getattr(res.reads[i], k) for i in range(n)) / n
    if mode < 8:
        print(f"{name}: {n} reads {ms:.1f} ms = {n / ms * 1e3:.0f} reads/s | per read kcycles: fwd {f('fen'):.0f} (materialising rows {f('rsn'):.0f}) replay+walk {f('rev_end_row'):.0f}", flush=True)
    else:
        print(f"{name}: {n} reads {ms:.1f} ms = {n / ms * 1e3:.0f} reads/s | per read kcycles: fwd {f('fen'):.0f} rev {f('rsn'):.0f} (materialising rows {f('displacement'):.0f}) pairs {f('rec_col'):.0f} replay+walk {f('rev_end_row'):.0f}", flush=True)
    al.close()
