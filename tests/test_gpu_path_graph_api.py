"""rg_set_path_graph: a prebuilt PathGraph (pathwise_graph.rs:10-18) handed to the device instead of GFA text.
The arrays come from the independent Python builder (oracle/pyref, transcribed from pathwise_graph.rs:135-248), so the test
also checks the device's own GFA flattening against it: both routes must give the same records, and the GFA route is the
one the oracle parity tests cover. Graphs: the hand-built ones of the reference's inline tests (pathwise_graph.rs:364-544)
and synthetic bubbles."""
import importlib.util
import os

import pytest

from recgraph_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("recgraph_pyref", os.path.join(ROOT, "oracle", "pyref", "recgraph_pyref.py"))
pyref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(pyref)

# pathwise_graph.rs:364-404 (two paths through a diamond) and :406-449 / :498-544 (three paths, two starts, two ends)
DIAMOND = "H\tVN:Z:1.0\nS\t1\tA\nS\t2\tT\nS\t3\tC\nS\t4\tG\nL\t1\t+\t2\t+\t0M\nL\t1\t+\t3\t+\t0M\nL\t2\t+\t4\t+\t0M\nL\t3\t+\t4\t+\t0M\n" \
          "P\tp1\t1+,2+,4+\t*\nP\tp2\t1+,3+,4+\t*\n"
MULTI = "H\tVN:Z:1.0\nS\t1\tA\nS\t2\tT\nS\t3\tC\nS\t4\tG\nS\t5\tA\nL\t1\t+\t3\t+\t0M\nL\t2\t+\t3\t+\t0M\nL\t2\t+\t4\t+\t0M\n" \
        "L\t3\t+\t5\t+\t0M\nL\t4\t+\t5\t+\t0M\nP\tp1\t1+,3+,5+\t*\nP\tp2\t2+,4+,5+\t*\nP\tp3\t2+,3+,5+\t*\n"


def _records(al, mode, reads):
    codes, off = al.pack_reads(reads)
    res = al.align_packed(mode, codes, off)
    out = []
    for i in range(res.n_reads):
        r = res.reads[i]
        runs = [(res.runs[r.run_off + k].row, res.runs[r.run_off + k].op_count) for k in range(r.n_runs + r.n_runs_rev)]
        out.append((r.status, r.score, r.score_f32, r.displacement, r.end_row, r.end_col, r.start_row, r.best_path, r.rev_best_path,
                    r.fen, r.rsn, r.rec_col, tuple(runs)))
    return out, al.format_gaf_all(mode, res, off)


def _both_routes(gfa_text, reads, modes):
    from recgraph_b200 import Aligner
    segs, paths = pyref.read_gfa(gfa_text)
    g = pyref.create_path_graph(segs, paths)
    a, b = Aligner(), Aligner()
    a.load_gfa_text(gfa_text)
    b.set_path_graph(g.lnz, g.nwp, g.pred, g.pn, [x if x <= g.P else g.P + 1 for x in g.alphas], g.P, g.ids)
    for mode in modes:
        ra, ta = _records(a, mode, reads)
        rb, tb = _records(b, mode, reads)
        assert ra == rb, f"mode {mode}"
        assert ta == tb, f"mode {mode}: GAF text"
    assert a.graph_info()[0] == b.graph_info()[0] and a.graph_info()[2] == b.graph_info()[2]


def test_reference_inline_test_graphs():
    _both_routes(DIAMOND, ["ATG", "ACG", "AG", "TTTT"], [4, 5, 6, 7, 8, 9])
    _both_routes(MULTI, ["ACA", "TGA", "TCA", "TGT", "GGGA"], [5, 7, 9])


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_synthetic_graphs(seed):
    g = synth.make_graph(900 + 300 * seed, 3 + 2 * seed, seed=seed)
    reads = synth.make_reads(g, 10, 120, err=0.04, seed=seed + 50, mosaic_breaks=1)
    _both_routes(g.gfa(), reads, [4, 5, 8, 9])


def test_lnz_graph_after_path_graph_drops_the_paths():
    """a path graph left by an earlier load must not survive rg_set_lnz_graph."""
    from recgraph_b200 import Aligner, RecGraphError
    al = Aligner()
    al.load_gfa_text(DIAMOND)
    al.align(5, ["ATG"])
    al.set_lnz_graph(list("$ACGTF"), {1, 5}, {1: [0], 5: [4]})
    with pytest.raises(RecGraphError):
        al.align(5, ["ACGT"])
    recs, _ = al.align(2, ["ACGT"])
    assert recs[0].status & ~1 == 0
