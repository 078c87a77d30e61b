"""Parity at BASELINE.json's full sizes (C2: 100 kbp graph, 1 kbp reads) through size-independent properties:
two independent device implementations agree, results do not depend on batch composition / order, CIGAR bookkeeping
is consistent, and a sample is compared with the oracle run with RGO_PRED32=1 (the graph has more than 65 535 rows,
outside the reference's 16-bit predecessor domain)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from recgraph_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def c2():
    g = synth.make_graph(100000, 8, seed=1)
    reads = synth.make_reads(g, 48, 1000, err=0.05, seed=3)
    return g, reads


def _records(al, mode, reads):
    codes, off = al.pack_reads(reads)
    res = al.align_packed(mode, codes, off)
    out = []
    for i in range(res.n_reads):
        r = res.reads[i]
        runs = [(res.runs[r.run_off + k].row, res.runs[r.run_off + k].op_count) for k in range(r.n_runs)]
        out.append((r.status, r.score, r.end_row, r.end_col, r.start_row, r.start_col, r.cells, tuple(runs)))
    return out


def test_c2_blocked_and_striped_kernels_agree_and_order_invariant(c2):
    from recgraph_b200 import Aligner
    g, reads = c2
    al = Aligner()
    al.load_gfa_text(g.gfa())
    al.set_scoring()
    a = _records(al, 2, reads[:24])
    perm = list(reversed(range(24)))
    b = _records(al, 2, [reads[i] for i in perm])
    assert [b[perm.index(i)] for i in range(24)] == a
    os.environ["RG_FORCE_STRIPED"] = "1"
    try:
        al2 = Aligner()
        al2.load_gfa_text(g.gfa())
        al2.set_scoring()
        c = _records(al2, 2, reads[:24])
    finally:
        del os.environ["RG_FORCE_STRIPED"]
    assert c == a
    # bookkeeping: read-consuming steps (D, d, L) cover the read from start_col to end_col; rows never increase
    for (status, score, end_row, end_col, start_row, start_col, cells, runs), rd in zip(a, reads):
        assert status & ~1 == 0
        consumed = sum(oc & 0x0fffffff for row, oc in runs if (oc >> 28) in (0, 1, 3))
        assert consumed == end_col - start_col and end_col == len(rd) and start_col == 0
        rows = [row for row, oc in runs]
        assert rows == sorted(rows, reverse=True)
        assert cells > 0


def test_c2_sample_against_oracle_pred32(c2):
    from recgraph_b200 import run_cli
    g, reads = c2
    with tempfile.TemporaryDirectory() as d:
        gfa, fa = os.path.join(d, "g.gfa"), os.path.join(d, "r.fa")
        open(gfa, "w").write(g.gfa())
        open(fa, "w").write(synth.fasta(reads[:3]))
        rc, out, err = run_cli(["-m", "2", fa, gfa])
        assert rc == 0, err
        from tests import oracle_lib
        oracle_lib.build()
        env = dict(os.environ, RGO_PRED32="1")
        r = subprocess.run([os.path.join(ROOT, "oracle", "_build", "recgraph_oracle"), "-m", "2", fa, gfa],
                           capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr
        assert out == r.stdout


def _cli_vs_oracle(mode, g, reads, extra=(), timeout=1500):
    from recgraph_b200 import run_cli
    from tests import oracle_lib
    with tempfile.TemporaryDirectory() as d:
        gfa, fa = os.path.join(d, "g.gfa"), os.path.join(d, "r.fa")
        open(gfa, "w").write(g.gfa())
        open(fa, "w").write(synth.fasta(reads))
        args = ["-m", str(mode)] + list(extra) + [fa, gfa]
        rc, out, err = run_cli(args)
        assert rc == 0, err
        oracle_lib.build()
        r = subprocess.run([os.path.join(ROOT, "oracle", "_build", "recgraph_oracle")] + args, capture_output=True,
                           text=True, timeout=timeout)
        assert r.returncode == 0, r.stderr
        assert out == r.stdout


def test_c3_mode5_sample_against_oracle():
    """BASELINE config 3 at full size: -m 5, 32 haplotype paths, 10 kbp graph, 2 kbp reads (a sample of the 10k reads;
    the oracle needs ~3 GB and tens of seconds per read for its n x L x P tensor)."""
    g = synth.make_graph(10000, 32, seed=1)
    reads = synth.make_reads(g, 2, 2000, err=0.05, seed=3)
    _cli_vs_oracle(5, g, reads)


def test_c4_mode9_sample_against_oracle():
    """BASELINE config 4 at full size: -m 9 (R=4, r=0.1, B=1), 64 paths, 5 kbp graph, 1 kbp reads copied from
    2-breakpoint path mosaics with 2 % errors (one read: best_alignment is O(n^2 L) on the CPU)."""
    g = synth.make_graph(5000, 64, seed=1)
    reads = synth.make_reads(g, 1, 1000, err=0.02, seed=3, mosaic_breaks=2)
    _cli_vs_oracle(9, g, reads, extra=["-R", "4", "-r", "0.1", "-B", "1"])
