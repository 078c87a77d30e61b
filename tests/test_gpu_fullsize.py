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


def _oracle_sharded(args, reads, d, gfa, procs, env=None, timeout=3000):
    """The single-threaded oracle over read shards, `procs` processes at a time (one per host core where memory allows);
    returns the concatenated stdout in input order. Read names keep their global index."""
    from tests import oracle_lib
    oracle_lib.build()
    exe = os.path.join(ROOT, "oracle", "_build", "recgraph_oracle")
    n = len(reads)
    per = max(1, (n + procs - 1) // procs)
    bounds = [(lo, min(n, lo + per)) for lo in range(0, n, per)]
    running = []
    for k, (lo, hi) in enumerate(bounds):
        fa = os.path.join(d, f"shard{k}.fa")
        with open(fa, "w") as f:
            f.write("".join(f">read{i}\n{reads[i]}\n" for i in range(lo, hi)))
        running.append(subprocess.Popen([exe] + list(args) + [fa, gfa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                        env=env))
    outs = []
    for p in running:
        so, se = p.communicate(timeout=timeout)
        assert p.returncode == 0, se
        outs.append(so)
    return "".join(outs)


def _cli_vs_oracle(mode, g, reads, extra=(), procs=1, env=None, strip_number_quirk=False):
    from recgraph_b200 import run_cli
    with tempfile.TemporaryDirectory() as d:
        gfa, fa = os.path.join(d, "g.gfa"), os.path.join(d, "r.fa")
        open(gfa, "w").write(g.gfa())
        open(fa, "w").write(synth.fasta(reads))
        args = ["-m", str(mode)] + list(extra)
        rc, out, err = run_cli(args + [fa, gfa])
        assert rc == 0, err
        exp = _oracle_sharded(args, reads, d, gfa, procs, env=env)
        if out != exp:
            a, b = out.splitlines(), exp.splitlines()
            for k, (x, y) in enumerate(zip(a, b)):
                assert x == y, f"first difference at line {k}:\n GPU: {x[:300]}\n REF: {y[:300]}"
            raise AssertionError(f"line count differs: {len(a)} vs {len(b)}")


def test_c2_sample_against_oracle_pred32(c2):
    """64 reads of BASELINE config 2 at full size (100 kbp graph: more than 65 535 rows, outside the reference's 16-bit
    predecessor domain, so the oracle runs with the truncation disabled), one oracle process per host core."""
    g, reads = c2
    g64 = reads[:48] + synth.make_reads(g, 16, 1000, err=0.05, seed=4)
    _cli_vs_oracle(2, g, g64, procs=min(16, os.cpu_count() or 1), env=dict(os.environ, RGO_PRED32="1"))


def test_c2_60kbp_against_literal_u16_oracle():
    """The reference-valid copy of config 2: a 60 kbp graph keeps lnz.len() <= 65 535, so the oracle runs
    LITERALLY (16-bit predecessors, bitfield_path.rs:39-44) — 64 reads of 1 kbp at 5 % error."""
    g = synth.make_graph(60000, 8, seed=1)
    assert g.n_chars + 2 <= 65535
    reads = synth.make_reads(g, 64, 1000, err=0.05, seed=3)
    _cli_vs_oracle(2, g, reads, procs=min(16, os.cpu_count() or 1))


def test_c3_mode5_sample_against_oracle():
    """BASELINE config 3 at full size: -m 5, 32 haplotype paths, 10 kbp graph, 2 kbp reads — 16 of the 10k reads (the
    oracle needs ~3 GB and tens of seconds per read for its n x L x P tensor: 4 processes at a time)."""
    g = synth.make_graph(10000, 32, seed=1)
    reads = synth.make_reads(g, 16, 2000, err=0.05, seed=3)
    _cli_vs_oracle(5, g, reads, procs=4)


def test_c4_mode9_sample_against_oracle():
    """BASELINE config 4 at full size: -m 9 (R=4, r=0.1, B=1), 64 paths, 5 kbp graph, 1 kbp reads copied from
    2-breakpoint path mosaics with 2 % errors — 8 reads (best_alignment is O(n^2 L) on the CPU), one process each."""
    g = synth.make_graph(5000, 64, seed=1)
    reads = synth.make_reads(g, 8, 1000, err=0.02, seed=3, mosaic_breaks=2)
    _cli_vs_oracle(9, g, reads, extra=["-R", "4", "-r", "0.1", "-B", "1"], procs=8)
