"""Host logic on a CPU box: the product's GFA ingestion + flattening (csrc/host_graph.cpp: LnzGraph of graph.rs:31-123 with
r-values and segment ids, PathGraph / reverse PathGraph / distances of pathwise_graph.rs:135-354) against the oracle's
builders, through the host-only diagnostics rg_debug_dump_lnz / rg_debug_dump_pathgraph (no device involved)."""
import ctypes
import os

import pytest

from recgraph_b200 import _lib, synth
from tests import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _take(lib, p):
    assert p
    s = ctypes.string_at(p).decode()
    lib.rg_free(p)
    return s


def _check(gfa_text):
    lib = _lib.load()
    raw = gfa_text.encode()
    assert _take(lib, lib.rg_debug_dump_lnz(raw, len(raw))) == O.dump_lnz(gfa_text)
    for rev in (False, True):
        assert _take(lib, lib.rg_debug_dump_pathgraph(raw, len(raw), int(rev))) == O.dump_pathgraph(gfa_text, reverse_graph=rev)


def test_example_graph():
    _check(open(os.path.join(ROOT, "tests", "golden", "example", "graph.gfa")).read())


@pytest.mark.parametrize("bp,paths,seed", [(300, 3, 1), (1200, 5, 21), (6000, 8, 22), (5000, 64, 1), (3000, 128, 7), (20000, 32, 3)])
def test_synthetic_graphs(bp, paths, seed):
    _check(synth.make_graph(bp, paths, seed=seed).gfa())


def test_malformed_gfa_is_an_error_not_a_crash():
    lib = _lib.load()
    raw = b"S\t1\tACGT\nL\t1\t+\t9\t+\t0M\n"
    assert _take(lib, lib.rg_debug_dump_lnz(raw, len(raw))).startswith("ERROR")


def _fasta(text):
    lib = _lib.load()
    r = _lib.Reads()
    err = ctypes.create_string_buffer(256)
    raw = text.encode()
    rc = lib.rg_read_fasta_text(raw, len(raw), ctypes.byref(r), err, len(err))
    if rc != 0:
        return rc, err.value.decode()
    out = []
    for i in range(r.n_reads):
        codes = bytes(r.codes[k] for k in range(r.off[i], r.off[i + 1]))
        out.append((r.names[i].decode(), "".join("ACGTN"[c] for c in codes)))
    lib.rg_free_reads(ctypes.byref(r))
    return 0, out


def test_fasta_reader_follows_sequences_rs():
    """sequences.rs:5-45: multi-line records are concatenated, '-' becomes 'N', letters are upper-cased, the name is the
    header line without '>', empty lines are skipped; a header without sequence is the reference's "wrong fasta file
    format" panic."""
    rc, reads = _fasta(">r1 some description\nacgt\nAC-T\n\n>r2\nNNNN\nacg\n")
    assert rc == 0
    assert reads == [("r1 some description", "ACGTACNT"), ("r2", "NNNNACG")]
    rc, msg = _fasta(">only_a_name\n>r2\nACGT\n")
    assert rc != 0 and "fasta" in msg.lower()
    rc, msg = _fasta(">r1\nACGU\n")  # outside A,C,G,T,N: the reference panics later on the score lookup; we refuse at the door
    assert rc != 0
