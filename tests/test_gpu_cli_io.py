"""CLI surface around the hot path: `-o` file semantics (utils.rs:200-219 with both numbering conventions: POA modes pass
i + 1, pathwise modes pass i — main.rs:260,268,311 — so the second record of a pathwise run TRUNCATES the file), the
api.rs no-gap entry points (f32 matrix with gap = X, api.rs:11-40,76-99), and the sharded multi-device driver (`--gpus N`:
same text as one device)."""
import os

import pytest

from recgraph_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example")
EX = [os.path.join(EXAMPLE, "reads.fa"), os.path.join(EXAMPLE, "graph.gfa")]


@pytest.mark.parametrize("mode", ["0", "2", "3", "4", "5", "9"])
def test_out_file_semantics(tmp_path, mode):
    from recgraph_b200 import run_cli
    from tests import oracle_lib
    a, b = tmp_path / "gpu.gaf", tmp_path / "ref.gaf"
    extra = ["-b", "50"] if mode in ("0", "2") else []
    rc, out, err = run_cli(["-m", mode, "-o", str(a)] + extra + EX)
    orc, oout, oerr = oracle_lib.run_cli(["-m", mode, "-o", str(b)] + extra + EX)
    assert rc == 0 and orc == 0, (err, oerr)
    assert out == oout            # warnings still go to stdout
    assert a.read_text() == b.read_text()
    n = len(a.read_text().splitlines())
    # POA modes: all 52 records; pathwise modes: the record of read 1 re-creates the file (number == 1), read 0's is lost
    assert n == (52 if mode in ("0", "2", "3") else 51)
    # an existing file is appended to unless number == 1
    a.write_text("stale\n")
    rc, out, err = run_cli(["-m", mode, "-o", str(a)] + extra + EX)
    assert rc == 0
    assert a.read_text().splitlines()[0] != "stale"


def _example_reads():
    names, seqs = [], []
    for ln in open(EX[0]):
        if ln.startswith(">"):
            names.append(ln[1:].strip())
        else:
            seqs.append(ln.strip())
    return names, seqs


def test_api_no_gap_entry_points():
    """align_global_no_gap / align_local_no_gap (api.rs:11-40,76-99): exec_simd with the f32 match/mismatch matrix whose
    gap-vs-char entries are X (score_matrix.rs:52-66, not 2X as on the command line) and bases_to_add = 0.1 * len. The CLI
    cannot express that matrix, so the expectation comes from the oracle's exec_simd through its own C entry point."""
    import ctypes
    import recgraph_b200 as rb
    from tests import oracle_lib
    names, seqs = _example_reads()
    al = rb.Aligner()
    al.load_gfa(EX[1])
    lib = oracle_lib.load()
    lib.rgo_api_no_gap.restype = ctypes.c_void_p
    lib.rgo_api_no_gap.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    gfa_text = open(EX[1]).read().encode()
    n_real = 0
    for k in range(8):
        for local in (0, 1):
            p = lib.rgo_api_no_gap(gfa_text, seqs[k].encode(), names[k].encode(), local)
            exp = ctypes.string_at(p).decode()
            lib.rgo_free(p)
            assert not exp.startswith("PANIC"), exp
            g = rb.align_local_no_gap(seqs[k], al, (names[k], k + 1)) if local else rb.align_global_no_gap(seqs[k], al, (names[k], k + 1))
            assert g.to_string() == exp.rstrip("\n").split("\n")[-1], (k, local)
            n_real += "\t" in exp and not exp.startswith("\t")
    assert n_real >= 8


def test_scalar_local_mode_on_the_device():
    """RG_MODE_LOCAL_SCALAR = local_poa::exec (local_poa.rs:181-255, the routine of hosts without AVX2): the reference's two
    inline vectors (local_poa.rs:304-338, 341-377) and GAF text against the oracle's restatement of that routine."""
    import ctypes
    import recgraph_b200 as rb
    from tests import oracle_lib
    from tests.test_oracle_golden import POA_CASES
    al = rb.Aligner()
    idx = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 5}
    seen = 0
    for variant, (lnz, nwp, preds), read, scores, o, e, bta, expected, cite in POA_CASES:
        if variant != 1:
            continue
        table = [[-1] * 6 for _ in range(6)]
        for (a, b), v in scores.items():
            table[idx[a]][idx[b]] = v
        table[5][5] = 0
        al.set_lnz_graph(lnz, nwp, preds)
        al.set_scoring(table=table)
        recs, _ = al.align(11, [read[1:]])
        assert recs[0].score == expected, cite
        seen += 1
    assert seen == 2
    lib = oracle_lib.load()
    lib.rgo_local_scalar_gaf.restype = ctypes.c_void_p
    lib.rgo_local_scalar_gaf.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    names, seqs = _example_reads()
    gfa_text = open(EX[1]).read().encode()
    al2 = rb.Aligner()
    al2.load_gfa(EX[1])
    for (m, x) in ((2, 4), (1, 1), (3, 2)):
        al2.set_scoring(match=m, mismatch=x)
        recs, text = al2.align(11, seqs[:12], names=names[:12])
        got = text.splitlines()
        for k in range(12):
            p = lib.rgo_local_scalar_gaf(gfa_text, seqs[k].encode(), names[k].encode(), m, -x)
            exp = ctypes.string_at(p).decode()
            lib.rgo_free(p)
            assert not exp.startswith("PANIC"), exp
            assert got[k] == exp.rstrip("\n").split("\n")[-1], (m, x, k)
    g = synth.make_graph(2500, 5, seed=9)
    reads = synth.make_reads(g, 10, 200, err=0.06, seed=10)
    al3 = rb.Aligner()
    al3.load_gfa_text(g.gfa())
    al3.set_scoring()
    recs, text = al3.align(11, reads)
    for k, ln in enumerate(text.splitlines()):
        p = lib.rgo_local_scalar_gaf(g.gfa().encode(), reads[k].encode(), f"read{k}".encode(), 2, -4)
        exp = ctypes.string_at(p).decode()
        lib.rgo_free(p)
        assert ln == exp.rstrip("\n").split("\n")[-1], k


def test_gpus_flag_single_device_is_identity():
    from recgraph_b200 import run_cli
    for mode in ("2", "5"):
        a = run_cli(["-m", mode] + EX)
        b = run_cli(["-m", mode, "--gpus", "1"] + EX)
        assert a[0] == 0 and a[:2] == b[:2]


@pytest.mark.parametrize("mode", ["0", "2", "5", "7", "9"])
def test_sharded_driver_two_devices(mode, tmp_path):
    """`recgraph --gpus 2`: contiguous cost-balanced shards, one context + graph replica per device, records in input order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from recgraph_b200 import run_cli
    g = synth.make_graph(3000, 6, seed=5)
    reads = synth.make_reads(g, 41, 300, err=0.05, seed=6, mosaic_breaks=1, exact_len=False)
    gfa, fa = tmp_path / "g.gfa", tmp_path / "r.fa"
    gfa.write_text(g.gfa())
    fa.write_text(synth.fasta(reads))
    one = run_cli(["-m", mode, str(fa), str(gfa)])
    two = run_cli(["-m", mode, "--gpus", "2", str(fa), str(gfa)])
    assert one[0] == 0 and two[0] == 0, (one[2], two[2])
    assert one[1] == two[1]
    o = tmp_path / "out.gaf"
    three = run_cli(["-m", mode, "--gpus", "2", "-o", str(o), str(fa), str(gfa)])
    assert three[0] == 0


def test_mode0_zero_band_amplitude(tmp_path):
    """b + f * L < 1 gives mode 0 rows whose band holds column 0 only: no cell is processed, best_col keeps its initial value
    `left` (global_abpoa.rs:80,160-162,222), the end cell is unset and the reference prints "band not enough for correct
    output" — or aligns when f * L reaches 1 for the longer read. The device took the arg-max of an empty set there until
    the last hour of round 2 (found by running -b 0 against the oracle). Same 40 graphs x 3 flag sets as that run."""
    import numpy as np
    from recgraph_b200 import run_cli, synth
    from tests import oracle_lib
    aligned = 0
    for seed in range(800000, 800040):
        rng = np.random.default_rng(seed)
        g = synth.make_graph(int(rng.integers(60, 300)), 3, seed=seed, mean_seg=int(rng.integers(3, 12)), p_snp=0.25, p_indel=0.15)
        reads = synth.make_reads(g, 2, int(rng.integers(10, 100)), err=float(rng.choice([0.0, 0.05, 0.2])), seed=seed + 1)
        gfa, fa = tmp_path / "g.gfa", tmp_path / "r.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        for extra in (["-b", "0", "-f", "0.0"], ["-b", "0", "-f", "0.01"], ["-b", "0", "-f", "0.0", "-s", "true"]):
            args = ["-m", "0"] + extra + [str(fa), str(gfa)]
            rc, out, err = run_cli(args)
            orc, oout, oerr = oracle_lib.run_cli(args)
            assert (rc, out) == (orc, oout), f"seed {seed} {extra}: rc {rc} vs {orc}\n GPU: {out[:300]}\n REF: {oout[:300]}"
            aligned += orc == 0
    assert aligned >= 40
