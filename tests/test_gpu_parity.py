"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle — byte-exact GAF."""
import os

import pytest

from recgraph_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example")

IMPLEMENTED = {0, 1, 2, 3}


def _both(args):
    from recgraph_b200 import run_cli
    from tests import oracle_lib
    rc, out, err = run_cli(args)
    orc, oout, oerr = oracle_lib.run_cli(args)
    return (rc, out, err), (orc, oout, oerr)


def _assert_same(args):
    (rc, out, err), (orc, oout, oerr) = _both(args)
    assert orc == 0, oerr
    assert rc == 0, err
    if out != oout:
        a, b = out.splitlines(), oout.splitlines()
        for k, (x, y) in enumerate(zip(a, b)):
            if x != y:
                raise AssertionError(f"{args}: first difference at line {k}:\n GPU: {x[:600]}\n REF: {y[:600]}")
        raise AssertionError(f"{args}: line count differs: {len(a)} vs {len(b)}")


@pytest.fixture(scope="module")
def synth_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    out = {}
    for name, (bp, paths, nreads, rlen, err, seed) in {
        "small": (1200, 5, 24, 150, 0.05, 21),
        "mid": (6000, 8, 16, 700, 0.05, 22),
        "short_reads": (800, 4, 20, 31, 0.1, 23),
        "long_read": (900, 3, 6, 1400, 0.03, 24),
    }.items():
        g = synth.make_graph(bp, paths, seed=seed)
        reads = synth.make_reads(g, nreads, rlen, err=err, seed=seed + 100)
        gfa, fa = d / f"{name}.gfa", d / f"{name}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        out[name] = (str(fa), str(gfa))
    return out


@pytest.mark.parametrize("extra", [[], ["-b", "50"], ["-b", "1000"], ["-b", "0", "-f", "0.2"], ["-O", "10", "-E", "1"],
                                   ["-O", "0", "-E", "3"], ["-M", "1", "-X", "1", "-O", "2", "-E", "1"],
                                   ["-t", "HOXD70", "-O", "400", "-E", "30"], ["-t", "HOXD55", "-O", "100", "-E", "20"]])
def test_mode2_example(extra):
    _assert_same(["-m", "2"] + extra + [os.path.join(EXAMPLE, "reads.fa"), os.path.join(EXAMPLE, "graph.gfa")])


@pytest.mark.parametrize("name", ["small", "mid", "short_reads", "long_read"])
@pytest.mark.parametrize("extra", [[], ["-b", "5", "-f", "0.1"], ["-b", "3000"], ["-O", "6", "-E", "1"]])
def test_mode2_synthetic(synth_files, name, extra):
    fa, gfa = synth_files[name]
    _assert_same(["-m", "2"] + extra + [fa, gfa])


def test_mode2_reference_unit_vectors():
    """The reference's inline tests (gap_global_abpoa.rs:465-756) through rg_set_lnz_graph on the GPU."""
    from recgraph_b200 import Aligner
    from tests.test_oracle_golden import POA_CASES
    al = Aligner()
    for variant, (lnz, nwp, preds), read, scores, o, e, bta, expected, cite in POA_CASES:
        if variant != 2:
            continue
        table = [[0] * 6 for _ in range(6)]
        idx = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 5}
        for (a, b), v in scores.items():
            table[idx[a]][idx[b]] = v
        al.set_lnz_graph(lnz, nwp, preds)
        al.set_scoring(table=table, gap_open=-o, gap_ext=-e, fixed_bta=bta)
        recs, _ = al.align(2, [read[1:]])
        assert recs[0].status & 4 == 0, cite
        assert recs[0].score == expected, cite


def test_empty_batch_and_errors():
    from recgraph_b200 import Aligner, RecGraphError
    import numpy as np
    al = Aligner()
    with pytest.raises(RecGraphError):
        al.align(2, ["ACGT"])  # no graph yet
    al.load_gfa(os.path.join(EXAMPLE, "graph.gfa"))
    res = al.align_packed(2, np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert res.n_reads == 0
    with pytest.raises(RecGraphError):
        al.align(2, ["ACGTXX"])


EX = [os.path.join(EXAMPLE, "reads.fa"), os.path.join(EXAMPLE, "graph.gfa")]


@pytest.mark.parametrize("extra", [[], ["-b", "50"], ["-b", "1000"], ["-b", "20", "-f", "0.3"], ["-M", "1", "-X", "1"],
                                   ["-b", "200", "-t", "HOXD70"], ["-b", "300", "-t", "HOXD55"], ["-b", "7"]])
def test_mode0_example(extra):
    """BASELINE config 1: `-m 0` on example/ (AVX2 semantics). Default flags give the degenerate
    "band not enough" records; larger -b gives real alignments."""
    _assert_same(["-m", "0"] + extra + EX)


@pytest.mark.parametrize("extra", [[], ["-M", "3", "-X", "2"], ["-t", "HOXD70"], ["-t", "HOXD55"], ["-M", "1", "-X", "5"]])
def test_mode1_example(extra):
    _assert_same(["-m", "1"] + extra + EX)


@pytest.mark.parametrize("extra", [[], ["-O", "10", "-E", "1"], ["-O", "0", "-E", "2"], ["-M", "1", "-X", "1", "-O", "1", "-E", "1"],
                                   ["-t", "HOXD70", "-O", "400", "-E", "30"]])
def test_mode3_example(extra):
    _assert_same(["-m", "3"] + extra + EX)


@pytest.mark.parametrize("mode", ["0", "1", "3"])
@pytest.mark.parametrize("name", ["small", "mid", "short_reads"])
@pytest.mark.parametrize("extra", [[], ["-b", "40", "-f", "0.1"]])
def test_modes013_synthetic(synth_files, mode, name, extra):
    if mode != "0" and extra:
        pytest.skip("band flags only matter for mode 0")
    fa, gfa = synth_files[name]
    _assert_same(["-m", mode] + extra + [fa, gfa])


def test_mode3_reference_unit_vectors():
    """gap_local_poa.rs:198-277 through rg_set_lnz_graph on the GPU."""
    from recgraph_b200 import Aligner
    from tests.test_oracle_golden import POA_CASES
    al = Aligner()
    for variant, (lnz, nwp, preds), read, scores, o, e, bta, expected, cite in POA_CASES:
        if variant != 3:
            continue
        table = [[0] * 6 for _ in range(6)]
        idx = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 5}
        for (a, b), v in scores.items():
            table[idx[a]][idx[b]] = v
        al.set_lnz_graph(lnz, nwp, preds)
        al.set_scoring(table=table, gap_open=-o, gap_ext=-e)
        recs, _ = al.align(3, [read[1:]])
        assert recs[0].score == expected, cite


@pytest.mark.parametrize("mode", ["4", "5"])
@pytest.mark.parametrize("extra", [[], ["-M", "1", "-X", "3"], ["-t", "HOXD70"]])
def test_pathwise_example(mode, extra):
    _assert_same(["-m", mode] + extra + EX)


@pytest.mark.parametrize("mode", ["4", "5"])
@pytest.mark.parametrize("name", ["small", "mid", "short_reads", "long_read"])
def test_pathwise_synthetic(synth_files, mode, name):
    fa, gfa = synth_files[name]
    _assert_same(["-m", mode, fa, gfa])


@pytest.mark.parametrize("mode", ["8", "9"])
@pytest.mark.parametrize("extra", [[], ["-R", "1", "-r", "0.05"], ["-B", "0.6"], ["-R", "0", "-r", "0"], ["-M", "1", "-X", "3", "-R", "2"]])
def test_recombination_example(mode, extra):
    _assert_same(["-m", mode] + extra + EX)


@pytest.fixture(scope="module")
def mosaic_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("mosaic")
    out = {}
    for name, (bp, paths, nreads, rlen, err, seed, breaks) in {
        "m1": (900, 6, 16, 200, 0.02, 31, 2),
        "m2": (1500, 10, 12, 350, 0.03, 32, 1),
        "m3": (600, 4, 10, 80, 0.0, 33, 3),
    }.items():
        g = synth.make_graph(bp, paths, seed=seed)
        reads = synth.make_reads(g, nreads, rlen, err=err, seed=seed + 100, mosaic_breaks=breaks)
        gfa, fa = d / f"{name}.gfa", d / f"{name}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        out[name] = (str(fa), str(gfa))
    return out


@pytest.mark.parametrize("mode", ["8", "9"])
@pytest.mark.parametrize("name", ["m1", "m2", "m3"])
@pytest.mark.parametrize("extra", [[], ["-R", "2", "-r", "0.01"]])
def test_recombination_mosaics(mosaic_files, mode, name, extra):
    fa, gfa = mosaic_files[name]
    _assert_same(["-m", mode] + extra + [fa, gfa])


def test_api_rs_entry_points_match_cli():
    """api.rs:43-72,102-128: library defaults (bases_to_add = 0.1 * len, o = -10, e = -6) == CLI with those flags."""
    import recgraph_b200 as rb
    from tests import oracle_lib
    fa, gfa = EX
    names, seqs = [], []
    for ln in open(fa):
        if ln.startswith(">"):
            names.append(ln[1:].strip())
        else:
            seqs.append(ln.strip())
    al = rb.Aligner()
    al.load_gfa(gfa)
    bta = int(len(seqs[0]) * 0.1)
    rc, out, _ = oracle_lib.run_cli(["-m", "2", "-O", "10", "-E", "6", "-b", str(bta), "-f", "0", fa, gfa])
    exp = [l for l in out.splitlines() if "\t" in l]
    rc, out3, _ = oracle_lib.run_cli(["-m", "3", "-O", "10", "-E", "6", fa, gfa])
    exp3 = [l for l in out3.splitlines() if "\t" in l]
    for k in range(4):
        g = rb.align_global_gap(seqs[k], al, (names[k], k + 1))
        assert g.to_string() == exp[k]
        g3 = rb.align_local_gap(seqs[k], al, (names[k], k + 1))
        assert g3.to_string() == exp3[k]
        assert g3.path == [int(x) for x in exp3[k].split("\t")[5].split(">") if x]


# ---- -s true: reverse-complement retries (main.rs:82-101 mode 0 via the scalar routine, 150-164 mode 1, 198-214 mode 2,
# 233-249 mode 3), reversed handle map and strand in the GAF record
@pytest.fixture(scope="module")
def strand_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("strand")
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    out = {}
    for name, (bp, paths, nreads, rlen, err, seed) in {"mix": (1500, 5, 20, 180, 0.04, 41), "mix_long": (5000, 6, 12, 600, 0.05, 42)}.items():
        g = synth.make_graph(bp, paths, seed=seed)
        reads = synth.make_reads(g, nreads, rlen, err=err, seed=seed + 100)
        reads = [r if k % 2 == 0 else "".join(comp[c] for c in reversed(r)) for k, r in enumerate(reads)]
        gfa, fa = d / f"{name}.gfa", d / f"{name}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        out[name] = (str(fa), str(gfa))
    return out


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("name", ["mix", "mix_long"])
@pytest.mark.parametrize("extra", [[], ["-b", "40", "-f", "0.1"]])
def test_amb_strand_synthetic(strand_files, mode, name, extra):
    fa, gfa = strand_files[name]
    _assert_same(["-m", str(mode), "-s", "true"] + extra + [fa, gfa])


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_amb_strand_example(mode):
    _assert_same(["-m", str(mode), "-s", "true", "-b", "50", os.path.join(EXAMPLE, "reads.fa"), os.path.join(EXAMPLE, "graph.gfa")])


def test_scalar_mode0_reference_unit_vectors():
    """The reference's inline tests of the scalar routine (global_abpoa.rs:577-754) through RG_MODE_GLOBAL_SCALAR."""
    from recgraph_b200 import Aligner
    from tests.test_oracle_golden import POA_CASES
    al = Aligner()
    idx = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 5}
    n = 0
    for variant, (lnz, nwp, preds), read, scores, o, e, bta, expected, cite in POA_CASES:
        if variant != 0:
            continue
        table = [[-1] * 6 for _ in range(6)]  # one gap score for all characters (the vectors only define A / C)
        for (a, b), v in scores.items():
            table[idx[a]][idx[b]] = v
        table[5][5] = 0
        al.set_lnz_graph(lnz, nwp, preds)
        al.set_scoring(table=table, fixed_bta=bta)
        recs, _ = al.align(10, [read[1:]])
        assert recs[0].status & 4 == 0, cite
        assert recs[0].score == expected, cite
        n += 1
    assert n == 4


# Packed 16-bit rows of the mode-2 kernel under scores at the edge of their domain (|score| <= 30): fast drift of the row
# maximum (re-basing), wide score spread (the range guard leaves the packed form), cheap gaps (long L / U runs).
@pytest.mark.parametrize("name", ["small", "mid"])
@pytest.mark.parametrize("extra", [["-M", "30", "-X", "30", "-O", "30", "-E", "30", "-b", "3000"],
                                   ["-M", "10", "-X", "20", "-O", "25", "-E", "5", "-b", "3000"],
                                   ["-M", "30", "-X", "1", "-O", "1", "-E", "1", "-b", "3000"],
                                   ["-M", "1", "-X", "30", "-O", "0", "-E", "1", "-b", "3000"],
                                   ["-M", "31", "-X", "4", "-O", "4", "-E", "2", "-b", "3000"]])
def test_mode2_packed_rows_score_extremes(synth_files, name, extra):
    fa, gfa = synth_files[name]
    _assert_same(["-m", "2"] + extra + [fa, gfa])


def test_mode2_packed_blocked32_and_striped_kernels_agree():
    """Three independent implementations of mode 2 (packed 16-bit rows, 32-bit blocked rows, striped kernel) produce the
    same records and run lists on the example, synthetic sets and read lengths at every lane-width boundary."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ab_s16", os.path.join(ROOT, "tools", "ab_s16.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main() == 0


@pytest.mark.parametrize("name", ["m0_b50", "m1", "m2_default", "m2_b50", "m2_s_true_b50", "m0_s_true_b50", "m3", "m4", "m5",
                                  "m8", "m9"])
def test_example_against_committed_fixtures(name):
    """The GPU command line against the golden GAF files under tests/golden/example/expected (written by
    tools/make_golden.py from the oracle): byte-identical stdout."""
    import importlib.util
    from recgraph_b200 import run_cli
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rc, out, err = run_cli(mg.CASES[name] + EX)
    assert rc == 0, err
    assert out == open(os.path.join(EXAMPLE, "expected", name + ".gaf")).read()
