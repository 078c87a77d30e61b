"""Double pin of the parts of the path the reference's own tests do not cover: the C++ oracle (oracle/*.cpp, the
checker of the CUDA path) against a SECOND, independent restatement written in Python from the Rust sources only
(oracle/pyref/recgraph_pyref.py) — pathwise DP of modes 4 / 5, `align` / `rev_align` / `absolute_scores` /
`best_alignment` of modes 8 / 9, `build_alignment`, the four `gaf_output_*` builders, path-length helpers and GAF text.
Byte-identical stdout on the example and on 220 random small graphs (single source and sink); the same for the
headline mode 2 and for modes 0 / 1 / 3 (the AVX2 routines the reference's CLI runs, with their GAF builders) on 100 random
graphs each with random band / gap / score flags, for the -s true flows of modes 0-3, and for the experimental affine
pathwise modes 6 / 7: every mode of the reference is read twice."""
import importlib.util
import os

import numpy as np
import pytest

from recgraph_b200 import synth
from tests import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example")
_spec = importlib.util.spec_from_file_location("recgraph_pyref", os.path.join(ROOT, "oracle", "pyref", "recgraph_pyref.py"))
pyref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(pyref)


def _oracle(mode, fa, gfa, extra=()):
    rc, out, err = oracle_lib.run_cli(["-m", str(mode)] + list(extra) + [fa, gfa])
    assert rc == 0, err
    return out


@pytest.mark.parametrize("mode", [4, 5, 8, 9])
def test_example_first_reads(mode, tmp_path):
    fa_text = open(os.path.join(EXAMPLE, "reads.fa")).read()
    gfa_text = open(os.path.join(EXAMPLE, "graph.gfa")).read()
    nreads = 2
    lines = fa_text.splitlines()
    fa = tmp_path / "r.fa"
    fa.write_text("\n".join(lines[:2 * nreads]) + "\n")
    got = pyref.run(mode, fa.read_text(), gfa_text)
    assert got == _oracle(mode, str(fa), os.path.join(EXAMPLE, "graph.gfa"))


def _case(seed):
    rng = np.random.default_rng(seed)
    bp = int(rng.integers(40, 130))
    paths = int(rng.integers(2, 7))
    g = synth.make_graph(bp, paths, seed=seed, mean_seg=int(rng.integers(3, 10)), p_snp=0.3, p_indel=0.15)
    rlen = int(rng.integers(8, 40))
    reads = synth.make_reads(g, 2, rlen, err=float(rng.choice([0.0, 0.05, 0.15])), seed=seed + 1, mosaic_breaks=int(rng.integers(0, 3)))
    return g, reads


@pytest.mark.parametrize("block", range(11))
def test_random_small_graphs(block, tmp_path):
    """20 graphs per block x modes 4, 5, 8, 9 (+ a second scoring for the recombination modes)."""
    for seed in range(1000 + 20 * block, 1020 + 20 * block):
        g, reads = _case(seed)
        gfa, fa = tmp_path / f"g{seed}.gfa", tmp_path / f"r{seed}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        for mode in (4, 5, 8, 9):
            got = pyref.run(mode, fa.read_text(), gfa.read_text())
            exp = _oracle(mode, str(fa), str(gfa))
            assert got == exp, f"seed {seed} mode {mode}:\n PY : {got[:500]}\n C++: {exp[:500]}"
        if seed % 4 == 0:
            for mode in (8, 9):
                got = pyref.run(mode, fa.read_text(), gfa.read_text(), match=1, mismatch=3, base_rec_cost=1, multi_rec_cost=0.05,
                                rec_band_width=0.7)
                exp = _oracle(mode, str(fa), str(gfa), ["-M", "1", "-X", "3", "-R", "1", "-r", "0.05", "-B", "0.7"])
                assert got == exp, f"seed {seed} mode {mode} (R=1 r=0.05 B=0.7)"


# ---------------------------------------------------------------------------------------------------------------- mode 2
def test_mode2_example(tmp_path):
    """gap_global_abpoa::exec + gaf_of_gap_abpoa + band_ampl_enough (the headline mode): example reads, default band and -b 50"""
    fa_text = open(os.path.join(EXAMPLE, "reads.fa")).read()
    gfa_text = open(os.path.join(EXAMPLE, "graph.gfa")).read()
    fa = tmp_path / "r.fa"
    fa.write_text("\n".join(fa_text.splitlines()[:12]) + "\n")
    for kw, extra in (({}, []), ({"extra_b": 50}, ["-b", "50"]), ({"gap_open": 10, "gap_ext": 1}, ["-O", "10", "-E", "1"])):
        got = pyref.run_mode2(fa.read_text(), gfa_text, **kw)
        assert got == _oracle(2, str(fa), os.path.join(EXAMPLE, "graph.gfa"), extra), extra


@pytest.mark.parametrize("block", range(4))
def test_mode2_random_small_graphs(block, tmp_path):
    for seed in range(3000 + 25 * block, 3025 + 25 * block):
        rng = np.random.default_rng(seed)
        g = synth.make_graph(int(rng.integers(60, 400)), 3, seed=seed, mean_seg=int(rng.integers(3, 12)), p_snp=0.25, p_indel=0.15)
        reads = synth.make_reads(g, 2, int(rng.integers(10, 120)), err=float(rng.choice([0.0, 0.05, 0.2])), seed=seed + 1)
        gfa, fa = tmp_path / f"g{seed}.gfa", tmp_path / f"r{seed}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        b, f = int(rng.integers(0, 6)), float(rng.choice([0.0, 0.01, 0.1, 0.5]))
        rc, exp, err = oracle_lib.run_cli(["-m", "2", "-b", str(b), "-f", str(f), str(fa), str(gfa)])
        try:
            got = pyref.run_mode2(fa.read_text(), gfa.read_text(), extra_b=b, extra_f=f)
        except (RuntimeError, IndexError) as ex:   # inputs on which the reference panics (an empty row that gets indexed, the 'u' trace code)
            assert rc == 101, f"seed {seed}: pyref says the reference panics ({ex}), the oracle exits with {rc}"
            continue
        assert rc == 0, err
        assert got == exp, f"seed {seed} -b {b} -f {f}:\n PY : {got[:400]}\n C++: {exp[:400]}"


# ------------------------------------------------------------------------------------------------- modes 0 / 1 / 3
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_poa_modes_example(mode, tmp_path):
    """exec_simd of modes 0 / 1 (AVX2 semantics, f32 path values decoded through their decimal text), gap_local_poa::exec,
    and their GAF builders: first example reads, default and one alternative parameter set per mode."""
    fa_text = open(os.path.join(EXAMPLE, "reads.fa")).read()
    gfa_text = open(os.path.join(EXAMPLE, "graph.gfa")).read()
    fa = tmp_path / "r.fa"
    fa.write_text("\n".join(fa_text.splitlines()[:8]) + "\n")
    alt = {0: ({"extra_b": 50}, ["-b", "50"]), 1: ({"match": 3, "mismatch": 2}, ["-M", "3", "-X", "2"]),
           3: ({"gap_open": 10, "gap_ext": 1}, ["-O", "10", "-E", "1"])}[mode]
    for kw, extra in (({}, []), alt):
        got = pyref.run_poa(mode, fa.read_text(), gfa_text, **kw)
        assert got == _oracle(mode, str(fa), os.path.join(EXAMPLE, "graph.gfa"), extra), (mode, extra)


@pytest.mark.parametrize("block", range(4))
def test_poa_modes_random_small_graphs(block, tmp_path):
    """25 graphs per block x modes 0 (random -b / -f), 1 and 3 (random -O / -E), random match / mismatch scores."""
    for seed in range(4000 + 25 * block, 4025 + 25 * block):
        rng = np.random.default_rng(seed)
        g = synth.make_graph(int(rng.integers(60, 300)), 3, seed=seed, mean_seg=int(rng.integers(3, 12)), p_snp=0.25, p_indel=0.15)
        reads = synth.make_reads(g, 2, int(rng.integers(10, 100)), err=float(rng.choice([0.0, 0.05, 0.2])), seed=seed + 1)
        gfa, fa = tmp_path / f"g{seed}.gfa", tmp_path / f"r{seed}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        M, X = int(rng.choice([1, 2, 5])), int(rng.choice([1, 4, 7]))
        b, f = int(rng.choice([1, 2, 5, 30])), float(rng.choice([0.0, 0.01, 0.1, 0.5]))
        O, E = int(rng.choice([0, 1, 4, 10])), int(rng.choice([1, 2, 5]))
        for mode, kw, extra in ((0, {"extra_b": b, "extra_f": f}, ["-b", str(b), "-f", str(f)]), (1, {}, []),
                                (3, {"gap_open": O, "gap_ext": E}, ["-O", str(O), "-E", str(E)])):
            rc, exp, err = oracle_lib.run_cli(["-m", str(mode), "-M", str(M), "-X", str(X)] + extra + [str(fa), str(gfa)])
            try:
                got = pyref.run_poa(mode, fa.read_text(), gfa.read_text(), match=M, mismatch=X, **kw)
            except (RuntimeError, IndexError, KeyError) as ex:   # inputs on which the reference panics
                assert rc == 101, f"seed {seed} mode {mode}: pyref says the reference panics ({ex!r}), the oracle exits with {rc}"
                continue
            assert rc == 0, f"seed {seed} mode {mode}: {err}"
            assert got == exp, f"seed {seed} mode {mode} -M {M} -X {X} {extra}:\n PY : {got[:400]}\n C++: {exp[:400]}"


# ------------------------------------------------------------------------------------------------- -s true (modes 0-3)
def _amb_case(seed, tmp_path):
    rng = np.random.default_rng(seed)
    g = synth.make_graph(int(rng.integers(60, 300)), 3, seed=seed, mean_seg=int(rng.integers(3, 12)), p_snp=0.25, p_indel=0.15)
    reads = synth.make_reads(g, 3, int(rng.integers(10, 100)), err=float(rng.choice([0.0, 0.05, 0.2])), seed=seed + 1)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    reads[1] = "".join(comp[c] for c in reversed(reads[1]))      # one read from the other strand
    gfa, fa = tmp_path / f"g{seed}.gfa", tmp_path / f"r{seed}.fa"
    gfa.write_text(g.gfa())
    fa.write_text(synth.fasta(reads))
    return fa, gfa, rng


@pytest.mark.parametrize("block", range(3))
def test_ambiguous_strand_random_small_graphs(block, tmp_path):
    """-s true: rev_and_compl, the scalar global_abpoa::exec + gaf_of_global_abpoa of mode 0's retry, the reversed handle
    map, strand '-', and each mode's selection rule (mode 1 keeps the lower score, mode 3 keeps strand '+')."""
    for seed in range(6000 + 20 * block, 6020 + 20 * block):
        fa, gfa, rng = _amb_case(seed, tmp_path)
        b, f = int(rng.choice([1, 5, 30, 200])), float(rng.choice([0.0, 0.01, 0.3]))
        O, E = int(rng.choice([0, 4, 10])), int(rng.choice([1, 2, 5]))
        for mode in (0, 1, 2, 3):
            extra = ["-s", "true"]
            kw = {}
            if mode in (0, 2):
                extra += ["-b", str(b), "-f", str(f)]
                kw.update(extra_b=b, extra_f=f)
            if mode in (2, 3):
                extra += ["-O", str(O), "-E", str(E)]
                kw.update(gap_open=O, gap_ext=E)
            rc, exp, err = oracle_lib.run_cli(["-m", str(mode)] + extra + [str(fa), str(gfa)])
            try:
                got = pyref.run_poa_amb(mode, fa.read_text(), gfa.read_text(), **kw)
            except (RuntimeError, IndexError, KeyError) as ex:
                assert rc == 101, f"seed {seed} mode {mode}: pyref says the reference panics ({ex!r}), the oracle exits with {rc}"
                continue
            assert rc == 0, f"seed {seed} mode {mode}: {err}"
            assert got == exp, f"seed {seed} mode {mode} {extra}:\n PY : {got[:500]}\n C++: {exp[:500]}"


# ------------------------------------------------------------------------------------------------- modes 6 / 7
@pytest.mark.parametrize("mode", [6, 7])
def test_gap_pathwise_example(mode, tmp_path):
    fa_text = open(os.path.join(EXAMPLE, "reads.fa")).read()
    gfa_text = open(os.path.join(EXAMPLE, "graph.gfa")).read()
    fa = tmp_path / "r.fa"
    fa.write_text("\n".join(fa_text.splitlines()[:4]) + "\n")
    for kw, extra in (({}, []), ({"gap_open": 10, "gap_ext": 1}, ["-O", "10", "-E", "1"])):
        got = pyref.run_gap_pathwise(mode, fa.read_text(), gfa_text, **kw)
        assert got == _oracle(mode, str(fa), os.path.join(EXAMPLE, "graph.gfa"), extra), (mode, extra)


@pytest.mark.parametrize("block", range(6))
def test_gap_pathwise_random_small_graphs(block, tmp_path):
    """The experimental affine pathwise modes: the delta-encoded tensors of pathwise_alignment_gap(_semi).rs with their
    `alphas[i]` / `alphas[p]` inconsistency, raw-entry gap tests of the builders, every-slot best_ending_node; 20 graphs
    per block x modes 6, 7, random -O / -E."""
    for seed in range(1000 + 20 * block, 1020 + 20 * block):
        g, reads = _case(seed)
        gfa, fa = tmp_path / f"g{seed}.gfa", tmp_path / f"r{seed}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        rng = np.random.default_rng(seed + 77)
        O, E = int(rng.choice([0, 1, 4, 10])), int(rng.choice([1, 2, 5]))
        for mode in (6, 7):
            rc, exp, err = oracle_lib.run_cli(["-m", str(mode), "-O", str(O), "-E", str(E), str(fa), str(gfa)])
            try:
                got = pyref.run_gap_pathwise(mode, fa.read_text(), gfa.read_text(), gap_open=O, gap_ext=E)
            except (RuntimeError, IndexError, KeyError) as ex:
                assert rc == 101, f"seed {seed} mode {mode}: pyref says the reference panics / hangs ({ex!r}), the oracle exits with {rc}"
                continue
            if rc == 101:      # the oracle stops at the first read the reference dies on; everything before must agree
                assert got.startswith(exp), f"seed {seed} mode {mode}: oracle died, pyref did not:\n PY : {got[:300]}\n C++: {exp[:300]}"
                raise AssertionError(f"seed {seed} mode {mode}: the oracle reports a reference panic, pyref completes")
            assert got == exp, f"seed {seed} mode {mode} -O {O} -E {E}:\n PY : {got[:400]}\n C++: {exp[:400]}"


# ------------------------------------------------------------------------------------------------- -t HOXD70 / HOXD55
def test_matrix_files_random_small_graphs(tmp_path):
    """score_matrix.rs:67-105 and every routine's (row, column) key order under a NON-symmetric matrix (HOXD70: (G, T) = -114,
    (T, G) = -144): 12 graphs x both matrices x modes 0-5, 7, 9."""
    sm = pyref.score_matrix_from_matrix_file("HOXD70")
    assert sm[("A", "A")] == 91 and sm[("T", "G")] == -144 and ("-", "-") not in sm           # score_matrix.rs:118-131
    assert pyref.score_matrix_from_matrix_file("HOXD55")[("T", "G")] == -90
    for seed in range(8000, 8012):
        g, reads = _case(seed)
        gfa, fa = tmp_path / f"g{seed}.gfa", tmp_path / f"r{seed}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        F, G = fa.read_text(), gfa.read_text()
        for mat in ("HOXD70", "HOXD55"):
            runs = [(0, lambda: pyref.run_poa(0, F, G, matrix=mat)), (1, lambda: pyref.run_poa(1, F, G, matrix=mat)),
                    (2, lambda: pyref.run_mode2(F, G, matrix=mat)), (3, lambda: pyref.run_poa(3, F, G, matrix=mat)),
                    (4, lambda: pyref.run(4, F, G, matrix=mat)), (5, lambda: pyref.run(5, F, G, matrix=mat)),
                    (7, lambda: pyref.run_gap_pathwise(7, F, G, matrix=mat)), (9, lambda: pyref.run(9, F, G, matrix=mat))]
            for mode, fn in runs:
                rc, exp, err = oracle_lib.run_cli(["-m", str(mode), "-t", mat, str(fa), str(gfa)])
                try:
                    got = fn()
                except (RuntimeError, IndexError, KeyError) as ex:
                    assert rc == 101, f"seed {seed} mode {mode} {mat}: pyref says the reference panics ({ex!r}), the oracle exits with {rc}"
                    continue
                assert rc == 0, f"seed {seed} mode {mode} {mat}: {err}"
                assert got == exp, f"seed {seed} mode {mode} -t {mat}:\n PY : {got[:400]}\n C++: {exp[:400]}"
