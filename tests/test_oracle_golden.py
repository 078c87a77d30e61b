"""Pins the CPU oracle against every known-answer test the reference ships.

Each case transcribes one `#[test]` of /root/reference/src (file:line cited per case): same graph, read,
score matrix, parameters and expected value. These are the reference's only golden vectors for the path.
"""
import re

import os

import pytest

from tests import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

AA = {("A", "A"): 1, ("A", "-"): -1, ("-", "A"): -1}
AC = dict(AA)
AC.update({("C", "C"): 1, ("-", "C"): -1, ("C", "-"): -1, ("C", "A"): -1, ("A", "C"): -1})
GAP_AC = {("A", "A"): 1, ("C", "C"): 1, ("C", "A"): -1, ("A", "C"): -1}

G_TEST1 = (list("$AAAAF"), {1, 5}, {1: [0], 5: [4]})
G_TEST2 = (list("$AACAAAF"), {1, 3, 4, 5, 7}, {1: [0], 3: [2], 4: [2], 5: [3, 4], 7: [6]})
G_STARTS = (list("$ACACCAAF"), {1, 2, 3, 4, 5, 8}, {1: [0], 2: [0], 3: [1, 2], 4: [1, 2], 5: [3, 4], 8: [7]})
G_ENDS = (list("$ACACCAACF"), {1, 2, 3, 4, 5, 7, 8, 9},
          {1: [0], 2: [0], 3: [1, 2], 4: [1, 2], 5: [3, 4], 7: [6], 8: [6], 9: [7, 8]})

# (variant, graph, read, scores, o, e, bta, expected, citation)
POA_CASES = [
    (0, G_TEST1, "$AAAA", AA, 0, 0, 100, 4, "global_abpoa.rs:577-610 test1"),
    (0, G_TEST2, "$AACAA", AC, 0, 0, 4, 5, "global_abpoa.rs:612-655 test2"),
    (0, G_STARTS, "$CACAA", AC, 0, 0, 4, 5, "global_abpoa.rs:657-702 multiple_starts"),
    (0, G_ENDS, "$CACAA", AC, 0, 0, 4, 5, "global_abpoa.rs:705-754 multiple_ends"),
    (2, G_TEST1, "$AAAA", {("A", "A"): 1}, -4, -1, 3, 4, "gap_global_abpoa.rs:465-498 test1"),
    (2, G_TEST2, "$AACAAC", GAP_AC, -4, -1, 3, 0, "gap_global_abpoa.rs:501-543 gap_correctly_considered"),
    (2, G_STARTS, "$CACAA", GAP_AC, -4, -1, 3, 5, "gap_global_abpoa.rs:546-589 multiple_starts"),
    (2, G_ENDS, "$CACAA", GAP_AC, -4, -1, 3, 5, "gap_global_abpoa.rs:592-639 multiple_ends"),
    (2, G_TEST2, "$AACAAC", GAP_AC, 0, -1, 5, 4, "gap_global_abpoa.rs:642-683 same_result_as_normal_if_o_0"),
    (2, (list("$ACACAF"), {1, 6}, {1: [0], 6: [5]}), "$AAA", GAP_AC, -100, -1, 10, -101,
     "gap_global_abpoa.rs:685-720 gap_open_only_once_if_penalty_high"),
    (2, (list("$AAAAAF"), {1, 6}, {1: [0], 6: [5]}), "$AAAAAAAAA", GAP_AC, -4, -1, 7, -3,
     "gap_global_abpoa.rs:722-756 sequence_longer_than_graph"),
]
LOC = {(a, b): (1 if a == b else -1) for a in "ACG-" for b in "ACG-"}
LOC3 = {(a, b): (1 if a == b else -1) for a in "ACG" for b in "ACG"}
G_LOC1 = (list("$GGCCCGGF"), {1, 8}, {1: [0], 8: [7]})
G_LOC2 = (list("$GGGCCCGGF"), {1, 6, 9}, {1: [0], 6: [3], 9: [8, 5]})
POA_CASES += [
    (1, G_LOC1, "$AACCCAA", LOC, 0, 0, 0, 3, "local_poa.rs:304-338 consider_substrings"),
    (1, G_LOC2, "$AACCCAA", LOC, 0, 0, 0, 2, "local_poa.rs:341-377 consider_best_predecessor"),
    (3, G_LOC1, "$AACCCAA", LOC3, -4, -2, 0, 3, "gap_local_poa.rs:198-235 consider_substrings"),
    (3, G_LOC2, "$AACCCAA", LOC3, -4, -2, 0, 2, "gap_local_poa.rs:238-277 consider_best_predecessor"),
]


@pytest.mark.parametrize("case", POA_CASES, ids=[c[-1] for c in POA_CASES])
def test_reference_poa_score_vectors(case):
    variant, (lnz, nwp, preds), read, scores, o, e, bta, expected, _ = case
    rc, score, _cells = O.poa_score(variant, lnz, nwp, preds, list(read), scores, o, e, bta)
    assert rc == 0
    assert score == expected


def _gfa(segs, edges, paths=()):
    lines = ["H\tVN:Z:1.0"]
    for i, s in enumerate(segs, 1):
        lines.append(f"S\t{i}\t{s}")
    for a, b in edges:
        lines.append(f"L\t{a}\t+\t{b}\t+\t0M")
    for k, p in enumerate(paths):
        lines.append(f"P\tp{k}\t" + ",".join(f"{x}+" for x in p) + "\t*")
    return "\n".join(lines) + "\n"


def _parse_dump(txt):
    assert not txt.startswith("PANIC"), txt
    d = {"pred": {}, "hofp": {}, "node": {}, "edge": {}}
    for ln in txt.splitlines():
        if ln.startswith("pred ") and ":" in ln:
            k, v = ln[5:].split(":")
            d["pred"][int(k)] = [int(x) for x in v.split()]
        elif ln.startswith("pred "):
            _, i, p, bits = ln.split()
            d["edge"].setdefault(int(i), {})[int(p)] = bits
        elif ln.startswith("hofp "):
            k, v = ln[5:].split(":")
            d["hofp"][int(k)] = v.strip()
        elif ln.startswith("node "):
            m = re.match(r"node (\d+): id=(\d+) alpha=(\d+) paths=([01]*)", ln)
            d["node"][int(m.group(1))] = (int(m.group(2)), int(m.group(3)), m.group(4))
        elif "=" in ln:
            k, v = ln.split("=", 1)
            d[k] = v
    return d


def test_graph_struct_correctly_created():
    """graph.rs:193-210"""
    d = _parse_dump(O.dump_lnz(_gfa(["A", "T", "C", "G"], [(1, 2), (2, 3), (3, 4)])))
    assert d["nwp"][1] == "1" and d["nwp"][5] == "1"
    assert d["pred"][1][0] == 0 and d["pred"][5][0] == 4
    assert d["lnz"] == "$ATCGF"


def test_rev_graph_struct_correctly_created():
    """graph.rs:212-229"""
    d = _parse_dump(O.dump_lnz(_gfa(["A", "T", "C", "G"], [(1, 2), (2, 3), (3, 4)]), amb_mode=True))
    assert d["nwp"][1] == "1" and d["nwp"][5] == "1"
    assert d["pred"][1][0] == 0 and d["pred"][5][0] == 4
    assert d["lnz"] == "$CGATF"


def test_handle_id_from_lnz_pos_and_sorted_handles():
    """graph.rs:231-259 (0-based handle counter there; segment ids here are 1-based)."""
    d = _parse_dump(O.dump_lnz(_gfa(["A", "TA", "CGG", "G", "TCCCC"], [(1, 2), (1, 3), (3, 4), (3, 5)])))
    assert [d["hofp"][i] for i in (1, 2, 4, 6, 7, 12)] == ["1", "2", "3", "3", "4", "5"]
    assert d["hofp"][0] == "-1"


PW_DIAMOND = (["A", "T", "C", "G"], [(1, 2), (1, 3), (2, 4), (3, 4)], [[1, 2, 4], [1, 3, 4]])
PW_MULTI = (["A", "B", "T", "C", "G", "H"], [(1, 3), (1, 4), (3, 5), (4, 5), (2, 6)],
            [[1, 3, 5], [1, 4, 5], [2, 6]])


def test_pathwise_graph_correctly_created():
    """pathwise_graph.rs:364-404"""
    d = _parse_dump(O.dump_pathgraph(_gfa(*PW_DIAMOND)))
    assert d["paths_number"] == "2"
    assert d["lnz"] == "$ATCGF"
    assert d["nwp"][2] == "1"
    assert d["node"][2][2] == "10"
    assert d["node"][0][2] == "11" and d["node"][5][2] == "11"


def test_multiple_starts_and_ends_pathwise():
    """pathwise_graph.rs:406-449"""
    d = _parse_dump(O.dump_pathgraph(_gfa(*PW_MULTI)))
    assert d["paths_number"] == "3"
    assert d["node"][3][2][:2] == "10"
    assert d["node"][0][2][:2] == "11" and d["node"][7][2][:2] == "11"


def test_reverse_pathwise_graph_correctly_created():
    """pathwise_graph.rs:452-495"""
    d = _parse_dump(O.dump_pathgraph(_gfa(*PW_DIAMOND), is_reversed=True))
    assert d["paths_number"] == "2"
    assert d["lnz"] == "$CGATF"
    assert d["nwp"][2] == "1"
    assert d["node"][2][2] == "01"
    assert d["node"][3][2] == "10"
    assert d["node"][0][2] == "11" and d["node"][5][2] == "11"


def test_pred_hash_struct():
    """pathwise_graph.rs:498-544"""
    d = _parse_dump(O.dump_pathgraph(_gfa(*PW_MULTI)))
    assert d["edge"][5] == {3: "100", 4: "010"}


def test_match_miss_matrix_correct():
    """score_matrix.rs:110-116"""
    import ctypes
    lib = O.load()
    pres = ctypes.c_int()
    assert lib.rgo_score_lookup(0, 10, -10, b"A", b"A", ctypes.byref(pres)) == 10 and pres.value
    assert lib.rgo_score_lookup(0, 10, -10, b"A", b"C", ctypes.byref(pres)) == -10 and pres.value
    assert lib.rgo_score_lookup(0, 10, -10, b"N", b"N", ctypes.byref(pres)) == -10 and pres.value
    lib.rgo_score_lookup(0, 10, -10, b"-", b"-", ctypes.byref(pres))
    assert not pres.value
    # main.rs:36 path: gap-vs-char is 2*X in the i32 builder, X in the api f32 builder (score_matrix.rs:42,58)
    assert lib.rgo_score_lookup(0, 2, -4, b"A", b"-", ctypes.byref(pres)) == -8
    assert lib.rgo_score_lookup(3, 2, -4, b"A", b"-", ctypes.byref(pres)) == -4


def test_hoxd_correct():
    """score_matrix.rs:118-130"""
    import ctypes
    lib = O.load()
    pres = ctypes.c_int()
    assert lib.rgo_score_lookup(1, 0, 0, b"A", b"A", ctypes.byref(pres)) == 91
    assert lib.rgo_score_lookup(1, 0, 0, b"T", b"G", ctypes.byref(pres)) == -144
    assert lib.rgo_score_lookup(2, 0, 0, b"A", b"A", ctypes.byref(pres)) == 91
    assert lib.rgo_score_lookup(2, 0, 0, b"T", b"G", ctypes.byref(pres)) == -90
    for which in (1, 2):
        lib.rgo_score_lookup(which, 0, 0, b"-", b"-", ctypes.byref(pres))
        assert not pres.value
        assert lib.rgo_score_lookup(which, 0, 0, b"A", b"-", ctypes.byref(pres)) == -200


def test_rev_and_compl():
    """sequences.rs:86-100"""
    assert O.rev_and_compl("$AAT") == "$ATT"
    assert O.rev_and_compl("$ATCGN") == "$NCGAT"


def test_band_half_width_is_f32():
    """main.rs:57: `(b + f * seq.len() as f32) as usize`; 0.01f32*100f32 rounds to 1.0; -b<0 saturates (F9)."""
    lib = O.load()
    assert lib.rgo_bases_to_add(1.0, 0.01, 100) == 2
    assert lib.rgo_bases_to_add(1.0, 0.01, 151) == 2
    assert lib.rgo_bases_to_add(1.0, 0.01, 1001) == 11
    assert lib.rgo_bases_to_add(-5.0, 0.01, 151) == 0


def test_f32_display_matches_rust():
    assert O.f32_display(186.0) == "186"
    assert O.f32_display(295.8) == "295.8"
    assert O.f32_display(-273.0) == "-273"
    assert O.f32_display(12.1) == "12.1"
    assert O.f32_display(128.4) == "128.4"


def test_simd_variants_agree_with_scalar_on_reference_vectors():
    """exec_simd has no reference test; on the tie-free vectors above it must reach the same optimum."""
    full = {(a, b): (1 if a == b else -1) for a in "ACGTN-" for b in "ACGTN-"}
    for variant, (lnz, nwp, preds), read, _s, _o, _e, _bta, expected, cite in POA_CASES[:4]:
        rc, score, _ = O.poa_score(10, lnz, nwp, preds, list(read), full, 0, 0, 100)
        assert rc == 0 and score == expected, cite


def test_affine_pathwise_modes_reduce_to_linear_when_gap_open_is_zero():
    """Modes 6 / 7 (pathwise_alignment_gap*.rs) ship no test in the reference. Self-consistency of the restatement:
    with o = 0 and e = the linear modes' gap score, the affine recurrences give the scores of modes 4 / 5 on the
    example; a negative o can only lower them; mode 7 reports the CIGAR of a 150-base read with 150 read-consuming ops."""
    import re
    gfa = open(os.path.join(ROOT, "tests", "golden", "example", "graph.gfa")).read()
    reads, cur = [], []
    for ln in open(os.path.join(ROOT, "tests", "golden", "example", "reads.fa")):
        if ln.startswith(">"):
            if cur:
                reads.append("".join(cur))
            cur = []
        else:
            cur.append(ln.strip().upper())
    reads.append("".join(cur))
    for r in reads[:6]:
        s4 = int(re.search(r"score: (-?\d+)", O.pathwise_one(gfa, r, 4)).group(1))
        s5 = int(re.search(r"score: (-?\d+)", O.pathwise_one(gfa, r, 5)).group(1))
        m6 = O.pathwise_one(gfa, r, 6, o=0, e=-8).split(" ")
        m7 = O.pathwise_one(gfa, r, 7, o=0, e=-8).split(" ")
        assert int(m6[0]) == s4 and int(m7[0]) == s5
        a6 = O.pathwise_one(gfa, r, 6, o=-4, e=-8).split(" ")
        a7 = O.pathwise_one(gfa, r, 7, o=-4, e=-2).split(" ")
        assert int(a6[0]) <= s4
        cig = a7[2].split("\t")[0]
        ops = re.findall(r"(\d+)([MXID])", cig)
        assert sum(int(n) for n, c in ops if c in "MXD") == len(r)  # I = graph-only steps (build_cigar, …_output.rs:471-556)


def test_oracle_reproduces_committed_fixtures():
    """tests/golden/example/expected/*.gaf were written by tools/make_golden.py (oracle output on the shipped example);
    any change of the restatement shows up here."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ex = os.path.join(ROOT, "tests", "golden", "example")
    for name, flags in mg.CASES.items():
        rc, out, err = O.run_cli(flags + [os.path.join(ex, "reads.fa"), os.path.join(ex, "graph.gfa")])
        assert rc == 0, (name, err)
        assert out == open(os.path.join(ex, "expected", name + ".gaf")).read(), name
