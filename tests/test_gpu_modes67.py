"""Modes 6 / 7 (affine-gap pathwise alignment, experimental in the reference: pathwise_alignment_gap.rs,
pathwise_alignment_gap_semi.rs, builders pathwise_alignment_output.rs:186-451) on the device against the oracle:
byte-identical stdout (CIGAR line + "Best path sequence i: p")."""
import os

import pytest

from recgraph_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example")
EX = [os.path.join(EXAMPLE, "reads.fa"), os.path.join(EXAMPLE, "graph.gfa")]


def _same(args):
    from recgraph_b200 import run_cli
    from tests import oracle_lib
    rc, out, err = run_cli(args)
    orc, oout, oerr = oracle_lib.run_cli(args)
    assert orc == 0, oerr
    assert rc == 0, err
    if out != oout:
        a, b = out.splitlines(), oout.splitlines()
        for k, (x, y) in enumerate(zip(a, b)):
            assert x == y, f"{args}: first difference at line {k}:\n GPU: {x[:400]}\n REF: {y[:400]}"
        raise AssertionError(f"{args}: line count differs: {len(a)} vs {len(b)}")


@pytest.mark.parametrize("mode", ["6", "7"])
@pytest.mark.parametrize("extra", [[], ["-O", "10", "-E", "1"], ["-O", "0", "-E", "2"], ["-M", "1", "-X", "3", "-O", "2", "-E", "2"],
                                   ["-t", "HOXD70", "-O", "400", "-E", "30"]])
def test_modes67_example(mode, extra):
    _same(["-m", mode] + extra + EX)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("m67")
    out = {}
    for name, (bp, paths, nreads, rlen, err, seed, breaks) in {
        "small": (1200, 5, 24, 150, 0.05, 21, 0),
        "mid": (4000, 8, 8, 500, 0.05, 22, 0),
        "short_reads": (800, 4, 20, 31, 0.1, 23, 0),
        "p40": (1500, 40, 8, 200, 0.04, 24, 1),
        "mosaic": (900, 6, 12, 200, 0.02, 31, 2),
    }.items():
        g = synth.make_graph(bp, paths, seed=seed)
        reads = synth.make_reads(g, nreads, rlen, err=err, seed=seed + 100, mosaic_breaks=breaks)
        gfa, fa = d / f"{name}.gfa", d / f"{name}.fa"
        gfa.write_text(g.gfa())
        fa.write_text(synth.fasta(reads))
        out[name] = (str(fa), str(gfa))
    return out


@pytest.mark.parametrize("mode", ["6", "7"])
@pytest.mark.parametrize("name", ["small", "mid", "short_reads", "p40", "mosaic"])
def test_modes67_synthetic(files, mode, name):
    fa, gfa = files[name]
    _same(["-m", mode, fa, gfa])


@pytest.mark.parametrize("name", ["m6", "m7"])
def test_modes67_committed_fixtures(name):
    import importlib.util
    from recgraph_b200 import run_cli
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rc, out, err = run_cli(mg.CASES[name] + EX)
    assert rc == 0, err
    assert out == open(os.path.join(EXAMPLE, "expected", name + ".gaf")).read()
