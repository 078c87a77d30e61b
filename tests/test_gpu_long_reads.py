"""Pathwise modes on reads of very different lengths in ONE batch: the score-transport kernel runs one launch per column-block
class (256 threads x 4 / 8 / 16 / 32 columns, and the 384-thread instance for reads above 8 191 bases), records and GAF
must come back in input order and equal the oracle's."""
import os

import numpy as np
import pytest

from recgraph_b200 import synth

pytestmark = pytest.mark.gpu


def _reads(g, lengths, seed):
    rng = np.random.default_rng(seed)
    out = []
    for k, ln in enumerate(lengths):
        src = g.path_sequence(g.paths[k % len(g.paths)]).decode()
        # longer than the graph: repeat the path sequence (the aligner must cope with reads longer than any path)
        s = (src * (ln // len(src) + 2))[int(rng.integers(0, len(src))):][:ln]
        a = list(s)
        for _ in range(max(1, ln // 40)):
            a[int(rng.integers(0, ln))] = "ACGT"[int(rng.integers(0, 4))]
        out.append("".join(a))
    return out


@pytest.mark.parametrize("mode", ["4", "5", "9"])
def test_mixed_read_lengths_pathwise(tmp_path, mode):
    from recgraph_b200 import run_cli
    from tests import oracle_lib
    g = synth.make_graph(420, 4, seed=77, mean_seg=8, p_snp=0.25, p_indel=0.1)
    lengths = [60, 9000, 300, 1100, 12000, 2500, 5000, 33, 8200, 1023, 1024, 4100]
    if mode == "9":
        lengths = [60, 9000, 300, 1100, 2500, 8200, 1024]
    reads = _reads(g, lengths, 5)
    gfa, fa = tmp_path / "g.gfa", tmp_path / "r.fa"
    gfa.write_text(g.gfa())
    fa.write_text(synth.fasta(reads))
    args = ["-m", mode, str(fa), str(gfa)]
    rc, out, err = run_cli(args)
    assert rc == 0, err
    orc, oout, oerr = oracle_lib.run_cli(args)
    assert orc == 0, oerr
    a, b = out.splitlines(), oout.splitlines()
    assert len(a) == len(b)
    for k, (x, y) in enumerate(zip(a, b)):
        assert x == y, f"read {k} (length {lengths[k]}): GPU {x[:200]} / REF {y[:200]}"
