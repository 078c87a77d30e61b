"""Mode 2 on graphs with LONG skip edges (deletion / SV edges that jump thousands of rows).

The first-column seed of gap_global_abpoa.rs:78-92 is `o + e * (best_p + 1)` where best_p is the smallest predecessor's
ROW INDEX: on a segment start reached by a skip edge of K rows it sits |e| * (K - 1) above the trend of the rest of the
row, and drops back on the next row. The packed 16-bit rows of k_gap_global_blk must leave the packed form for such
rows (the seed does not fit their window); these tests compare them with the oracle and with the 32-bit-only build
of the same kernel (RG_NO_S16)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def skip_graph(rows_skipped, seed, tail=120, head=60, seg=64):
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rs(k):
        return bases[rng.integers(0, 4, size=k)].tobytes().decode()

    segs = [rs(head)]
    made = 0
    while made < rows_skipped:
        segs.append(rs(seg))
        made += seg
    segs.append(rs(tail))
    lines = ["H\tVN:Z:1.0"]
    n = len(segs)
    for i, s in enumerate(segs, 1):
        lines.append(f"S\t{i}\t{s}")
    for i in range(1, n):
        lines.append(f"L\t{i}\t+\t{i + 1}\t+\t0M")
    lines.append(f"L\t1\t+\t{n}\t+\t0M")  # the skip edge: row of segment n's first base has min predecessor = end of segment 1
    lines.append("P\tfull\t" + ",".join(f"{i}+" for i in range(1, n + 1)) + "\t*")
    lines.append(f"P\tdel\t1+,{n}+\t*")
    reads = []
    for k in range(6):
        src = segs[0] + segs[-1] if k % 2 == 0 else segs[-3] + segs[-2] + segs[-1]
        a = list(src[: 150])
        for _ in range(5):
            a[int(rng.integers(0, len(a)))] = "ACGT"[int(rng.integers(0, 4))]
        reads.append("".join(a))
    return "\n".join(lines) + "\n", reads


@pytest.mark.parametrize("rows,extra", [(17000, []), (30000, []), (1300, ["-O", "30", "-E", "30", "-M", "30", "-X", "30"]),
                                        (6000, ["-O", "4", "-E", "8"]), (20000, ["-b", "300"])])
def test_mode2_long_skip_edge(tmp_path, rows, extra):
    from recgraph_b200 import run_cli
    from tests import oracle_lib
    gfa_text, reads = skip_graph(rows, seed=rows)
    gfa, fa = tmp_path / "g.gfa", tmp_path / "r.fa"
    gfa.write_text(gfa_text)
    fa.write_text("".join(f">r{i}\n{r}\n" for i, r in enumerate(reads)))
    args = ["-m", "2"] + extra + [str(fa), str(gfa)]
    rc, out, err = run_cli(args)
    assert rc == 0, err
    orc, oout, oerr = oracle_lib.run_cli(args)
    assert orc == 0, oerr
    assert out == oout
    os.environ["RG_NO_S16"] = "1"
    try:
        rc2, out2, _ = run_cli(args)
    finally:
        del os.environ["RG_NO_S16"]
    assert rc2 == 0 and out2 == out
