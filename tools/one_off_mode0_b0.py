"""mode 0 (and the -s retry's scalar routine) with a zero-width band amplitude, device vs oracle on random small graphs"""
import os, sys, tempfile, pathlib, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recgraph_b200 import synth, run_cli
from tests import oracle_lib
d = pathlib.Path(tempfile.mkdtemp())
bad = n = ok = 0
for seed in range(800000, 800000 + (int(sys.argv[1]) if len(sys.argv) > 1 else 120)):
    rng = np.random.default_rng(seed)
    g = synth.make_graph(int(rng.integers(60, 300)), 3, seed=seed, mean_seg=int(rng.integers(3, 12)), p_snp=0.25, p_indel=0.15)
    reads = synth.make_reads(g, 2, int(rng.integers(10, 100)), err=float(rng.choice([0.0, 0.05, 0.2])), seed=seed + 1)
    (d / "g.gfa").write_text(g.gfa()); (d / "r.fa").write_text(synth.fasta(reads))
    for extra in (["-b", "0", "-f", "0.0"], ["-b", "0", "-f", "0.01"], ["-b", "0", "-f", "0.0", "-s", "true"]):
        args = ["-m", "0"] + extra + [str(d / "r.fa"), str(d / "g.gfa")]
        rc, out, err = run_cli(args); orc, oout, oerr = oracle_lib.run_cli(args)
        n += 1; ok += orc == 0
        if rc != orc or out != oout:
            bad += 1
            if bad <= 5: print("DIFF", seed, extra, rc, orc, repr(out[:120]), repr(oout[:120]), err[-100:], oerr[-100:])
print("cases", n, "oracle-ok", ok, "differences", bad)
