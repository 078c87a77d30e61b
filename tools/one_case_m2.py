import sys, os, numpy as np, tempfile
sys.path.insert(0, os.getcwd())
from recgraph_b200 import synth, run_cli
from tests import oracle_lib
seed=int(sys.argv[1])
rng = np.random.default_rng(seed)
g = synth.make_graph(int(rng.integers(60, 400)), 3, seed=seed, mean_seg=int(rng.integers(3, 12)), p_snp=0.25, p_indel=0.15)
reads = synth.make_reads(g, 2, int(rng.integers(10, 120)), err=float(rng.choice([0.0, 0.05, 0.2])), seed=seed + 1)
d=tempfile.mkdtemp()
open(d+'/g.gfa','w').write(g.gfa()); open(d+'/r.fa','w').write(synth.fasta(reads))
b, f = int(rng.integers(0, 6)), float(rng.choice([0.0, 0.01, 0.1, 0.5]))
args=["-m","2","-b",str(b),"-f",str(f),d+'/r.fa',d+'/g.gfa']
print(args, [len(r) for r in reads])
rc,out,err=run_cli(args); orc,oout,oerr=oracle_lib.run_cli(args)
print("device rc",rc, "oracle rc", orc, "same out", out==oout)
print(err[-200:]); print(out[:300]); print(oout[:300])
