import sys, os, importlib.util, tempfile, pathlib, numpy as np, time
sys.path.insert(0, os.getcwd())
from recgraph_b200 import synth
from tests import oracle_lib
spec=importlib.util.spec_from_file_location("pyref","oracle/pyref/recgraph_pyref.py"); pyref=importlib.util.module_from_spec(spec); spec.loader.exec_module(pyref)
d=pathlib.Path(tempfile.mkdtemp())
bad=0; n=0; t0=time.time()
def cmp(what, args, fn):
    global bad, n
    n+=1
    rc,exp,err=oracle_lib.run_cli(args)
    try:
        got=fn()
    except (RuntimeError, IndexError, KeyError, AssertionError) as ex:
        if rc!=101:
            bad+=1; print("MISMATCH(panic)", what, repr(ex)[:100], rc, flush=True)
        return
    if rc!=0:
        bad+=1; print("MISMATCH(rc)", what, rc, err[-120:], flush=True); return
    if got!=exp:
        bad+=1; print("MISMATCH(out)", what, flush=True)
base=int(sys.argv[1]); count=int(sys.argv[2])
for seed in range(base, base+count):
    rng=np.random.default_rng(seed)
    bp=int(rng.integers(40,200)); P=int(rng.integers(2,8))
    g=synth.make_graph(bp,P,seed=seed,mean_seg=int(rng.integers(3,12)),p_snp=0.3,p_indel=0.15)
    reads=synth.make_reads(g,2,int(rng.integers(8,70)),err=float(rng.choice([0.0,0.05,0.15])),seed=seed+1,mosaic_breaks=int(rng.integers(0,3)))
    gfa,fa=d/"g.gfa",d/"r.fa"; gfa.write_text(g.gfa()); fa.write_text(synth.fasta(reads))
    F,G=fa.read_text(),gfa.read_text()
    M,X=int(rng.choice([1,2,5])),int(rng.choice([1,3,4,7]))
    O,E=int(rng.choice([0,1,4,10])),int(rng.choice([1,2,5]))
    b,f=int(rng.choice([0,1,2,5,30])),float(rng.choice([0.0,0.01,0.1,0.5]))
    R,r,B=int(rng.choice([0,1,4,9])),float(rng.choice([0.0,0.05,0.1,0.4])),float(rng.choice([0.4,0.7,1.0]))
    mx=["-M",str(M),"-X",str(X)]
    MAT=str(rng.choice(["none","none","HOXD70","HOXD55"]))
    if MAT!="none": mx=["-t",MAT]
    for mode in (4,5):
        cmp((seed,mode), ["-m",str(mode)]+mx+[str(fa),str(gfa)], lambda: pyref.run(mode,F,G,match=M,mismatch=X,matrix=MAT))
    for mode in (8,9):
        cmp((seed,mode,R,r,B), ["-m",str(mode)]+mx+["-R",str(R),"-r",str(r),"-B",str(B),str(fa),str(gfa)], lambda: pyref.run(mode,F,G,match=M,mismatch=X,matrix=MAT,base_rec_cost=R,multi_rec_cost=r,rec_band_width=B))
    for mode in (6,7):
        cmp((seed,mode,O,E), ["-m",str(mode)]+mx+["-O",str(O),"-E",str(E),str(fa),str(gfa)], lambda: pyref.run_gap_pathwise(mode,F,G,match=M,mismatch=X,matrix=MAT,gap_open=O,gap_ext=E))
    cmp((seed,2,b,f,O,E), ["-m","2"]+mx+["-b",str(b),"-f",str(f),"-O",str(O),"-E",str(E),str(fa),str(gfa)], lambda: pyref.run_mode2(F,G,match=M,mismatch=X,matrix=MAT,gap_open=O,gap_ext=E,extra_b=b,extra_f=f))
    cmp((seed,0,b,f), ["-m","0"]+mx+["-b",str(b),"-f",str(f),str(fa),str(gfa)], lambda: pyref.run_poa(0,F,G,match=M,mismatch=X,matrix=MAT,extra_b=b,extra_f=f))
    cmp((seed,1), ["-m","1"]+mx+[str(fa),str(gfa)], lambda: pyref.run_poa(1,F,G,match=M,mismatch=X,matrix=MAT))
    cmp((seed,3,O,E), ["-m","3"]+mx+["-O",str(O),"-E",str(E),str(fa),str(gfa)], lambda: pyref.run_poa(3,F,G,match=M,mismatch=X,matrix=MAT,gap_open=O,gap_ext=E))
    for mode in (0,1,2,3):
        kw={}; ex=["-s","true"]
        if mode in (0,2): ex+=["-b",str(max(b,1)),"-f",str(f)]; kw.update(extra_b=max(b,1),extra_f=f)
        if mode in (2,3): ex+=["-O",str(O),"-E",str(E)]; kw.update(gap_open=O,gap_ext=E)
        cmp((seed,mode,"amb"), ["-m",str(mode)]+mx+ex+[str(fa),str(gfa)], lambda: pyref.run_poa_amb(mode,F,G,match=M,mismatch=X,matrix=MAT,**kw))
print("done", n, "comparisons", bad, "mismatches", round(time.time()-t0), "s", flush=True)
