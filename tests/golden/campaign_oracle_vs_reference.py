#!/usr/bin/env python
"""Randomised campaign: the float64 restatement (oracle/nem_oracle.c) against the UNMODIFIED
reference (oracle/_ref/nem_ref_harness) on random shapes, graphs, K = 2..5, both algorithms, both
update orders, all dispersion / proportion models and random initial parameters.  Needs
/root/reference (build container only); not collected by pytest.

    python tests/golden/campaign_oracle_vs_reference.py SEED CASES

End of round 1 (seeds 1-3, 140 cases, 101 comparable): every difference falls under a documented
gate (DESIGN.md sections 3 and 6) --
  * beta*ctx beyond the double exp range: the reference's posterior is NaN (cases skipped here,
    pinned by tests/test_oracle_vs_reference.py::test_reference_context_overflow_is_gated_not_copied);
  * exact score ties (equal proportions, one dispersion, unweighted graph): the count-based density
    keeps them tied, the reference's in-order float32 sums break them by rounding noise, and the
    trajectories part;
  * fuzzy nem: float32 log-densities move the iteration at which max|dt| < 1e-8 is met by one or
    two, and a weighted median sitting exactly on n/2 flips a centre between 1 and 1/2;
  * one random input crashes the reference harness itself (SIGSEGV), reported and skipped.
"""
import sys, os, tempfile, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import nemo
from pangenomenem_b200 import synth
def rel_close(a,b,tol):
    a=np.asarray(a,dtype=np.float64); b=np.asarray(b,dtype=np.float64)
    return bool(np.all(np.abs(a-b)<=tol*np.maximum(np.abs(b),1e-30)))
nemo.build(ref=True); assert nemo.have_ref()
rng=np.random.default_rng(int(sys.argv[1]))
bad=0; ran=0
for it in range(int(sys.argv[2])):
    n=int(rng.integers(300,4000)); d=int(rng.integers(8,140)); k=int(rng.choice([2,3,3,4,5]))
    graph=str(rng.choice(["pangenome","random","chain"])); weighted=bool(rng.integers(0,2))
    algo=str(rng.choice(["ncem","ncem","nem"])); disp=str(rng.choice(["s__","sk_","s_d","skd"])); prop=str(rng.choice(["pk","p_"]))
    update=str(rng.choice(["seq","seq","para"])); beta=float(rng.choice([0.0,0.3,0.5,1.0,2.0]))
    pg=synth.make_pangenome(n,d,seed=int(rng.integers(0,1<<30)),graph=graph,weighted=weighted)
    # random initial theta: constant or per-genome centres in {0,.5,1}, dispersions in (0.05,0.5)
    p=rng.dirichlet(np.ones(k)*5).astype(np.float32)
    if rng.random()<0.5: cen=np.repeat(rng.choice([0.,.5,1.],size=k)[:,None],d,axis=1)
    else: cen=rng.choice([0.,.5,1.],size=(k,d))
    if rng.random()<0.5: ds=np.repeat(rng.uniform(0.05,0.5,size=k)[:,None],d,axis=1)
    else: ds=rng.uniform(0.05,0.5,size=(k,d))
    cen=cen.astype(np.float32); ds=ds.astype(np.float32)
    kw=dict(algo=algo,beta=beta,disp=disp,prop=prop,update=update,it_max=int(rng.choice([3,8,30])))
    with tempfile.TemporaryDirectory() as tmp:
        base=os.path.join(tmp,'nem_file')
        synth.write_nem_files(base,pg,m_text=synth.m_line(1,p,cen,ds),weighted_flag=1 if weighted else 0)
        hp=None
        try:
            r=nemo.run_ref_harness(base,os.path.join(tmp,'out'),k=k,tie="first",**kw)
        except Exception as e:
            print("harness error",e); continue
        # theta exactly as the reference reads the .m text
        th=synth.read_m_text(open(base+'.m').read(),k,d) if hasattr(synth,'read_m_text') else None
    if r["status"]!=0 or r["density_zero"]:
        continue
    # the reference reads proportions as float, last = 1 - sum
    pr=p.copy(); acc=np.float32(1.0)
    for q in range(k-1): acc=np.float32(acc-pr[q])
    pr[k-1]=acc
    o=nemo.Problem(pg.x,pg.row_ptr,pg.col,pg.wgt,k=k,**kw).fit(pr,cen,ds)
    ran+=1
    msgs=[]
    # known gates: beta*ctx beyond the double exp range (reference NaN)
    lab=o.label if algo=="ncem" else o.t.argmax(axis=1)
    ctx=np.zeros((n,k)); np.add.at(ctx,(np.repeat(np.arange(n),np.diff(pg.row_ptr)),lab[pg.col]),pg.wgt)
    if (beta*ctx.max(axis=1)>700).any():
        skipped_over=globals().get('skipped_over',0)+1; globals()['skipped_over']=skipped_over; ran-=1; continue
    if o.iters!=r["iters"] or o.converged!=r["converged"]: msgs.append(f"iters {o.iters}/{r['iters']} conv {o.converged}/{r['converged']}")
    if not np.array_equal(o.center,r["center"]): msgs.append(f"centres differ at {(o.center!=r['center']).sum()}")
    if algo=="ncem":
        nd=int((o.label!=r["cm"].argmax(axis=1)).sum())
        if nd>o.n_ties: msgs.append(f"labels differ {nd} (ties {o.n_ties})")
        if not rel_close(o.disp,r["disp"],1e-5): msgs.append("disp")
        if not rel_close(o.prop,r["prop"],1e-5): msgs.append("prop")
    else:
        if np.abs(o.t-r["cm"]).max()>5e-3: msgs.append(f"t maxdiff {np.abs(o.t-r['cm']).max():.2e}")
    if msgs:
        bad+=1; print(it,dict(n=n,d=d,k=k,graph=graph,w=weighted,**kw),msgs)
print("ran",ran,"bad",bad,"skipped for exp overflow",globals().get("skipped_over",0))
