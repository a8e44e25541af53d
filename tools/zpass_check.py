import sys, os
sys.path.insert(0,'/root/repo')
import numpy as np, synth
from sternheimergw_b200 import Context
syn=synth.preset("si64"); c=Context(0); c.install_system(syn); kq=syn.kpairs[0].kq
rng=np.random.default_rng(5); nvec=int(sys.argv[1]) if len(sys.argv)>1 else 37
psi=np.zeros((kq.npwx,nvec),complex,order="F"); psi[:kq.npw]=rng.standard_normal((kq.npw,nvec))+1j*rng.standard_normal((kq.npw,nvec))
om=rng.standard_normal(nvec)+0.3j
out={}
for rep in range(3):
    for v in ("0","1","2","3"):
        os.environ["SGW_ZPASS"]=v
        o=c.linear_op(0,om,kq.alpha_pv,psi)
        if v=="0": ref=o
        else:
            d=np.abs(o-ref); bad=np.argwhere(d>0)
            print(rep, v, "equal" if bad.size==0 else ("DIFF n=%d max=%.3e cols=%s rows=%s"%(len(bad), d.max(), sorted(set(bad[:,1].tolist()))[:10], bad[:5,0].tolist())))
