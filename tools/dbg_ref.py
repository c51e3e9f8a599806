import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import numpy as np
import reflib
from helpers import *
from oracle import oracle as orc
from cufinufft_b200 import cufinufft, gpuarray

def run(nufft_type, modes, M, tol, dtype, dist, opts, nudge=False):
    dim=len(modes); shape=tuple(modes)[::-1]; cd=cdtype(dtype)
    pts=make_points(M,dim,dtype,seed=21,dist=dist)
    kp,nf,bs,_=orc.plan_params(nufft_type,modes,tol,dtype,gpu_method=opts.get('gpu_method'))
    # points whose stencil start is exact (reference reads ker[ns] there)
    hit=np.zeros(M,bool)
    for d in range(dim):
        x=pts[d]; pi=dtype(np.pi)
        shift=np.where(x<-pi,1.5,np.where(x>=pi,-0.5,0.5))
        xr=((x.astype(np.float64)*0.159154943091895336+shift)*nf[d]).astype(dtype).astype(np.float64)
        t=xr-kp.ns/2.0
        hit|= (np.ceil(t)==t)
    print('case',nufft_type,modes,M,tol,dtype.__name__,dist,opts,'ns',kp.ns,'exact-stencil points:',hit.sum())
    if nudge:
        pts=[p[~hit].copy() for p in pts]; M=pts[0].size
    dev=[gpuarray.to_gpu(p) for p in pts]
    ours=cufinufft(nufft_type,shape,eps=tol,dtype=dtype,**opts); ours.set_pts(*dev[::-1])
    ref=reflib.RefPlan(nufft_type,modes,tol,dtype,**opts); ref.set_pts(dev)
    rng=np.random.default_rng(5)
    if nufft_type==1:
        data=make_strengths(M,dtype)[0]
        c=gpuarray.to_gpu(data); fo=gpuarray.zeros(shape,cd); fr=gpuarray.zeros(shape,cd)
        ours.execute(c,fo); ref.execute(c,fr); a=fo.get(); b=fr.get()
        idx=rng.integers(0,int(np.prod(modes)),40)
        ex=orc.dirft1_sampled(pts,data,modes,1,idx); ga=a.ravel()[idx]; gb=b.ravel()[idx]
    else:
        data=make_modes_data(modes,dtype)[0]
        fk=gpuarray.to_gpu(data); co=gpuarray.zeros((M,),cd); cr=gpuarray.zeros((M,),cd)
        ours.execute(co,fk); ref.execute(cr,fk); a=co.get(); b=cr.get()
        idx=rng.integers(0,M,40)
        ex=orc.dirft2_sampled(pts,data,modes,-1,idx); ga=a[idx]; gb=b[idx]
    sc=np.abs(ex).max()
    print('   nan ours %d ref %d | rel-l2 ours-vs-ref %.3e | max err vs direct: ours %.3e ref %.3e'%(
        np.isnan(a).sum(), np.isnan(b).sum(), rel_l2(a,b), np.abs(ga-ex).max()/sc, np.abs(gb-ex).max()/sc))
    if nufft_type==2:
        d=np.abs(a-b); w=np.argsort(d)[-5:]; print('   worst pts', w, d[w], 'hit?', hit[w] if not nudge else '')
    ref.destroy(); ours.destroy()

for nudge in (False, True):
    run(1,(64,64,64),1_000_000,1e-5,np.float32,'cluster',dict(gpu_method=2),nudge)
    run(1,(512,512),262_144,1e-4,np.float32,'uniform',dict(gpu_method=2),nudge)
    run(2,(512,512),262_144,1e-4,np.float32,'uniform',dict(gpu_method=1),nudge)
    run(1,(1000,1000),1_000_000,1e-3,np.float32,'uniform',dict(gpu_method=2),nudge)
