#!/usr/bin/env python
"""numpy replay of k_feat_nn's streaming order on oracle descriptors of a voxelised KITTI-shaped pair: record-level filter
tests with and without norm-ordered tile skipping, refine volume, tiles fetched per 128-query block (quoted in
profiles/README.md and DESIGN.md section 6).  Run from the repo root: python profiles/emulate_feat_nn_tiles.py"""
import sys,time
sys.path.insert(0,'fast-lio-sam-qn_b200'); sys.path.insert(0,'.')
import numpy as np
from b200reg import synth
from oracle import oracle
s,d,T=synth.make_pair(2000,100000,100000,mode="quatro",voxel=0.3)
_,_,fs=oracle.fpfh(s); _,_,fd=oracle.fpfh(d)
def norms(F): return np.sqrt((F.reshape(-1,3,11).astype(np.float64)**2).sum(2)).astype(np.float32)
def expand10(v):
    v=v.astype(np.uint32)
    v=(v|(v<<16))&0x030000FF; v=(v|(v<<8))&0x0300F00F; v=(v|(v<<4))&0x030C30C3; v=(v|(v<<2))&0x09249249
    return v
def code(N):
    q=np.minimum(1023,(N*10.23).astype(np.int64))
    return (expand10(q[:,2])<<2)|(expand10(q[:,1])<<1)|expand10(q[:,0])
def prep(F):
    ok=(F!=0).any(1); F=F[ok]; N=norms(F); c=code(N); o=np.argsort(c,kind='stable'); return F[o],N[o],c[o]
Q,QN,qc=prep(fs); B,BN,bc=prep(fd)
thr2=35.0**2
TILE=64
nt=(len(B)+TILE-1)//TILE
bmin=np.array([BN[t*TILE:(t+1)*TILE].min(0) for t in range(nt)]); bmax=np.array([BN[t*TILE:(t+1)*TILE].max(0) for t in range(nt)])
print("queries",len(Q),"base",len(B),"tiles",nt)
best=np.full(len(Q),thr2,np.float32)
tests_old=0; tests_new=0; refines=0; tile_visits=0; block_tile_loads=0
# per block of 128 queries: start tile
nblk=(len(Q)+127)//128
Bd=B.astype(np.float64); Qd=Q.astype(np.float64)
for blk in range(nblk):
    q0=blk*128; q1=min(len(Q),q0+128)
    mid=qc[min(len(Q)-1,q0+64)]
    t0=min(nt-1,np.searchsorted(bc,mid)//TILE)
    order=list(range(t0,nt))+list(range(t0-1,-1,-1))
    bq=best[q0:q1]; qn=QN[q0:q1]
    for t in order:
        bound=bq*1.0001+1e-3
        e=np.maximum(np.maximum(bmin[t]-qn,qn-bmax[t]),0); lb=(e*e).sum(1)
        need=lb<=bound
        tests_old+=(q1-q0)*min(TILE,len(B)-t*TILE)
        if not need.any(): continue
        block_tile_loads+=1
        idx=np.nonzero(need)[0]
        tile_visits+=len(idx)
        bn=BN[t*TILE:(t+1)*TILE]
        tests_new+=len(idx)*len(bn)
        er=qn[idx][:,None,:]-bn[None,:,:]; lbr=(er*er).sum(2)
        pas=lbr<=bound[idx][:,None]
        refines+=pas.sum()
        # exact distances for passing (approx: whole tile), update best with tile min over passing
        D=((Qd[q0:q1][idx][:,None,:]-Bd[t*TILE:(t+1)*TILE][None,:,:])**2).sum(2)
        D=np.where(pas,D,np.inf)
        bq[idx]=np.minimum(bq[idx],D.min(1).astype(np.float32))
    best[q0:q1]=bq
tot=len(Q)*len(B)
print("record tests old %.3g new %.3g ratio %.3f"%(tests_old,tests_new,tests_new/tests_old))
print("refines %.3g (%.2f%% of pairs)"%(refines,100*refines/tot))
print("per-query tile visits avg %.1f of %d; block tile loads avg %.1f of %d"%(tile_visits/len(Q),nt,block_tile_loads/nblk,nt))
