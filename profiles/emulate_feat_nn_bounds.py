#!/usr/bin/env python
"""Survivor fraction of candidate lower bounds for the 33-D matcher (block norms, DCT projections + residual norms, PCA
projections + residual) against the true within-bound fraction, on oracle descriptors (quoted in DESIGN.md sections 6 and 9).
Run from the repo root: python profiles/emulate_feat_nn_bounds.py"""
import sys
sys.path.insert(0,'fast-lio-sam-qn_b200'); sys.path.insert(0,'.')
import numpy as np
from b200reg import synth
from oracle import oracle
s,d,T=synth.make_pair(2000,100000,100000,mode="quatro",voxel=0.3)
_,_,fs=oracle.fpfh(s); _,_,fd=oracle.fpfh(d)
ok=lambda F:F[(F!=0).any(1)].astype(np.float64)
A=ok(fs); B=ok(fd)
D=np.maximum((A*A).sum(1)[:,None]+(B*B).sum(1)[None,:]-2*A@B.T,0)
best=D.min(1); bound=best*1.0001+1e-3
print("true within bound: %.4f"%((D<=bound[:,None]).mean()))
def blocknorm(F): return np.sqrt((F.reshape(-1,3,11)**2).sum(2))
def lb_feats(fa,fb):  # plain euclid in feature space
    return np.maximum((fa*fa).sum(1)[:,None]+(fb*fb).sum(1)[None,:]-2*fa@fb.T,0)
def report(name,fa,fb):
    L=lb_feats(fa,fb); print("%-28s survivors %.4f"%(name,(L<=bound[:,None]).mean()))
# (a) block norms: lb = sum (|a_k|-|b_k|)^2 == euclid on norm vectors
report("block norms (3)",blocknorm(A),blocknorm(B))
# DCT basis per block
def dct_basis(n=11):
    k=np.arange(n)[:,None]; i=np.arange(n)[None,:]
    M=np.cos(np.pi*(i+0.5)*k/n)*np.sqrt(2.0/n); M[0]/=np.sqrt(2); return M  # rows orthonormal
M=dct_basis()
def proj_feats(F,rows):  # rows: list of (block, dctindex)
    U=np.zeros((len(rows),33))
    for r,(b,k) in enumerate(rows): U[r,b*11:(b+1)*11]=M[k]
    P=F@U.T
    res=np.sqrt(np.maximum((F*F).sum(1)-(P*P).sum(1),0))
    return np.concatenate([P,res[:,None]],1)
for name,rows in [("dct1 x3 + resid (4)",[(b,1) for b in range(3)]),("dct1,2 x3 + resid (7)",[(b,k) for b in range(3) for k in (1,2)]),("dct1,2,3 x3 + resid (10)",[(b,k) for b in range(3) for k in (1,2,3)])]:
    report(name,proj_feats(A,rows),proj_feats(B,rows))
# per-block residual variant: dct1,2 per block + per-block residual norms (9)
def proj_blockres(F,ks):
    out=[]
    for b in range(3):
        Fb=F[:,b*11:(b+1)*11]; P=Fb@M[ks].T
        out.append(P); out.append(np.sqrt(np.maximum((Fb*Fb).sum(1)-(P*P).sum(1),0))[:,None])
    return np.concatenate(out,1)
report("dct1 + block resid (6)",proj_blockres(A,[1]),proj_blockres(B,[1]))
report("dct1,2 + block resid (9)",proj_blockres(A,[1,2]),proj_blockres(B,[1,2]))
# PCA
C=np.cov(np.concatenate([A,B]).T); w,V=np.linalg.eigh(C); V=V[:,::-1]
mu=np.concatenate([A,B]).mean(0)
for m in (3,7,11):
    U=V[:,:m]
    def pf(F):
        P=(F-mu)@U; res=np.sqrt(np.maximum(((F-mu)**2).sum(1)-(P*P).sum(1),0)); return np.concatenate([P,res[:,None]],1)
    report("pca%d + resid (%d)"%(m,m+1),pf(A),pf(B))
