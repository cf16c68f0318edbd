#!/usr/bin/env python
"""CPU emulation of a 4-wide view of the LBVH (each record = the grandchildren of a binary node) against the binary walk:
records fetched, box tests and leaves per 15-NN / 1-NN query.  Run from the repo root.  Outcome: half the dependent record
fetches for the same box tests -- but on the GPU the 1-NN phases got slower (see profiles/README.md)."""
import sys
sys.path.insert(0,'profiles'); sys.path.insert(0,'fast-lio-sam-qn_b200')
import numpy as np
import emulate_knn_seeding as E
from b200reg import synth
src,dst,T=synth.make_pair(1000,100000,100000)
P,nodes,root=E.build(src[:,:3].astype(np.float64))
# 4-ary view: children of node = grandchildren (or child if child is leaf)
def kids4(ref):
    lo0,hi0,r0,lo1,hi1,r1=nodes[ref]
    out=[]
    for (lo,hi,r) in ((lo0,hi0,r0),(lo1,hi1,r1)):
        if r>=0:
            a=nodes[r]; out.append((a[0],a[1],a[2])); out.append((a[3],a[4],a[5]))
        else: out.append((lo,hi,r))
    return out
def search4(q,k,seeds):
    best=[]
    def insert(pos):
        d2=float(((P[pos]-q)**2).sum())
        if len(best)<k: best.append((d2,pos)); best.sort()
        elif (d2,pos)<best[-1]: best[-1]=(d2,pos); best.sort()
    seen=set()
    for s in seeds:
        if s not in seen: seen.add(s); insert(s)
    worst=lambda: best[-1][0] if len(best)==k else np.inf
    visits=boxes=leaves=0
    stack=[(root,0.0)]
    while stack:
        ref,dn=stack.pop()
        if dn>worst(): continue
        if ref<0:
            leaves+=1
            c=-1-ref; a,cnt=c>>4,c&15
            for pos in range(a,a+cnt):
                if pos not in seen: seen.add(pos); insert(pos)
            continue
        visits+=1
        ks=[(E.box_d2(q,lo,hi),r) for lo,hi,r in kids4(ref)]
        boxes+=len(ks)
        ks.sort(key=lambda x:-x[0])   # far first onto the stack -> near popped first
        for d,r in ks:
            if not (d>worst()): stack.append((r,d))
    return visits,boxes,leaves,[p for _,p in best]
n=len(P); rng=np.random.default_rng(0)
qs=np.concatenate([np.arange(s,s+32) for s in rng.integers(0,n-64,16)])
def window(i,w=15):
    lo=max(0,i-w//2); hi=min(n-1,lo+w-1); lo=max(0,hi-(w-1)); return list(range(lo,hi+1))
E.K=15
b2=np.array([E.search(P,nodes,root,P[i],window(i),set())[:2] for i in qs])
w4=[];ok=True
for i in qs:
    v,b,l,res=search4(P[i],15,window(i)); w4.append((v,b,l)); ok&=sorted(res)==sorted(E.search(P,nodes,root,P[i],window(i),set())[3])
print("15-NN binary: records %.1f boxes %.1f leaves %.1f"%(b2[:,0].mean()/2,b2[:,0].mean(),b2[:,1].mean()))
print("15-NN 4-ary : records %.1f boxes %.1f leaves %.1f same=%s"%(*np.mean(w4,0),ok))
Tt,tn,tr=E.build(dst[:,:3].astype(np.float64))
P,nodes,root=Tt,tn,tr
S=src[:,:3].astype(np.float64)
E.K=1
qi=rng.integers(0,len(S),512)
b1=np.array([E.search(P,nodes,root,S[i],[],set())[:2] for i in qi])
w1=[search4(S[i],1,[])[:3] for i in qi]
print("1-NN binary: records %.1f boxes %.1f leaves %.1f"%(b1[:,0].mean()/2,b1[:,0].mean(),b1[:,1].mean()))
print("1-NN 4-ary : records %.1f boxes %.1f leaves %.1f"%tuple(np.mean(w1,0)))
# variant: nearest first, the other three pushed unsorted (slot order)
def search4u(q,k,seeds):
    best=[]
    def insert(pos):
        d2=float(((P[pos]-q)**2).sum())
        if len(best)<k: best.append((d2,pos)); best.sort()
        elif (d2,pos)<best[-1]: best[-1]=(d2,pos); best.sort()
    seen=set()
    for s in seeds:
        if s not in seen: seen.add(s); insert(s)
    worst=lambda: best[-1][0] if len(best)==k else np.inf
    visits=leaves=0
    stack=[(root,0.0)]
    while stack:
        ref,dn=stack.pop()
        if dn>worst(): continue
        while ref>=0:
            visits+=1
            ks=[(E.box_d2(q,lo,hi),r) for lo,hi,r in kids4(ref)]
            m=min(range(len(ks)),key=lambda t:ks[t][0])
            for t,(d,r) in enumerate(ks):
                if t!=m and not (d>worst()): stack.append((r,d))
            dn,ref=ks[m]
            if dn>worst(): ref=None; break
        if ref is not None and ref<0:
            leaves+=1
            c=-1-ref; a,cnt=c>>4,c&15
            for pos in range(a,a+cnt):
                if pos not in seen: seen.add(pos); insert(pos)
    return visits,leaves,[p for _,p in best]
w1u=[search4u(S[i],1,[])[:2] for i in qi]
print("1-NN 4-ary nearest-first, rest unsorted: records %.1f leaves %.1f"%tuple(np.mean(w1u,0)))
# seeded variant (previous correspondence ~ true NN known): seed with true NN
w1s=[]
for i in qi:
    nn=E.search(P,nodes,root,S[i],[],set())[3][0]
    w1s.append(search4u(S[i],1,[nn])[:2])
print("1-NN 4-ary seeded with the true NN (iteration >= 2): records %.1f leaves %.1f"%tuple(np.mean(w1s,0)))
b1s=[]
for i in qi:
    nn=E.search(P,nodes,root,S[i],[],set())[3][0]
    b1s.append(E.search(P,nodes,root,S[i],[nn],set())[:2])
print("1-NN binary seeded with the true NN: records %.1f leaves %.1f"%(np.mean(b1s,0)[0]/2,np.mean(b1s,0)[1]))
