"""Shared-memory bank-conflict model of the cell path's LDS.64 gathers: builds the byte lists of
a liquid-like configuration the way csrc/cells.cu does (rows, K = 4 bins, sorted slots, lists in slot order) and
counts the wavefronts of every gather instruction (two half-warps x max addresses per bank pair) for
  plain    lists as stored (padded to the warp maximum)                    -> 4.08 wavefronts, 24.6 evaluations
  aligned  each of the three row sections padded to its own warp maximum   -> 3.36 wavefronts, 36.1 evaluations
  range    contiguous [first, last] slot ranges per row section            -> 3.32 wavefronts, 44.1 evaluations
(ncu measures 4.1 for the kernel; ideal = 2.0).   python scripts/bank_conflict_model.py"""
import numpy as np
rng=np.random.default_rng(0)
n=256; N=n*n; rho=0.8; box=np.sqrt(N/rho); a=box/n
g=(np.stack(np.meshgrid(np.arange(n),np.arange(n),indexing='ij'),-1).reshape(-1,2)+0.5)*a
R=np.mod(g+rng.normal(0,0.18,g.shape)*a,box)       # liquid-like disorder
rl=3.0
nrows=int(box/rl); nbx=int(box/(rl/4)); hy=box/nrows; wx=box/nbx
row=np.minimum((R[:,1]/hy).astype(int),nrows-1); b=np.minimum((R[:,0]/wx).astype(int),nbx-1)
order=np.lexsort((np.arange(N),b,row)); R=R[order]; row=row[order]; b=b[order]
cell=row*nbx+b
cs=np.searchsorted(cell,np.arange(nrows*nbx+1))
def wavefronts(idx):  # idx: (32,) window slot indices (or -1 sentinel -> use 252+lane%4)
    tot=0
    for h in range(2):
        sub=idx[h*16:(h+1)*16]
        addr=np.unique(sub)
        banks=addr%16
        tot+=np.bincount(banks,minlength=16).max()
    return tot
res={'plain':[], 'aligned':[], 'evals_plain':[], 'evals_aligned':[], 'range':[], 'evals_range':[]}
units=0
for r in range(2,nrows-2):
    s_row=cs[r*nbx]; e_row=cs[(r+1)*nbx]
    for i0 in range(((s_row+31)//32)*32, e_row-32, 32*7):
        lanes=np.arange(i0,i0+32)
        if row[lanes[0]]!=row[lanes[-1]]: continue
        bf,bl=b[lanes[0]],b[lanes[-1]]
        if bf<5 or bl>nbx-6: continue
        lists=[[],[],[]]
        ws=[cs[(r+k-1)*nbx+bf-4]&~1 for k in range(3)]
        per=[[None]*32 for k in range(3)]
        for li,i in enumerate(lanes):
            for k in range(3):
                s=cs[(r+k-1)*nbx+b[i]-4]; e=cs[(r+k-1)*nbx+b[i]+5]
                d=R[s:e]-R[i]; r2=(d*d).sum(1)
                nb=np.nonzero((r2<rl*rl)&(np.arange(s,e)!=i))[0]+s-ws[k]+k*84
                per[k][li]=nb
        units+=1
        # plain: concatenated, padded to warp max with sentinel
        full=[np.concatenate([per[k][li] for k in range(3)]) for li in range(32)]
        L=max(len(f) for f in full); L=((L+3)//4)*4
        w=0
        for t in range(L):
            idx=np.array([full[li][t] if t<len(full[li]) else 252+li%4 for li in range(32)])
            w+=wavefronts(idx)
        res['plain'].append(w/L); res['evals_plain'].append(L)
        # aligned: each section padded to its warp max
        w=0; Lt=0
        for k in range(3):
            Lk=max(len(per[k][li]) for li in range(32))
            for t in range(Lk):
                idx=np.array([per[k][li][t] if t<len(per[k][li]) else 252+li%4 for li in range(32)])
                w+=wavefronts(idx)
            Lt+=Lk
        res['aligned'].append(w/Lt); res['evals_aligned'].append(Lt)
        # contiguous ranges [first,last] per section, aligned
        w=0; Lt=0
        for k in range(3):
            rng_=[(per[k][li].min(), per[k][li].max()) if len(per[k][li]) else (0,-1) for li in range(32)]
            Lk=max(hi-lo+1 for lo,hi in rng_)
            for t in range(Lk):
                idx=np.array([rng_[li][0]+t if rng_[li][0]+t<=rng_[li][1] else 252+li%4 for li in range(32)])
                w+=wavefronts(idx)
            Lt+=Lk
        res['range'].append(w/Lt); res['evals_range'].append(Lt)
print("units",units)
for k in res: print(k, np.mean(res[k]))
