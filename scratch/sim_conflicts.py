import numpy as np, time
rng=np.random.default_rng(0)
def cost(codes):  # codes [n][8]; groups of 32 consecutive
    n=len(codes)//32*32
    c=codes[:n].reshape(-1,32,8)
    tot=0
    for g in c:
        for j in range(8):
            col=g[:,j]
            # distinct addresses per bank
            u=np.unique(col)
            tot+=np.bincount(u%32,minlength=32).max()
    return tot/(n//32*8)
def greedy(codes,W=128,p=2):
    n=len(codes)
    remaining=list(range(n))
    order=[]
    while len(remaining)>=32:
        load=np.zeros((8,32),np.int32)   # distinct-address count per bank (approx: count all)
        seen=[set() for _ in range(8)]
        for s in range(32):
            pool=remaining[:W]
            cc=codes[pool]            # [W][8]
            b=cc%32
            # incremental cost: sum_j load[j][bank]^p  (prefer empty banks); zero if address already present
            inc=np.zeros(len(pool))
            for j in range(8):
                l=load[j][b[:,j]].astype(float)
                dup=np.array([c in seen[j] for c in cc[:,j]])
                inc+=np.where(dup,0,(l+1)**p-l**p)
            k=int(np.argmin(inc))
            idx=pool[k]
            order.append(idx)
            for j in range(8):
                if codes[idx,j] not in seen[j]:
                    seen[j].add(codes[idx,j]); load[j][codes[idx,j]%32]+=1
            remaining.pop(k)
    order+=remaining
    return codes[order]
codes=rng.integers(0,256,size=(2720,8))
print('random', cost(codes))
for W in (64,256,1024):
    t=time.time(); r=greedy(codes,W); print('greedy W',W,cost(r),time.time()-t)

def greedy_chunked(codes,CH=512,p=2):
    out=[]
    for b in range(0,len(codes),CH):
        ch=codes[b:b+CH]
        out.append(greedy(ch,W=CH,p=p) if len(ch)>=32 else ch)
    return np.concatenate(out)
for CH in (256,512,1024):
    t=time.time(); r=greedy_chunked(codes,CH); print('chunked',CH,cost(r),time.time()-t)
