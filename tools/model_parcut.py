import random, sys
NONE=1<<30
INF=(NONE,1<<40)
def seq_merge(ids, look):
    ids=list(ids)
    while True:
        best=NONE; bi=-1
        for i in range(len(ids)-1):
            r=look(ids[i],ids[i+1])
            if r<best: best=r; bi=i
        if best==NONE: return ids
        ids[bi:bi+2]=[best]
def par(ids0, look, P=32):
    m=len(ids0); idv=list(ids0); alive=[True]*m
    nxt=[i+1 for i in range(m)]; prv=[i-1 for i in range(m)]
    rk=[look(idv[i],idv[i+1]) if i+1<m else NONE for i in range(m)]
    c=(m+P-1)//P
    if c>1: c|=1
    rounds=0; merges=0
    while True:
        props=[]
        for t in range(P):
            lo=min(t*c,m); hi=min(lo+c,m)
            m1=INF; m2=INF
            for x in range(lo,hi):
                if rk[x]==NONE: continue
                v=(rk[x],x)
                m2=min(m2,max(m1,v)); m1=min(m1,v)
            props.append((m1,m2))
        if all(p[0]==INF for p in props): break
        rounds+=1
        info=[None]*P; claim={}
        vs=[]
        for t,(m1,m2) in enumerate(props):
            if m1==INF: continue
            r,x=m1; j=nxt[x]; q=prv[x]; k=nxt[j]
            R=look(r,idv[k]) if k<m else NONE
            L=look(idv[q],r) if q>=0 else NONE
            Lk=(L,q) if L!=NONE else INF
            Rk=(R,x) if R!=NONE else INF
            cc=min(m2,Lk,Rk)
            nextkey=(m1[0],m1[1]+1)
            v=max(nextkey,cc)
            info[t]=dict(x=x,j=j,q=q,k=k,r=r,L=L,R=R,key=m1,v=v)
            for part in (x,j):
                claim[part]=min(claim.get(part,INF),m1)
        cut=INF
        for t in range(P):
            I=info[t]
            if I is None: continue
            cut=min(cut,I['v'])
            lowest=INF
            for part in (I['q'],I['x'],I['j'],I['k']):
                if part in claim: lowest=min(lowest,claim[part])
            if lowest<I['key']: cut=min(cut,I['key'])
        for t in range(P):
            I=info[t]
            if I is None or not (I['key']<cut): continue
            merges+=1
            x,j,q,k=I['x'],I['j'],I['q'],I['k']
            idv[x]=I['r']; alive[j]=False; rk[j]=NONE; rk[x]=I['R']; nxt[x]=k
            if k<m: prv[k]=x
            if q>=0: rk[q]=I['L']
    return [idv[i] for i in range(m) if alive[i]], rounds, merges
def test(seed):
    rng=random.Random(seed)
    A=rng.choice([2,3,3,5])
    toks=[(a,) for a in range(A)]
    seen=set(toks)
    target=rng.randint(8,60)
    tries=0
    while len(toks)<target and tries<1000:
        tries+=1
        a=rng.choice(toks); b=rng.choice(toks); t=a+b
        if len(t)<=12 and t not in seen: seen.add(t); toks.append(t)
    order=toks[A:]
    if rng.random()<0.7: rng.shuffle(order)
    toks=toks[:A]+order
    idx={t:i for i,t in enumerate(toks)}
    def look(a,b): return idx.get(toks[a]+toks[b],NONE)
    n=rng.randint(2,300)
    kind=rng.randint(0,2)
    if kind==0: s=[rng.randrange(A) for _ in range(n)]
    elif kind==1: s=[rng.randrange(A)]*n
    else:
        per=[rng.randrange(A) for _ in range(rng.randint(2,4))]; s=(per*n)[:n]
    a=seq_merge(s,look); b,r,mg=par(s,look,P=rng.choice([4,8,32,128,256]))
    return a==b,(s,toks,a,b),r,mg
bad=0; R=0; M=0
N=int(sys.argv[1]) if len(sys.argv)>1 else 3000
for seed in range(N):
    ok,info,r,mg=test(seed)
    R+=r; M+=mg
    if not ok:
        bad+=1
        if bad<3: print(seed,info)
print('bad',bad,'merges/round',M/max(R,1))

def big(P, n=3500, A=52, seed=1):
    rng=random.Random(seed)
    # vocab: all pairs random rank, some triples
    pairs={}
    rank=A
    allp=[(a,b) for a in range(A) for b in range(A)]
    rng.shuffle(allp)
    toks={}
    for (a,b) in allp[:1500]:
        toks[(a,b)]=rank; rank+=1
    # triples: (pairtoken, letter) random 3000
    ids=list(toks.values())
    for _ in range(3000):
        toks[(rng.choice(ids),rng.randrange(A))]=rank; rank+=1
        toks[(rng.randrange(A),rng.choice(ids))]=rank; rank+=1
    def look(a,b): return toks.get((a,b),NONE)
    s=[rng.randrange(A) for _ in range(n)]
    b,r,mg=par(s,look,P=P)
    return r,mg
for P in (32,128,256,512,1024):
    r,mg=big(P)
    print(P,'rounds',r,'merges',mg,'per round',mg/r)
