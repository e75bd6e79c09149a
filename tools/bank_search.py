import itertools, random
TZ=5
def ndx(a): return (a>>1)&1
def ndy(a): return (a>>2)&1
def ndz(a): return (a^(a>>1))&1
def cost(CSTR, SS, offs, rowp, colp):
    tot=0; worst=0
    for parity in (0,1):
      for t in (0,1):
        for half in (0,1):
            banks={}
            for l in range(half*16, half*16+16):
                a=rowp[l>>2]; q=l&3; b=colp[2*q+t]
                col=(ndy(a))*TZ+ndz(a)
                rx=ndx(b)-ndx(a); ry=ndy(b)-ndy(a); rz=ndz(b)-ndz(a)
                if ndx(a)==0: sg = (0 if parity==0 else 1) if rx==0 else 2
                else: sg = (1 if parity==0 else 0) if rx==0 else 3
                s9=(ry+1)*3+rz+1
                addr=offs[sg]+col*CSTR+s9*SS
                banks.setdefault(addr%16,set()).add((sg,addr))
            m=max(len(v) for v in banks.values()); tot+=m; worst=max(worst,m)
    return tot/8
random.seed(1)
best=(9,None)
ident=list(range(8))
results=[]
for CSTR,SS in ()  and ((81,3),(9,1),(10,1),(11,1),(12,1),(13,1),(17,1),(27,3),(28,3),(29,3),(83,3),(85,3)):
    bestc=(9,)
    for trial in range(300):
        rp=ident[:]; cp=ident[:]
        random.shuffle(rp); random.shuffle(cp)
        offs=[0]+[random.randrange(16) for _ in range(3)]
        c=cost(CSTR,SS,offs,rp,cp)
        # hill climb
        improved=True
        while improved:
            improved=False
            for i in range(8):
                for j in range(i+1,8):
                    for which in (0,1):
                        P=(rp if which==0 else cp)
                        P[i],P[j]=P[j],P[i]
                        c2=cost(CSTR,SS,offs,rp,cp)
                        if c2<c: c=c2; improved=True
                        else: P[i],P[j]=P[j],P[i]
            for k in (1,2,3):
                for v in range(16):
                    old=offs[k]; offs[k]=v
                    c2=cost(CSTR,SS,offs,rp,cp)
                    if c2<c: c=c2; improved=True
                    else: offs[k]=old
        if c<bestc[0]: bestc=(c,CSTR,SS,tuple(offs),tuple(rp),tuple(cp))
        if c<=1.0: break
    print(bestc, flush=True)
print("same-perm search")
for TZv in (5,7):
    TZ=TZv
    for CSTR,SS in ((81,3),(83,3),(85,3),(87,3),(89,3),(82,3),(84,3)):
        bestc=(9,)
        for rp in itertools.permutations(range(8)):
            if rp[0]>3: continue
            for offs in ((0,0,8,8),(0,8,0,8),(0,4,12,8),(0,7,8,15),(0,1,2,3),(0,8,4,12),(0,0,4,4),(0,2,8,10)):
                c=cost(CSTR,SS,offs,rp,rp)
                if c<bestc[0]: bestc=(c,CSTR,SS,offs,rp)
            if bestc[0]<=1.0: break
        print("TZ",TZ,bestc,flush=True)
