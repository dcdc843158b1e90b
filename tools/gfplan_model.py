from gfplan_mat import *
import itertools, random
def subset(j,i): return (j & ~i)==0
def transform(C,L):
    R=len(C); K=len(C[0]); B=1<<L
    Rb=R//B*B; Kb=K//B*B   # leftover rows/cols untransformed
    # A = H^{(x)L} on aligned blocks of outputs, same on inputs; C' = A C B (B involutory)
    def Hrow(n,i):  # row i of block-diag H on n elements (blocks of size B; leftover identity)
        blk=i//B
        if i>=n//B*B: return [1 if j==i else 0 for j in range(n)]
        return [1 if (j//B==blk and subset(j%B,i%B)) else 0 for j in range(n)]
    A=[Hrow(R,i) for i in range(R)]; Bm=[Hrow(K,i) for i in range(K)]
    AC=[[0]*K for _ in range(R)]
    for i in range(R):
        for j in range(R):
            if A[i][j]:
                for c in range(K): AC[i][c]^=C[j][c]
    # y = Bm x ; x = Bm y ; C x = C Bm y -> (C Bm)[r][c] = XOR_j C[r][j] Bm[j][c]
    Cp=[[0]*K for _ in range(R)]
    for r in range(R):
        for j in range(K):
            for c in range(K):
                if Bm[j][c]: Cp[r][c]^=AC[r][j]
    return Cp,A,Bm
def order_inputs(K,L):
    B=1<<L; nb=K//B
    o=[]
    for typ in sorted(range(B), key=lambda t:-bin(t).count('1')):
        for b in range(nb): o.append(b*B+typ)
    o+=list(range(nb*B,K))
    return o
def cost(Cp,L,GS,K,R,verbose=False):
    order=order_inputs(K,L)
    groups=[order[i:i+GS] for i in range(0,K,GS)]
    B=1<<L
    tcost=(K//B)*(L*B//2) + (R//B)*(L*B//2)   # butterflies (2-input XORs; ptxas merges some)
    combos=set(); steps=0
    for r in range(R):
        top=max([c.bit_length() for c in Cp[r]]+[0])-1
        if top<0: continue
        for plane in range(top,-1,-1):
            n=0
            for gi,g in enumerate(groups):
                idx=tuple(t for t in g if (Cp[r][t]>>plane)&1)
                if idx:
                    n+=1
                    if len(idx)>1: combos.add((gi,idx))
            if plane==top: steps+= (n-1+1)//2 if n>1 else 0
            else: steps+= 3 + (max(n-1,0)+1)//2
    # combos cost: each multi-input combo ~1 op if built incrementally (upper bound len-1)
    ccost=sum(1 if len(i)<=3 else 2 for _,i in combos)
    if verbose: print("   transform",tcost,"combos",ccost,"steps",steps)
    return tcost+ccost+steps
def best(C,name):
    R=len(C);K=len(C[0])
    res=[]
    for L in range(0,4):
        if (1<<L)>R or (1<<L)>K: break
        Cp,_,_=transform(C,L)
        for GS in (3,4):
            res.append((cost(Cp,L,GS,K,R),L,GS))
    res.sort()
    base=[c for c in res if c[1]==0]
    print(name,"best",res[0],"base",min(base), "all",sorted(set((c,l) for c,l,g in res)))
    return res[0]
def decode_rows(k,m,missing):
    M=coding(k,m); present=[i for i in range(k+m) if i not in missing][:k]
    sub=[M[i] for i in present]; invm=matinv(sub)
    rows=[]
    for ms in missing:
        if ms<k: rows.append(invm[ms])
        else:
            rows.append([ __import__('functools').reduce(lambda x,y:x^y,[mul(M[ms][t],invm[t][c]) for t in range(k)]) for c in range(k)])
    return rows
if __name__=="__main__":
    for k,m in [(12,4),(16,4),(8,8),(8,4),(4,2),(6,2),(2,2),(10,4),(7,5),(8,3),(5,3),(14,2)]:
        best(coding(k,m)[k:],"enc(%d,%d)"%(k,m))
    best(decode_rows(12,4,[0,1,2,3]),"dec(12,4){0,1,2,3}")
    best(decode_rows(12,4,[1,5,12,15]),"dec(12,4){1,5,12,15}")
    best(decode_rows(12,4,[0]),"dec(12,4){0}")
    best(decode_rows(16,4,[0,7,16,19]),"dec(16,4){0,7,16,19}")
    best(decode_rows(16,4,[0,1,2,3]),"dec(16,4){0,1,2,3}")
    best(decode_rows(8,8,[0,1,2,3,4,5,6,7]),"dec(8,8){0..7}")
    Cp,_,_=transform(coding(12,4)[12:],2)
    for r in Cp: print(' '.join('%02x'%v for v in r))
    cost(Cp,2,3,12,4,True)
