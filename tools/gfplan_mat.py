import itertools
exp=[0]*512; log=[0]*256; x=1
for i in range(255):
    exp[i]=x; log[x]=i; x<<=1
    if x&0x100: x^=0x11d
for i in range(255,512): exp[i]=exp[i-255]
def mul(a,b): return 0 if a==0 or b==0 else exp[log[a]+log[b]]
def inv(a): return exp[255-log[a]]
def pw(a,n): return 1 if n==0 else (0 if a==0 else exp[(log[a]*n)%255])
def matinv(m):
    n=len(m); a=[row[:]+[1 if i==j else 0 for j in range(n)] for i,row in enumerate(m)]
    for r in range(n):
        if a[r][r]==0:
            for rb in range(r+1,n):
                if a[rb][r]: a[r],a[rb]=a[rb],a[r]; break
        s=inv(a[r][r]); a[r]=[mul(v,s) for v in a[r]]
        for r2 in range(n):
            if r2!=r and a[r2][r]:
                f=a[r2][r]; a[r2]=[v^mul(f,w) for v,w in zip(a[r2],a[r])]
    return [row[n:] for row in a]
def coding(k,m):
    vm=[[pw(r,c) for c in range(k)] for r in range(k+m)]
    top=matinv(vm[:k])
    return [[ __import__('functools').reduce(lambda x,y:x^y,[mul(vm[r][t],top[t][c]) for t in range(k)]) for c in range(k)] for r in range(k+m)]
if __name__=="__main__":
    M=coding(12,4)[12:]
    for r in M: print(' '.join('%02x'%v for v in r))
    for b in range(8):
        print(b,[sum((v>>b)&1 for v in r) for r in M])
