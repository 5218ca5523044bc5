import numpy as np
lane=np.arange(32)
def shfl(v, src): return v[src]
def nform(t,h):
    src=(lane & ~3) | (2*h + ((lane&3)>>1)); a=shfl(t[0],src); b=shfl(t[1],src); return np.where(lane&1,b,a)
def tform(t,h):
    src=4*(4*h+(lane&3))+(lane>>3); a=shfl(t[0],src); b=shfl(t[1],src); return np.where((lane>>2)&1,b,a)
def to_tile(M):  # 8x8 -> C layout
    r=lane>>2; c0=2*(lane&3); return [M[r,c0].copy(), M[r,c0+1].copy()]
def from_tile(t):
    M=np.zeros((8,8)); r=lane>>2; c0=2*(lane&3); M[r,c0]=t[0]; M[r,c0+1]=t[1]; return M
def dmma(c,a,b):
    # A[row=lane/4][col=lane%4]=a ; B[row=lane%4][col=lane/4]=b
    A=np.zeros((8,4)); B=np.zeros((4,8)); A[lane>>2,lane&3]=a; B[lane&3,lane>>2]=b
    C=from_tile(c)+A@B; return to_tile(C)
def tile_spd_inverse(t):
    t=[t[0].copy(),t[1].copy()]; r=lane>>2; c0=2*(lane&3); ok=True
    for p in range(8):
        comp=t[1] if p&1 else t[0]
        d=shfl(comp, np.full(32,4*p+(p>>1))); cr=shfl(comp,(lane&~3)|(p>>1))
        pc0=shfl(t[0],4*p+(lane&3)); pc1=shfl(t[1],4*p+(lane&3))
        ok=ok and (d[0]>0); pinv=1/d; crp=cr*pinv
        nx=t[0]-crp*pc0; ny=t[1]-crp*pc1
        nx=np.where(r==p,pc0*pinv,nx); ny=np.where(r==p,pc1*pinv,ny)
        nx=np.where(c0==p, np.where(r==p,-pinv,crp), nx)
        ny=np.where(c0+1==p, np.where(r==p,-pinv,crp), ny)
        t=[nx,ny]
    return [-t[0],-t[1]], ok
rng=np.random.default_rng(0)
X=rng.standard_normal((8,20)); S=X@X.T+np.eye(8)
P,ok=tile_spd_inverse(to_tile(S)); print("tile inv err", np.abs(from_tile(P)-np.linalg.inv(S)).max(), ok)
# block sweep
NB=7; n=8*NB
X=rng.standard_normal((n,70)); B=np.eye(n)+X@X.T*0.1
tix=lambda i,j: i*(i+1)//2+j
A=[None]*(NB*(NB+1)//2)
for i in range(NB):
    for j in range(i+1): A[tix(i,j)]=to_tile(B[8*i:8*i+8,8*j:8*j+8])
for kb in range(NB):
    P,ok=tile_spd_inverse(A[tix(kb,kb)]); assert ok
    Pn=[tform(P,0),tform(P,1)]; Pa=[nform(P,0),nform(P,1)]
    Vold={}
    for m in range(NB):
        if m==kb: continue
        Vold[m]=[nform(A[tix(m,kb)],h) if m>kb else tform(A[tix(kb,m)],h) for h in (0,1)]
    for m in range(NB):
        if m==kb: continue
        T=[np.zeros(32),np.zeros(32)]
        if m>kb:
            T=dmma(T,Vold[m][0],Pn[0]); T=dmma(T,Vold[m][1],Pn[1]); A[tix(m,kb)]=T
        else:
            T=dmma(T,Pa[0],Vold[m][0]); T=dmma(T,Pa[1],Vold[m][1]); A[tix(kb,m)]=T
    for i in range(NB):
        if i==kb: continue
        Tn=[-(nform(A[tix(i,kb)],h) if i>kb else tform(A[tix(kb,i)],h)) for h in (0,1)]
        for j in range(i+1):
            if j==kb: continue
            t=A[tix(i,j)]; t=dmma(t,Tn[0],Vold[j][0]); t=dmma(t,Tn[1],Vold[j][1]); A[tix(i,j)]=t
    A[tix(kb,kb)]=[-P[0],-P[1]]
R=np.zeros((n,n))
for i in range(NB):
    for j in range(i+1):
        R[8*i:8*i+8,8*j:8*j+8]=from_tile(A[tix(i,j)])
        if i!=j: R[8*j:8*j+8,8*i:8*i+8]=from_tile(A[tix(i,j)]).T
print("block sweep err", np.abs(-R-np.linalg.inv(B)).max())
