"""32-lane NumPy emulation of the DMMA fragment logic used by the E-step r x r phase (csrc/estep_seg.cu,
factor_variance_dmma): Gram matrix G' diag(w) G, blocked sweep, variance diag(G Minv G').  Validates the index math
against plain NumPy before it runs on the GPU."""
import numpy as np

lane = np.arange(32)
r_, c0_ = lane >> 2, 2 * (lane & 3)


def dmma(c, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[lane >> 2, lane & 3] = a
    B[lane & 3, lane >> 2] = b
    C[r_, c0_] = c[0]; C[r_, c0_ + 1] = c[1]
    C = C + A @ B
    return [C[r_, c0_].copy(), C[r_, c0_ + 1].copy()]


def nform(t, h):
    src = (lane & ~3) | (2 * h + ((lane & 3) >> 1))
    return np.where(lane & 1, t[1][src], t[0][src])


def tform(t, h):
    src = 4 * (4 * h + (lane & 3)) + (lane >> 3)
    return np.where((lane >> 2) & 1, t[1][src], t[0][src])


def tile_inv(t):
    t = [t[0].copy(), t[1].copy()]
    for p in range(8):
        comp = t[1] if p & 1 else t[0]
        d = comp[np.full(32, 4 * p + (p >> 1))]; cr = comp[(lane & ~3) | (p >> 1)]
        pc0 = t[0][4 * p + (lane & 3)]; pc1 = t[1][4 * p + (lane & 3)]
        pinv = 1 / d; crp = cr * pinv
        nx = t[0] - crp * pc0; ny = t[1] - crp * pc1
        nx = np.where(r_ == p, pc0 * pinv, nx); ny = np.where(r_ == p, pc1 * pinv, ny)
        nx = np.where(c0_ == p, np.where(r_ == p, -pinv, crp), nx)
        ny = np.where(c0_ + 1 == p, np.where(r_ == p, -pinv, crp), ny)
        t = [nx, ny]
    return [-t[0], -t[1]]


tix = lambda i, j: i * (i + 1) // 2 + j
rng = np.random.default_rng(1)
W, nc = 50, 11
NB = (nc + 7) // 8
ldg = nc | 1
G = np.zeros((W, ldg)); G[:, :nc] = rng.standard_normal((W, nc)) * 0.3
w = rng.random(W)


def gload(t, c):      # predicated fragment load of the compact factor
    ok = (t < W) & (c < nc)
    return np.where(ok, G[np.minimum(t, W - 1), np.minimum(c, ldg - 1)], 0.0)


# ---- Gram: A-operand element [i = 8 ti + lane/4][t = 4 k + lane%4] = G[t][i] w[t]; B-operand [t][j = 8 tj + lane/4] ----
A = [[np.zeros(32), np.zeros(32)] for _ in range(NB * (NB + 1) // 2)]
for k in range((W + 3) // 4):
    t = 4 * k + (lane & 3)
    wt = np.where(t < W, w[np.minimum(t, W - 1)], 0.0)
    g = [gload(t, 8 * b + (lane >> 2)) for b in range(NB)]
    for i in range(NB):
        for j in range(i + 1):
            A[tix(i, j)] = dmma(A[tix(i, j)], g[i] * wt, g[j])
for i in range(NB):      # + identity (also on the padding so that the sweep stays well defined)
    A[tix(i, i)][0] = A[tix(i, i)][0] + (r_ == c0_)
    A[tix(i, i)][1] = A[tix(i, i)][1] + (r_ == c0_ + 1)
ref = np.eye(8 * NB); ref[:nc, :nc] += G[:, :nc].T @ (w[:, None] * G[:, :nc])
got = np.zeros((8 * NB, 8 * NB))
for i in range(NB):
    for j in range(i + 1):
        blk = np.zeros((8, 8)); blk[r_, c0_] = A[tix(i, j)][0]; blk[r_, c0_ + 1] = A[tix(i, j)][1]
        got[8 * i:8 * i + 8, 8 * j:8 * j + 8] = blk
        got[8 * j:8 * j + 8, 8 * i:8 * i + 8] = blk.T
print("gram err", np.abs(got - ref).max())
# ---- blocked sweep (same as the H-step kernel) ----
for kb in range(NB):
    P = tile_inv(A[tix(kb, kb)])
    Pt = [tform(P, 0), tform(P, 1)]; Pn = [nform(P, 0), nform(P, 1)]
    V = {m: [nform(A[tix(m, kb)], h) if m > kb else tform(A[tix(kb, m)], h) for h in (0, 1)] for m in range(NB) if m != kb}
    for m in V:
        T = [np.zeros(32), np.zeros(32)]
        if m > kb:
            T = dmma(T, V[m][0], Pt[0]); T = dmma(T, V[m][1], Pt[1]); A[tix(m, kb)] = T
        else:
            T = dmma(T, Pn[0], V[m][0]); T = dmma(T, Pn[1], V[m][1]); A[tix(kb, m)] = T
    for i in V:
        Tn = [-(nform(A[tix(i, kb)], h) if i > kb else tform(A[tix(kb, i)], h)) for h in (0, 1)]
        for j in range(i + 1):
            if j == kb:
                continue
            A[tix(i, j)] = dmma(dmma(A[tix(i, j)], Tn[0], V[j][0]), Tn[1], V[j][1])
    A[tix(kb, kb)] = [-P[0], -P[1]]
# ---- store M = -Minv to a row-major SMEM image with ld = 8 NB + 4 (both triangles) ----
ldm = 8 * NB + 4
M = np.zeros((8 * NB, ldm))
for i in range(NB):
    for j in range(i + 1):
        M[8 * i + r_, 8 * j + c0_] = A[tix(i, j)][0]; M[8 * i + r_, 8 * j + c0_ + 1] = A[tix(i, j)][1]
        M[8 * j + c0_, 8 * i + r_] = A[tix(i, j)][0]; M[8 * j + c0_ + 1, 8 * i + r_] = A[tix(i, j)][1]
print("inverse err", np.abs(-M[:nc, :nc] - np.linalg.inv(ref[:nc, :nc])).max())
# ---- variance: T = G M (row tiles of 8 bins), v_t = - sum_j T[t][j] G[t][j] ----
v = np.zeros(W)
Bop = [[M[4 * k + (lane & 3), 8 * jt + (lane >> 2)] for jt in range(NB)] for k in range(2 * NB)]   # [c][j]
for tt in range((W + 7) // 8):
    trow = 8 * tt + (lane >> 2)
    aop = [gload(trow, 4 * k + (lane & 3)) for k in range(2 * NB)]
    acc = np.zeros(32)
    for jt in range(NB):
        T = [np.zeros(32), np.zeros(32)]
        for k in range(2 * NB):
            T = dmma(T, aop[k], Bop[k][jt])
        acc = acc + T[0] * gload(trow, 8 * jt + c0_) + T[1] * gload(trow, 8 * jt + c0_ + 1)
    acc = acc + acc[lane ^ 1]; acc = acc + acc[lane ^ 2]
    sel = ((lane & 3) == 0) & (trow < W)
    v[trow[sel]] = -acc[sel]
vref = np.einsum("ti,ij,tj->t", G[:, :nc], np.linalg.inv(ref[:nc, :nc]), G[:, :nc])
print("variance err", np.abs(v - vref).max())
