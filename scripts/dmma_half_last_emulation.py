"""32-lane emulation of the H-step block sweep (hstep_dmma.cu) on a W x W matrix padded with the identity to 8 NB:
checks that leaving out the second k4 half of every product of the LAST sweep step (HALF_LAST, W - 8 (NB - 1) <= 4)
changes nothing -- the operands it would have multiplied are exactly zero -- and that the result is -B^-1.
Same lane-level helpers as scripts/dmma_block_sweep_emulation.py.  Run: python scripts/dmma_half_last_emulation.py"""
import numpy as np

lane = np.arange(32)


def shfl(v, src):
    return v[src]


def nform(t, h):
    src = (lane & ~3) | (2 * h + ((lane & 3) >> 1))
    return np.where(lane & 1, shfl(t[1], src), shfl(t[0], src))


def tform(t, h):
    src = 4 * (4 * h + (lane & 3)) + (lane >> 3)
    return np.where((lane >> 2) & 1, shfl(t[1], src), shfl(t[0], src))


def to_tile(M):
    r, c0 = lane >> 2, 2 * (lane & 3)
    return [M[r, c0].copy(), M[r, c0 + 1].copy()]


def from_tile(t):
    M = np.zeros((8, 8))
    r, c0 = lane >> 2, 2 * (lane & 3)
    M[r, c0], M[r, c0 + 1] = t[0], t[1]
    return M


def dmma(c, a, b):
    A, B = np.zeros((8, 4)), np.zeros((4, 8))
    A[lane >> 2, lane & 3] = a
    B[lane & 3, lane >> 2] = b
    return to_tile(from_tile(c) + A @ B)


def tile_spd_inverse(t):
    t = [t[0].copy(), t[1].copy()]
    r, c0 = lane >> 2, 2 * (lane & 3)
    for p in range(8):
        comp = t[1] if p & 1 else t[0]
        d = shfl(comp, np.full(32, 4 * p + (p >> 1)))
        cr = shfl(comp, (lane & ~3) | (p >> 1))
        pc0, pc1 = shfl(t[0], 4 * p + (lane & 3)), shfl(t[1], 4 * p + (lane & 3))
        assert d[0] > 0
        pinv = 1 / d
        crp = cr * pinv
        nx, ny = t[0] - crp * pc0, t[1] - crp * pc1
        nx, ny = np.where(r == p, pc0 * pinv, nx), np.where(r == p, pc1 * pinv, ny)
        nx = np.where(c0 == p, np.where(r == p, -pinv, crp), nx)
        ny = np.where(c0 + 1 == p, np.where(r == p, -pinv, crp), ny)
        t = [nx, ny]
    return [-t[0], -t[1]]


def tix(i, j):
    return i * (i + 1) // 2 + j


def sweep(B, NB, half_last):
    """Returns (-B^-1 assembled from the tiles, number of DMMA issued, max |operand| that half_last leaves out)."""
    A = [None] * (NB * (NB + 1) // 2)
    for i in range(NB):
        for j in range(i + 1):
            A[tix(i, j)] = to_tile(B[8 * i:8 * i + 8, 8 * j:8 * j + 8])
    n_dmma, skipped = 0, 0.0
    for kb in range(NB):
        hi = not (kb == NB - 1 and half_last)
        P = tile_spd_inverse(A[tix(kb, kb)])
        Pt, Pn = [tform(P, 0), tform(P, 1)], [nform(P, 0), nform(P, 1)]
        V = {}
        for m in range(NB):
            if m != kb:
                V[m] = [nform(A[tix(m, kb)], h) if m > kb else tform(A[tix(kb, m)], h) for h in (0, 1)]
                if kb == NB - 1:
                    skipped = max(skipped, np.abs(V[m][1]).max())
        for m in range(NB):
            if m == kb:
                continue
            T = [np.zeros(32), np.zeros(32)]
            for h in ((0, 1) if hi else (0,)):
                T = dmma(T, V[m][h], Pt[h]) if m > kb else dmma(T, Pn[h], V[m][h])
                n_dmma += 1
            A[tix(m, kb) if m > kb else tix(kb, m)] = T
        for i in range(NB):
            if i == kb:
                continue
            Tn = [-(nform(A[tix(i, kb)], h) if i > kb else tform(A[tix(kb, i)], h)) for h in (0, 1)]
            if kb == NB - 1:
                skipped = max(skipped, np.abs(Tn[1]).max())
            for j in range(i + 1):
                if j == kb:
                    continue
                for h in ((0, 1) if hi else (0,)):
                    A[tix(i, j)] = dmma(A[tix(i, j)], Tn[h], V[j][h])
                    n_dmma += 1
        A[tix(kb, kb)] = [-P[0], -P[1]]
    n = 8 * NB
    R = np.zeros((n, n))
    for i in range(NB):
        for j in range(i + 1):
            R[8 * i:8 * i + 8, 8 * j:8 * j + 8] = from_tile(A[tix(i, j)])
            if i != j:
                R[8 * j:8 * j + 8, 8 * i:8 * i + 8] = from_tile(A[tix(i, j)]).T
    return R, n_dmma, skipped


def main():
    rng = np.random.default_rng(0)
    for W in (50, 49, 52, 44, 9, 12):
        NB = (W + 7) // 8
        n = 8 * NB
        assert W - 8 * (NB - 1) <= 4
        X = rng.standard_normal((W, 70))
        d = np.sqrt(rng.random(W) * 3)
        B = np.eye(n)
        B[:W, :W] += d[:, None] * (0.1 * X @ X.T) * d[None, :]          # I + d K d on the real rows, identity padding
        full, n_full, skipped = sweep(B, NB, half_last=False)
        half, n_half, _ = sweep(B, NB, half_last=True)
        err = np.abs(-full[:W, :W] - np.linalg.inv(B[:W, :W])).max()
        print("W=%d NB=%d: DMMA %d -> %d, operands left out max |.| = %g, identical = %s, |-R - inv(B)| = %.2e"
              % (W, NB, n_full, n_half, skipped, np.array_equal(full, half), err))
        assert skipped == 0.0 and np.array_equal(full, half) and err < 1e-12


if __name__ == "__main__":
    main()
