"""Lane-level emulator of the warp-per-instance FP64 tensor-core KKT kernel (csrc/kkt_wp_kernels.cuh).

Every "register" is a numpy array of 32 lanes; mma884 / shfl follow the PTX semantics of mma.sync.m8n8k4.f64 and
shfl.sync.idx.  The functions below are written the way the CUDA kernel is written (same fragment layouts, same
operation order), so that the fragment algebra — C fragments reused as A / B operands, 2 x 2 block inverse, selection
-matrix transposes, physical index maps — is checked on the CPU against the oracle before any GPU time is spent.
Run: python tools/emu/kkt_wp_emu.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

LANE = np.arange(32)
G = LANE >> 2
Q = LANE & 3


def mma884(d0, d1, a, b):
    """D[g][2q+e] += sum_k A[g][k] B[k][c]; lane (g,q) holds a = A[g][q], b = B[q][g], d_e = D[g][2q+e]."""
    A = a.reshape(8, 4)                      # A[g][k]
    B = b.reshape(8, 4).T                    # b lane (c,k) = B[k][c]  ->  B[k][c]
    D = A @ B                                # 8 x 8
    d0 = d0 + D[G, 2 * Q]
    d1 = d1 + D[G, 2 * Q + 1]
    return d0, d1


def shfl(v, src):
    return v[src]


# ------------------------------------------------------------------ physical index maps
class Map:
    def __init__(self, n, m):
        assert n in (8, 12) and 1 <= m <= 4
        self.n, self.m = n, m

    def zmap(self, pos):
        """physical position -> index into z = [x; u]; -1 = unused"""
        n, m = self.n, self.m
        if pos < 8:
            return pos
        j = pos - 8
        if j % 2 == 0:
            return 8 + j // 2 if 8 + j // 2 < n else -1
        return n + (j - 1) // 2 if (j - 1) // 2 < m else -1

    def xmap(self, pos):
        """physical position -> state / constraint-row index; -1 = pad"""
        z = self.zmap(pos)
        return z if 0 <= z < self.n else -1


def tiles_from_dense(M16):
    """16 x 16 physical matrix -> tiles[rt][ct][e] (C-fragment layout)."""
    return [[[M16[8 * rt + G, 8 * ct + 2 * Q + e].copy() for e in range(2)] for ct in range(2)] for rt in range(2)]


def dense_from_tiles(T):
    M = np.zeros((16, 16))
    for rt in range(2):
        for ct in range(2):
            for e in range(2):
                M[8 * rt + G, 8 * ct + 2 * Q + e] = T[rt][ct][e]
    return M


def zeros_tiles():
    return [[[np.zeros(32), np.zeros(32)] for _ in range(2)] for _ in range(2)]


# ------------------------------------------------------------------ products  X * Z'
def product(X, Z, ksteps, acc=None, sign=1.0, tiles=((0, 0), (0, 1), (1, 0), (1, 1))):
    """out[rt][ct] (+)= sign * sum over contraction steps of X[rt][kt][e] (as A) x Z[ct][kt][e] (as B)  = X Z'.
    ksteps: list of (kt, e) contraction steps: full z space = 4 steps, x space = 3 steps (odd slots of tile 1 pad)."""
    out = zeros_tiles() if acc is None else acc
    for (rt, ct) in tiles:
        d0, d1 = out[rt][ct]
        for (kt, e) in ksteps:
            d0, d1 = mma884(d0, d1, sign * X[rt][kt][e], Z[ct][kt][e])
        out[rt][ct] = [d0, d1]
    return out


KZ = [(0, 0), (0, 1), (1, 0), (1, 1)]
KX = [(0, 0), (0, 1), (1, 0)]


def prod8(X, Z, sign=1.0, acc=None):
    """8 x 8 tiles: acc (+)= sign * X Z'."""
    d0, d1 = (np.zeros(32), np.zeros(32)) if acc is None else acc
    for e in range(2):
        d0, d1 = mma884(d0, d1, sign * X[e], Z[e])
    return [d0, d1]


def transpose8(X, scale=1.0, acc=None):
    """acc (+)= scale * X' through two selection-matrix MMAs (a C fragment read as a B fragment is the transpose)."""
    d0, d1 = (np.zeros(32), np.zeros(32)) if acc is None else acc
    s0 = np.where(G == 2 * Q, scale, 0.0)
    s1 = np.where(G == 2 * Q + 1, scale, 0.0)
    d0, d1 = mma884(d0, d1, s0, X[0])
    d0, d1 = mma884(d0, d1, s1, X[1])
    return [d0, d1]


def gj8c(a, skip_odd=False):
    """In-place Gauss-Jordan inverse of an SPD 8 x 8 tile in C-fragment layout (no pivoting).  skip_odd: the odd
    rows / columns are identity padding (x space, tile 1).  Returns (inverse, bad pivot index or 0, lo, hi)."""
    a0, a1 = a[0].copy(), a[1].copy()
    bad = 0
    pivs = []
    for kk in range(8):
        if skip_odd and (kk & 1):
            continue
        prow0 = shfl(a0, 4 * kk + Q)
        prow1 = shfl(a1, 4 * kk + Q)
        src = a1 if kk & 1 else a0
        pcol = shfl(src, 4 * G + (kk >> 1))
        piv = shfl(src, np.full(32, 4 * kk + (kk >> 1)))
        pivs.append(piv[0])
        if not (piv[0] > 0) and bad == 0:
            bad = kk + 1
        p = 1.0 / piv
        f = pcol * p
        rowk = G == kk
        c0 = (2 * Q) == kk
        c1 = (2 * Q + 1) == kk
        n0 = np.where(rowk, np.where(c0, p, prow0 * p), np.where(c0, -f, a0 - f * prow0))
        n1 = np.where(rowk, np.where(c1, p, prow1 * p), np.where(c1, -f, a1 - f * prow1))
        a0, a1 = n0, n1
    return [a0, a1], bad, pivs


def inv16(M, x_space):
    """Inverse of a symmetric positive definite 16 x 16 physical matrix (identity on its pad slots) by 2 x 2 block
    elimination on 8 x 8 tiles: 12 MMAs + two 8 x 8 Gauss-Jordan inverses."""
    I00, bad0, p0 = gj8c(M[0][0])
    T1 = prod8(I00, M[1][0])                   # I00 M01   (M10 = M01')
    T1t = prod8(M[1][0], I00)                  # M10 I00 = T1'
    S = prod8(M[1][0], T1t, sign=-1.0, acc=[M[1][1][0].copy(), M[1][1][1].copy()])   # M11 - M10 I00 M01
    N11, bad1, p1 = gj8c(S, skip_odd=x_space)
    N01 = prod8(T1, N11, sign=-1.0)            # -T1 N11
    N10 = prod8(N11, T1, sign=-1.0)            # -N11 T1'
    N00 = prod8(N01, T1, sign=-1.0, acc=[I00[0].copy(), I00[1].copy()])   # I00 + T1 N11 T1'
    bad = bad0 if bad0 else (8 + bad1 if bad1 else 0)
    return [[N00, N01], [N10, N11]], bad, p0 + p1


# ------------------------------------------------------------------ vectors (16 physical slots, in "shared memory")
def matvec_row(M, x):
    """out[8 rt + g] = sum_c M[8rt+g][c] x[c]  (each lane sums its 4 columns, then the quad reduces by shuffles)."""
    out = np.zeros(16)
    for rt in range(2):
        s = np.zeros(32)
        for ct in range(2):
            for e in range(2):
                s = s + M[rt][ct][e] * x[8 * ct + 2 * Q + e]
        s = s + shfl(s, LANE ^ 1)
        s = s + shfl(s, LANE ^ 2)
        out[8 * rt + G] = s          # every lane of the quad holds the sum; lane q == 0 stores it
    return out


def matvec_col(M, x):
    """out[c] = sum_r M[r][c] x[r]  (M' x): each lane multiplies by x[row], the 8 row groups reduce by shuffles."""
    out = np.zeros(16)
    for ct in range(2):
        for e in range(2):
            s = np.zeros(32)
            for rt in range(2):
                s = s + M[rt][ct][e] * x[8 * rt + G]
            s = s + shfl(s, LANE ^ 4)
            s = s + shfl(s, LANE ^ 8)
            s = s + shfl(s, LANE ^ 16)
            out[8 * ct + 2 * Q + e] = s
    return out


# ------------------------------------------------------------------ the solve
def solve_instance(prob, i, soc=False):
    """One instance of a KKT problem dict (p = [n, ps, ..., ps, n] with ps <= 4 stage rows on the interior knots,
    any Hessian mode, structural D2)."""
    n, m, N = prob["n"], prob["m"], prob["N"]
    mp = Map(n, m)
    zpos = [mp.zmap(p) for p in range(16)]
    xpos = [mp.xmap(p) for p in range(16)]
    xmask = np.array([1.0 if v >= 0 else 0.0 for v in xpos])
    padx = np.diag(1.0 - xmask)               # identity on the pad slots of x space

    def phys_H(k):
        w = n + (m if k < N - 1 else 0)
        H = np.zeros((w, w))
        if soc:
            H = np.eye(w)
        else:
            H[:n, :n] = prob["Q"][i, k]
            if k < N - 1:
                H[n:, n:] = prob["R"][i, k]
                if prob.get("Hux") is not None and int(prob.get("hess_mode", 1)) == 0:
                    H[n:, :n] = prob["Hux"][i, k]
                    H[:n, n:] = prob["Hux"][i, k].T
        M = np.eye(16)
        for a in range(16):
            for b in range(16):
                za, zb = zpos[a], zpos[b]
                if 0 <= za < w and 0 <= zb < w:
                    M[a, b] = H[za, zb]
                elif a != b:
                    M[a, b] = 0.0
        return M

    def phys_rows_z(Y):
        """(rows, w) matrix with rows indexed by state/constraint index, columns by z -> 16 x 16 physical"""
        M = np.zeros((16, 16))
        for a in range(16):
            for b in range(16):
                r, z = xpos[a], zpos[b]
                if r >= 0 and 0 <= z < Y.shape[1] and r < Y.shape[0]:
                    M[a, b] = Y[r, z]
        return M

    def phys_vec_z(v):
        o = np.zeros(16)
        for a in range(16):
            if 0 <= zpos[a] < len(v):
                o[a] = v[zpos[a]]
        return o

    def phys_vec_x(v):
        o = np.zeros(16)
        for a in range(16):
            if xpos[a] >= 0:
                o[a] = v[xpos[a]]
        return o

    def state_part(Hi):
        """x-space restriction of a z-space symmetric matrix: odd slots of tile 1 -> identity padding."""
        D = dense_from_tiles(Hi)
        D = D * np.outer(xmask, xmask) + padx
        return tiles_from_dense(D)

    recs = []
    Cp = zeros_tiles()
    dp = np.zeros(16)
    info = 0
    spread = 0.0
    for k in range(N):
        first, last = k == 0, k == N - 1
        gz = np.zeros(16) if soc else phys_vec_z(np.concatenate([prob["q"][i, k], prob["r"][i, k]]) if not last else prob["q"][i, k])
        if last:
            Fd = prob["C"][k][i]                       # C_N (n x n)
            dvec = phys_vec_x(prob["c"][k][i])
        else:
            Fd = np.concatenate([prob["A"][i, k], prob["B"][i, k]], axis=1)
            dvec = phys_vec_x(prob["d"][i, k])
        F = tiles_from_dense(phys_rows_z(Fd))
        Hi, bad, piv = inv16(tiles_from_dense(phys_H(k)), x_space=last)
        if bad and not info:
            info = (k + 1) * 1000 + bad
        hg = matvec_row(Hi, gz)
        TF = product(F, Hi, KZ)                        # F Hi   (Hi symmetric)
        Gm = product(TF, F, KZ, tiles=((0, 0), (0, 1), (1, 1)))   # F Hi F' (upper tiles)
        rho = matvec_row(F, hg) - dvec
        if first:
            Cd = prob["C"][0][i]
            Cc = tiles_from_dense(phys_rows_z(Cd))
            TC = product(Cc, Hi, KZ)
            Sig = product(TC, Cc, KZ)                  # C Hi C'
            Sd = dense_from_tiles(Sig) + padx
            Sig = tiles_from_dense(Sd)
            T = product(TF, Cc, KZ, sign=-1.0)         # -(F Hi C')
            y = (matvec_row(Cc, hg) - phys_vec_x(prob["c"][0][i])) * xmask
        else:
            Qi = state_part(Hi)
            Sig = [[[Cp[rt][ct][e] + Qi[rt][ct][e] for e in range(2)] for ct in range(2)] for rt in range(2)]
            T = TF                                     # state columns are selected by the 3-step contraction
            y = (dp - hg) * xmask
        Si, bad, piv = inv16(Sig, x_space=True)
        if bad and not info:
            info = (1000 + 100 + bad) if first else (k * 1000 + 200 + bad)
        spread = max(spread, max(piv) / min(piv))
        Z = product(T, Si, KX)                         # T Si
        v = matvec_row(Si, y) * xmask
        X = product(Z, T, KX, sign=-1.0, acc=Gm, tiles=((0, 0), (0, 1), (1, 1)))   # G - Z T'
        P00 = transpose8(X[0][0], 0.5, acc=[0.5 * X[0][0][0], 0.5 * X[0][0][1]])
        P11 = transpose8(X[1][1], 0.5, acc=[0.5 * X[1][1][0], 0.5 * X[1][1][1]])
        P10 = transpose8(X[0][1], 1.0)
        Cp = [[P00, X[0][1]], [P10, P11]]
        Tv = matvec_row(T, v)
        dp = (rho + Tv) * xmask
        stage = None
        ps = int(prob["p"][k]) if not (first or last) else 0
        if ps:
            # stage rows C (ps x w), c: vectors only.  tc_j = Hi C_j' (z space), B = C Hi C', E_j = F tc_j (x space),
            # D_j = -(tc_j)_x;  eliminate lam_{k-1}: sd_j = Si D_j, B' = B - D'Si D, E'_j = E_j + T_x sd_j,
            # c'_j = (C hg - c)_j - sd_j'y;  eliminate mu_k: Bi = B'^-1, Cp -= E'' Bi E', dp -= E'' Bi c'.
            Cd = prob["C"][k][i]
            cz = [phys_vec_z(Cd[j]) for j in range(ps)]
            tc = [matvec_row(Hi, cz[j]) for j in range(ps)]
            Bm = np.array([[cz[j] @ tc[jp] for jp in range(ps)] for j in range(ps)])
            Ej = [matvec_row(F, tc[j]) * xmask for j in range(ps)]
            Dj = [-tc[j] * xmask for j in range(ps)]
            sd = [matvec_row(Si, Dj[j]) * xmask for j in range(ps)]
            ct = np.array([cz[j] @ hg - prob["c"][k][i][j] for j in range(ps)])
            Bp = Bm - np.array([[Dj[j] @ sd[jp] for jp in range(ps)] for j in range(ps)])
            Ep = [Ej[j] + matvec_row(T, sd[j]) * xmask for j in range(ps)]
            cp_ = ct - np.array([sd[j] @ y for j in range(ps)])
            Bi = np.linalg.inv(Bp)              # ps <= 4: every lane does it redundantly in the kernel
            if not np.all(np.linalg.eigvalsh(Bp) > 0) and not info:
                info = (k + 1) * 1000 + 100 + 1
            Wj = [sum(Bi[j, jp] * Ep[jp] for jp in range(ps)) for j in range(ps)]
            Cd_ = dense_from_tiles(Cp)
            for j in range(ps):
                Cd_ -= np.outer(Ep[j], Wj[j])
            Cd_ = 0.5 * (Cd_ + Cd_.T)
            Cp = tiles_from_dense(Cd_)
            bc = Bi @ cp_
            dp = dp - sum(bc[j] * Ep[j] for j in range(ps))
            stage = (sd, Bi, Ep, cp_, cz)
        recs.append((Z, v, stage))
    # last block: mu_N' = Cp^-1 dp
    Cd = dense_from_tiles(Cp) + padx
    Bl, bad, piv = inv16(tiles_from_dense(Cd), x_space=True)
    if bad and not info:
        info = N * 1000 + 100 + bad
    xcur = matvec_row(Bl, dp) * xmask

    NN = N * n + (N - 1) * m
    psm = int(prob["p"][1]) if N > 2 else 0
    P = 2 * n + (N - 1) * n + (N - 2) * psm
    dz, mult = np.zeros(NN), np.zeros(P)

    def lam_off(kk):     # offset of lam_kk (0-based knot), kk = 0 .. N-2
        return n + kk * (n + psm)
    xinv = [p for p in range(16) if xpos[p] >= 0]
    mult[P - n:] = -xcur[xinv]
    for k in range(N - 1, -1, -1):
        first, last = k == 0, k == N - 1
        Z, v, stage = recs[k]
        xprev = (v + matvec_col(Z, xcur)) * xmask      # x_{k-1} = v_k + Z_k' x_k
        xi = None
        if stage is not None:
            sd, Bi, Ep, cp_, cz = stage
            xi = Bi @ (cp_ - np.array([Ep[j] @ xcur for j in range(len(sd))]))   # mu_k' = B'^-1 (c' - E' x_k)
            xprev = (xprev - sum(xi[j] * sd[j] for j in range(len(sd)))) * xmask
        gz = np.zeros(16) if soc else phys_vec_z(np.concatenate([prob["q"][i, k], prob["r"][i, k]]) if not last else prob["q"][i, k])
        Fd = prob["C"][k][i] if last else np.concatenate([prob["A"][i, k], prob["B"][i, k]], axis=1)
        F = tiles_from_dense(phys_rows_z(Fd))
        res = gz - matvec_col(F, xcur)                 # g + D1' lam_k,  lam = -x
        if xi is not None:
            res = res - sum(xi[j] * cz[j] for j in range(len(xi)))   # C' mu_k, mu = -xi
        if not first:
            res = res + xprev                          # D2' lam_{k-1} = +x_{k-1} on the state slots
        else:
            Cc = tiles_from_dense(phys_rows_z(prob["C"][0][i]))
            res = res - matvec_col(Cc, xprev)
        Hi, _, _ = inv16(tiles_from_dense(phys_H(k)), x_space=last)
        z = -matvec_row(Hi, res)
        w = n + (0 if last else m)
        for ppos in range(16):
            zz = zpos[ppos]
            if 0 <= zz < w:
                dz[k * (n + m) + zz] = z[ppos]
        if first:
            mult[:n] = -xprev[xinv]
        else:
            mult[lam_off(k - 1): lam_off(k - 1) + n] = -xprev[xinv]
        if xi is not None:
            mult[lam_off(k - 1) + n: lam_off(k - 1) + n + psm] = -xi
        xcur = xprev
    return dz, mult, info, spread


def main():
    import oracle
    from lqr_b200 import problems
    oracle.build()
    worst = 0.0
    for (n, m, N, b, hess, mid_p) in [(12, 4, 9, 2, 1, 0), (12, 4, 30, 2, 2, 0), (8, 4, 12, 2, 1, 0), (12, 2, 12, 1, 1, 0),
                                      (8, 1, 14, 1, 1, 0), (12, 4, 5, 1, 1, 0), (12, 4, 12, 2, 0, 0), (12, 4, 12, 2, 1, 1),
                                      (12, 4, 15, 1, 1, 2), (8, 4, 12, 1, 0, 1), (12, 4, 12, 1, 1, 3)]:
        prob = problems.random_lqr_kkt(n, m, N, b, seed=n + N, mid_p=mid_p, hess_mode=hess)
        if hess == 2:
            for key in ("Q", "R"):
                prob[key] = prob[key] * np.eye(prob[key].shape[-1])
        for soc in (False, True):
            dzo, lamo, io = oracle.kkt_solve(prob, soc=soc)
            for i in range(b):
                dz, lam, info, spread = solve_instance(prob, i, soc=soc)
                e = max(np.linalg.norm(dz - dzo[i]) / np.linalg.norm(dzo[i]), np.linalg.norm(lam - lamo[i]) / np.linalg.norm(lamo[i]))
                worst = max(worst, e)
                print(f"n={n} m={m} N={N} hess={hess} mid_p={mid_p} soc={soc} inst={i}: rel err {e:.2e} info {info} log2 pivot ratio {np.log2(spread):.1f}")
    assert worst < 1e-9, worst
    print("EMULATOR_OK worst", worst)


if __name__ == "__main__":
    main()
