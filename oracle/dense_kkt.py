"""CPU ORACLE (test infrastructure): the global-matrix form of the KKT solve.

Plays the role of ``src/sparse_solver.jl:267-292`` and of the dense check
``[H D'; D 0] \\ [-g; -d]`` at ``test/cholesky_solve.jl:42-44``: assemble the
global Hessian ``H`` (``src/jacobian_blocks.jl:73-89``), gradient ``g``
(``src/cholesky_solver.jl:289-304``) and linearised constraints ``D, d``
(``src/conblocks.jl:100-113``: row groups ``[C_1; D_1; C_2; D_2; …; C_N]``,
D_k carrying D1_k in knot k's columns and D2_{k+1} in knot k+1's), then solve
the saddle-point system with a sparse LU and extended-precision (x87 long
double) iterative refinement, so that a 1e-10 verdict is about the kernel under
test and not about this oracle.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def assemble(prob: dict, i: int = 0, soc: bool = False):
    """Global (H, g, D, d) of instance ``i`` of a math-order KKT problem dict."""
    n, m, N = prob["n"], prob["m"], prob["N"]
    p = np.asarray(prob["p"])
    mode = int(prob.get("hess_mode", 1))
    NN = N * n + (N - 1) * m
    P = int(p.sum()) + (N - 1) * n
    H = sp.lil_matrix((NN, NN))
    g = np.zeros(NN)
    D = sp.lil_matrix((P, NN))
    d = np.zeros(P)
    zoff, roff = 0, 0
    for k in range(N):
        w = n + (m if k < N - 1 else 0)
        ix = slice(zoff, zoff + n)
        iu = slice(zoff + n, zoff + w)
        if soc:  # Ginv=false: H = I, g = 0 (src/jacobian_blocks.jl:233-239)
            H[zoff:zoff + w, zoff:zoff + w] = np.eye(w)
        else:
            Q = prob["Q"][i, k]
            if mode == 2:
                Q = np.diag(np.diag(Q))
            H[ix, ix] = Q
            g[ix] = prob["q"][i, k]
            if k < N - 1:
                R = prob["R"][i, k]
                if mode == 2:
                    R = np.diag(np.diag(R))
                H[iu, iu] = R
                g[iu] = prob["r"][i, k]
                if mode == 0 and prob.get("Hux") is not None:
                    H[iu, ix] = prob["Hux"][i, k]
                    H[ix, iu] = prob["Hux"][i, k].T
        ps = int(p[k])
        if ps:
            D[roff:roff + ps, zoff:zoff + w] = prob["C"][k][i]
            d[roff:roff + ps] = prob["c"][k][i]
            roff += ps
        if k < N - 1:
            D[roff:roff + n, ix] = prob["A"][i, k]
            D[roff:roff + n, iu] = prob["B"][i, k]
            wn = n + (m if k + 1 < N - 1 else 0)
            if prob.get("D2") is not None:
                D[roff:roff + n, zoff + w:zoff + w + wn] = prob["D2"][k][i]
            else:
                D[roff:roff + n, zoff + w:zoff + w + n] = -np.eye(n)
            d[roff:roff + n] = prob["d"][i, k]
            roff += n
        zoff += w
    return sp.csc_matrix(H), g, sp.csc_matrix(D), d


def _ld_matvec(Kcoo, x):
    out = np.zeros(Kcoo.shape[0], dtype=np.longdouble)
    np.add.at(out, Kcoo.row, Kcoo.data.astype(np.longdouble) * x[Kcoo.col])
    return out


def solve_refined(H, g, D, d, iters: int = 6):
    """Solve [H D';D 0][dz;lam] = -[g;d] (test/cholesky_solve.jl:42), refined in long double."""
    NN, P = H.shape[0], D.shape[0]
    Kmat = sp.bmat([[H, D.T], [D, None]], format="csc")
    rhs = -np.concatenate([g, d])
    lu = spla.splu(Kmat)
    Kcoo = Kmat.tocoo()
    x = lu.solve(rhs).astype(np.longdouble)
    rhs_ld = rhs.astype(np.longdouble)
    for _ in range(iters):
        r = rhs_ld - _ld_matvec(Kcoo, x)
        x = x + lu.solve(np.asarray(r, dtype=np.float64)).astype(np.longdouble)
    x = np.asarray(x, dtype=np.float64)
    return x[:NN], x[NN:]


def kkt_truth(prob: dict, i: int = 0, soc: bool = False):
    """(dz*, lam*) of instance i in the reference's orderings (Primals / [μ1;λ1;…;μN])."""
    H, g, D, d = assemble(prob, i, soc)
    return solve_refined(H, g, D, d)


def kkt_residuals(prob: dict, i: int, dz, lam, soc: bool = False):
    """The two residual norms the reference asserts (test/cholesky_solve.jl:39-40), relative form
    of SURVEY §8d: ||H dz + g + D'lam|| / max(1,||g||), ||D dz + d|| / max(1,||d||)."""
    H, g, D, d = assemble(prob, i, soc)
    rs = H @ dz + g + D.T @ lam
    rp = D @ dz + d
    return (np.linalg.norm(rs) / max(1.0, np.linalg.norm(g)),
            np.linalg.norm(rp) / max(1.0, np.linalg.norm(d)))


def riccati_as_kkt(prob: dict):
    """Re-express a Riccati problem (LTV, affine cost) as the KKT problem with only the
    initial-condition and dynamics constraints (that equality is the parity
    test for config 2).  c_1 = x_1 - x0 with the step taken from z = 0, so dz IS the trajectory."""
    n, m, N = prob["n"], prob["m"], prob["N"]
    b = prob["x0"].shape[0]
    lti = bool(prob.get("lti", False))

    def knots(a):
        a = np.asarray(a)
        return np.repeat(a[:, None], N - 1, axis=1) if lti else a
    Q = np.concatenate([knots(prob["Q"]), prob["Qf"][:, None]], axis=1)
    qk = knots(prob["q"]) if prob.get("q") is not None else np.zeros((b, N - 1, n))
    qf = prob["qf"] if prob.get("qf") is not None else np.zeros((b, n))
    q = np.concatenate([qk, qf[:, None]], axis=1)
    r = knots(prob["r"]) if prob.get("r") is not None else np.zeros((b, N - 1, m))
    p = np.zeros(N, dtype=np.int32)
    p[0] = n
    C0 = np.zeros((b, n, n + m))
    C0[:, :, :n] = np.eye(n)
    Cs = [C0] + [np.zeros((b, 0, n + (m if k < N - 1 else 0))) for k in range(1, N)]
    cs = [-prob["x0"]] + [np.zeros((b, 0)) for _ in range(1, N)]
    return dict(n=n, m=m, N=N, p=p, hess_mode=1, Q=Q, R=knots(prob["R"]), Hux=None, q=q, r=r,
                A=knots(prob["A"]), B=knots(prob["B"]), d=np.zeros((b, N - 1, n)), D2=None,
                C=Cs, c=cs)
