"""Host-side mirror of the LQR.jl interface for the hot path, batched.

Same names, argument meaning and error behaviour as the reference; Julia's ``f!`` is spelled ``f_``.
Every object carries a leading batch axis (a single instance is batch = 1) and every numerical call
goes through the C ABI (``include/lqrb200.h``) to the CUDA kernels — there is no CPU path here.

  reference (file:line)                                   mirror
  ------------------------------------------------------  -----------------------------------------
  LQRProblem, size, num_vars   src/lqr_problem.jl:1-25     LQRProblem, size, num_vars
  Primals                      src/lqr_problem.jl:46-73    Primals
  DPSolver, solve!             src/dynamic_programming.jl  DPSolver, LQRSolution, solve_
  rollout!                     src/least_squares.jl:195    rollout_
  BlockCholesky, cholesky!,    src/block_cholesky.jl:19-   BlockCholesky, cholesky_, ldiv_, ldiv
    ldiv!, \\                     101
  InvertedQuadratic,           src/block_cholesky.jl:107-  InvertedQuadratic, update_cost_,
    update_cholesky!, gradient   159                         update_cholesky_, gradient
  ConstraintBlock(s), dims,    src/conblocks.jl:36-113     ConstraintBlock, ConstraintBlocks, dims,
    copy_blocks!, num_constraints                            copy_blocks_, num_constraints
  build_shur_factors, calculate_shur_factors!, cholesky!(U,F), forward/backward_substitution!,
    calculate_primals!         src/jacobian_blocks.jl:155-286, src/cholesky_solve.jl:28-143,
                               src/cholesky_solver.jl:185-236   same names (fused on device, see
                                                                CholeskySolver docstring)
  CholeskySolver, _solve!, solve!, step!, residual, second_order_correction!, get_step,
    get_multipliers, get_residual, get_linearized_constraints, get_cost_expansion
                               src/cholesky_solver.jl:39-363    CholeskySolver and functions below
"""
from __future__ import annotations

import numpy as np

from . import _lib, ops
from ._lib import HESS_BLOCKDIAG, HESS_DENSE, HESS_DIAG, Handle, LqrbError

__all__ = [
    "LQRProblem", "Primals", "LQRSolution", "DPSolver", "LeastSquaresSolver", "solve_", "rollout_", "size", "num_vars",
    "BlockCholesky", "cholesky_", "ldiv_", "ldiv", "InvertedQuadratic", "update_cost_", "update_cholesky_",
    "gradient", "ConstraintBlock", "ConstraintBlocks", "dims", "copy_blocks_", "gen_con_inds", "num_constraints",
    "CholeskySolver", "build_shur_factors", "calculate_shur_factors_", "forward_substitution_",
    "backward_substitution_", "calculate_primals_", "residual", "second_order_correction_", "get_step",
    "get_multipliers", "get_residual", "get_linearized_constraints", "get_cost_expansion", "step_",
    "get_shur_factors", "get_cholesky", "copy_shur_factors_",
    "Handle", "LqrbError", "HESS_DENSE", "HESS_BLOCKDIAG", "HESS_DIAG",
]


def _b(a, nd):
    """Promote an array with `nd` trailing problem axes to a batched (batch, ...) float64 array."""
    a = np.asarray(a, dtype=np.float64)
    return a[None] if a.ndim == nd else a


# ===================================================================== LQRProblem / Primals
class LQRProblem:
    """Time-invariant LQR data Qf,Q,R,A,B,x0,u0,tf,N (src/lqr_problem.jl:1-11), batched; pass per-knot
    arrays (an extra axis of length N-1 after the batch axis) for the LTV generalisation, and q,r,qf
    for affine cost terms.  Diagonal Q/R may be given as vectors."""

    def __init__(self, Qf, Q, R, A, B, x0, u0=None, tf=1.0, N=2, q=None, r=None, qf=None):
        A = np.asarray(A, dtype=np.float64)
        self.ltv = A.ndim == 4
        nd = 3 if self.ltv else 2
        self.A = _b(A, nd)
        self.B = _b(B, nd)
        n, m = self.A.shape[-1], self.B.shape[-1]

        def mat(M, k, ndv):
            M = np.asarray(M, dtype=np.float64)
            if M.shape[-1] == k and (M.ndim < 2 or M.shape[-2] != k):  # a Diagonal given as a vector
                M = M[..., None] * np.eye(k)
            return _b(M, ndv)
        self.Q, self.R = mat(Q, n, nd), mat(R, m, nd)
        self.Qf = mat(Qf, n, 2)
        self.x0 = _b(x0, 1)
        self.u0 = None if u0 is None else _b(u0, 1)
        self.q = None if q is None else _b(q, nd - 1)
        self.r = None if r is None else _b(r, nd - 1)
        self.qf = None if qf is None else _b(qf, 1)
        self.tf, self.N = float(tf), int(N)
        self.batch = max(a.shape[0] for a in (self.A, self.B, self.Q, self.R, self.Qf, self.x0))
        for name in ("A", "B", "Q", "R", "Qf", "x0", "q", "r", "qf"):
            a = getattr(self, name)
            if a is not None and a.shape[0] != self.batch:
                if a.shape[0] != 1:
                    raise ValueError(f"{name}: batch axis {a.shape[0]} does not match {self.batch}")
                setattr(self, name, np.broadcast_to(a, (self.batch,) + a.shape[1:]).copy())
        if self.ltv and self.A.shape[1] != self.N - 1:
            raise ValueError("LTV arrays need N-1 knots")

    def as_dict(self):
        n, m, N = size(self)
        z = lambda a, shp: np.zeros(shp) if a is None else a  # noqa: E731
        kn = (self.batch, N - 1) if self.ltv else (self.batch,)
        return dict(n=n, m=m, N=N, lti=not self.ltv, A=self.A, B=self.B, Q=self.Q, R=self.R,
                    q=z(self.q, kn + (n,)), r=z(self.r, kn + (m,)), Qf=self.Qf,
                    qf=z(self.qf, (self.batch, n)), x0=self.x0)


def size(prob):
    """Base.size(prob) = (n, m, N)  (src/lqr_problem.jl:21; src/cholesky_solver.jl:105)."""
    if isinstance(prob, LQRProblem):
        return prob.A.shape[-1], prob.B.shape[-1], prob.N
    return prob.n, prob.m, prob.N


def num_vars(prob):
    """N*n + (N-1)*m  (src/lqr_problem.jl:22-25; src/cholesky_solver.jl:104)."""
    n, m, N = size(prob)
    return _lib.num_vars(n, m, N)


class Primals:
    """One flat vector Z = [x1;u1;x2;u2;...;xN] per instance with per-knot views X[k], U[k]
    (src/lqr_problem.jl:46-73)."""

    def __init__(self, n, m, N, tf=1.0, batch=1):
        self.n, self.m, self.N, self.tf = n, m, N, tf
        self.Z = np.zeros((batch, _lib.num_vars(n, m, N)))

    @property
    def X(self):
        n, m, N = self.n, self.m, self.N
        return [self.Z[:, k * (n + m): k * (n + m) + n] for k in range(N)]

    @property
    def U(self):
        n, m, N = self.n, self.m, self.N
        return [self.Z[:, k * (n + m) + n: (k + 1) * (n + m)] for k in range(N - 1)]

    @property
    def X_(self):
        return np.stack(self.X, axis=1)

    @property
    def U_(self):
        return np.stack(self.U, axis=1)

    def copy(self):
        out = Primals(self.n, self.m, self.N, self.tf, self.Z.shape[0])
        out.Z[:] = self.Z
        return out

    def __add__(self, other):
        out = self.copy()
        out.Z += other.Z
        return out

    def __rmul__(self, a):
        out = self.copy()
        out.Z *= a
        return out


class LQRSolution:
    """K, X, U containers (the reference exports this name but never defines it; src/LQR.jl:19,
    used at src/dynamic_programming.jl:54)."""

    def __init__(self, prob: LQRProblem):
        n, m, N = size(prob)
        b = prob.batch
        self.K = np.zeros((b, N - 1, m, n))
        self.d = np.zeros((b, N - 1, m))   # affine feed-forward (zero for the reference's form)
        self.X = np.zeros((b, N, n))
        self.U = np.zeros((b, N - 1, m))
        self.info = np.zeros(b, dtype=np.int32)


# ===================================================================== Riccati
class DPSolver:
    """DPSolver(prob): Riccati workspace (src/dynamic_programming.jl:2-23).  Here the workspace is the
    library handle; P, PA, PB, ... live in registers / shared memory on the device."""

    def __init__(self, prob: LQRProblem, handle: Handle | None = None, device: int = 0):
        self.n, self.m, self.N = size(prob)
        self.handle = handle or ops.default_handle(device)


class LeastSquaresSolver:
    """LeastSquaresSolver(prob) (src/least_squares.jl:1-59): the condensed form of the unconstrained LTI problem —
    block-Toeplitz T, (T'QT + R) U = -T'Q L x0 by Cholesky, then rollout!.  The Toeplitz matrices live in a device
    workspace (lqrb_lsq_solve_f64); O((N m)^3) per instance, for short horizons and as a cross-check of the Riccati
    path (test/least_squares.jl:38)."""

    def __init__(self, prob: LQRProblem, handle: Handle | None = None, device: int = 0):
        if prob.ltv or prob.q is not None or prob.r is not None or prob.qf is not None:
            raise LqrbError("LeastSquaresSolver takes the reference's time-invariant LQRProblem without affine terms")
        self.n, self.m, self.N = size(prob)
        self.handle = handle or ops.default_handle(device)
        self.info = np.zeros(prob.batch, dtype=np.int32)


def solve_(sol, solver, prob=None):
    """solve!(sol, solver::DPSolver, prob) (src/dynamic_programming.jl:54-72): backward Riccati pass then
    forward rollout; solve!(sol::Primals, solver::LeastSquaresSolver, prob) (src/least_squares.jl:158-190); or
    solve!(solver::CholeskySolver) (src/cholesky_solver.jl:109-120) when called with a CholeskySolver."""
    if isinstance(sol, CholeskySolver):
        return sol.solve_()
    if isinstance(solver, LeastSquaresSolver):
        n, m, N = solver.n, solver.m, solver.N
        Z = np.zeros((prob.batch, _lib.num_vars(n, m, N)))
        ops.lsq_solve(solver.handle, n, m, N, prob.batch, ops.cm(prob.A), ops.cm(prob.B), ops.cm(prob.Q), ops.cm(prob.R),
                      ops.cm(prob.Qf), ops.f64(prob.x0), Z, solver.info)
        if isinstance(sol, Primals):
            sol.Z[:] = Z
        else:
            sol.X[:], sol.U[:] = ops.split_primals(Z, n, m, N)
            sol.info[:] = solver.info
        return sol
    X, U, K, kff, info = ops.riccati_solve_problem(prob.as_dict(), want_gains=True, handle=solver.handle)
    sol.X[:], sol.U[:], sol.K[:], sol.d[:], sol.info[:] = X, U, K, kff, info
    return sol


def rollout_(X, U, prob: LQRProblem, handle: Handle | None = None):
    """rollout!: X[1]=x0; X[k+1] = A X[k] + B U[k]  (src/least_squares.jl:195-202)."""
    h = handle or ops.default_handle()
    n, m, N = size(prob)
    b = prob.batch
    Xo = np.zeros((b, N, n))
    ops.rollout(h, n, m, N, b, 0 if prob.ltv else _lib.FLAG_LTI, ops.cm(prob.A), ops.cm(prob.B),
                ops.f64(prob.x0), ops.f64(np.broadcast_to(U, (b, N - 1, m))), Xo)
    X[...] = Xo
    return X


# ===================================================================== BlockCholesky
class BlockCholesky:
    """Cholesky of M = [A C'; C B] in three modes (src/block_cholesky.jl:19-52): dense (whole-matrix
    potrf), block_diag (C = 0, separate potrf), diag (Diagonal storage: keeps the inverse).  `M` holds
    the factor exactly like the reference's `chol.M` (upper triangle; for diag the reciprocals)."""

    def __init__(self, n, m, batch=1, diag=False, block_diag=False, uplo="U", handle=None):
        if uplo != "U":
            raise ValueError("only uplo='U' is supported (the reference's default, src/block_cholesky.jl:42)")
        self.n, self.m, self.batch = n, m, batch
        self.diag, self.block_diag, self.uplo = bool(diag), bool(block_diag) or bool(diag), uplo
        self.mode = HESS_DIAG if diag else (HESS_BLOCKDIAG if block_diag else HESS_DENSE)
        self.M = np.zeros((batch, n + m, n + m))     # math order
        self.info = np.zeros(batch, dtype=np.int32)
        self.handle = handle or ops.default_handle()

    @property
    def U(self):
        """chol.F.U"""
        return np.triu(self.M)


def cholesky_(chol, A, B=None, C=None):
    """cholesky!(chol, A, B[, C]) (src/block_cholesky.jl:55-91); or cholesky!(U, F) on Schur blocks
    (src/cholesky_solve.jl:28-33) when called with a CholeskySolver's block lists."""
    if isinstance(chol, _ShurBlocks):
        return chol.solver._stage("cholesky")      # lqrb_kkt_factor_f64
    n, m, b = chol.n, chol.m, chol.batch
    A = np.broadcast_to(_b(A, 2), (b, n, n))
    B = np.zeros((b, m, m)) if B is None else np.broadcast_to(_b(B, 2), (b, m, m))
    Cc = None
    if C is not None and chol.mode == HESS_DENSE:   # block_diag ignores C (:56-57)
        Cc = np.broadcast_to(_b(C, 2), (b, m, n))
    Mcm = np.zeros((b, n + m, n + m))
    ops.block_cholesky(chol.handle, n, m, b, chol.mode, ops.cm(A), ops.cm(B), ops.cm(Cc), Mcm, chol.info)
    chol.M[:] = np.swapaxes(Mcm, -1, -2)
    return chol


def ldiv_(chol: BlockCholesky, b):
    """ldiv!(chol, b): in-place solve (src/block_cholesky.jl:93-96); b is (batch, w) or (batch, w, nrhs)."""
    w = chol.n + chol.m
    bb = np.asarray(b)
    vec = bb.ndim == 2
    buf = np.ascontiguousarray(np.swapaxes(bb.reshape(chol.batch, w, -1), -1, -2), dtype=np.float64)
    ops.block_ldiv(chol.handle, chol.n, chol.m, chol.batch, chol.mode, ops.cm(chol.M), buf.shape[1], buf)
    out = np.swapaxes(buf, -1, -2)
    b[...] = out[..., 0] if vec else out
    return b


def ldiv(chol: BlockCholesky, b):
    """chol \\ b (src/block_cholesky.jl:98-101)."""
    return ldiv_(chol, np.array(b, dtype=np.float64, copy=True))


class InvertedQuadratic:
    """chol(H_k) + gradient (q, r) of one knot's cost expansion (src/block_cholesky.jl:107-124)."""

    def __init__(self, n, m, batch=1, diag=False, block_diag=True, handle=None):
        self.chol = BlockCholesky(n, m, batch, diag=diag, block_diag=block_diag, handle=handle)
        self.q = np.zeros((batch, n))
        self.r = np.zeros((batch, m))


def update_cost_(icost: InvertedQuadratic, Q, R, q, r, H=None):
    """update_cost!(icost, cost) (src/block_cholesky.jl:145-153)."""
    if icost.chol.block_diag or H is None:
        cholesky_(icost.chol, Q, R)
    else:
        cholesky_(icost.chol, Q, R, H)
    icost.q[:] = q
    if icost.chol.m:
        icost.r[:] = r
    return icost


def update_cholesky_(chols, costs):
    """update_cholesky!(chols, obj) over the horizon (src/block_cholesky.jl:155-159); `costs[k]` is a
    dict(Q=, R=, q=, r=[, H=])."""
    for icost, c in zip(chols, costs):
        update_cost_(icost, c["Q"], c.get("R"), c["q"], c.get("r"), c.get("H"))


def gradient(icost: InvertedQuadratic):
    """[q; r], or q at the terminal knot (src/block_cholesky.jl:126-132)."""
    return np.concatenate([icost.q, icost.r], axis=1) if icost.chol.m else icost.q.copy()


# ===================================================================== ConstraintBlock
class ConstraintBlock:
    """Y = [D2; C; D1] (rows n1+p+n2, width w), y = [c; d], with views D2, C, D1, c, d that alias Y and y
    (src/conblocks.jl:36-72)."""

    def __init__(self, n1, p, n2, w, batch=1):
        self.y = np.zeros((batch, p + n2))
        self.Y = np.zeros((batch, n1 + p + n2, w))
        self.D2 = self.Y[:, :n1]
        self.C = self.Y[:, n1:n1 + p]
        self.D1 = self.Y[:, n1 + p:]
        self.c = self.y[:, :p]
        self.d = self.y[:, p:]
        self.res = np.zeros((batch, w))


def dims(block):
    """(n1, p, n2)  (src/conblocks.jl:98; src/jacobian_blocks.jl:171)."""
    return block.D2.shape[1], block.C.shape[1], block.D1.shape[1]


def ConstraintBlocks(n, m, N, p, batch=1):
    """Per-knot block sizes for dynamics-coupled problems (src/conblocks.jl:74-96): n2 = n for k < N
    (the dynamics rows starting at k), n1 = n for k > 1, p = stage rows, w = n + m*(k<N).
    D2 is initialised to the structural [-I 0] (test/cartpole.jl:34-42)."""
    blocks = []
    for k in range(N):
        blk = ConstraintBlock(n if k > 0 else 0, int(p[k]), n if k < N - 1 else 0, n + (m if k < N - 1 else 0), batch)
        if k > 0:
            blk.D2[:, :, :n] = -np.eye(n)
        blocks.append(blk)
    return blocks


def gen_con_inds(cons, N, structure="by_knotpoint"):
    """gen_con_inds(conSet, structure) (src/conblocks.jl:122-166): index ranges of every constraint at every knot it
    applies to, in the concatenated constraint vector.  `cons` stands in for TrajOptCore's ConstraintList (un-vendored):
    a list of (length, knots) with `knots` a range of 0-based knot indices.  Returns cons_inds[i][j] = range for
    constraint i at its j-th knot.  structure: "by_constraint" | "by_knotpoint" | "by_block" (per-knot offsets, the order
    BlockConstraintSet uses, :186)."""
    out = [[range(0, 0) for _ in knots] for _, knots in cons]
    if structure == "by_constraint":
        idx = 0
        for i, (ln, knots) in enumerate(cons):
            for j, _ in enumerate(knots):
                out[i][j] = range(idx, idx + ln)
                idx += ln
    elif structure == "by_knotpoint":
        idx = 0
        for k in range(N):
            for i, (ln, knots) in enumerate(cons):
                if k in knots:
                    out[i][k - knots[0]] = range(idx, idx + ln)
                    idx += ln
    elif structure == "by_block":
        idx = [0] * N
        for k in range(N):
            for i, (ln, knots) in enumerate(cons):
                if k in knots:
                    out[i][k - knots[0]] = range(idx[k], idx[k] + ln)
                    idx[k] += ln
    else:
        raise ValueError(f"unknown structure {structure!r}")
    return out


def num_constraints(blocks):
    return sum(b.y.shape[1] for b in blocks)


def copy_blocks_(D, d, blocks, i=0):
    """copy_blocks!(D, d, blocks) (src/conblocks.jl:100-113) for instance i: row offset advances by n1+p,
    so D1_k and D2_{k+1} share rows."""
    off1 = off2 = 0
    for blk in blocks:
        n1, p, n2 = dims(blk)
        w = blk.Y.shape[2]
        D[off1:off1 + n1 + p + n2, off2:off2 + w] = blk.Y[i]
        d[off1 + n1:off1 + n1 + p + n2] = blk.y[i]
        off1 += n1 + p
        off2 += w
    return D, d


# ===================================================================== CholeskySolver
class _ShurBlocks:
    """The Vector{BlockUpperTriangular3} the reference passes between the five steps
    (src/jacobian_blocks.jl:95-169).  On the device the block rows are records kept by the handle
    (lqrb_kkt_factor_f64); this object names them for the step functions and gives dense read-back:
    ``kind`` = "shur" (the unfactored S blocks + h) or "chol" (the block rows of U)."""

    def __init__(self, solver, kind):
        self.solver = solver
        self.kind = kind

    def dense(self):
        """(M, h): dense S (kind "shur") or U (kind "chol") of every instance, read back from the device."""
        S, h, U = self.solver._dense_factors()
        return (S, h) if self.kind == "shur" else (U, h)


def build_shur_factors(conSet_or_solver, uplo="U"):
    """build_shur_factors(conSet, :U) (src/jacobian_blocks.jl:155-169)."""
    if uplo not in ("U", ":U"):
        raise ValueError("only the upper variant has substitution methods (src/cholesky_solve.jl:145-168)")
    return _ShurBlocks(conSet_or_solver, "shur")


class CholeskySolver:
    """Batched CholeskySolver (src/cholesky_solver.jl:39-86) over linearised block data.

    The reference fills its blocks through un-vendored TrajOptCore calls (update!, :155-164); this
    mirror takes the linearised data directly: `Jinv`-side cost blocks (Q, R, Hux, q, r with a
    hess_mode = BlockCholesky mode) and `conSet.blocks`-side constraint blocks (A, B, d, C, c, p[, D2]).

    _solve_() runs the whole chain of src/cholesky_solver.jl:166-182 in ONE fused kernel launch.  The five step
    functions of the reference's script (test/cholesky_solve.jl:14-35) are real device steps too, in two launches
    on the general kernel: calculate_shur_factors_ + cholesky_ = lqrb_kkt_factor_f64 (the handle keeps U; the
    unfactored S and U can be read back with get_shur_factors / get_cholesky), forward_substitution_ +
    backward_substitution_ + calculate_primals_ = lqrb_kkt_solve_factored_f64 (any number of right-hand sides
    against the kept factor, see solve_factored_).
    """

    def __init__(self, prob: dict, handle: Handle | None = None, device: int = 0):
        self.handle = handle or ops.default_handle(device)
        self.prob = prob
        self.flat = ops.kkt_flatten(prob)
        self.n, self.m, self.N = prob["n"], prob["m"], prob["N"]
        self.batch = self.flat["batch"]
        self.p = np.asarray(prob["p"], dtype=np.int32)
        self.shur_blocks = _ShurBlocks(self, "shur")
        self.chol_blocks = _ShurBlocks(self, "chol")
        NN, P = _lib.num_vars(self.n, self.m, self.N), _lib.num_cons(self.n, self.N, self.p)
        self.dZ = np.zeros((self.batch, NN))
        self.lam = np.zeros((self.batch, P))
        self.res = np.zeros((self.batch, NN))
        self.info = np.zeros(self.batch, dtype=np.int32)
        self._state = "new"
        self._Ginv = True

    # ---- data refresh (the L3 update! of the reference happens in the caller)
    def update_(self, **arrays):
        """Replace any of Q,R,Hux,q,r,A,B,d,C,c with new linearised data (update!, :155-164)."""
        self.prob.update(arrays)
        self.flat = ops.kkt_flatten(self.prob)
        self._state = "new"

    # ---- fused solve
    def _run(self, soc: bool):
        f = self.flat
        ops.kkt_solve(self.handle, self.n, self.m, self.N, self.batch, f["p"], f["hess_mode"],
                      _lib.FLAG_SOC if soc else 0, f["Q"], f["R"], f["Hux"], f["q"], f["r"], f["A"], f["B"],
                      f["d"], f["D2"], f["C"], f["c"], self.dZ, self.lam, self.res, self.info)
        self._state = "solved"

    def _solve_(self):
        """_solve!(solver) (src/cholesky_solver.jl:166-182)."""
        self._Ginv = True
        self._run(False)
        return self

    # ---- the reference's five steps as two device launches
    def _flags(self):
        return 0 if self._Ginv else _lib.FLAG_SOC

    def factor_(self):
        """calculate_shur_factors! + cholesky!(chol_blocks, shur_blocks): the handle keeps the block rows of U."""
        f = self.flat
        ops.kkt_factor(self.handle, self.n, self.m, self.N, self.batch, f["p"], f["hess_mode"], self._flags(),
                       f["Q"], f["R"], f["Hux"], f["A"], f["B"], f["D2"], f["C"], self.info)
        self._state = "factored"
        return self

    def solve_factored_(self, q=None, r=None, d=None, c=None):
        """forward_substitution! + backward_substitution! + calculate_primals! with the kept factor.  Without
        arguments the right-hand side is the solver's own (q, r, d, c); pass new ones to re-use the factor
        (src/cholesky_solver.jl:246,259-263).  c is the list over knots of (batch, p_k) arrays."""
        if self._state not in ("factored", "solved_factored"):
            self.factor_()
        f = self.flat
        cflat = f["c"] if c is None else np.ascontiguousarray(
            np.concatenate([np.asarray(ck, dtype=np.float64).reshape(self.batch, -1) for ck in c], axis=1))
        ops.kkt_solve_factored(self.handle, self.n, self.m, self.N, self.batch, f["p"], f["hess_mode"],
                               f["D2"] is not None, self._flags(), ops.f64(f["q"] if q is None else q),
                               ops.f64(f["r"] if r is None else r), ops.f64(f["d"] if d is None else d), cflat,
                               self.dZ, self.lam, self.res, None)
        self._state = "solved_factored"
        return self

    def _stage(self, step):
        if step == "shur":
            self._state = "staged"
        elif step == "cholesky":
            self.factor_()
        elif self._state != "solved_factored":
            self.solve_factored_()
        return self

    def _dense_factors(self):
        """Dense (S, h, U) of every instance computed by the DEVICE path (lqrb_kkt_get_shur_f64), math order."""
        f = self.flat
        P = self.lam.shape[1]
        S, U = np.zeros((self.batch, P, P)), np.zeros((self.batch, P, P))
        h = np.zeros((self.batch, P))
        ops.kkt_get_shur(self.handle, self.n, self.m, self.N, self.batch, f["p"], f["hess_mode"], self._flags(),
                         f["Q"], f["R"], f["Hux"], f["q"], f["r"], f["A"], f["B"], f["d"], f["D2"], f["C"], f["c"],
                         S, h, U, None)
        return S, h, np.swapaxes(U, -1, -2).copy()     # S is symmetric; U comes back column-major

    def second_order_correction_(self, d=None, c=None):
        """second_order_correction! (src/cholesky_solver.jl:254-273): the same chain with Ginv=false,
        i.e. dz^ = -D'(DD')^-1 d, on freshly evaluated constraint values (d, c) if given."""
        if d is not None or c is not None:
            upd = {}
            if d is not None:
                upd["d"] = d
            if c is not None:
                upd["c"] = c
            self.update_(**upd)
        self._Ginv = False
        self._run(True)
        return self.dZ

    def solve_(self):
        raise LqrbError("solve!(::CholeskySolver) needs the nonlinear model: use lqr_b200.sqp.DubinsSQP "
                        "(the on-device SQP driver) — linearisation is third-party code in the reference")


def calculate_shur_factors_(shur, Jinv=None, blocks=None, Ginv=True):
    """calculate_shur_factors!(F, Jinv, blocks[, Ginv]) (src/jacobian_blocks.jl:220-229).  On the device S is
    formed block row by block row inside the factor launch (cholesky_ below); this call fixes Ginv and marks the
    data as staged.  get_shur_factors(solver) returns the S and h the device forms."""
    shur.solver._Ginv = bool(Ginv)
    return shur.solver._stage("shur")


def forward_substitution_(chol):
    """forward_substitution!(chol) (src/cholesky_solve.jl:93-117) — first half of lqrb_kkt_solve_factored_f64."""
    return chol.solver._stage("forward")


def backward_substitution_(chol):
    """backward_substitution!(chol) (src/cholesky_solve.jl:119-143) — second half of the same launch."""
    return chol.solver._stage("backward")


def calculate_primals_(dZ, Jinv=None, chol=None, blocks=None):
    """calculate_primals!(dZ, Jinv, chol, blocks) (src/cholesky_solver.jl:185-199)."""
    chol.solver._stage("primals")
    dZ[...] = chol.solver.dZ
    return dZ


def get_shur_factors(solver: CholeskySolver):
    """get_shur_factors (src/cholesky_solver.jl:333-341): dense (S, h, lam) per instance, S and h as the DEVICE forms
    them (copy_shur_factors!, src/jacobian_blocks.jl:173-211), lam = the multipliers of the last solve."""
    S, h, _ = solver._dense_factors()
    return S, h, solver.lam


def get_cholesky(solver: CholeskySolver):
    """get_cholesky (src/cholesky_solver.jl:352-359): dense upper-triangular U with U'U = S, from the device."""
    return solver._dense_factors()[2]


def copy_shur_factors_(S, h, lam, blocks: _ShurBlocks):
    """copy_shur_factors!(S, h, lam, F) (src/jacobian_blocks.jl:173-180): dense image of the block rows `blocks`
    (solver.shur_blocks -> S, solver.chol_blocks -> U) of every instance; lam gets the kept multipliers."""
    M, hv = blocks.dense()
    S[...] = M
    h[...] = hv
    lam[...] = blocks.solver.lam
    return S, h, lam


def step_(solver: CholeskySolver):
    """One QP step of step! (src/cholesky_solver.jl:122-153) on already-linearised data: _solve! only;
    the merit update / line search of the reference are third-party (TO.line_search)."""
    return solver._solve_()


def residual(solver: CholeskySolver, recalculate=False):
    """residual(solver; recalculate) (src/cholesky_solver.jl:238-252): norm over knots of ||res_k||, res_k =
    D1'lam_k + C'mu_k + D2'lam_{k-1} + g_k, per instance.  recalculate=True evaluates it on the device from the
    solver's CURRENT blocks (after update_) and the multipliers kept from the last solve — the reference
    re-linearises first (:240-246), here the caller has already pushed the new blocks with update_()."""
    n, m, N = solver.n, solver.m, solver.N
    if recalculate:
        f = solver.flat
        norms = np.zeros(solver.batch)
        ops.kkt_residual(solver.handle, n, m, N, solver.batch, f["p"], 0 if solver._Ginv else _lib.FLAG_SOC,
                         f["q"], f["r"], f["A"], f["B"], f["D2"], f["C"], solver.lam, solver.res, norms)
        return norms
    r = solver.res
    body = r[:, :(N - 1) * (n + m)].reshape(solver.batch, N - 1, n + m)
    per_knot = np.concatenate([np.linalg.norm(body, axis=2), np.linalg.norm(r[:, (N - 1) * (n + m):], axis=1)[:, None]], axis=1)
    return np.linalg.norm(per_knot, axis=1)


def second_order_correction_(solver: CholeskySolver, **kw):
    return solver.second_order_correction_(**kw)


def get_step(solver: CholeskySolver):
    """get_step (src/cholesky_solver.jl:299-313): the flat primal step, Primals order."""
    return solver.dZ


def get_multipliers(solver: CholeskySolver):
    """get_multipliers (src/cholesky_solver.jl:343-350): flat [mu_1; lam_1; mu_2; ...; mu_N]."""
    return solver.lam


def get_residual(solver: CholeskySolver):
    """get_residual (src/cholesky_solver.jl:315-331)."""
    return solver.res


def get_linearized_constraints(solver: CholeskySolver, i=0):
    """Dense (D, d) of instance i (src/cholesky_solver.jl:278-285 -> copy_blocks!, src/conblocks.jl:100-113)."""
    pr, n, m, N = solver.prob, solver.n, solver.m, solver.N
    blocks = ConstraintBlocks(n, m, N, solver.p, 1)
    for k, blk in enumerate(blocks):
        blk.C[0] = pr["C"][k][i]
        blk.c[0] = pr["c"][k][i]
        if k < N - 1:
            blk.D1[0, :, :n] = pr["A"][i, k]
            blk.D1[0, :, n:] = pr["B"][i, k]
            blk.d[0] = pr["d"][i, k]
        if k > 0 and pr.get("D2") is not None:
            blk.D2[0] = pr["D2"][k - 1][i]
    P, NN = num_constraints(blocks), _lib.num_vars(n, m, N)
    return copy_blocks_(np.zeros((P, NN)), np.zeros(P), blocks, 0)


def get_cost_expansion(solver: CholeskySolver, i=0):
    """Dense (H, g) of instance i (src/cholesky_solver.jl:287-297; build_H!, src/jacobian_blocks.jl:73-89)."""
    pr, n, m, N = solver.prob, solver.n, solver.m, solver.N
    NN = _lib.num_vars(n, m, N)
    H, g = np.zeros((NN, NN)), np.zeros(NN)
    mode = int(pr.get("hess_mode", HESS_BLOCKDIAG))
    for k in range(N):
        o = k * (n + m)
        Q = pr["Q"][i, k]
        H[o:o + n, o:o + n] = np.diag(np.diag(Q)) if mode == HESS_DIAG else Q
        g[o:o + n] = pr["q"][i, k]
        if k < N - 1:
            R = pr["R"][i, k]
            H[o + n:o + n + m, o + n:o + n + m] = np.diag(np.diag(R)) if mode == HESS_DIAG else R
            g[o + n:o + n + m] = pr["r"][i, k]
            if mode == HESS_DENSE and pr.get("Hux") is not None:
                H[o + n:o + n + m, o:o + n] = pr["Hux"][i, k]
                H[o:o + n, o + n:o + n + m] = pr["Hux"][i, k].T
    return H, g
