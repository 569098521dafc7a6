"""CPU ORACLE — test infrastructure, not product code.

ctypes binding of ``oracle/lqr_oracle.c`` (the C restatement of the LQR.jl hot
path) plus ``oracle.dense_kkt`` (an independent refined sparse KKT solve that
plays the role of ``src/sparse_solver.jl:267-292`` / ``test/cholesky_solve.jl:42``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(``lqr.jl_b200``) never does.

Parity pin status: the reference keeps no golden vectors and Julia is absent,
so the oracle is pinned by the reference's own test identities (see
``tests/test_oracle.py``), not by reference outputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblqr_oracle.so")
_lib = None

HESS_DENSE, HESS_BLOCKDIAG, HESS_DIAG = 0, 1, 2
FLAG_SOC, FLAG_LTI = 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lqr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liblqr_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.lqro_riccati_work_doubles.restype = C.c_size_t
        _lib.lqro_num_threads.restype = C.c_int
    return _lib


def _p(a):
    if a is None:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def num_threads() -> int:
    return int(lib().lqro_num_threads())


# --------------------------------------------------------------------------
# Riccati.  Arrays are instance-major; within an instance every small matrix
# is column-major (Julia order), so numpy shapes read reversed:
#   A: (batch, N-1, n, n)  with A[b,k,j,i] = A_k[i,j]   (i.e. stored transposed)
# To keep call sites readable the helpers below take *mathematical* numpy
# arrays (A[b,k] is the n x n matrix) and do the transposition here.
# --------------------------------------------------------------------------
def _cm(a, nd=2):
    """math-order (..., r, c) -> column-major contiguous buffer."""
    if a is None:
        return None
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))


def riccati(prob: dict, nthreads: int = 0):
    """Batched Riccati on a problem dict (see lqr_b200.problems.riccati_*).

    Returns X (batch,N,n), U (batch,N-1,m), K (batch,N-1,m,n), kff (batch,N-1,m), info.
    """
    n, m, N = prob["n"], prob["m"], prob["N"]
    lti = bool(prob.get("lti", False))
    A, B, Q, R = (_cm(prob[k]) for k in ("A", "B", "Q", "R"))
    q, r = _f64(prob.get("q")), _f64(prob.get("r"))
    Qf, qf, x0 = _cm(prob["Qf"]), _f64(prob.get("qf")), _f64(prob["x0"])
    batch = x0.shape[0]
    X = np.zeros((batch, N, n))
    U = np.zeros((batch, N - 1, m))
    K = np.zeros((batch, N - 1, n, m))  # column-major m x n per knot
    kff = np.zeros((batch, N - 1, m))
    info = np.zeros(batch, dtype=np.int32)
    lib().lqro_riccati_batch(C.c_int(n), C.c_int(m), C.c_int(N), C.c_int(FLAG_LTI if lti else 0),
                             C.c_long(batch), _p(A), _p(B), _p(Q), _p(R), _p(q), _p(r), _p(Qf),
                             _p(qf), _p(x0), _p(X), _p(U), _p(K), _p(kff), _p(info),
                             C.c_int(nthreads))
    return X, U, np.swapaxes(K, -1, -2).copy(), kff, info


def riccati_raw(n, m, N, flags, batch, A, B, Q, R, q, r, Qf, qf, x0, X, U, K, kff, info, nthreads=0):
    """Zero-copy call on column-major instance-major buffers (bench cpu_baseline leg)."""
    lib().lqro_riccati_batch(C.c_int(n), C.c_int(m), C.c_int(N), C.c_int(flags), C.c_long(batch),
                             _p(A), _p(B), _p(Q), _p(R), _p(q), _p(r), _p(Qf), _p(qf), _p(x0),
                             _p(X), _p(U), _p(K), _p(kff), _p(info), C.c_int(nthreads))


# --------------------------------------------------------------------------
# KKT chain.  Problem dict (math-order numpy, batch leading):
#   Q (b,N,n,n) R (b,N-1,m,m) Hux (b,N-1,m,n)|None q (b,N,n) r (b,N-1,m)
#   A (b,N-1,n,n) B (b,N-1,n,m) d (b,N-1,n)
#   p: int32[N]; C: list over knots of (b,p_k,w_k); c: list of (b,p_k)
#   D2: None or list over k=1..N-1 of (b,n,w_{k})
# --------------------------------------------------------------------------
def kkt_flatten(prob: dict):
    """math-order dict -> the column-major instance-major flat buffers of the C ABI."""
    n, m, N = prob["n"], prob["m"], prob["N"]
    b = prob["q"].shape[0]
    p = np.ascontiguousarray(prob["p"], dtype=np.int32)
    Cf = [np.swapaxes(np.asarray(Ck, dtype=np.float64), -1, -2).reshape(b, -1)
          for Ck in prob["C"]]
    cf = [np.asarray(ck, dtype=np.float64).reshape(b, -1) for ck in prob["c"]]
    Cflat = np.ascontiguousarray(np.concatenate(Cf, axis=1)) if Cf else np.zeros((b, 0))
    cflat = np.ascontiguousarray(np.concatenate(cf, axis=1)) if cf else np.zeros((b, 0))
    D2 = prob.get("D2")
    if D2 is not None:
        D2 = np.ascontiguousarray(np.concatenate(
            [np.swapaxes(np.asarray(x, dtype=np.float64), -1, -2).reshape(b, -1) for x in D2], axis=1))
    return dict(n=n, m=m, N=N, batch=b, p=p, hess_mode=int(prob.get("hess_mode", HESS_BLOCKDIAG)),
                Q=_cm(prob["Q"]), R=_cm(prob["R"]), Hux=_cm(prob.get("Hux")),
                q=_f64(prob["q"]), r=_f64(prob["r"]), A=_cm(prob["A"]), B=_cm(prob["B"]),
                d=_f64(prob["d"]), D2=D2, C=Cflat, c=cflat)


def kkt_solve(prob: dict, soc: bool = False, nthreads: int = 0, want_res: bool = False):
    """Batched reference-algorithm KKT solve.  Returns dz (b,NN), mult (b,P), info[, res]."""
    f = prob if "batch" in prob else kkt_flatten(prob)
    n, m, N, b = f["n"], f["m"], f["N"], f["batch"]
    NN = N * n + (N - 1) * m
    P = int(f["p"].sum()) + (N - 1) * n
    dz = np.zeros((b, NN))
    mult = np.zeros((b, P))
    res = np.zeros((b, NN)) if want_res else None
    info = np.zeros(b, dtype=np.int32)
    lib().lqro_kkt_batch(C.c_int(n), C.c_int(m), C.c_int(N), C.c_int(f["hess_mode"]),
                         C.c_int(FLAG_SOC if soc else 0), _p(f["p"]), C.c_long(b), _p(f["Q"]),
                         _p(f["R"]), _p(f["Hux"]), _p(f["q"]), _p(f["r"]), _p(f["A"]), _p(f["B"]),
                         _p(f["d"]), _p(f["D2"]), _p(f["C"]), _p(f["c"]), _p(dz), _p(mult),
                         _p(res), _p(info), C.c_int(nthreads))
    return (dz, mult, info, res) if want_res else (dz, mult, info)


def kkt_solve_dense_outputs(prob: dict, i: int = 0, soc: bool = False):
    """Single instance i with the dense S, h, U extractors (get_shur_factors / get_cholesky)."""
    f = prob if "batch" in prob else kkt_flatten(prob)
    n, m, N = f["n"], f["m"], f["N"]
    NN = N * n + (N - 1) * m
    P = int(f["p"].sum()) + (N - 1) * n
    dz, mult, res = np.zeros(NN), np.zeros(P), np.zeros(NN)
    S, h, U = np.zeros((P, P)), np.zeros(P), np.zeros((P, P))

    def row(a):
        return None if a is None else np.ascontiguousarray(a[i])
    keep = [row(f[k]) for k in ("Q", "R", "Hux", "q", "r", "A", "B", "d", "D2", "C", "c")]
    info = lib().lqro_kkt_solve_flat(C.c_int(n), C.c_int(m), C.c_int(N), C.c_int(f["hess_mode"]),
                                     C.c_int(FLAG_SOC if soc else 0), _p(f["p"]),
                                     *[_p(a) for a in keep], _p(dz), _p(mult), _p(res), _p(S),
                                     _p(h), _p(U))
    # C wrote column-major P x P: transpose to math order
    return dict(dz=dz, mult=mult, res=res, S=S.T.copy(), h=h, U=U.T.copy(), info=int(info))


def block_cholesky(n, m, mode, A, B, Cc=None):
    """src/block_cholesky.jl:55-91.  Returns (M math-order (w,w), info)."""
    w = n + m
    M = np.zeros((w, w))
    info = lib().lqro_block_cholesky(C.c_int(n), C.c_int(m), C.c_int(mode), _p(_cm(A)), _p(_cm(B)),
                                     _p(_cm(Cc)), _p(M))
    return M.T.copy(), int(info)


def block_ldiv(n, m, mode, M, b):
    """src/block_cholesky.jl:93-96; b is (w,) or (w,nrhs) math order."""
    w = n + m
    bb = np.asarray(b, dtype=np.float64)
    vec = bb.ndim == 1
    buf = np.array(bb.reshape(w, -1).T, dtype=np.float64, order="C", copy=True)  # column-major w x nrhs
    lib().lqro_block_ldiv(C.c_int(n), C.c_int(m), C.c_int(mode), _p(_cm(M)), C.c_int(buf.shape[0]),
                          _p(buf), C.c_int(w))
    out = buf.T.copy()
    return out[:, 0] if vec else out
