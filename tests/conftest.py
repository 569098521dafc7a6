import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def handle():
    """A live lqrb handle on cuda:0.  No fallback: fails if the library or the GPU is missing."""
    from lqr_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


@pytest.fixture(scope="session")
def build_abi_smoke():
    """Compiles tests/abi_smoke.c (plain C99 caller of the ABI) against include/lqrb200.h and the in-tree library."""
    import subprocess

    def build(out_dir):
        from lqr_b200 import _lib
        exe = os.path.join(out_dir, "abi_smoke")
        lib_dir = os.path.dirname(_lib.LIB_PATH)
        subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", lib_dir, "-llqrb200", "-lm",
                        f"-Wl,-rpath,{lib_dir}"], check=True, capture_output=True, text=True)
        return exe
    return build


def condensed_least_squares_gradient(A, B, Q, R, Qf, x0, U):
    """The reference's condensed form (src/least_squares.jl: build_toeplitz, buildAb!): x_{2..N} = T u + L x0 with
    T block-Toeplitz (T[i,j] = A^(i-j) B) and L[i] = A^i, cost |Hx (T u + L x0)|^2 + u' Hu u.  Returns the
    gradient  T' Qbar (T U + L x0) + Rbar U  whose vanishing is test/least_squares.jl:38."""
    import numpy as np
    n, m = B.shape
    K = U.shape[0]  # N - 1
    T = np.zeros((K * n, K * m))
    L = np.zeros((K * n, n))
    Ap = np.eye(n)
    pw = [np.eye(n)]
    for _ in range(K):
        pw.append(A @ pw[-1])
    for i in range(K):
        L[i * n:(i + 1) * n] = pw[i + 1]
        for j in range(i + 1):
            T[i * n:(i + 1) * n, j * m:(j + 1) * m] = pw[i - j] @ B
    Qbar = np.zeros((K * n, K * n))
    for i in range(K):
        Qbar[i * n:(i + 1) * n, i * n:(i + 1) * n] = Q if i < K - 1 else Qf
    Rbar = np.kron(np.eye(K), R)
    u = U.reshape(-1)
    return T.T @ (Qbar @ (T @ u + L @ x0)) + Rbar @ u, np.linalg.norm(T.T @ (Qbar @ (L @ x0)))
