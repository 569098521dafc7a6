"""GPU: the C ABI driven by a plain C99 program (tests/abi_smoke.c) — no Python on the calling side — on the
reference's cartpole fixture (config 1, test/problems.jl:58-88); its output file is checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from lqr_b200 import _lib, ops, problems

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_caller_solves_the_cartpole_fixture(tmp_path, oracle_mod, build_abi_smoke):
    prob = problems.cartpole_fixture()
    f = ops.kkt_flatten(prob)
    n, m, N, b = f["n"], f["m"], f["N"], f["batch"]
    src = tmp_path / "problem.bin"
    with open(src, "wb") as fh:
        np.array([n, m, N, b, f["hess_mode"], f["C"].shape[1], f["c"].shape[1], 0], dtype=np.int32).tofile(fh)
        f["p"].tofile(fh)
        for k in ("Q", "R", "q", "r", "A", "B", "d", "C", "c"):
            np.ascontiguousarray(f[k], dtype=np.float64).tofile(fh)
    exe = build_abi_smoke(str(tmp_path))
    out = tmp_path / "result.bin"
    r = subprocess.run([exe, str(src), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI_SMOKE_OK" in r.stdout, (r.stdout, r.stderr)
    assert "kkt_tpi<4,1" in r.stdout and "riccati_tpi<4,1>" in r.stdout, r.stdout
    NN, P = _lib.num_vars(n, m, N), _lib.num_cons(n, N, f["p"])
    raw = np.fromfile(out, dtype=np.float64)
    dz, mult = raw[:NN * b].reshape(b, NN), raw[NN * b:(NN + P) * b].reshape(b, P)
    dzo, lamo, _ = oracle_mod.kkt_solve(prob)
    rel = lambda a, c: np.linalg.norm(a - c) / np.linalg.norm(c)  # noqa: E731
    assert rel(dz, dzo) <= 1e-10 and rel(mult, lamo) <= 1e-10
    # the LTI Riccati leg against the oracle's reference-form DPSolver
    Z = raw[(2 * NN + P) * b:]
    pr = dict(n=n, m=m, N=N, lti=True, A=prob["A"][:1, 0], B=prob["B"][:1, 0], Q=prob["Q"][:1, 0], R=prob["R"][:1, 0],
              q=None, r=None, Qf=prob["Q"][:1, N - 1], qf=None, x0=0.1 * np.arange(1, n + 1)[None])
    Xo, Uo, _, _, _ = oracle_mod.riccati(pr)
    X, U = ops.split_primals(Z[None], n, m, N)
    assert rel(X, Xo) <= 1e-10 and rel(U, Uo) <= 1e-10
