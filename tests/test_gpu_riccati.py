"""GPU parity: batched Riccati (C ABI -> CUDA) against the CPU oracle and the refined KKT truth.

Tolerance (BASELINE.json north_star): relative solution error <= 1e-10 in FP64."""
import numpy as np
import pytest

from lqr_b200 import _lib, ops, problems

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _check(prob, handle, oracle_mod, tol=TOL):
    X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=handle)
    Xo, Uo, Ko, kffo, infoo = oracle_mod.riccati(prob)
    assert (info == 0).all() and (infoo == 0).all()
    b = X.shape[0]
    worst = max(max(_rel(X[i], Xo[i]), _rel(U[i], Uo[i]), _rel(K[i], Ko[i]), _rel(kff[i], kffo[i]))
                for i in range(b))
    assert worst <= tol, (worst, handle.last_kernel)
    return X, U


@pytest.mark.parametrize("batch", [1, 31, 32, 33, 257])
def test_cartpole_ltv_vs_oracle(handle, oracle_mod, batch):
    prob = problems.riccati_cartpole_batch(batch, seed=batch)
    _check(prob, handle, oracle_mod)
    assert handle.last_kernel.startswith("riccati_tpi<4,1>")


def test_cartpole_vs_refined_kkt_truth(handle, oracle_mod):
    """config 2 parity: Riccati == KKT solution with only init + dynamics constraints."""
    from oracle import dense_kkt
    prob = problems.riccati_cartpole_batch(8, seed=0)
    X, U = _check(prob, handle, oracle_mod)
    kp = dense_kkt.riccati_as_kkt(prob)
    n, m, N = 4, 1, 101
    for i in range(8):
        zt, _ = dense_kkt.kkt_truth(kp, i)
        Xt, Ut = ops.split_primals(zt[None], n, m, N)
        assert _rel(X[i], Xt[0]) <= TOL and _rel(U[i], Ut[0]) <= TOL


@pytest.mark.parametrize("n,m,N", [(2, 1, 11), (3, 2, 201), (4, 2, 50), (6, 3, 31), (2, 2, 20), (3, 1, 33), (3, 3, 25), (4, 3, 30),
                                   (5, 1, 40), (5, 2, 31), (5, 3, 22), (6, 1, 35), (6, 2, 41)])
@pytest.mark.parametrize("lti", [False, True])
def test_tpi_sizes(handle, oracle_mod, n, m, N, lti):
    prob = problems.random_lqr_riccati(n, m, N, 70, seed=n * 10 + m, lti=lti)
    _check(prob, handle, oracle_mod)
    assert handle.last_kernel.startswith("riccati_tpi")


@pytest.mark.parametrize("n,m,N,batch", [(7, 2, 20, 9), (10, 3, 101, 10), (7, 7, 15, 5), (32, 7, 12, 3), (40, 12, 11, 2)])
def test_cooperative_sizes(handle, oracle_mod, n, m, N, batch):
    prob = problems.random_lqr_riccati(n, m, N, batch, seed=n)
    handle.set_option("riccati_pad", 0)  # (without it these sizes are embedded in a tuned size class, next test)
    try:
        _check(prob, handle, oracle_mod)
    finally:
        handle.set_option("riccati_pad", 1)
    assert handle.last_kernel.startswith("riccati_coop")


@pytest.mark.parametrize("lti", [False, True])
@pytest.mark.parametrize("n,m,N,batch,kern", [(7, 2, 20, 9, "riccati_dmma<8,2>"), (10, 3, 101, 33, "riccati_dmma<12,3>"),
                                              (5, 4, 30, 7, "riccati_dmma<8,4>"), (11, 1, 40, 5, "riccati_dmma<12,1>"),
                                              (7, 7, 15, 5, "riccati_cta_dmma<16,8>"), (14, 7, 25, 34, "riccati_cta_dmma<16,8>"),
                                              (20, 6, 21, 4, "riccati_cta_dmma<24,8>"), (32, 7, 12, 3, "riccati_cta_dmma<32,8>"),
                                              (40, 12, 11, 2, "riccati_cta_dmma<48,16>"), (60, 10, 9, 2, "riccati_cta_dmma<64,16>")])
def test_sizes_without_a_tuned_kernel_are_padded_into_one(handle, oracle_mod, n, m, N, batch, kern, lti):
    """A size with no tensor-core kernel of its own is embedded in the next tuned size class (pad states and controls that
    stay exactly zero); X, U, K, kff are the original problem's."""
    prob = problems.random_lqr_riccati(n, m, N, batch, seed=3 * n + m, lti=lti)
    _check(prob, handle, oracle_mod)
    assert handle.last_kernel.startswith(kern) and handle.last_kernel.endswith(f"<- ({n},{m}) padded"), handle.last_kernel


@pytest.mark.parametrize("n,m,N,batch", [(12, 4, 101, 10), (12, 4, 2, 3), (12, 4, 7, 133), (8, 4, 60, 9), (12, 1, 40, 5),
                                         (8, 1, 33, 6), (12, 4, 1001, 4), (12, 2, 50, 7), (12, 3, 41, 34), (8, 2, 30, 5),
                                         (8, 3, 33, 33), (12, 3, 2, 3)])
@pytest.mark.parametrize("lti", [False, True])
def test_dmma_sizes(handle, oracle_mod, n, m, N, batch, lti):
    """warp-per-instance FP64 tensor-core kernel (config 5a shape and its siblings)."""
    prob = problems.random_lqr_riccati(n, m, N, batch, seed=n + m, lti=lti)
    _check(prob, handle, oracle_mod)
    assert handle.last_kernel.startswith("riccati_dmma")


@pytest.mark.parametrize("n,m,N,batch", [(32, 8, 12, 3), (64, 16, 11, 2), (64, 16, 2, 3), (64, 16, 101, 5), (32, 8, 300, 7),
                                         (16, 8, 40, 5), (24, 8, 30, 4), (48, 16, 25, 3), (16, 16, 30, 5), (24, 16, 21, 4), (32, 16, 26, 3)])
@pytest.mark.parametrize("lti", [False, True])
def test_cta_dmma_sizes(handle, oracle_mod, n, m, N, batch, lti):
    """CTA-per-instance FP64 tensor-core kernel (config 5b shape and its smaller sibling)."""
    prob = problems.random_lqr_riccati(n, m, N, batch, seed=n + m, lti=lti)
    _check(prob, handle, oracle_mod)
    assert handle.last_kernel.startswith("riccati_cta_dmma")


def test_cta_dmma_matches_cooperative_kernel(handle):
    prob = problems.random_lqr_riccati(64, 16, 60, 6, seed=78)
    X1, U1, K1, k1, i1 = ops.riccati_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith("riccati_cta_dmma")
    handle.set_option("riccati_variant", 2)
    try:
        X2, U2, K2, k2, i2 = ops.riccati_solve_problem(prob, handle=handle)
        assert handle.last_kernel.startswith("riccati_coop")
    finally:
        handle.set_option("riccati_variant", 0)
    assert (i1 == 0).all() and (i2 == 0).all()
    assert _rel(X1, X2) <= 1e-11 and _rel(U1, U2) <= 1e-11 and _rel(K1, K2) <= 1e-11 and _rel(k1, k2) <= 1e-11


def test_cta_dmma_info_reports_nonpositive_pivot(handle):
    prob = problems.random_lqr_riccati(64, 16, 20, 4, seed=2)
    prob["R"][1] = -1e6 * np.eye(16)
    _, _, _, _, info = ops.riccati_solve_problem(prob, handle=handle)
    assert info[1] != 0 and (np.delete(info, 1) == 0).all()


def test_dmma_matches_cooperative_kernel(handle):
    """Two independent CUDA implementations of the same recursion agree to rounding."""
    prob = problems.random_lqr_riccati(12, 4, 301, 40, seed=77)
    X1, U1, K1, k1, i1 = ops.riccati_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith("riccati_dmma")
    handle.set_option("riccati_variant", 2)
    try:
        X2, U2, K2, k2, i2 = ops.riccati_solve_problem(prob, handle=handle)
        assert handle.last_kernel.startswith("riccati_coop")
    finally:
        handle.set_option("riccati_variant", 0)
    assert (i1 == 0).all() and (i2 == 0).all()
    assert _rel(X1, X2) <= 1e-11 and _rel(U1, U2) <= 1e-11 and _rel(K1, K2) <= 1e-11 and _rel(k1, k2) <= 1e-11


def test_dmma_info_reports_nonpositive_pivot(handle):
    prob = problems.random_lqr_riccati(12, 4, 50, 9, seed=1)
    prob["R"][3] = -1e6 * np.eye(4)
    _, _, _, _, info = ops.riccati_solve_problem(prob, handle=handle)
    assert info[3] != 0 and (np.delete(info, 3) == 0).all()


def test_no_affine_terms_reference_form(handle, oracle_mod):
    """The reference DPSolver form: LTI, no q/r/qf (src/dynamic_programming.jl:54-72)."""
    prob = problems.random_lqr_riccati(4, 1, 101, 40, seed=5, lti=True)
    prob["q"] = prob["r"] = prob["qf"] = None
    _check(prob, handle, oracle_mod)


@pytest.mark.parametrize("D,N,kern", [(2, 21, "riccati_tpi<4,2"), (3, 31, "riccati_tpi<6,3"), (4, 26, "riccati_dmma<8,4")])
def test_condensed_least_squares_identity(handle, D, N, kern):
    """test/least_squares.jl:38: the controls of the Riccati pass make the gradient of the reference's condensed
    least-squares form vanish (an identity that does not go through the oracle).  DoubleIntegrator(D) numbers."""
    from tests.conftest import condensed_least_squares_gradient
    n, m, dt = 2 * D, D, 2.0 / (N - 1)
    A = np.eye(n); A[:D, D:] = dt * np.eye(D)
    B = np.vstack([0.5 * dt * dt * np.eye(D), dt * np.eye(D)])
    Q = np.diag(np.concatenate([10.0 * np.ones(D), np.ones(D)])); R = 0.1 * np.eye(m); Qf = 10 * Q
    rng = np.random.default_rng(D)
    x0 = np.concatenate([np.ones((3, D)), np.zeros((3, D))], axis=1) + 0.1 * rng.standard_normal((3, n))
    prob = dict(n=n, m=m, N=N, lti=True, A=A[None].repeat(3, 0), B=B[None].repeat(3, 0), Q=Q[None].repeat(3, 0),
                R=R[None].repeat(3, 0), q=None, r=None, Qf=Qf[None].repeat(3, 0), qf=None, x0=x0)
    X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith(kern) and (info == 0).all()
    for i in range(3):
        grad, scale = condensed_least_squares_gradient(A, B, Q, R, Qf, x0[i], U[i])
        assert np.abs(grad).max() < 1e-11 * max(1.0, scale)


def test_device_resident_path_matches_host_path(handle):
    import torch
    prob = problems.riccati_cartpole_batch(100, seed=3)
    f = ops.riccati_flatten(prob)
    n, m, N, b = 4, 1, 101, 100
    L = _lib.riccati_layout(n, m, N)
    ldb = _lib.padded_batch(b)
    dev = {k: torch.from_numpy(f[k]).cuda() for k in ("A", "B", "Q", "R", "q", "r", "Qf", "qf", "x0")}
    knots = torch.zeros(ldb * L.knot_count * L.rows_per_knot, dtype=torch.float64, device="cuda")
    term = torch.zeros(ldb * L.term_rows, dtype=torch.float64, device="cuda")
    Zp = torch.zeros(ldb * L.z_rows, dtype=torch.float64, device="cuda")
    Z = torch.zeros(b, L.z_rows, dtype=torch.float64, device="cuda")
    handle.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ops.riccati_pack(handle, n, m, N, b, 0, dev["A"], dev["B"], dev["Q"], dev["R"], dev["q"], dev["r"],
                         dev["Qf"], dev["qf"], dev["x0"], knots, term)
        ops.riccati_solve_packed(handle, n, m, N, b, 0, knots, term, Zp)
        ops.unpack_rows(handle, L.z_rows, b, ops.riccati_tile_width(handle, n, m), Zp, Z)
        torch.cuda.synchronize()
    finally:
        handle.set_stream(None)
    X, U, _, _, _ = ops.riccati_solve_problem(prob, handle=handle)
    Xd, Ud = ops.split_primals(Z.cpu().numpy(), n, m, N)
    assert np.array_equal(Xd, X) and np.array_equal(Ud, U)


def test_info_reports_nonpositive_pivot(handle):
    prob = problems.riccati_cartpole_batch(40, seed=9)
    prob["R"][7] = -1e6          # E = R + B'PB becomes negative at instance 7
    _, _, _, _, info = ops.riccati_solve_problem(prob, handle=handle)
    assert info[7] != 0 and (np.delete(info, 7) == 0).all()


def test_linearity_full_size(handle):
    """Size-independent property at the BASELINE batch (65,536): with q=r=qf=0 the trajectory is linear
    in x0, so solve(2*x0) == 2*solve(x0) up to rounding."""
    b = 65536
    base = problems.riccati_cartpole_batch(256, seed=11)
    rep = b // 256
    prob = {k: (np.tile(v, (rep,) + (1,) * (v.ndim - 1)) if isinstance(v, np.ndarray) else v)
            for k, v in base.items()}
    prob["q"] = prob["r"] = prob["qf"] = None
    rng = np.random.default_rng(0)
    prob["x0"] = rng.standard_normal((b, 4))
    X1, U1, _, _, info = ops.riccati_solve_problem(prob, want_gains=False, handle=handle)
    prob["x0"] = 2.0 * prob["x0"]
    X2, U2, _, _, _ = ops.riccati_solve_problem(prob, want_gains=False, handle=handle)
    assert (info == 0).all()
    assert np.abs(X2 - 2 * X1).max() <= 1e-9 * max(1.0, np.abs(X1).max())
    assert np.abs(U2 - 2 * U1).max() <= 1e-9 * max(1.0, np.abs(U1).max())


def test_rollout_entry_point(handle, oracle_mod):
    prob = problems.riccati_cartpole_batch(50, seed=2)
    X, U, _, _, _ = ops.riccati_solve_problem(prob, handle=handle)
    f = ops.riccati_flatten(prob)
    Xr = np.zeros((50, 101, 4))
    ops.rollout(handle, 4, 1, 101, 50, 0, f["A"], f["B"], f["x0"], np.ascontiguousarray(U), Xr)
    assert _rel(Xr, X) <= 1e-12 and handle.last_kernel == "rollout_warp"


@pytest.mark.parametrize("n,m,N,b,lti,kern", [(12, 4, 60, 37, False, "rollout_warp"), (7, 3, 20, 5, True, "rollout_warp"),
                                              (32, 8, 9, 3, False, "rollout_warp"), (40, 8, 7, 3, False, "rollout_tpi")])
def test_rollout_kernels(handle, n, m, N, b, lti, kern):
    """rollout! (src/least_squares.jl:195-202) against numpy for both kernels (warp per instance for n <= 32)."""
    prob = problems.random_lqr_riccati(n, m, N, b, seed=n, lti=lti)
    rng = np.random.default_rng(1)
    U = rng.standard_normal((b, N - 1, m))
    f = ops.riccati_flatten(prob)
    X = np.zeros((b, N, n))
    ops.rollout(handle, n, m, N, b, _lib.FLAG_LTI if lti else 0, f["A"], f["B"], f["x0"], U, X)
    assert handle.last_kernel == kern
    Xr = np.zeros_like(X)
    Xr[:, 0] = prob["x0"]
    for k in range(N - 1):
        Ak = prob["A"] if lti else prob["A"][:, k]
        Bk = prob["B"] if lti else prob["B"][:, k]
        Xr[:, k + 1] = np.einsum("bij,bj->bi", Ak, Xr[:, k]) + np.einsum("bij,bj->bi", Bk, U[:, k])
    assert _rel(X, Xr) <= 1e-13


@pytest.mark.parametrize("n,m,N,batch,kern", [(12, 4, 201, 1027, "riccati_dmma"), (64, 16, 31, 301, "riccati_cta_dmma")])
def test_tuned_kernels_closed_loop_at_scale(handle, n, m, N, batch, kern):
    """Size-independent properties on a batch that spans many CTAs (odd tail included): the returned trajectory
    satisfies the dynamics with the returned gains (u = -K x - kff, x+ = A x + B u) for EVERY instance, and
    doubling (x0, q, r, qf) doubles the solution."""
    prob = problems.random_lqr_riccati(n, m, N, batch, seed=21)
    X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith(kern) and (info == 0).all()
    u_cl = -np.einsum("bkij,bkj->bki", K, X[:, :-1]) - kff
    x_next = np.einsum("bkij,bkj->bki", prob["A"], X[:, :-1]) + np.einsum("bkij,bkj->bki", prob["B"], U)
    s = max(1.0, np.abs(X).max())
    assert np.abs(U - u_cl).max() <= 1e-10 * max(1.0, np.abs(U).max())
    assert np.abs(X[:, 1:] - x_next).max() <= 1e-10 * s
    assert np.abs(X[:, 0] - prob["x0"]).max() == 0.0
    p2 = dict(prob)
    for key in ("x0", "q", "r", "qf"):
        p2[key] = 2.0 * prob[key]
    X2, U2, _, _, _ = ops.riccati_solve_problem(p2, want_gains=False, handle=handle)
    assert np.abs(X2 - 2 * X).max() <= 1e-9 * s and np.abs(U2 - 2 * U).max() <= 1e-9 * max(1.0, np.abs(U).max())


@pytest.mark.parametrize("n,m,N,kern", [(4, 1, 401, "riccati_tpi<4,1>"), (6, 3, 401, "riccati_tpi<6,3>"), (12, 4, 401, "riccati_dmma<12,4>"),
                                        (10, 3, 401, "riccati_dmma<12,3>"), (64, 16, 401, "riccati_cta_dmma<64,16>"),
                                        (20, 6, 401, "riccati_cta_dmma<24,8>")])
def test_gain_converges_to_scipy_dare(handle, n, m, N, kern):
    """Independent of the oracle: the first gain of a long LTI horizon against scipy.linalg.solve_discrete_are
    (tests/test_oracle.py::test_riccati_gain_converges_to_scipy_dare holds the oracle to the same number)."""
    import scipy.linalg as sl
    prob = problems.dare_lti_riccati(n, m, N, 5, seed=n)
    X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=handle)
    assert (info == 0).all() and handle.last_kernel.startswith(kern), handle.last_kernel
    for i in range(5):
        A, B, Q, R = prob["A"][i], prob["B"][i], prob["Q"][i], prob["R"][i]
        P = sl.solve_discrete_are(A, B, Q, R)
        Kd = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
        rho = np.abs(np.linalg.eigvals(A - B @ Kd)).max()
        tol = max(1e-9, 100.0 * rho ** (2 * (N - 1)))
        assert np.linalg.norm(K[i, 0] - Kd) / np.linalg.norm(Kd) <= tol, (i, rho, tol)


@pytest.mark.parametrize("n,m,N,kern", [(6, 3, 201, "riccati_tpi<6,3>"), (12, 4, 201, "riccati_dmma<12,4>"), (10, 3, 201, "riccati_dmma<12,3>"),
                                        (64, 16, 201, "riccati_cta_dmma<64,16>"), (20, 6, 201, "riccati_cta_dmma<24,8>"),
                                        (9, 5, 201, "riccati_cta_dmma<16,8>")])
def test_open_loop_unstable_systems(handle, oracle_mod, n, m, N, kern):
    """Where the kernels are deliberately NOT the reference's operation order: `P_ = Q + A'PA - A'PB K`
    (src/dynamic_programming.jl:46-51) does not symmetrise P, and with an open-loop unstable A (spectral radius 1.2-1.3)
    the antisymmetric rounding residue grows by |A|^2 per knot until potrf fails or the gains are wrong — the CPU oracle,
    which follows the reference, shows exactly that on these inputs.  Every kernel family symmetrises the cost-to-go
    exactly each knot and matches scipy's DARE gain."""
    import scipy.linalg as sl
    prob = problems.dare_lti_riccati(n, m, N, 4, seed=1, unstable=True)
    X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=handle)
    assert (info == 0).all() and handle.last_kernel.startswith(kern), handle.last_kernel
    worst_oracle = 0.0
    Xo, Uo, Ko, kffo, infoo = oracle_mod.riccati(prob)
    for i in range(4):
        A, B, Q, R = prob["A"][i], prob["B"][i], prob["Q"][i], prob["R"][i]
        P = sl.solve_discrete_are(A, B, Q, R)
        Kd = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
        rho = np.abs(np.linalg.eigvals(A - B @ Kd)).max()
        tol = max(1e-9, 100.0 * rho ** (2 * (N - 1)))
        assert np.linalg.norm(K[i, 0] - Kd) / np.linalg.norm(Kd) <= tol, (i, rho, tol)
        worst_oracle = max(worst_oracle, np.linalg.norm(Ko[i, 0] - Kd) / np.linalg.norm(Kd) if infoo[i] == 0 else 1.0)
    assert worst_oracle > 1e-6  # the reference's own order has lost the answer on at least one of these instances
