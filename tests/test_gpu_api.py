"""GPU tests of the reference-interface mirror (lqr_b200.*): same call sequences as the reference's
scripts, checked with the identities those scripts state."""
import numpy as np
import pytest
import torch

import lqr_b200 as LQR
from lqr_b200 import problems

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_dpsolver_like_test_dp(handle, oracle_mod):
    """test/dp.jl: prob -> DPSolver -> solve!(sol, solver, prob); X[1] == x0 and the closed loop holds."""
    pr = problems.random_lqr_riccati(4, 1, 101, 16, seed=1, lti=True)
    prob = LQR.LQRProblem(pr["Qf"], pr["Q"], pr["R"], pr["A"], pr["B"], pr["x0"], N=101)
    solver = LQR.DPSolver(prob, handle=handle)
    sol = LQR.LQRSolution(prob)
    LQR.solve_(sol, solver, prob)
    assert (sol.info == 0).all() and np.array_equal(sol.X[:, 0], pr["x0"])
    u = -np.einsum("bkij,bkj->bki", sol.K, sol.X[:, :-1])
    assert _rel(u, sol.U) < 1e-12
    xn = np.einsum("bij,bkj->bki", pr["A"], sol.X[:, :-1]) + np.einsum("bij,bkj->bki", pr["B"], sol.U)
    assert _rel(xn, sol.X[:, 1:]) < 1e-12
    pr["q"] = pr["r"] = pr["qf"] = None
    Xo, Uo, Ko, _, _ = oracle_mod.riccati(pr)
    assert _rel(sol.X, Xo) < 1e-10 and _rel(sol.K, Ko) < 1e-10
    X2 = np.zeros_like(sol.X)
    LQR.rollout_(X2, sol.U, prob, handle=handle)
    assert _rel(X2, sol.X) < 1e-12


def test_block_cholesky_like_reference_test(handle):
    """test/block_cholesky.jl:11-67 (n=10, m=5, C = 1e-3*rand), batched over 7 instances."""
    rng = np.random.default_rng(0)
    n, m, b = 10, 5, 7
    A = rng.random((b, n, n)); A = np.einsum("bki,bkj->bij", A, A)
    B = rng.random((b, m, m)); B = np.einsum("bki,bkj->bij", B, B)
    C = rng.random((b, m, n)) * 1e-3
    rhs = rng.random((b, n + m))
    M = np.block([[A, np.swapaxes(C, 1, 2)], [C, B]])
    chol = LQR.BlockCholesky(n, m, batch=b, handle=handle)
    assert not chol.block_diag                                                     # :20
    LQR.cholesky_(chol, A, B, C)
    assert (chol.info == 0).all()
    for i in range(b):
        assert np.allclose(chol.U[i], np.linalg.cholesky(M[i]).T, rtol=1e-9, atol=1e-11)  # :24
    x = LQR.ldiv(chol, rhs)
    assert _rel(x, np.linalg.solve(M, rhs[..., None])[..., 0]) < 1e-8               # :26
    b1 = rhs.copy(); LQR.ldiv_(chol, b1)
    assert np.array_equal(b1, x)                                                   # :27-29
    # block diagonal (:36-49)
    chol = LQR.BlockCholesky(n, m, batch=b, block_diag=True, handle=handle)
    assert chol.block_diag
    LQR.cholesky_(chol, A, B)
    Md = M.copy(); Md[:, n:, :n] = 0; Md[:, :n, n:] = 0
    for i in range(b):
        assert np.allclose(chol.U[i], np.linalg.cholesky(Md[i]).T, rtol=1e-9, atol=1e-11)
    assert _rel(LQR.ldiv(chol, rhs), np.linalg.solve(Md, rhs[..., None])[..., 0]) < 1e-8
    # diagonal (:56-67): stores the inverse
    Ad, Bd = rng.random((b, n)) + 0.1, rng.random((b, m)) + 0.1
    chol = LQR.BlockCholesky(n, m, batch=b, diag=True, handle=handle)
    LQR.cholesky_(chol, Ad[..., None] * np.eye(n), Bd[..., None] * np.eye(m))
    dd = np.concatenate([Ad, Bd], axis=1)
    assert np.allclose(np.diagonal(chol.M, axis1=1, axis2=2), 1.0 / dd)
    assert np.allclose(LQR.ldiv(chol, rhs), rhs / dd)
    # potrf info on an indefinite block (src/cholesky_solve.jl:1-3)
    Abad = A.copy(); Abad[3] = -np.eye(n)
    chol = LQR.BlockCholesky(n, m, batch=b, block_diag=True, handle=handle)
    LQR.cholesky_(chol, Abad, B)
    assert chol.info[3] == 1 and (np.delete(chol.info, 3) == 0).all()


def test_inverted_quadratic_update(handle):
    """InvertedQuadratic / update_cost! / gradient (src/block_cholesky.jl:107-153)."""
    rng = np.random.default_rng(1)
    n, m, b = 3, 2, 4
    Q = np.einsum("bki,bkj->bij", *(2 * [rng.random((b, n, n))])) + np.eye(n)
    R = np.einsum("bki,bkj->bij", *(2 * [rng.random((b, m, m))])) + np.eye(m)
    q, r = rng.random((b, n)), rng.random((b, m))
    ic = LQR.InvertedQuadratic(n, m, batch=b, handle=handle)
    LQR.update_cholesky_([ic], [dict(Q=Q, R=R, q=q, r=r)])
    g = LQR.gradient(ic)
    assert np.array_equal(g, np.concatenate([q, r], axis=1))
    H = np.zeros((b, n + m, n + m)); H[:, :n, :n] = Q; H[:, n:, n:] = R
    assert _rel(LQR.ldiv(ic.chol, g), np.linalg.solve(H, g[..., None])[..., 0]) < 1e-10


def test_cholesky_solver_step_sequence_like_reference_script(handle):
    """test/cholesky_solve.jl:7-44, the step-by-step script, on DoubleIntegrator(3,101)."""
    prob = problems.double_integrator_fixture()
    solver = LQR.CholeskySolver(prob, handle=handle)
    LQR.calculate_shur_factors_(solver.shur_blocks, None, None)        # :14
    D, d = LQR.get_linearized_constraints(solver)                       # :17
    H, g = LQR.get_cost_expansion(solver)
    Sd, rd, _ = LQR.get_shur_factors(solver)                            # :18  S, h as the DEVICE forms them
    S = D @ np.linalg.solve(H, D.T)
    r = D @ np.linalg.solve(H, g) - d
    assert _rel(Sd[0], S) < 1e-12                                       # :19  S ≈ D*(H\D')
    assert _rel(rd[0], r) < 1e-12                                       # :20  D*(H\g) - d ≈ r
    LQR.cholesky_(solver.chol_blocks, solver.shur_blocks)               # :22  lqrb_kkt_factor_f64
    assert handle.last_kernel.startswith("kkt_coop") and (solver.info == 0).all()
    U = LQR.get_cholesky(solver)[0]                                     # :23  block rows of U from the device
    assert np.array_equal(U, np.triu(U))
    assert _rel(U, np.linalg.cholesky(Sd[0]).T) < 1e-9                  # :24  cholesky(S).U ≈ U
    assert _rel(U.T @ U, Sd[0]) < 1e-12                                 # :25  U'U ≈ S
    LQR.forward_substitution_(solver.chol_blocks)                       # :28  lqrb_kkt_solve_factored_f64
    LQR.backward_substitution_(solver.chol_blocks)                      # :29
    lam = LQR.get_multipliers(solver)[0]                                # :30
    dZ = np.zeros_like(solver.dZ)
    LQR.calculate_primals_(dZ, None, solver.chol_blocks, None)          # :34
    dZ = LQR.get_step(solver)[0]                                        # :35
    S2, h2, lam2 = np.zeros_like(Sd), np.zeros_like(rd), np.zeros_like(solver.lam)
    LQR.copy_shur_factors_(S2, h2, lam2, solver.shur_blocks)            # src/jacobian_blocks.jl:173-180
    assert np.array_equal(S2, Sd) and np.array_equal(lam2, solver.lam)
    assert _rel(lam, -np.linalg.solve(S, r)) < 1e-7                     # :31
    assert _rel(dZ, -np.linalg.solve(H, D.T @ lam + g)) < 1e-10         # :36
    assert np.linalg.norm(D @ dZ + d) < 1e-10                           # :39
    assert np.linalg.norm(H @ dZ + g + D.T @ lam) < 1e-10               # :40
    NN, P = H.shape[0], D.shape[0]
    sol = np.linalg.solve(np.block([[H, D.T], [D, np.zeros((P, P))]]), -np.concatenate([g, d]))  # :42
    assert _rel(dZ, sol[:NN]) < 1e-8 and _rel(lam, sol[NN:]) < 1e-7     # :43-44
    # residual(solver) = ||g + D'lam|| knot-wise (src/cholesky_solver.jl:238-252; test/cholesky_comp.jl:45)
    res = LQR.residual(solver)[0]
    assert abs(res - np.linalg.norm(g + D.T @ lam)) < 1e-9 * max(1.0, res)
    assert _rel(LQR.get_residual(solver)[0], g + D.T @ lam) < 1e-10
    # _solve! gives the same thing in one fused launch on the tuned kernel (:166-182)
    s2 = LQR.CholeskySolver(prob, handle=handle)._solve_()
    assert handle.last_kernel.startswith("kkt_tpi<6,3")
    assert _rel(s2.dZ, solver.dZ) < 1e-10 and _rel(s2.lam, solver.lam) < 1e-9
    assert _rel(s2.res, solver.res) < 1e-9
    # second_order_correction!: dz^ = -D'(DD')^-1 d (:254-273); cond(DD') ~ 1e7 here
    dzh = LQR.second_order_correction_(solver)[0]
    assert _rel(dzh, -D.T @ np.linalg.solve(D @ D.T, d)) < 1e-6 or np.linalg.norm(d) == 0


@pytest.mark.parametrize("m,mid_p,explicit_D2", [(2, 0, False), (3, 2, False), (2, 1, True)])
def test_residual_recalculate_matches_dense(handle, m, mid_p, explicit_D2):
    """residual(solver; recalculate=true) (src/cholesky_solver.jl:238-252): calc_residual! with the KEPT
    multipliers on re-evaluated blocks.  Checked against the dense g + D'lam of the get_* extractors
    (test/cholesky_comp.jl:45) — relative 1e-12: it is one dot product per entry."""
    prob = problems.random_lqr_kkt(4, m, 9, 5, seed=11, mid_p=mid_p, explicit_D2=explicit_D2)
    solver = LQR.CholeskySolver(prob, handle=handle)._solve_()
    res_solve = solver.res.copy()
    # same blocks: the recalculated residual is the one the solve reported
    nrm = LQR.residual(solver, recalculate=True)
    assert _rel(solver.res, res_solve) < 1e-12
    assert _rel(nrm, np.linalg.norm(res_solve, axis=1)) < 1e-12
    assert _rel(nrm, LQR.residual(solver)) < 1e-12
    # new blocks, kept multipliers (what step! evaluates after the line search, :126-134)
    rng = np.random.default_rng(5)
    lam = solver.lam.copy()
    solver.update_(A=prob["A"] + 0.01 * rng.standard_normal(prob["A"].shape),
                   q=prob["q"] + 0.1 * rng.standard_normal(prob["q"].shape))
    solver.lam[...] = lam
    nrm = LQR.residual(solver, recalculate=True)
    for i in (0, 4):
        D, _ = LQR.get_linearized_constraints(solver, i)
        _, g = LQR.get_cost_expansion(solver, i)
        want = g + D.T @ lam[i]
        assert _rel(solver.res[i], want) < 1e-12
        assert abs(nrm[i] - np.linalg.norm(want)) < 1e-12 * np.linalg.norm(want)


def test_residual_entry_device_pointers_and_soc(handle):
    """lqrb_kkt_residual_f64 on device-resident arrays; LQRB_FLAG_SOC drops the gradient (Ginv = false,
    src/cholesky_solver.jl:229-231) and accepts NULL q, r; argument errors are LAPACK-style."""
    from lqr_b200 import _lib, ops
    prob = problems.random_lqr_kkt(3, 2, 12, 40, seed=4, mid_p=1)
    f = ops.kkt_flatten(prob)
    n, m, N, b = 3, 2, 12, 40
    NN, P = _lib.num_vars(n, m, N), _lib.num_cons(n, N, f["p"])
    rng = np.random.default_rng(0)
    lam = rng.standard_normal((b, P))
    host_res, host_nrm = np.zeros((b, NN)), np.zeros(b)
    ops.kkt_residual(handle, n, m, N, b, f["p"], 0, f["q"], f["r"], f["A"], f["B"], f["D2"], f["C"], lam, host_res, host_nrm)
    dev = {k: torch.from_numpy(np.ascontiguousarray(f[k])).cuda() for k in ("q", "r", "A", "B", "C")}
    dlam = torch.from_numpy(lam).cuda()
    dres, dnrm = torch.zeros((b, NN), dtype=torch.float64, device="cuda"), torch.zeros(b, dtype=torch.float64, device="cuda")
    ops.kkt_residual(handle, n, m, N, b, f["p"], 0, dev["q"], dev["r"], dev["A"], dev["B"], None, dev["C"], dlam, dres, dnrm)
    handle.synchronize()
    assert np.array_equal(dres.cpu().numpy(), host_res) and np.array_equal(dnrm.cpu().numpy(), host_nrm)
    # SOC: no gradient
    soc_res = np.zeros((b, NN))
    ops.kkt_residual(handle, n, m, N, b, f["p"], _lib.FLAG_SOC, None, None, f["A"], f["B"], f["D2"], f["C"], lam, soc_res, None)
    g = np.concatenate([np.concatenate([prob["q"][:, :N - 1], prob["r"]], axis=2).reshape(b, -1), prob["q"][:, N - 1]], axis=1)
    assert _rel(host_res - soc_res, g) < 1e-12
    with pytest.raises(LQR.LqrbError):
        ops.kkt_residual(handle, n, m, N, b, f["p"], 0, None, None, f["A"], f["B"], None, f["C"], lam, soc_res, None)
    with pytest.raises(LQR.LqrbError):
        ops.kkt_residual(handle, n, m, N, b, f["p"], 0, f["q"], f["r"], f["A"], f["B"], None, f["C"], lam, None, None)


def test_edge_cases_empty_batch_shortest_horizon_and_argument_errors(handle, oracle_mod):
    """Empty batch is a no-op; N = 2 is the shortest horizon (one control knot + the terminal knot); argument
    errors come back as negative LAPACK-style codes (never a crash, never a silent fallback)."""
    from lqr_b200 import _lib, ops
    # ---- batch = 0: every entry returns 0 and touches nothing
    prob = problems.random_lqr_kkt(4, 2, 6, 1, seed=1)
    f = ops.kkt_flatten(prob)
    e = np.zeros((0, 1))
    ops.kkt_solve(handle, 4, 2, 6, 0, f["p"], f["hess_mode"], 0, e, e, None, e, e, e, e, e, None, e, e, e, e, None, None)
    ops.kkt_residual(handle, 4, 2, 6, 0, f["p"], 0, e, e, e, e, None, e, e, e, None)
    rp = problems.random_lqr_riccati(4, 2, 6, 1, seed=1)
    g = ops.riccati_flatten(rp)
    ops.riccati(handle, 4, 2, 6, 0, 0, e, e, e, e, e, e, e, e, e, e, None, None, None)
    # ---- shortest horizons: N = 2 needs m >= n to be feasible with init + goal rows (cooperative kernel);
    # the thread-per-instance and tuned families at their shortest feasible N
    for n, m, N in [(2, 2, 2), (3, 3, 2), (3, 2, 3), (4, 1, 5), (12, 4, 4)]:
        p2 = problems.random_lqr_kkt(n, m, N, 3, seed=n)
        dz, lam, info = ops.kkt_solve_problem(p2, handle=handle)
        dzo, lamo, infoo = oracle_mod.kkt_solve(p2)
        assert (info == 0).all() and (infoo == 0).all(), (n, m, N, handle.last_kernel, info)
        assert _rel(dz, dzo) < 1e-9 and _rel(lam, lamo) < 1e-9, (n, m, N, handle.last_kernel)
    for n, m in [(2, 2), (4, 1), (12, 4)]:
        r2 = problems.random_lqr_riccati(n, m, 2, 3, seed=n)
        X, U, K, kff, info = ops.riccati_solve_problem(r2, handle=handle)
        Xo, Uo, Ko, kffo, _ = oracle_mod.riccati(r2)
        assert (info == 0).all() and _rel(X, Xo) < 1e-10 and _rel(U, Uo) < 1e-10 and _rel(K, Ko) < 1e-10
    # an over-determined problem (N = 2, m < n: 12 rows on 9 unknowns) is REPORTED, by the kernel and by the oracle
    bad = problems.random_lqr_kkt(4, 1, 2, 3, seed=4)
    _, _, info = ops.kkt_solve_problem(bad, handle=handle)
    _, _, infoo = oracle_mod.kkt_solve(bad)
    assert (info != 0).all() and (infoo != 0).all()
    # ---- argument errors
    dz, lam = np.zeros((1, _lib.num_vars(4, 2, 6))), np.zeros((1, _lib.num_cons(4, 6, f["p"])))
    def code(call):
        with pytest.raises(LQR.LqrbError) as ei:
            call()
        return str(ei.value)
    assert "code -4" in code(lambda: ops.kkt_solve(handle, 4, 2, 1, 1, f["p"], 1, 0, f["Q"], f["R"], None, f["q"], f["r"],
                                                   f["A"], f["B"], f["d"], None, f["C"], f["c"], dz, lam))
    assert "code -3" in code(lambda: ops.kkt_solve(handle, 4, 5, 6, 1, f["p"], 1, 0, f["Q"], f["R"], None, f["q"], f["r"],
                                                   f["A"], f["B"], f["d"], None, f["C"], f["c"], dz, lam))
    assert "code -7" in code(lambda: ops.kkt_solve(handle, 4, 2, 6, 1, f["p"], 7, 0, f["Q"], f["R"], None, f["q"], f["r"],
                                                   f["A"], f["B"], f["d"], None, f["C"], f["c"], dz, lam))
    bad_p = f["p"].copy(); bad_p[2] = 99
    assert "code -6" in code(lambda: ops.kkt_solve(handle, 4, 2, 6, 1, bad_p, 1, 0, f["Q"], f["R"], None, f["q"], f["r"],
                                                   f["A"], f["B"], f["d"], None, f["C"], f["c"], dz, lam))
    assert "code -20" in code(lambda: ops.kkt_solve(handle, 4, 2, 6, 1, f["p"], 1, 0, f["Q"], f["R"], None, f["q"], f["r"],
                                                    f["A"], f["B"], f["d"], None, f["C"], f["c"], None, lam))
    # the handle is still usable after the errors
    dz2, lam2, info2 = ops.kkt_solve_problem(prob, handle=handle)
    assert (info2 == 0).all()


@pytest.mark.parametrize("n,m,N,b,mid_p,hess,d2x", [(4, 1, 12, 5, 0, 2, False), (6, 3, 9, 3, 1, 1, False),
                                                    (5, 2, 10, 4, 1, 0, True), (12, 4, 10, 3, 0, 1, False),
                                                    (3, 3, 2, 2, 0, 1, False)])
def test_factor_once_solve_many(handle, oracle_mod, n, m, N, b, mid_p, hess, d2x):
    """lqrb_kkt_factor_f64 once, lqrb_kkt_solve_factored_f64 for several right-hand sides; each
    must equal the fused solve of the same data (oracle and the tuned / cooperative kernels)."""
    prob = problems.random_lqr_kkt(n, m, N, b, seed=17 + n, mid_p=mid_p, hess_mode=hess, explicit_D2=d2x)
    solver = LQR.CholeskySolver(prob, handle=handle).factor_()
    assert (solver.info == 0).all()
    rng = np.random.default_rng(5)
    for trial in range(3):
        if trial:
            prob = dict(prob, q=rng.standard_normal(prob["q"].shape), r=rng.standard_normal(prob["r"].shape),
                        d=rng.standard_normal(prob["d"].shape), c=[rng.standard_normal(ck.shape) for ck in prob["c"]])
        solver.solve_factored_(q=prob["q"], r=prob["r"], d=prob["d"], c=prob["c"])
        dzo, lamo, _, reso = oracle_mod.kkt_solve(prob, want_res=True)
        assert _rel(solver.dZ, dzo) < 1e-10 and _rel(solver.lam, lamo) < 1e-10
        assert np.abs(solver.res - reso).max() < 1e-10 * max(1.0, np.abs(reso).max())
    # a solve against a factor of another shape is refused
    other = LQR.CholeskySolver(problems.random_lqr_kkt(n, m, N + 1, b, seed=1, mid_p=mid_p, hess_mode=hess), handle=handle)
    other._state = "factored"
    with pytest.raises(LQR.LqrbError):
        other.solve_factored_()


def test_soc_with_kept_factor(handle, oracle_mod):
    """second_order_correction!'s chain (Ginv = false) through the factor / solve split: S = D D' is factored once
    and re-used for two sets of constraint values (src/cholesky_solver.jl:259-263)."""
    prob = problems.cartpole_fixture()
    solver = LQR.CholeskySolver(prob, handle=handle)
    solver._Ginv = False
    solver.factor_()
    rng = np.random.default_rng(2)
    for _ in range(2):
        d = 1e-3 * rng.standard_normal(prob["d"].shape)
        c = [1e-3 * rng.standard_normal(ck.shape) for ck in prob["c"]]
        solver.solve_factored_(d=d, c=c)
        dzo, lamo, _ = oracle_mod.kkt_solve(dict(prob, d=d, c=c), soc=True)
        assert _rel(solver.dZ, dzo) < 1e-9 and _rel(solver.lam, lamo) < 1e-9


@pytest.mark.parametrize("name", ["cartpole", "dubins_mid"])
def test_dense_extractors_on_other_shapes(handle, name):
    """test/constraint_blocks.jl:70-72,122-125 identities (S ≈ D(H\\D'), U'U ≈ S) on the device's block rows."""
    prob = problems.cartpole_fixture(31) if name == "cartpole" else problems.dubins_kkt_batch(3, seed=4, N=17, mid_p=1)
    solver = LQR.CholeskySolver(prob, handle=handle)._solve_()
    S, h, lam = LQR.get_shur_factors(solver)
    U = LQR.get_cholesky(solver)
    for i in range(S.shape[0]):
        D, d = LQR.get_linearized_constraints(solver, i)
        H, g = LQR.get_cost_expansion(solver, i)
        assert _rel(S[i], D @ np.linalg.solve(H, D.T)) < 1e-11
        assert _rel(h[i], D @ np.linalg.solve(H, g) - d) < 1e-11
        assert _rel(U[i].T @ U[i], S[i]) < 1e-12
        assert _rel(lam[i], -np.linalg.solve(S[i], h[i])) < 1e-6


@pytest.mark.parametrize("n,m,N,b", [(4, 1, 101, 3), (6, 3, 31, 4), (2, 1, 12, 5)])
def test_least_squares_solver_matches_riccati(handle, n, m, N, b):
    """solve!(sol::Primals, ::LeastSquaresSolver, prob) (src/least_squares.jl:158-190) on the device: the condensed
    Cholesky solve lands on the Riccati solution (two independent device algorithms, same optimum), and its controls
    zero the reference's own gradient identity A'(AU + b) + R U = 0 (test/least_squares.jl:38)."""
    from tests.conftest import condensed_least_squares_gradient
    pr = problems.random_lqr_riccati(n, m, N, b, seed=5 + n, lti=True)
    prob = LQR.LQRProblem(pr["Qf"], pr["Q"], pr["R"], pr["A"], pr["B"], pr["x0"], N=N)
    sol = LQR.Primals(n, m, N, batch=b)
    lsq = LQR.LeastSquaresSolver(prob, handle=handle)
    LQR.solve_(sol, lsq, prob)
    assert handle.last_kernel.startswith("lsq_solve") and (lsq.info == 0).all()
    ref = LQR.LQRSolution(prob)
    LQR.solve_(ref, LQR.DPSolver(prob, handle=handle), prob)
    assert _rel(sol.X_, ref.X) < 1e-9 and _rel(sol.U_, ref.U) < 1e-9
    for i in range(b):
        grad, scale = condensed_least_squares_gradient(pr["A"][i], pr["B"][i], pr["Q"][i], pr["R"][i], pr["Qf"][i],
                                                       pr["x0"][i], sol.U_[i])
        assert np.abs(grad).max() < 1e-10 * max(1.0, scale)
    # an affine or time-varying problem is refused (the reference's solver is LTI, src/least_squares.jl:61-104)
    ltv = problems.random_lqr_riccati(n, m, N, b, seed=1)
    with pytest.raises(LQR.LqrbError):
        LQR.LeastSquaresSolver(LQR.LQRProblem(ltv["Qf"], ltv["Q"], ltv["R"], ltv["A"], ltv["B"], ltv["x0"], N=N))
