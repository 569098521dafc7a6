"""GPU: the conditioning hole of the explicit-inverse KKT kernels (kkt_hw2, kkt_cta) is closed on the DEFAULT
dispatch.  tools/stress_scales.py as tests: cost blocks rescaled by 10^-3 .. 10^5, every size class with a tuned kernel.

Tolerance, stated per case: the relative error of (dz, multipliers) against the extended-precision global KKT solve
(oracle/dense_kkt.py) must be <= max(1e-10, 4 x the error the reference's own operation order makes on the same
input) — the second term only matters where the PROBLEM is ill-conditioned (Q 10^5 lighter than R: the CPU oracle
itself is 1e-7 off) and no block-Cholesky solver, the reference included, can reach 1e-10."""
import numpy as np
import pytest

from lqr_b200 import ops, problems

pytestmark = pytest.mark.gpu

SCALES = [(1.0, 1.0), (1e3, 1e-3), (1e-3, 1e3), (1e5, 1.0), (1.0, 1e-5), (1e2, 1e-2), (1e4, 1e-1), (79.0, 1e-3)]
SIZES = [(12, 4, 40, 4, "kkt_hw<12,4"), (8, 4, 30, 4, "kkt_hw<8,4"), (64, 16, 12, 2, "kkt_cta_dmma<64,16"),
         (24, 8, 20, 2, "kkt_cta_dmma<24,8")]


def _err(prob, dz, lam):
    from oracle import dense_kkt
    e = 0.0
    for i in range(dz.shape[0]):
        zt, lt = dense_kkt.kkt_truth(prob, i)
        e = max(e, np.linalg.norm(dz[i] - zt) / np.linalg.norm(zt), np.linalg.norm(lam[i] - lt) / np.linalg.norm(lt))
    return e


@pytest.mark.parametrize("qs,rs", SCALES)
@pytest.mark.parametrize("n,m,N,b,kern", SIZES)
def test_default_dispatch_under_rescaled_costs(handle, oracle_mod, n, m, N, b, kern, qs, rs):
    prob = problems.random_lqr_kkt(n, m, N, b, seed=200 + int(np.log10(qs) * 7 + np.log10(rs)), mid_p=0, hess_mode=1)
    prob["Q"] = prob["Q"] * qs
    prob["R"] = prob["R"] * rs
    dz, lam, info = ops.kkt_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith(kern) and (info == 0).all()
    bits, resolved = handle.kkt_last_condition(b)
    dzo, lamo, _ = oracle_mod.kkt_solve(prob)
    e, eo = _err(prob, dz, lam), _err(prob, dzo, lamo)
    assert e <= max(1e-10, 4.0 * eo), (e, eo, bits.tolist(), resolved, handle.last_kernel)
    # flagged instances went through the Cholesky-based kernel, the others did not
    assert resolved == int((bits >= 6).sum())
    assert ("+kkt_coop[" in handle.last_kernel) == (resolved > 0)


def test_condition_estimate_and_switch(handle):
    """kkt_refine = 0 leaves the tuned kernel alone; the estimate grows with the spread of the cost scaling."""
    prob = problems.random_lqr_kkt(8, 4, 30, 4, seed=3, mid_p=0, hess_mode=1)
    ops.kkt_solve_problem(prob, handle=handle)
    b0, r0 = handle.kkt_last_condition(4)
    prob["Q"] = prob["Q"] * 1e3
    prob["R"] = prob["R"] * 1e-3
    dz1, lam1, _ = ops.kkt_solve_problem(prob, handle=handle)
    b1, r1 = handle.kkt_last_condition(4)
    assert r0 == 0 and r1 == 4 and (b1 > b0).all() and (b0 >= 0).all()
    handle.set_option("kkt_refine", 0)
    try:
        dz2, lam2, _ = ops.kkt_solve_problem(prob, handle=handle)
        assert "+kkt_coop" not in handle.last_kernel
    finally:
        handle.set_option("kkt_refine", 1)
    assert not np.array_equal(dz1, dz2)


@pytest.mark.parametrize("n,m,N,kind", [(12, 4, 30, "kkt"), (10, 3, 30, "kkt"), (14, 7, 12, "kkt"), (10, 3, 40, "riccati"),
                                        (20, 6, 25, "riccati")])
def test_host_path_with_uneven_chunks(handle, oracle_mod, n, m, N, kind):
    """Host buffers go through two streams in chunks; the last chunk is smaller than the others.  Everything a chunk
    allocates from inside (re-solve lists and records, padded arrays) has one allocation per stream, so the chunks in
    flight cannot overlap whatever their sizes: 70 instances in chunks of 32 (32 + 32 + 6), half of them rescaled so
    that the re-solve path runs in every chunk."""
    b = 70
    handle.set_option("host_chunk", 32)
    try:
        if kind == "kkt":
            prob = problems.random_lqr_kkt(n, m, N, b, seed=31 + n, mid_p=0, hess_mode=1)
            prob["Q"][::2] *= 1e3
            prob["R"][::2] *= 1e-3
            dz, lam, info = ops.kkt_solve_problem(prob, handle=handle)
            dzo, lamo, infoo = oracle_mod.kkt_solve(prob)
            assert (info == 0).all() and (infoo == 0).all()
            e, eo = _err(prob, dz, lam), _err(prob, dzo, lamo)
            assert e <= max(1e-10, 4.0 * eo), (e, eo, handle.last_kernel)
            # the same call in one chunk gives the same bits
            handle.set_option("host_chunk", 0)
            dz1, lam1, _ = ops.kkt_solve_problem(prob, handle=handle)
            assert np.array_equal(dz, dz1) and np.array_equal(lam, lam1)
        else:
            prob = problems.random_lqr_riccati(n, m, N, b, seed=31 + n)
            X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=handle)
            assert "padded" in handle.last_kernel
            Xo, Uo, Ko, kffo, _ = oracle_mod.riccati(prob)
            for a, c in ((X, Xo), (U, Uo), (K, Ko), (kff, kffo)):
                assert np.linalg.norm(a - c) / np.linalg.norm(c) <= 1e-10
            handle.set_option("host_chunk", 0)
            X1, U1, K1, kff1, _ = ops.riccati_solve_problem(prob, handle=handle)
            assert np.array_equal(X, X1) and np.array_equal(U, U1) and np.array_equal(K, K1)
    finally:
        handle.set_option("host_chunk", 0)


def test_general_kernel_grid_stride_with_narrow_groups(handle, oracle_mod):
    """The general KKT kernel at 4 lanes per instance caps its grid and strides over the instances; the groups of one warp
    must make the same number of trips (idle ones shadow the last instance).  310,001 tiny problems with explicit D2."""
    b = 310001
    prob = problems.random_lqr_kkt(3, 2, 4, b, seed=77, mid_p=0, hess_mode=1, explicit_D2=True)
    dz, lam, info = ops.kkt_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith("kkt_coop<G=4>")
    dzo, lamo, infoo = oracle_mod.kkt_solve(prob)
    assert (info == 0).all() and (infoo == 0).all()
    ez = np.linalg.norm(dz - dzo, axis=1) / np.linalg.norm(dzo, axis=1)
    el = np.linalg.norm(lam - lamo, axis=1) / np.linalg.norm(lamo, axis=1)
    assert ez.max() <= 1e-9 and el.max() <= 1e-9 and np.median(ez) <= 1e-13, (ez.max(), el.max())


@pytest.mark.parametrize("n,m,N,mid_p,kern", [(10, 3, 12, 1, "kkt_wp_dmma<12,4"), (14, 7, 12, 0, "kkt_cta_dmma<16,8")])
def test_padded_path_in_several_chunks(handle, oracle_mod, n, m, N, mid_p, kern):
    """A small scratch budget makes the padded path (expand -> tuned kernel -> compact) run in several chunks."""
    b = 8001
    prob = problems.random_lqr_kkt(n, m, N, b, seed=5 + n, mid_p=mid_p, hess_mode=1)
    handle.set_option("scratch_budget_mb", 8)
    try:
        dz, lam, info, res = ops.kkt_solve_problem(prob, want_res=True, handle=handle)
    finally:
        handle.set_option("scratch_budget_mb", 49152)
    assert handle.last_kernel.startswith(kern) and "padded" in handle.last_kernel
    dzo, lamo, infoo, reso = oracle_mod.kkt_solve(prob, want_res=True)
    assert (info == 0).all() and (infoo == 0).all()
    ez = np.linalg.norm(dz - dzo, axis=1) / np.linalg.norm(dzo, axis=1)
    el = np.linalg.norm(lam - lamo, axis=1) / np.linalg.norm(lamo, axis=1)
    assert np.median(ez) <= 1e-11 and np.median(el) <= 1e-11 and ez.max() <= 1e-8 and el.max() <= 1e-8, (ez.max(), el.max())
    assert np.abs(res - reso).max() <= 1e-8 * max(1.0, np.abs(reso).max())
