"""GPU: the device-resident split (pack -> solve_packed -> unpack) at every size class, i.e. both tile widths
(T = 32 thread-per-instance, T = 1 records), against the host-pointer call; and the host-pointer path of the
cooperative KKT kernel over several chunks on the handle's two copy streams (its global workspace is per stream)."""
import numpy as np
import pytest

from lqr_b200 import _lib, ops, problems

pytestmark = pytest.mark.gpu


def _dev(f, keys):
    import torch
    return {k: (None if f[k] is None else torch.from_numpy(np.ascontiguousarray(f[k])).cuda()) for k in keys}


@pytest.mark.parametrize("n,m,N,b,tile", [(4, 1, 30, 70, 32), (12, 4, 25, 37, 1), (64, 16, 9, 5, 1), (7, 2, 12, 33, 1), (5, 2, 12, 33, 32)])
def test_riccati_device_resident_all_size_classes(handle, n, m, N, b, tile):
    import torch
    prob = problems.random_lqr_riccati(n, m, N, b, seed=n + N)
    f = ops.riccati_flatten(prob)
    assert ops.riccati_tile_width(handle, n, m) == tile
    L = _lib.riccati_layout(n, m, N)
    ldb = _lib.padded_batch(b)
    names = ("A", "B", "Q", "R", "q", "r", "Qf", "qf", "x0")
    dev = _dev(f, names)
    z = lambda rows: torch.zeros(ldb * rows, dtype=torch.float64, device="cuda")  # noqa: E731
    knots, term, Zp, gains = z(L.knot_count * L.rows_per_knot), z(L.term_rows), z(L.z_rows), z(L.gain_rows)
    Z = torch.zeros(b, L.z_rows, dtype=torch.float64, device="cuda")
    K = torch.zeros(b, N - 1, n, m, dtype=torch.float64, device="cuda")
    kff = torch.zeros(b, N - 1, m, dtype=torch.float64, device="cuda")
    Z2 = torch.zeros_like(Z)
    info = torch.zeros(b, dtype=torch.int32, device="cuda")
    handle.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ops.riccati_pack(handle, n, m, N, b, 0, *[dev[k] for k in names], knots, term)
        ops.riccati_solve_packed(handle, n, m, N, b, 0, knots, term, Zp, gains, info)
        ops.riccati_unpack(handle, n, m, N, b, Zp, gains, Z, K, kff)
        ops.unpack_rows(handle, L.z_rows, b, tile, Zp, Z2)      # the generic entry with the queried tile width
        torch.cuda.synchronize()
    finally:
        handle.set_stream(None)
    X, U, Kh, kffh, ih = ops.riccati_solve_problem(prob, handle=handle)
    Xd, Ud = ops.split_primals(Z.cpu().numpy(), n, m, N)
    assert (ih == 0).all() and int(info.abs().max()) == 0
    assert np.array_equal(Xd, X) and np.array_equal(Ud, U)
    assert torch.equal(Z, Z2)
    assert np.array_equal(np.swapaxes(K.cpu().numpy(), -1, -2), Kh) and np.array_equal(kff.cpu().numpy(), kffh)


@pytest.mark.parametrize("n,m,N,b,mid_p,tile", [(3, 2, 21, 70, 0, 32), (12, 4, 25, 37, 0, 1), (64, 16, 9, 5, 0, 1),
                                                (12, 4, 10, 9, 2, 1), (10, 3, 14, 33, 1, 1), (14, 7, 12, 6, 0, 1), (5, 2, 15, 40, 1, 32)])
def test_kkt_device_resident_all_size_classes(handle, n, m, N, b, mid_p, tile):
    import torch
    prob = problems.random_lqr_kkt(n, m, N, b, seed=n + N, mid_p=mid_p, hess_mode=1)
    f = ops.kkt_flatten(prob)
    p, hm = f["p"], f["hess_mode"]
    assert ops.kkt_tile_width(handle, n, m, N, p, hm) == tile
    NN, P = _lib.num_vars(n, m, N), _lib.num_cons(n, N, p)
    rows = _lib.kkt_data_rows(n, m, N, p, hm)
    ldb = _lib.padded_batch(b)
    names = ("Q", "R", "Hux", "q", "r", "A", "B", "d", "D2", "C", "c")
    dev = _dev(f, names)
    z = lambda r: torch.zeros(ldb * r, dtype=torch.float64, device="cuda")  # noqa: E731
    data, dzp, mp, rp = z(rows), z(NN), z(P), z(NN)
    dz = torch.zeros(b, NN, dtype=torch.float64, device="cuda")
    mult = torch.zeros(b, P, dtype=torch.float64, device="cuda")
    res = torch.zeros(b, NN, dtype=torch.float64, device="cuda")
    info = torch.zeros(b, dtype=torch.int32, device="cuda")
    handle.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ops.kkt_pack(handle, n, m, N, b, p, hm, *[dev[k] for k in names], data)
        ops.kkt_solve_packed(handle, n, m, N, b, p, hm, False, 0, data, dzp, mp, rp, info)
        ops.kkt_unpack(handle, n, m, N, b, p, hm, False, dzp, mp, rp, dz, mult, res)
        torch.cuda.synchronize()
    finally:
        handle.set_stream(None)
    dzh, lamh, ih, resh = ops.kkt_solve_problem(prob, want_res=True, handle=handle)
    assert (ih == 0).all() and int(info.abs().max()) == 0
    assert np.array_equal(dz.cpu().numpy(), dzh) and np.array_equal(mult.cpu().numpy(), lamh)
    assert np.array_equal(res.cpu().numpy(), resh)


def test_unpack_rows_rejects_unknown_tile(handle):
    import torch
    a = torch.zeros(64, dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.LqrbError):
        ops.unpack_rows(handle, 2, 32, 7, a, a.clone())


def test_cooperative_global_workspace_is_per_stream(handle):
    """n = 48, m = 16 on the cooperative kernel needs the global-memory workspace; with host pointers the batch is
    cut into chunks that alternate over two streams, so two of these kernels run at once.  The result must equal
    the single-stream device-pointer path bit for bit."""
    import torch
    n, m, N, b = 48, 16, 6, 8
    prob = problems.random_lqr_kkt(n, m, N, b, seed=21, mid_p=0, hess_mode=1)
    rep = 14                                                     # 112 instances: several 32-instance chunks
    big = {k: (np.tile(v, (rep,) + (1,) * (v.ndim - 1)) if isinstance(v, np.ndarray) and v.ndim > 1 else v)
           for k, v in prob.items()}
    big["C"] = [np.tile(c, (rep, 1, 1)) for c in prob["C"]]
    big["c"] = [np.tile(c, (rep, 1)) for c in prob["c"]]
    f = ops.kkt_flatten(big)
    B = f["batch"]
    NN, P = _lib.num_vars(n, m, N), _lib.num_cons(n, N, f["p"])
    handle.set_option("kkt_variant", 2)
    handle.set_option("host_chunk", 32)
    try:
        dz1, lam1, i1 = ops.kkt_solve_problem(f, handle=handle)
        assert "gmem-ws" in handle.last_kernel, handle.last_kernel
        names = ("Q", "R", "Hux", "q", "r", "A", "B", "d", "D2", "C", "c")
        dev = _dev(f, names)
        dz = torch.zeros(B, NN, dtype=torch.float64, device="cuda")
        mult = torch.zeros(B, P, dtype=torch.float64, device="cuda")
        info = torch.zeros(B, dtype=torch.int32, device="cuda")
        ops.kkt_solve(handle, n, m, N, B, f["p"], f["hess_mode"], 0, *[dev[k] for k in names], dz, mult, None, info)
        handle.synchronize()
    finally:
        handle.set_option("kkt_variant", 0)
        handle.set_option("host_chunk", 0)
    assert (i1 == 0).all() and int(info.abs().max()) == 0
    assert np.array_equal(dz1, dz.cpu().numpy()) and np.array_equal(lam1, mult.cpu().numpy())
    assert np.array_equal(dz1[:b], dz1[b:2 * b])
