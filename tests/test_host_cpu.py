"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/lqrb200.h declares, the
host-side mirror of the reference interface behaves, and the multi-rank plumbing works under gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import lqr_b200
from lqr_b200 import _lib, dist, ops, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "lqrb200.h")).read()
    declared = set(re.findall(r"\b(lqrb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lqrb_context"}
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    # and really from the in-tree shared object
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in declared:
        assert f" T {s}" in out, s


def test_layout_queries_need_no_gpu():
    assert _lib.lib().lqrb_version() == 100
    assert _lib.padded_batch(1) == 32 and _lib.padded_batch(65536) == 65536 and _lib.padded_batch(33) == 64
    L = _lib.riccati_layout(4, 1, 101)
    assert (L.rows_per_knot, L.knot_count, L.term_rows, L.z_rows, L.gain_rows) == (36, 100, 18, 504, 500)
    assert _lib.riccati_layout(4, 1, 101, _lib.FLAG_LTI).knot_count == 1
    # knot records of the warp-per-instance size class travel by 16-byte bulk copies: odd lengths get one padding row
    for n, m in [(12, 4), (12, 3), (12, 2), (8, 3), (8, 2), (8, 1)]:
        raw = n * n + n * m + n * (n + 1) // 2 + m * (m + 1) // 2 + n + m
        assert _lib.riccati_layout(n, m, 11).rows_per_knot == raw + raw % 2
    assert _lib.riccati_layout(7, 3, 11).rows_per_knot == 49 + 21 + 28 + 6 + 7 + 3  # other sizes: no padding
    # SURVEY §8d algorithmic bytes of config 2: 8 * (rows in + rows out) = 32,976
    assert 8 * (L.rows_per_knot * L.knot_count + L.term_rows + L.z_rows) == 32976
    p = np.array([3] + [0] * 199 + [3], dtype=np.int32)
    assert _lib.num_vars(3, 2, 201) == 1003 and _lib.num_cons(3, 201, p) == 606
    rows = _lib.kkt_data_rows(3, 2, 201, p, _lib.HESS_BLOCKDIAG)
    # first knot 9+5+15+3+15+3 = 50, 199 mid knots x (H 6+3, g 5, D1 15, d 3 = 32), last knot 6+3+9+3 = 21
    assert rows == 50 + 199 * 32 + 21
    assert _lib.kkt_data_rows(3, 2, 201, p, _lib.HESS_DIAG) < rows < _lib.kkt_data_rows(3, 2, 201, p, _lib.HESS_DENSE)


def test_no_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(_lib.LqrbError, match="no fallback"):
        _lib.Handle(0)
    with pytest.raises(_lib.LqrbError):
        ops.riccati_solve_problem(problems.riccati_cartpole_batch(2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lqr.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "lqr_oracle" not in src, f


def test_primals_layout():
    """Z = [x1;u1;...;xN] with aliasing views (src/lqr_problem.jl:46-73; test/view_knotpoint.jl)."""
    Z = lqr_b200.Primals(3, 2, 5, batch=2)
    assert Z.Z.shape == (2, 3 * 5 + 2 * 4)
    Z.X[1][:] = 7.0
    Z.U[3][:] = -1.0
    assert (Z.Z[:, 5:8] == 7.0).all() and (Z.Z[:, 18:20] == -1.0).all()
    W = 2.0 * (Z + Z)
    assert (W.Z == 4 * Z.Z).all() and W.Z is not Z.Z


def test_lqr_problem_sizes():
    pr = problems.random_lqr_riccati(6, 3, 31, 4, lti=True)
    prob = lqr_b200.LQRProblem(pr["Qf"], pr["Q"], pr["R"], pr["A"], pr["B"], pr["x0"], N=31)
    assert lqr_b200.size(prob) == (6, 3, 31) and lqr_b200.num_vars(prob) == 31 * 6 + 30 * 3
    assert not prob.ltv and prob.batch == 4
    dprob = lqr_b200.LQRProblem(np.ones(6), np.ones(6), np.ones(3), pr["A"][0], pr["B"][0], pr["x0"][0], N=31)
    assert dprob.Q.shape == (1, 6, 6) and np.array_equal(dprob.Q[0], np.eye(6))


def test_constraint_block_shapes_match_reference_test():
    """test/constraint_blocks.jl:25-32 on DoubleIntegrator(3,101): n=6, m=3."""
    n, m, N = 6, 3, 101
    prob = problems.double_integrator_fixture()
    blocks = lqr_b200.ConstraintBlocks(n, m, N, prob["p"])
    assert blocks[0].Y.shape[1:] == (2 * n, n + m)
    assert lqr_b200.dims(blocks[0]) == (0, n, n)
    assert blocks[N - 1].Y.shape[1:] == (2 * n, n)
    assert lqr_b200.dims(blocks[N - 1]) == (n, n, 0)
    assert lqr_b200.dims(blocks[N // 2]) == (n, 1, n)
    # views alias the block storage (:40-59)
    blocks[1].D1[:] = 3.0
    blocks[1].c[:] = 1.5
    assert (blocks[1].Y[:, n + 1:] == 3.0).all() and (blocks[1].y[:, :1] == 1.5).all()
    assert lqr_b200.num_constraints(blocks) == _lib.num_cons(n, N, prob["p"])


def test_dense_extractors_match_oracle_assembly():
    """get_linearized_constraints / get_cost_expansion (src/cholesky_solver.jl:278-297) against the
    independent global assembly used by the oracle."""
    from oracle import dense_kkt

    class Dummy(lqr_b200.CholeskySolver):
        def __init__(self, prob):   # no handle: extractors are host-only
            self.prob, self.n, self.m, self.N = prob, prob["n"], prob["m"], prob["N"]
            self.p = np.asarray(prob["p"], dtype=np.int32)
    for prob in (problems.double_integrator_fixture(), problems.random_lqr_kkt(4, 2, 7, 2, mid_p=1, hess_mode=0, explicit_D2=True)):
        s = Dummy(prob)
        D, d = lqr_b200.get_linearized_constraints(s, 0)
        H, g = lqr_b200.get_cost_expansion(s, 0)
        Ho, go, Do, do = dense_kkt.assemble(prob, 0)
        assert np.allclose(D, Do.toarray()) and np.allclose(d, do) and np.allclose(H, Ho.toarray()) and np.allclose(g, go)


def test_flatten_is_column_major_instance_major():
    prob = problems.random_lqr_kkt(3, 2, 4, 2, seed=0, mid_p=1)
    f = ops.kkt_flatten(prob)
    assert f["A"].shape == (2, 3, 3, 3) and f["A"][1, 2, 0, 1] == prob["A"][1, 2, 1, 0]
    assert f["C"].shape == (2, 3 * 5 + 2 * 1 * 5 + 3 * 3) and f["c"].shape == (2, 3 + 2 + 3)
    Z = np.arange(2 * (3 * 4 + 2 * 3), dtype=float).reshape(2, -1)
    X, U = ops.split_primals(Z, 3, 2, 4)
    assert X.shape == (2, 4, 3) and U.shape == (2, 3, 2) and X[0, 1, 0] == 5 and U[0, 0, 0] == 3 and X[0, 3, 0] == 15


def test_batch_slices_partition():
    for batch, world in [(65536, 8), (100, 3), (31, 2), (1, 4), (262144, 7)]:
        edges = [dist.batch_slice(batch, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == batch
        for (a, b), (c, d) in zip(edges, edges[1:]):
            assert b == c and a <= b and (b % 32 == 0 or b == batch)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from lqr_b200 import dist as ld
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
lo, hi = ld.batch_slice(1000, rank, 2)
ms = 10.0 if rank == 0 else 25.0
thr = ld.aggregate_throughput(hi - lo, ms)
mx = ld.max_over_ranks([ms, float(rank)])
assert abs(thr - 1000 / 25e-3) < 1e-6, thr
assert mx == [25.0, 1.0], mx
dist.barrier()
dist.destroy_process_group()
print("ok", rank, lo, hi)
"""


def test_two_rank_gloo_timing_protocol(tmp_path):
    """The N>1 path of bench.py: slices + barrier + max-over-ranks time + summed units, world_size 2."""
    port = 29600 + os.getpid() % 300
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert sorted(o.split()[2] for o, _ in outs) == ["0", "512"]


def test_header_is_plain_c_and_the_c_caller_links(tmp_path, build_abi_smoke):
    """include/lqrb200.h is valid C99 (no C++-isms, no torch types) and tests/abi_smoke.c — the non-Python caller
    of the ABI — compiles against it with -Wall -Wextra -Werror and links the in-tree library (run on the GPU by
    tests/test_gpu_abi.py)."""
    exe = build_abi_smoke(str(tmp_path))
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "liblqrb200.so" in out and "not found" not in out.split("liblqrb200.so")[1].splitlines()[0]


def test_julia_shim_binds_every_header_symbol():
    """lqr.jl_b200/julia/LQRB200.jl ccalls every exported entry of include/lqrb200.h, with the argument count the
    header declares (the shim is static — Julia is not installed — so this is the check that it cannot drift)."""
    hdr = open(os.path.join(ROOT, "include", "lqrb200.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {}
    for mm in re.finditer(r"\b(lqrb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr_nc, flags=re.S):
        args = mm.group(2).strip()
        decl[mm.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    assert set(decl) == set(_lib.EXPORTS)
    jl = open(os.path.join(ROOT, "lqr.jl_b200", "julia", "LQRB200.jl")).read()
    bound = {}
    for mm in re.finditer(r"ccall\(\(:(lqrb_[a-z0-9_]+), lib\),\s*([A-Za-z0-9{}]+),\s*\(([^)]*)\)", jl, flags=re.S):
        types = [t for t in mm.group(3).replace("\n", " ").split(",") if t.strip()]
        bound.setdefault(mm.group(1), set()).add(len(types))
    missing = sorted(set(decl) - set(bound))
    assert not missing, f"the Julia shim does not bind: {missing}"
    for name, counts in bound.items():
        assert name in decl, f"{name} is not in the header"
        assert counts == {decl[name]}, (name, counts, decl[name])
    # the reference's own names are exported
    for name in ("LQRProblem", "DPSolver", "solve!", "BlockCholesky", "InvertedQuadratic", "update_cholesky!",
                 "ConstraintBlock", "dims", "build_shur_factors", "calculate_shur_factors!", "forward_substitution!",
                 "backward_substitution!", "calculate_primals!", "CholeskySolver", "_solve!", "step!", "residual",
                 "get_step", "get_multipliers", "get_shur_factors", "get_cholesky", "copy_shur_factors!", "rollout!",
                 "second_order_correction!", "num_vars"):
        assert re.search(r"\b" + re.escape(name) + r"(?![A-Za-z0-9_])", jl.split("export", 1)[1]), name


def test_gen_con_inds_orderings():
    """gen_con_inds (src/conblocks.jl:122-166) on the DoubleIntegrator constraint list of test/problems.jl:39-48:
    goal (n rows at knot N), a 1-row plane constraint on knots 2..N-1, the dynamics (n rows on knots 1..N-1)."""
    n, N = 6, 5
    cons = [(n, range(N - 1, N)), (1, range(1, N - 1)), (n, range(0, N - 1))]
    byc = lqr_b200.gen_con_inds(cons, N, "by_constraint")
    assert byc[0][0] == range(0, 6) and byc[1][0] == range(6, 7) and byc[1][2] == range(8, 9) and byc[2][0] == range(9, 15)
    byk = lqr_b200.gen_con_inds(cons, N, "by_knotpoint")
    # knot 0: dynamics; knot 1: plane then dynamics; ...; knot 4: goal
    assert byk[2][0] == range(0, 6) and byk[1][0] == range(6, 7) and byk[2][1] == range(7, 13) and byk[0][0] == range(27, 33)
    assert sum(len(r) for rows in byk for r in rows) == n + (N - 2) + n * (N - 1)
    byb = lqr_b200.gen_con_inds(cons, N, "by_block")
    # per-knot offsets: at an interior knot the stage row comes first, then the dynamics rows
    assert byb[1][0] == range(0, 1) and byb[2][1] == range(1, 7) and byb[2][0] == range(0, 6) and byb[0][0] == range(0, 6)
    with pytest.raises(ValueError):
        lqr_b200.gen_con_inds(cons, N, "nope")
