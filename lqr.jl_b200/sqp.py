"""Batched Dubins SQP: the fixed-count outer loop of solve!(::CholeskySolver) (src/cholesky_solver.jl:109-153)
with the globalisation spec of src/sqp.jl:72-94, run entirely on the device (lqrb_sqp_dubins_f64)."""
from __future__ import annotations

import numpy as np

from . import _lib, ops


class DubinsSQP:
    """solver = DubinsSQP(x0, xf, N=11, tf=3.0); solver.solve_(Z0) -> Z.

    Attributes after solve_(): Z (batch, NN) Primals order, feas_p / feas_d per instance (the two numbers
    step! @shows, src/cholesky_solver.jl:129-134), iters per instance, kkt_solves (total)."""

    def __init__(self, x0, xf, N=11, tf=3.0, Q=1e-2, R=1e-2, Qf=100.0, iters=10, line_search=True,
                 eps_p=1e-5, eps_d=1e-5, handle=None, device=0):
        self.x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        self.xf = np.ascontiguousarray(np.atleast_2d(xf), dtype=np.float64)
        self.batch = self.x0.shape[0]
        self.opts = dict(N=N, iters=iters, dt=tf / (N - 1), q_diag=Q, r_diag=R, qf_diag=Qf, eps_p=eps_p,
                         eps_d=eps_d, line_search=int(line_search))
        self.handle = handle or ops.default_handle(device)
        self.NN = _lib.num_vars(3, 2, N)

    def solve_(self, Z0):
        self.Z = np.ascontiguousarray(np.broadcast_to(Z0, (self.batch, self.NN)), dtype=np.float64).copy()
        self.feas_p = np.zeros(self.batch)
        self.feas_d = np.zeros(self.batch)
        self.iters = np.zeros(self.batch, dtype=np.int32)
        self.kkt_solves = ops.sqp_dubins(self.handle, self.batch, self.opts, self.x0, self.xf, self.Z,
                                         self.feas_p, self.feas_d, self.iters)
        return self.Z
