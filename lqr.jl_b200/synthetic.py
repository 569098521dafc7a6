"""Seeded synthetic instances generated ON THE DEVICE, chunk by chunk, in the ABI's instance-major column-major
layout (what LQR.jl would hold) — the inputs of bench.py at the BASELINE.json batch sizes, where a host-side
numpy generator would need tens of GB of host memory and minutes of PCIe time.

Same distributions as the numpy generators of ``problems.py`` (SURVEY §8d "Concrete synthetic input"); every
instance is distinct (one torch.Generator per chunk, seeded with (seed, chunk index)).  ``to_math`` turns the
first instances of a chunk back into the math-order numpy dict the parity checks feed to the CPU oracle.
Data generation is not the hot path: plain torch ops are used here on purpose.
"""
from __future__ import annotations

import numpy as np

from . import problems
from ._lib import HESS_BLOCKDIAG


def _gen(seed: int, chunk: int, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(1_000_003 * int(seed) + int(chunk) + 1)
    return g


def _randn(g, *shape):
    import torch
    return torch.randn(*shape, generator=g, device=g.device, dtype=torch.float64)


def _rand(g, *shape):
    import torch
    return torch.rand(*shape, generator=g, device=g.device, dtype=torch.float64)


def _spd(g, lead, k, scale, ridge):
    """L'L + ridge I (symmetric by construction, so its column-major image is itself)."""
    import torch
    L = _randn(g, *lead, k, k) * scale
    S = torch.matmul(L.transpose(-1, -2), L)
    S = 0.5 * (S + S.transpose(-1, -2))
    return S + ridge * torch.eye(k, device=g.device, dtype=torch.float64)


def _cm(a):
    """math order (..., rows, cols) -> column-major buffer (..., cols, rows), contiguous."""
    return a.transpose(-1, -2).contiguous()


# ------------------------------------------------------------------ Riccati
RICCATI_NAMES = ("A", "B", "Q", "R", "q", "r", "Qf", "qf", "x0")


def riccati_cartpole_chunk(count, seed, chunk, N=101, device="cuda"):
    """Config 2 (problems.riccati_cartpole_batch): the cartpole fixture's A, B with 5 % relative noise per knot,
    Q_k = L'L + 1e-2 I, R_k in [0.05, 0.2], q, r ~ N(0,1), Qf = 100 I + symmetric noise, x0 ~ N(0,1)."""
    import torch
    g = _gen(seed, chunk, device)
    base = problems.cartpole_fixture(N)
    n, m = 4, 1
    A0 = torch.from_numpy(base["A"][0, 0]).to(device)
    B0 = torch.from_numpy(base["B"][0, 0]).to(device)
    A = A0 * (1.0 + 0.05 * _randn(g, count, N - 1, n, n))
    B = B0 * (1.0 + 0.05 * _randn(g, count, N - 1, n, m))
    Q = _spd(g, (count, N - 1), n, 0.1, 1e-2)
    R = 0.05 + 0.15 * _rand(g, count, N - 1, m, m)
    S = _randn(g, count, n, n)
    Qf = 100.0 * torch.eye(n, device=device, dtype=torch.float64) + 0.5 * (S + S.transpose(-1, -2))
    return dict(n=n, m=m, N=N, batch=count, A=_cm(A), B=_cm(B), Q=Q.contiguous(), R=R.contiguous(),
                q=_randn(g, count, N - 1, n), r=_randn(g, count, N - 1, m), Qf=Qf.contiguous(),
                qf=_randn(g, count, n), x0=_randn(g, count, n))


def random_riccati_chunk(n, m, N, count, seed, chunk, dt=0.01, device="cuda"):
    """Configs 5a-R / 5b-R (problems.random_lqr_riccati): A_k = I + dt J_k, J ~ N(0,1)/sqrt(n); B_k ~ N(0,1) dt;
    SPD Q_k, R_k; affine q, r."""
    import torch
    g = _gen(seed, chunk, device)
    A = torch.eye(n, device=device, dtype=torch.float64) + dt / np.sqrt(n) * _randn(g, count, N - 1, n, n)
    B = dt * _randn(g, count, N - 1, n, m)
    return dict(n=n, m=m, N=N, batch=count, A=_cm(A), B=_cm(B),
                Q=_spd(g, (count, N - 1), n, 1.0 / np.sqrt(n), 1e-1).contiguous(),
                R=_spd(g, (count, N - 1), m, 1.0 / np.sqrt(m), 1e-1).contiguous(),
                q=_randn(g, count, N - 1, n), r=_randn(g, count, N - 1, m),
                Qf=_spd(g, (count,), n, 1.0 / np.sqrt(n), 1.0).contiguous(), qf=_randn(g, count, n),
                x0=_randn(g, count, n))


def riccati_to_math(f, first=0, count=64):
    """First instances of a device chunk -> the math-order numpy dict of problems.py (for the CPU oracle)."""
    s = slice(first, first + count)
    h = {k: f[k][s].cpu().numpy() for k in RICCATI_NAMES}
    sw = lambda a: np.ascontiguousarray(np.swapaxes(a, -1, -2))  # noqa: E731
    return dict(n=f["n"], m=f["m"], N=f["N"], lti=False, A=sw(h["A"]), B=sw(h["B"]), Q=sw(h["Q"]), R=sw(h["R"]),
                q=h["q"], r=h["r"], Qf=sw(h["Qf"]), qf=h["qf"], x0=h["x0"])


# ------------------------------------------------------------------ KKT (init + dynamics + goal)
KKT_NAMES = ("Q", "R", "Hux", "q", "r", "A", "B", "d", "D2", "C", "c")


def _init_goal(g, count, n, m, N, c_scale=0.1):
    """p = [n, 0, ..., 0, n]: C_1 = [I 0] (x_1 = x_0), C_N = I (goal); column-major flat C, c."""
    import torch
    w = n + m
    C1 = torch.zeros(count, w, n, device=g.device, dtype=torch.float64)  # column-major n x w
    C1[:, :n, :] = torch.eye(n, device=g.device, dtype=torch.float64)
    CN = torch.eye(n, device=g.device, dtype=torch.float64).expand(count, n, n)
    C = torch.cat([C1.reshape(count, -1), CN.reshape(count, -1)], dim=1).contiguous()
    c = (c_scale * _randn(g, count, 2 * n)).contiguous()
    p = np.zeros(N, dtype=np.int32)
    p[0] = p[-1] = n
    return p, C, c


def random_kkt_chunk(n, m, N, count, seed, chunk, dt=0.01, device="cuda"):
    """Configs 5a-K / 5b-K (problems.random_lqr_kkt, mid_p = 0, block-diagonal Hessian)."""
    import torch
    g = _gen(seed, chunk, device)
    A = torch.eye(n, device=device, dtype=torch.float64) + dt / np.sqrt(n) * _randn(g, count, N - 1, n, n)
    B = dt * _randn(g, count, N - 1, n, m)
    if n <= 8:
        B = B + 0.1 * _randn(g, count, N - 1, n, m)
    p, C, c = _init_goal(g, count, n, m, N)
    return dict(n=n, m=m, N=N, batch=count, p=p, hess_mode=HESS_BLOCKDIAG,
                Q=_spd(g, (count, N), n, 1.0 / np.sqrt(n), 1e-1).contiguous(),
                R=_spd(g, (count, N - 1), m, 1.0 / np.sqrt(m), 1e-1).contiguous(), Hux=None,
                q=_randn(g, count, N, n), r=_randn(g, count, N - 1, m), A=_cm(A), B=_cm(B),
                d=0.1 * _randn(g, count, N - 1, n), D2=None, C=C, c=c)


def dubins_kkt_chunk(count, seed, chunk, N=201, dt=0.015, device="cuda"):
    """Config 3 (problems.dubins_kkt_batch): RK3 linearisation of the Dubins car about a random smooth
    (v, omega, theta) trajectory; block-diagonal SPD cost (Q ~ 1e-2, R ~ 1e-2, Qf ~ 100); init + goal rows."""
    import torch
    n, m = 3, 2
    g = _gen(seed, chunk, device)
    t = torch.linspace(0.0, 1.0, N - 1, device=device, dtype=torch.float64)
    ph = 2 * np.pi * _rand(g, count, 2, 1)
    v = 1.0 + 0.3 * torch.sin(2 * np.pi * t + ph[:, 0])
    om = 0.8 * torch.sin(2 * np.pi * t + ph[:, 1])
    th = (2 * np.pi * _rand(g, count, 1) - np.pi) + torch.cumsum(om * dt, dim=1)
    # analytic RK3 Jacobians (problems.dubins_rk3_jacobians): theta evolves linearly inside a step
    th2, th3 = th + 0.5 * dt * om, th + dt * om
    cb = (torch.cos(th) + 4 * torch.cos(th2) + torch.cos(th3)) / 6
    sb = (torch.sin(th) + 4 * torch.sin(th2) + torch.sin(th3)) / 6
    dc = (-4 * torch.sin(th2) * 0.5 * dt - torch.sin(th3) * dt) / 6
    ds = (4 * torch.cos(th2) * 0.5 * dt + torch.cos(th3) * dt) / 6
    A = torch.zeros(count, N - 1, n, n, device=device, dtype=torch.float64)
    B = torch.zeros(count, N - 1, n, m, device=device, dtype=torch.float64)
    A[..., 0, 0] = A[..., 1, 1] = A[..., 2, 2] = 1.0
    A[..., 0, 2] = -dt * v * sb
    A[..., 1, 2] = dt * v * cb
    B[..., 0, 0] = dt * cb
    B[..., 1, 0] = dt * sb
    B[..., 0, 1] = dt * v * dc
    B[..., 1, 1] = dt * v * ds
    B[..., 2, 1] = dt
    Q = _spd(g, (count, N), n, 0.05, 1e-2)
    Q[:, -1] = _spd(g, (count,), n, 1.0, 100.0)
    p, C, c = _init_goal(g, count, n, m, N)
    return dict(n=n, m=m, N=N, batch=count, p=p, hess_mode=HESS_BLOCKDIAG, Q=Q.contiguous(),
                R=_spd(g, (count, N - 1), m, 0.05, 1e-2).contiguous(), Hux=None, q=_randn(g, count, N, n),
                r=_randn(g, count, N - 1, m), A=_cm(A), B=_cm(B), d=0.01 * _randn(g, count, N - 1, n), D2=None,
                C=C, c=c)


def kkt_to_math(f, first=0, count=64):
    """First instances of a device KKT chunk (init + goal pattern) -> math-order numpy dict for the oracle."""
    n, m, N = f["n"], f["m"], f["N"]
    s = slice(first, first + count)
    sw = lambda a: np.ascontiguousarray(np.swapaxes(a[s].cpu().numpy(), -1, -2))  # noqa: E731
    w = n + m
    Cf, cf = f["C"][s].cpu().numpy(), f["c"][s].cpu().numpy()
    b = Cf.shape[0]
    C1 = np.swapaxes(Cf[:, :n * w].reshape(b, w, n), -1, -2)
    CN = np.swapaxes(Cf[:, n * w:].reshape(b, n, n), -1, -2)
    Cs = [C1] + [np.zeros((b, 0, w)) for _ in range(N - 2)] + [CN]
    cs = [cf[:, :n]] + [np.zeros((b, 0)) for _ in range(N - 2)] + [cf[:, n:]]
    return dict(n=n, m=m, N=N, p=f["p"].copy(), hess_mode=f["hess_mode"], Q=sw(f["Q"]), R=sw(f["R"]), Hux=None,
                q=f["q"][s].cpu().numpy(), r=f["r"][s].cpu().numpy(), A=sw(f["A"]), B=sw(f["B"]),
                d=f["d"][s].cpu().numpy(), D2=None, C=Cs, c=cs)


# ------------------------------------------------------------------ Dubins SQP (config 4)
def dubins_turn90_device(count, seed, N=201, tf=3.0, device="cuda"):
    """Config 4 (problems.dubins_turn90): x0 = 0, xf ~ [1.5, 1.5, pi/2] + N(0, 0.1^2), initial guess = the u = 0.1
    rollout.  Returns Z0 (count, NN) in Primals order, x0, xf (device tensors) and the options dict."""
    import torch
    n, m = 3, 2
    g = _gen(seed, 0, device)
    Z1, _, _, o = problems.dubins_turn90(1, N=N, tf=tf, seed=seed)   # the initial guess does not depend on xf
    Z0 = torch.from_numpy(Z1).to(device).expand(count, Z1.shape[1]).contiguous()
    x0 = torch.zeros(count, n, device=device, dtype=torch.float64)
    xf = torch.tensor([1.5, 1.5, np.pi / 2], device=device, dtype=torch.float64) + 0.1 * _randn(g, count, n)
    return Z0, x0, xf.contiguous(), o
