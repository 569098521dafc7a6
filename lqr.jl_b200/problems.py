"""Seeded synthetic problem fixtures (host side, numpy only).

These stand in for the reference's ``test/problems.jl`` fixtures and for the five
BASELINE.json configs.  The reference builds its problems through
un-vendored packages (RobotZoo / RobotDynamics / TrajOptCore), so the model
dynamics below are re-derived from first principles and are NOT reference-pinned;
what the solver consumes is the linearised block data these functions emit.

All arrays are "math order" numpy with the batch axis leading:
``A[b, k]`` is the n x n matrix of instance b, knot k (0-based).
"""
from __future__ import annotations

import numpy as np

HESS_DENSE, HESS_BLOCKDIAG, HESS_DIAG = 0, 1, 2


# ------------------------------------------------------------------ dynamics
def cartpole_dynamics(x, u, mc=1.0, mp=0.2, l=0.5, g=9.81):
    """Cart-pole manipulator equations, x=[pos, theta, vel, omega] (RobotZoo's model; unpinned)."""
    s, c = np.sin(x[1]), np.cos(x[1])
    Hm = np.array([[mc + mp, mp * l * c], [mp * l * c, mp * l * l]])
    Cm = np.array([[0.0, -mp * x[3] * l * s], [0.0, 0.0]])
    G = np.array([0.0, mp * g * l * s])
    Bm = np.array([1.0, 0.0])
    qdd = -np.linalg.solve(Hm, Cm @ x[2:] + G - Bm * u[0])
    return np.concatenate([x[2:], qdd])


def dubins_dynamics(x, u):
    """Dubins car: x=[px, py, theta], u=[v, omega]."""
    return np.array([u[0] * np.cos(x[2]), u[0] * np.sin(x[2]), u[1]])


def rk3(f, x, u, dt):
    """Explicit RK3 as the reference's tests integrate (test/cartpole.jl:38, RobotDynamics RK3)."""
    k1 = f(x, u) * dt
    k2 = f(x + k1 / 2, u) * dt
    k3 = f(x - k1 + 2 * k2, u) * dt
    return x + (k1 + 4 * k2 + k3) / 6


def linearize_fd(f, x, u, dt, eps=1e-6):
    """Central-difference Jacobians of the RK3 map (inputs are just data; accuracy is immaterial)."""
    n, m = len(x), len(u)
    A = np.zeros((n, n))
    B = np.zeros((n, m))
    for j in range(n):
        e = np.zeros(n)
        e[j] = eps
        A[:, j] = (rk3(f, x + e, u, dt) - rk3(f, x - e, u, dt)) / (2 * eps)
    for j in range(m):
        e = np.zeros(m)
        e[j] = eps
        B[:, j] = (rk3(f, x, u + e, dt) - rk3(f, x, u - e, dt)) / (2 * eps)
    return A, B


# ------------------------------------------------------- KKT problem helpers
def _init_goal_blocks(b, n, m, N, c_init, c_goal, mid_C=None, mid_c=None):
    """Stage-constraint blocks: x_1 = x0 at knot 1 (C=[I 0]), optional mid rows, goal at N (C=I)."""
    pm = 0 if mid_C is None else mid_C.shape[-2]
    p = np.full(N, pm, dtype=np.int32)
    p[0] = n
    p[-1] = n
    Cs, cs = [], []
    C0 = np.zeros((b, n, n + m))
    C0[:, :, :n] = np.eye(n)
    Cs.append(C0)
    cs.append(c_init)
    for k in range(1, N - 1):
        if pm:
            Cs.append(mid_C[:, k])
            cs.append(mid_c[:, k])
        else:
            Cs.append(np.zeros((b, 0, n + m)))
            cs.append(np.zeros((b, 0)))
    Cs.append(np.broadcast_to(np.eye(n), (b, n, n)).copy())
    cs.append(c_goal)
    return p, Cs, cs


def cartpole_fixture(N=101):
    """Config 1: the reference's Cartpole fixture (test/problems.jl:58-88) linearised about the
    u=0.01 rollout: Q=1e-2 I, Qf=100 I, R=0.1 I, tf=5, x0=0, xf=[0,pi,0,0]; init + dynamics + goal
    constraints; cost expansion scaled by dt on non-terminal knots (test/sparse_solver.jl:67-72)."""
    n, m = 4, 1
    tf = 5.0
    dt = tf / (N - 1)
    Qd, Qfd, Rd = 1e-2 * np.ones(n), 100.0 * np.ones(n), 0.1 * np.ones(m)
    x0 = np.zeros(n)
    xf = np.array([0.0, np.pi, 0.0, 0.0])
    X = np.zeros((N, n))
    U = np.full((N - 1, m), 0.01)
    X[0] = x0
    A = np.zeros((N - 1, n, n))
    B = np.zeros((N - 1, n, m))
    for k in range(N - 1):
        X[k + 1] = rk3(cartpole_dynamics, X[k], U[k], dt)
        A[k], B[k] = linearize_fd(cartpole_dynamics, X[k], U[k], dt)
    Q = np.zeros((N, n, n))
    q = np.zeros((N, n))
    for k in range(N - 1):
        Q[k] = np.diag(Qd * dt)
        q[k] = Qd * (X[k] - xf) * dt
    Q[N - 1] = np.diag(Qfd)
    q[N - 1] = Qfd * (X[N - 1] - xf)
    R = np.tile(np.diag(Rd * dt), (N - 1, 1, 1))
    r = Rd * U * dt
    p, Cs, cs = _init_goal_blocks(1, n, m, N, (X[0] - x0)[None], (X[N - 1] - xf)[None])
    return dict(n=n, m=m, N=N, p=p, hess_mode=HESS_DIAG, Q=Q[None], R=R[None], Hux=None, q=q[None],
                r=r[None], A=A[None], B=B[None], d=np.zeros((1, N - 1, n)), D2=None, C=Cs, c=cs,
                X=X, U=U, dt=dt, xf=xf)


def double_integrator_fixture(D=3, N=101, seed=1, dense_cost=False):
    """The reference's DoubleIntegrator(D,N) fixture (test/problems.jl:14-56): n=2D, m=D,
    Q=diag(10*1_D,1_D), R=0.1 I, Qf=10Q, tf=2, x0=[1_D;0_D], xf=0, a random p=max(D-2,1)-row
    plane constraint on knots 2..N-1, goal at N, init + dynamics constraints; zero-control rollout."""
    rng = np.random.default_rng(seed)
    n, m = 2 * D, D
    tf = 2.0
    dt = tf / (N - 1)
    A1 = np.eye(n)
    A1[:D, D:] = dt * np.eye(D)
    B1 = np.vstack([0.5 * dt * dt * np.eye(D), dt * np.eye(D)])
    Qm = np.diag(np.concatenate([10.0 * np.ones(D), np.ones(D)]))
    Rm = 0.1 * np.eye(m)
    mode = HESS_DIAG
    if dense_cost:
        Lq, Lr = rng.random((n, n)), rng.random((m, m))
        Qm, Rm = Lq.T @ Lq + np.eye(n), Lr.T @ Lr + np.eye(m)
        mode = HESS_BLOCKDIAG
    x0 = np.concatenate([np.ones(D), np.zeros(D)])
    xf = np.zeros(n)
    X = np.tile(x0, (N, 1))          # rollout with u = 0 and zero initial velocity
    U = np.zeros((N - 1, m))
    pm = max(D - 2, 1)
    Ac = rng.random((pm, n))
    Q = np.zeros((N, n, n))
    q = np.zeros((N, n))
    for k in range(N - 1):
        Q[k] = Qm * dt
        q[k] = Qm @ (X[k] - xf) * dt
    Q[N - 1] = 10 * Qm
    q[N - 1] = 10 * Qm @ (X[N - 1] - xf)
    R = np.tile(Rm * dt, (N - 1, 1, 1))
    r = (U @ Rm.T) * dt
    mid_C = np.zeros((1, N, pm, n + m))
    mid_C[0, :, :, :n] = Ac
    mid_c = np.zeros((1, N, pm))
    mid_c[0] = X @ Ac.T
    p, Cs, cs = _init_goal_blocks(1, n, m, N, (X[0] - x0)[None], (X[N - 1] - xf)[None], mid_C, mid_c)
    return dict(n=n, m=m, N=N, p=p, hess_mode=mode, Q=Q[None], R=R[None], Hux=None, q=q[None],
                r=r[None], A=np.tile(A1, (1, N - 1, 1, 1)), B=np.tile(B1, (1, N - 1, 1, 1)),
                d=np.zeros((1, N - 1, n)), D2=None, C=Cs, c=cs)


# ----------------------------------------------------------- batched configs
def _spd(rng, shape, k, scale, ridge):
    L = rng.standard_normal(shape + (k, k)) * scale
    return np.einsum("...ki,...kj->...ij", L, L) + ridge * np.eye(k)


def riccati_cartpole_batch(batch=65536, seed=0, N=101, dtype=np.float64):
    """Config 2: per instance the config-1 A_k,B_k perturbed by N(0,0.05^2) relative
    noise per knot (LTV), Q_k = L'L + 1e-2 I with L~N(0,0.1^2), R_k in [0.05,0.2], q,r~N(0,1),
    Qf = 100 I + symmetric noise, x0~N(0,1)."""
    base = cartpole_fixture(N)
    n, m = 4, 1
    rng = np.random.default_rng(seed)
    A = base["A"][0][None] * (1.0 + 0.05 * rng.standard_normal((batch, N - 1, n, n)))
    B = base["B"][0][None] * (1.0 + 0.05 * rng.standard_normal((batch, N - 1, n, m)))
    Q = _spd(rng, (batch, N - 1), n, 0.1, 1e-2)
    R = rng.uniform(0.05, 0.2, (batch, N - 1, m, m))
    q = rng.standard_normal((batch, N - 1, n))
    r = rng.standard_normal((batch, N - 1, m))
    S = rng.standard_normal((batch, n, n))
    Qf = 100.0 * np.eye(n) + 0.5 * (S + np.swapaxes(S, -1, -2))
    qf = rng.standard_normal((batch, n))
    x0 = rng.standard_normal((batch, n))
    return dict(n=n, m=m, N=N, lti=False, A=A, B=B, Q=Q, R=R, q=q, r=r, Qf=Qf, qf=qf, x0=x0)


def random_lqr_riccati(n, m, N, batch, seed=3, dt=0.01, lti=False):
    """Configs 5a/5b Riccati form: A_k = I + dt*J_k, J~N(0,1)/sqrt(n); B_k~N(0,1)*dt;
    SPD Q_k, R_k; affine q, r."""
    rng = np.random.default_rng(seed)
    kn = () if lti else (N - 1,)
    A = np.eye(n) + dt * rng.standard_normal((batch,) + kn + (n, n)) / np.sqrt(n)
    B = dt * rng.standard_normal((batch,) + kn + (n, m))
    Q = _spd(rng, (batch,) + kn, n, 1.0 / np.sqrt(n), 1e-1)
    R = _spd(rng, (batch,) + kn, m, 1.0 / np.sqrt(m), 1e-1)
    q = rng.standard_normal((batch,) + kn + (n,))
    r = rng.standard_normal((batch,) + kn + (m,))
    Qf = _spd(rng, (batch,), n, 1.0 / np.sqrt(n), 1.0)
    qf = rng.standard_normal((batch, n))
    x0 = rng.standard_normal((batch, n))
    return dict(n=n, m=m, N=N, lti=lti, A=A, B=B, Q=Q, R=R, q=q, r=r, Qf=Qf, qf=qf, x0=x0)


def dare_lti_riccati(n, m, N, batch, seed=1, unstable=False):
    """LTI problem with a well-damped optimal closed loop (A = 0.9 I + 0.2 J / sqrt(n), B ~ N(0,1) / sqrt(n), SPD Q, R, no
    affine terms, Qf = Q): the gain of the first knot of a long horizon converges to the gain of the discrete algebraic
    Riccati equation, which scipy.linalg.solve_discrete_are computes independently (tests).  ``unstable``: an open-loop
    unstable A — the recursion in the reference's form (no symmetrisation of P) then breaks down, the kernels do not."""
    rng = np.random.default_rng(seed)
    if unstable:  # open-loop spectral radius 1.2-1.3, strong actuation
        A = np.eye(n) + 0.3 * rng.standard_normal((batch, n, n)) / np.sqrt(n)
        B = rng.standard_normal((batch, n, m))
    else:
        A = 0.9 * np.eye(n) + 0.2 * rng.standard_normal((batch, n, n)) / np.sqrt(n)
        B = rng.standard_normal((batch, n, m)) / np.sqrt(n)
    L = rng.standard_normal((batch, n, n))
    Q = np.einsum("bij,bkj->bik", L, L) / n + np.eye(n)
    L = rng.standard_normal((batch, m, m))
    R = np.einsum("bij,bkj->bik", L, L) / m + 0.5 * np.eye(m)
    return dict(n=n, m=m, N=N, lti=True, A=A, B=B, Q=Q, R=R, q=np.zeros((batch, n)), r=np.zeros((batch, m)), Qf=Q.copy(),
                qf=np.zeros((batch, n)), x0=rng.standard_normal((batch, n)))


def random_lqr_kkt(n, m, N, batch, seed=3, dt=0.01, mid_p=0, hess_mode=HESS_BLOCKDIAG,
                   explicit_D2=False):
    """Configs 5a-K/5b-K and generic test problems: random LTV dynamics, SPD cost blocks, init + goal
    equality constraints, optional ``mid_p`` random stage rows on knots 2..N-1."""
    rng = np.random.default_rng(seed)
    A = np.eye(n) + dt * rng.standard_normal((batch, N - 1, n, n)) / np.sqrt(n)
    B = dt * rng.standard_normal((batch, N - 1, n, m)) + (0.1 if n <= 8 else 0.0) * \
        rng.standard_normal((batch, N - 1, n, m))
    Q = _spd(rng, (batch, N), n, 1.0 / np.sqrt(n), 1e-1)
    R = _spd(rng, (batch, N - 1), m, 1.0 / np.sqrt(m), 1e-1)
    Hux = None
    if hess_mode == HESS_DENSE:
        Hux = 0.05 / (np.sqrt(m) + np.sqrt(n)) * rng.standard_normal((batch, N - 1, m, n))  # keeps H > 0
    q = rng.standard_normal((batch, N, n))
    r = rng.standard_normal((batch, N - 1, m))
    d = 0.1 * rng.standard_normal((batch, N - 1, n))
    mid_C = mid_c = None
    if mid_p:
        mid_C = rng.standard_normal((batch, N, mid_p, n + m))
        mid_c = 0.1 * rng.standard_normal((batch, N, mid_p))
    p, Cs, cs = _init_goal_blocks(batch, n, m, N, 0.1 * rng.standard_normal((batch, n)),
                                  0.1 * rng.standard_normal((batch, n)), mid_C, mid_c)
    D2 = None
    if explicit_D2:
        D2 = []
        for k in range(1, N):
            w = n + (m if k < N - 1 else 0)
            blk = np.zeros((batch, n, w))
            blk[:, :, :n] = -np.eye(n) + 0.05 * rng.standard_normal((batch, n, n))
            blk[:, :, n:] = 0.05 * rng.standard_normal((batch, n, w - n))
            D2.append(blk)
    return dict(n=n, m=m, N=N, p=p, hess_mode=hess_mode, Q=Q, R=R, Hux=Hux, q=q, r=r, A=A, B=B,
                d=d, D2=D2, C=Cs, c=cs)


def dubins_rk3_jacobians(th, v, om, dt):
    """Vectorised RK3 linearisation of the Dubins car about (theta, v, omega) arrays of equal shape.
    Returns A (...,3,3), B (...,3,2) by differentiating the three RK3 stages analytically."""
    def f(th_, v_):
        return np.stack([v_ * np.cos(th_), v_ * np.sin(th_)], -1)
    # theta evolves linearly: stage angles
    th1, th2, th3 = th, th + 0.5 * dt * om, th + dt * om   # (x - k1 + 2k2)[theta] = th + dt*om
    c = (np.cos(th1) + 4 * np.cos(th2) + np.cos(th3)) / 6
    s = (np.sin(th1) + 4 * np.sin(th2) + np.sin(th3)) / 6
    shape = th.shape
    A = np.zeros(shape + (3, 3))
    B = np.zeros(shape + (3, 2))
    A[..., 0, 0] = A[..., 1, 1] = A[..., 2, 2] = 1.0
    A[..., 0, 2] = -dt * v * s
    A[..., 1, 2] = dt * v * c
    B[..., 0, 0] = dt * c
    B[..., 1, 0] = dt * s
    # d/d omega of the stage angles: 0, dt/2, dt
    dc = (-4 * np.sin(th2) * 0.5 * dt - np.sin(th3) * dt) / 6
    ds = (4 * np.cos(th2) * 0.5 * dt + np.cos(th3) * dt) / 6
    B[..., 0, 1] = dt * v * dc
    B[..., 1, 1] = dt * v * ds
    B[..., 2, 1] = dt
    return A, B


def dubins_kkt_batch(batch=262144, seed=1, N=201, dt=0.015, mid_p=0):
    """Config 3: n=3, m=2; A_k,B_k = RK3 linearisation about a random smooth
    (v_k, omega_k, theta_k) trajectory; block-diagonal SPD H_k (Q~1e-2, R~1e-2, Qf~100), g~N(0,1),
    init + goal equalities (p_1 = p_N = 3), optional p=1 mid rows (test/problems.jl:39-43)."""
    n, m = 3, 2
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 1.0, N - 1)
    ph = rng.uniform(0, 2 * np.pi, (batch, 2, 1))
    v = 1.0 + 0.3 * np.sin(2 * np.pi * t + ph[:, 0])
    om = 0.8 * np.sin(2 * np.pi * t + ph[:, 1])
    th = rng.uniform(-np.pi, np.pi, (batch, 1)) + np.cumsum(om * dt, axis=1)
    A, B = dubins_rk3_jacobians(th, v, om, dt)
    Q = _spd(rng, (batch, N), n, 0.05, 1e-2)
    Q[:, -1] = _spd(rng, (batch,), n, 1.0, 100.0)
    R = _spd(rng, (batch, N - 1), m, 0.05, 1e-2)
    q = rng.standard_normal((batch, N, n))
    r = rng.standard_normal((batch, N - 1, m))
    d = 0.01 * rng.standard_normal((batch, N - 1, n))
    mid_C = mid_c = None
    if mid_p:
        mid_C = rng.standard_normal((batch, N, mid_p, n + m))
        mid_c = 0.1 * rng.standard_normal((batch, N, mid_p))
    p, Cs, cs = _init_goal_blocks(batch, n, m, N, 0.1 * rng.standard_normal((batch, n)),
                                  0.1 * rng.standard_normal((batch, n)), mid_C, mid_c)
    return dict(n=n, m=m, N=N, p=p, hess_mode=HESS_BLOCKDIAG, Q=Q, R=R, Hux=None, q=q, r=r, A=A,
                B=B, d=d, D2=None, C=Cs, c=cs)


def dubins_turn90(batch, N=11, tf=3.0, seed=2):
    """Config 4: x0 = 0, xf ~ [1.5, 1.5, pi/2] + N(0, 0.1^2), initial guess = the u = 0.1 rollout
    (closed-form RK3 of the Dubins car).  Returns Z0 (batch, NN) in Primals order, x0, xf and the options."""
    n, m = 3, 2
    rng = np.random.default_rng(seed)
    dt = tf / (N - 1)
    o = dict(N=N, iters=10, dt=dt, q_diag=1e-2, r_diag=1e-2, qf_diag=100.0, eps_p=1e-5, eps_d=1e-5, line_search=1)
    x0 = np.zeros((batch, n))
    xf = np.array([1.5, 1.5, np.pi / 2]) + 0.1 * rng.standard_normal((batch, n))
    U = np.full((batch, N - 1, m), 0.1)
    X = np.zeros((batch, N, n))
    for k in range(N - 1):
        th, v, om = X[:, k, 2], U[:, k, 0], U[:, k, 1]
        th2, th3 = th + 0.5 * dt * om, th + dt * om
        cb = (np.cos(th) + 4 * np.cos(th2) + np.cos(th3)) / 6
        sb = (np.sin(th) + 4 * np.sin(th2) + np.sin(th3)) / 6
        X[:, k + 1] = np.stack([X[:, k, 0] + dt * v * cb, X[:, k, 1] + dt * v * sb, th + dt * om], -1)
    Z = np.zeros((batch, N * n + (N - 1) * m))
    body = Z[:, :(N - 1) * (n + m)].reshape(batch, N - 1, n + m)
    body[:, :, :n], body[:, :, n:] = X[:, :-1], U
    Z[:, (N - 1) * (n + m):] = X[:, -1]
    return Z, x0, xf, o
