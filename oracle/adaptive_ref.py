"""CPU oracle for the ADAPTIVE embedded Runge–Kutta solvers: Dormand–Prince 5(4) (the reference's default) and the
other adaptive strings of webui.py:110 — bosh3, fehlberg2, adaptive_heun (TEST INFRASTRUCTURE ONLY).

The reference's default ``solver=None`` means ``torchdiffeq.odeint(..., method='dopri5', rtol=1e-5,
atol=1e-5)`` (models/flow_matching.py:54).  torchdiffeq is a third-party package, unpinned in the
reference's requirements.txt:14 and ABSENT here, so this file restates its published algorithm
(Dormand & Prince 1980 tableau with Shampine's embedded 4th-order weights and dense-output midpoint
coefficients; Hairer's initial-step heuristic; I-controller with safety 0.9, ifactor 10, dfactor 0.2;
RMS mixed error norm over ALL elements; evaluation at the requested times by a 4th-order interpolant).
PARITY UNPINNED against torchdiffeq itself: nothing in the reference pins these semantics and the package is
not installable here.  What IS pinned (tests/test_adaptive.py): the 5th-order tableau against scipy's RK45 (an
independent implementation of the same published method), the solution of a nonlinear system against
scipy's adaptive RK45 at equal tolerances, and a closed-form ODE.  The embedded error weights are Shampine's
variant as published in torchdiffeq (last coefficient -1/60), which scipy does not share.
"""
from __future__ import annotations

import torch

# Butcher tableau (7 stages, FSAL)
ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
C_ERROR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
           -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]
SAFETY, IFACTOR, DFACTOR, ORDER = 0.9, 10.0, 0.2, 5


class Tableau:
    """One embedded Runge-Kutta pair as torchdiffeq's rk_common._ButcherTableau carries it (+ c_mid and the order)."""
    def __init__(self, alpha, beta, c_sol, c_error, c_mid, order):
        self.alpha, self.beta, self.c_sol, self.c_error, self.c_mid, self.order = alpha, beta, c_sol, c_error, c_mid, order
        self.stages = len(alpha)                 # new derivative evaluations per step
        # "this property (true for Dormand-Prince) lets us save a few FLOPs": the last stage input is the solution
        self.sol_is_last_stage = c_sol[-1] == 0 and list(c_sol[:-1]) == list(beta[-1])


TABLEAUS = {
    # torchdiffeq dopri5.py (Dormand & Prince 1980, Shampine's embedded weights)
    "dopri5": Tableau(ALPHA, BETA, C_SOL, C_ERROR, C_MID, 5),
    # torchdiffeq bosh3.py (Bogacki & Shampine 1989)
    "bosh3": Tableau([1 / 2, 3 / 4, 1.], [[1 / 2], [0., 3 / 4], [2 / 9, 1 / 3, 4 / 9]], [2 / 9, 1 / 3, 4 / 9, 0.],
                     [2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8], [0., 0.5, 0., 0.], 3),
    # torchdiffeq fehlberg2.py (Fehlberg 2(1))
    "fehlberg2": Tableau([1 / 2, 1.0], [[1 / 2], [1 / 256, 255 / 256]], [1 / 512, 255 / 256, 1 / 512],
                         [-1 / 512, 0, 1 / 512], [0., 0.5, 0.], 2),
    # torchdiffeq adaptive_heun.py (Heun-Euler 2(1))
    "adaptive_heun": Tableau([1.], [[1.]], [0.5, 0.5], [0.5, -0.5], [0.5, 0.], 2),
}


def rms_norm(x: torch.Tensor) -> float:
    return float(x.double().pow(2).mean().sqrt())


def select_initial_step(f, t0: float, y0, f0, rtol, atol, order: int = ORDER) -> float:
    """torchdiffeq misc._select_initial_step, called with ``order - 1`` by the adaptive solvers: exponent 1/order."""
    scale = atol + y0.abs() * rtol
    d0, d1 = rms_norm(y0 / scale), rms_norm(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = f(torch.tensor(t0 + h0, dtype=torch.float32), y1)
    d2 = rms_norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / order)
    return min(100 * h0, h1)


def odeint_adaptive(f, y0: torch.Tensor, method: str = "dopri5", t_end: float = 1.0, rtol: float = 1e-5, atol: float = 1e-5,
                    max_steps: int = 100000):
    """Integrates dy/dt = f(t, y) from 0 and returns (y(t_end), stats) with the embedded pair ``method``.  Time is
    carried in float64 (as torchdiffeq does), the state in float32; f receives t as a 0-dim float32 tensor.
    Restates rk_common.RKAdaptiveStepsizeODESolver: _runge_kutta_step (stage k_{i+1} at t1 exactly when alpha_i == 1; the
    LAST stage derivative becomes the next step's f0 for every tableau), _compute_error_ratio (RMS mixed norm),
    _optimal_step_size (safety 0.9, ifactor 10, dfactor 0.2, exponent 1/order), _interp_fit / _interp_evaluate."""
    tb = TABLEAUS[method]
    t0 = 0.0
    f0 = f(torch.tensor(t0, dtype=torch.float32), y0)
    dt = select_initial_step(f, t0, y0, f0, rtol, atol, tb.order)
    n_acc = n_rej = 0
    y = y0
    interp = None
    while True:
        if n_acc + n_rej >= max_steps:
            raise RuntimeError("max_steps exceeded")
        t1 = t0 + dt
        k = [f0]
        for i in range(tb.stages):
            yi = y
            for j, b in enumerate(tb.beta[i]):
                if b != 0:
                    yi = yi + (dt * b) * k[j]
            ti = t1 if tb.alpha[i] == 1.0 else t0 + tb.alpha[i] * dt
            k.append(f(torch.tensor(ti, dtype=torch.float32), yi))
        if tb.sol_is_last_stage:
            y1 = yi                                           # last stage input is the solution
        else:
            y1 = y + sum((dt * c) * kk for c, kk in zip(tb.c_sol, k) if c != 0)
        err = sum((dt * c) * kk for c, kk in zip(tb.c_error, k) if c != 0)
        tol = atol + rtol * torch.max(y.abs(), y1.abs())
        ratio = rms_norm(err / tol)
        accept = ratio <= 1.0
        if ratio == 0:
            factor = IFACTOR
        else:
            dfac = 1.0 if ratio < 1 else DFACTOR
            factor = min(IFACTOR, max(SAFETY / ratio ** (1.0 / tb.order), dfac))
        if accept:
            n_acc += 1
            y_mid = y + sum((dt * c) * kk for c, kk in zip(tb.c_mid, k) if c != 0)
            interp = (y, y1, y_mid, k[0], k[-1], dt, t0, t1)
            y, f0, t0 = y1, k[-1], t1
        else:
            n_rej += 1
        dt = dt * factor
        if accept and t0 >= t_end:
            break
    ya, yb, ym, fa, fb, h, ta, tb_ = interp
    a = 2 * h * (fb - fa) - 8 * (yb + ya) + 16 * ym
    b = h * (5 * fa - 3 * fb) + 18 * ya + 14 * yb - 32 * ym
    c = h * (fb - 4 * fa) - 11 * ya - 5 * yb + 16 * ym
    d = h * fa
    x = (t_end - ta) / (tb_ - ta)
    out = ya + x * (d + x * (c + x * (b + x * a)))
    return out, dict(accepted=n_acc, rejected=n_rej, nfe=2 + tb.stages * (n_acc + n_rej))


def odeint_dopri5(f, y0: torch.Tensor, t_end: float = 1.0, rtol: float = 1e-5, atol: float = 1e-5, max_steps: int = 1000):
    return odeint_adaptive(f, y0, "dopri5", t_end, rtol, atol, max_steps)
