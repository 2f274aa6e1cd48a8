"""Dormand-Prince 5(4) with scipy.integrate.solve_ivp's controller, on device tensors.

The reference's `Bridge.ode_sampler_int` (fdbm/bridge.py:115-140) flattens the complex state to numpy and lets
`scipy.integrate.solve_ivp(method='RK45')` drive the backbone from the host.  Here the same scheme (scipy
`_ivp/rk.py`: tableau, FSAL, local extrapolation; `_ivp/common.py`: `select_initial_step`, RMS error norm; SAFETY 0.9,
factors 0.2 ... 10, error exponent -1/5) keeps the state and the seven stage derivatives on the GPU: stage
combinations are `fdbm_lincomb` launches, the error norm is `fdbm_rk_error_norm` + one 8-byte read-back per attempted
step (the controller is inherently sequential).  The ODE right-hand side is  f(t, x) = w_x(t) x + w_s(t) D(x, y, t) +
w_y(t) y  with scalar weights from `path.ode_weights`.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr

_C = [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0]
_A = [[], [1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9], [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
      [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]]
_B = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]
_E = [-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40]
SAFETY, MIN_FACTOR, MAX_FACTOR, ERR_EXP = 0.9, 0.2, 10.0, -1.0 / 5.0


def _lincomb(out, terms):
    """out = sum c_k * t_k over (coefficient, complex tensor) pairs; zero coefficients are dropped."""
    terms = [(c, t) for c, t in terms if c != 0.0]
    lib = _lib.load()
    srcs = (C.c_void_p * len(terms))(*[ptr(t) for _, t in terms])
    coefs = (C.c_float * len(terms))(*[float(c) for c, _ in terms])
    check(lib.fdbm_lincomb(ptr(out), srcs, coefs, len(terms), out.numel() * 2, current_stream()), "fdbm_lincomb")
    return out


class _Norm:
    def __init__(self, like):
        self.dev = torch.zeros(1, dtype=torch.float64, device=like.device)
        self.n = like.numel()

    def __call__(self, terms, y, y_new, rtol, atol) -> float:
        lib = _lib.load()
        terms = [(c, t) for c, t in terms if c != 0.0]
        srcs = (C.c_void_p * len(terms))(*[ptr(t) for _, t in terms])
        coefs = (C.c_float * len(terms))(*[float(c) for c, _ in terms])
        check(lib.fdbm_rk_error_norm(srcs, coefs, len(terms), ptr(y), ptr(y_new), float(rtol), float(atol), self.n, ptr(self.dev),
                                     current_stream()), "fdbm_rk_error_norm")
        return math.sqrt(float(self.dev.item()) / self.n)          # the controller's one read-back per attempted step


def integrate_rk45(flow, x0, y, t0, t_bound, rtol, atol, max_nfev=100000):
    """Integrate from t0 to t_bound.  `flow(t, x) -> (D, (w_x, w_s, w_y))` evaluates the backbone; returns the final state."""
    direction = 1.0 if t_bound >= t0 else -1.0
    norm = _Norm(x0)
    nfev = [0]
    stage = torch.empty_like(x0)

    def fun(t, x, out):
        nfev[0] += 1
        if nfev[0] > max_nfev:
            raise RuntimeError(f"ode_sampler_int: more than {max_nfev} backbone evaluations (stiff path? the SB path's ODE is "
                               "singular at t = T)")
        D, (wx, ws, wy) = flow(t, x)
        return _lincomb(out, [(wx, x), (ws, D), (wy, y)])

    xcur = x0.clone()
    K = [torch.empty_like(x0) for _ in range(7)]
    x_new = torch.empty_like(x0)
    fun(t0, xcur, K[0])
    # ---- select_initial_step (scipy/integrate/_ivp/common.py), order = 4
    interval = abs(t_bound - t0)
    d0 = norm([(1.0, xcur)], xcur, xcur, rtol, atol)
    d1 = norm([(1.0, K[0])], xcur, xcur, rtol, atol)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval)
    _lincomb(stage, [(1.0, xcur), (h0 * direction, K[0])])
    fun(t0 + h0 * direction, stage, K[1])
    d2 = norm([(1.0, K[1]), (-1.0, K[0])], xcur, xcur, rtol, atol) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    h_abs = min(100 * h0, h1, interval)
    t = t0
    while direction * (t - t_bound) < 0:
        min_step = 10 * abs(float(np.nextafter(t, direction * np.inf)) - t)
        h_abs = max(h_abs, min_step)
        accepted, rejected = False, False
        while not accepted:
            if h_abs < min_step:
                raise RuntimeError("ode_sampler_int: required step size is less than spacing between numbers")
            h = h_abs * direction
            t_new = t + h
            if direction * (t_new - t_bound) > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = abs(h)
            for s in range(1, 6):                                             # rk_step (scipy/integrate/_ivp/rk.py)
                _lincomb(stage, [(1.0, xcur)] + [(a * h, K[j]) for j, a in enumerate(_A[s])])
                fun(t + _C[s] * h, stage, K[s])
            _lincomb(x_new, [(1.0, xcur)] + [(b * h, K[j]) for j, b in enumerate(_B)])
            fun(t + h, x_new, K[6])
            err = norm([(e * h, K[j]) for j, e in enumerate(_E)], xcur, x_new, rtol, atol)
            if err < 1:
                factor = MAX_FACTOR if err == 0 else min(MAX_FACTOR, SAFETY * err ** ERR_EXP)
                if rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                accepted = True
            else:
                h_abs *= max(MIN_FACTOR, SAFETY * err ** ERR_EXP)
                rejected = True
        t = t_new
        xcur, x_new = x_new, xcur
        K[0], K[6] = K[6], K[0]                                               # first-same-as-last
    integrate_rk45.last_nfev = nfev[0]
    return xcur
