"""Flow-matching / Schroedinger-bridge paths and samplers behind the reference's Bridge API
(fdbm/bridge.py:14-113, 187-385).

Host side only computes the scalar schedule (a handful of fp32 torch ops on [1]-sized CPU tensors,
in the reference's own op order so the coefficient table is bit-identical to bridge.py:308-337,
373-385); every tensor-sized operation -- prior sample, per-step update, the backbone -- runs in
libfdbm_b200 kernels.  With a fdbm_b200 backbone the whole N-step loop is one CUDA graph
(`fdbm_sampler_run`); with any other callable `model(xt, y, t)` the loop stays in Python and only
the update `x <- wx*x + ws*D + w3*{y|z}` is the fused kernel.

Out of scope: `ode_int` (host-driven scipy RK45) and `pc` samplers.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import FDBM_STEP, check, current_stream, ptr
from .registry import BridgeRegistry


class Bridge:
    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--N", type=int, default=5, help="The number of steps during sampling. 5 by default.")
        parser.add_argument("--T", type=float, default=1.0, help="The total time duration of the path. 1.0 by default.")
        parser.add_argument("--sampler_type", type=str, default="ode_ei", choices=["ode_ei", "sde_ei"],
                            help="The sampler type to use. 'ode_ei' by default.")
        parser.add_argument("--sampling_eps", type=float, default=1e-4, help="The minimum process time for sampling.")
        return parser

    def __init__(self, path, N=5, T=1.0, sampler_type="ode_ei", sampling_eps=1e-4, noise="torch", seed=0, **kwargs):
        self.path = BridgeRegistry.get_by_name(path)(T=T, **kwargs)
        self.N = N
        self.T = T
        self.sampler_type = sampler_type
        self.noise = noise            # "torch": z from torch.randn_like (reference RNG stream); "philox": in-kernel
        self.seed = seed
        self._calls = 0
        if self.path.sampling_direction == "forward":
            self.start_time, self.end_time = sampling_eps, self.path.T
        else:
            self.start_time, self.end_time = self.path.T, sampling_eps

    # ---- path maths (tiny, host) ---------------------------------------------------------------
    def _std(self, t):
        return self.path.sigma_t(t)

    def probability_path(self, s, y, t):
        """bridge.py:40-43 (training-time x_t mean/std; tensor arithmetic stays in torch)."""
        a_t, b_t, sigma_t = self.path.path_param(t)
        return a_t[:, None, None, None] * s + b_t[:, None, None, None] * y, sigma_t

    def score_fn(self, t, x, s, y):
        mean, sigma = self.probability_path(s, y, t)
        return -(x - mean) / (sigma[:, None, None, None] ** 2 + 1e-8)

    def time_grid(self) -> torch.Tensor:
        """bridge.py:70."""
        return torch.linspace(self.start_time, self.end_time, self.N + 1)

    def coefficient_table(self, sampler_type=None) -> torch.Tensor:
        """CPU fp32 [N,3]: (w_x, w_s, w_y) per step for ode_ei, (w_x, w_s, w_z) for sde_ei with the
        last w_z forced to 0 (bridge.py:105-106)."""
        st = sampler_type or self.sampler_type
        ts = self.time_grid()
        rows = []
        t_prev = ts[0] * torch.ones(1)
        for t in ts[1:]:
            t_cur = t * torch.ones(1)
            if st == "ode_ei":
                w = self.path.sampling_param_ode_ei(t_cur, t_prev, 1, "cpu")
            elif st == "sde_ei":
                w = list(self.path.sampling_param_sde_ei(t_cur, t_prev, 1, "cpu"))
                if t == ts[-1]:
                    w[2] = torch.zeros_like(w[2])
            else:
                raise NotImplementedError(f"sampler_type '{st}' is outside the accelerated path (ode_ei, sde_ei)")
            rows.append(torch.stack([w[0][0], w[1][0], w[2][0]]))
            t_prev = t_cur
        return torch.stack(rows).float().contiguous()

    # ---- tensor-sized work (kernels) -----------------------------------------------------------
    def _next_offset(self) -> int:
        self._calls += 1
        return self._calls << 20

    def prior_sampling(self, y: torch.Tensor) -> torch.Tensor:
        """bridge.py:45-49: x_start = b(t0) y + sigma(t0) z."""
        _, b, sig = self.path.path_param(self.start_time * torch.ones(1))
        b, sig = float(b[0]), float(sig[0])
        y = y.contiguous()
        x = torch.empty_like(y)
        # the reference draws z even when sigma(t0) = 0 (SB); keep torch's RNG stream in step with it
        z = torch.randn_like(y) if self.noise == "torch" else None
        check(_lib.load().fdbm_prior_sample(ptr(y), ptr(z), b, sig, self.seed, self._next_offset(), y.numel(), ptr(x),
                                            current_stream()), "fdbm_prior_sample")
        return x

    def sampler(self, model, y, **kwargs):
        if self.sampler_type == "ode_ei":
            return self.ode_sampler_ei(model, y, **kwargs)
        if self.sampler_type == "sde_ei":
            return self.sde_sampler_ei(model, y, **kwargs)
        raise NotImplementedError(f"sampler_type '{self.sampler_type}' is outside the accelerated path (ode_ei, sde_ei)")

    def ode_sampler_ei(self, model, y, **kwargs):
        """bridge.py:66-87."""
        return self._run(model, y, "ode_ei")

    def sde_sampler_ei(self, model, y, **kwargs):
        """bridge.py:89-113."""
        return self._run(model, y, "sde_ei")

    def _run(self, model, y, st):
        with torch.no_grad():
            y = y.contiguous()
            table = self.coefficient_table(st)                                     # host, fp32 [N,3]
            times = self.time_grid()[:-1].float().contiguous()                     # host: t_prev of every step
            xt = self.prior_sampling(y)
            kind = FDBM_STEP[st]
            noise = None
            if st == "sde_ei" and self.noise == "torch":
                noise = torch.stack([torch.randn_like(xt) for _ in range(self.N)])
            net = getattr(model, "dnn", model)
            if hasattr(net, "run_sampler"):                                         # fused, graph-captured loop
                net.run_sampler(y, xt, times, table, kind, noise, self.seed + self._next_offset())
                return xt
            lib = _lib.load()
            B = xt.shape[0]
            table, times = table.to(y.device), times.to(y.device)
            for i in range(self.N):
                est = model(xt, y, times[i] * torch.ones(B, device=y.device)).contiguous()
                third = y if st == "ode_ei" else (noise[i] if noise is not None else None)
                check(lib.fdbm_bridge_step(ptr(xt), ptr(est), ptr(third), ptr(table[i]), kind, self.seed + self._calls,
                                           i + 1, xt.numel(), current_stream()), "fdbm_bridge_step")
            return xt


class ProbabilityPath:
    def __init__(self, T=1.0):
        self.T = T


@BridgeRegistry.register("sb")
class ProbabilityPathSB(ProbabilityPath):
    """bridge.py:187-337."""

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--noise_schedule", type=str, default="bb", choices=["gmax", "vp", "ve", "bb"])
        parser.add_argument("--k", type=float, default=2.6)
        parser.add_argument("--c", type=float, default=0.4)
        parser.add_argument("--beta_0", type=float, default=0.01)
        parser.add_argument("--beta_1", type=float, default=20.0)
        parser.add_argument("--rho", type=float, default=1.0)
        parser.add_argument("--diffusion_coeff_mode", type=str, default="g", choices=["g", "ode"])
        return parser

    def __init__(self, noise_schedule="bb", k=2.6, c=0.4, beta_0=0.01, beta_1=20.0, rho=1.0, N=5, eps=1e-8,
                 **ignored_kwargs):
        super().__init__()                 # T is deliberately not forwarded, as in bridge.py:201
        self.noise_schedule, self.k, self.c = noise_schedule, k, c
        self.beta_0, self.beta_1, self.rho, self.N, self.eps = beta_0, beta_1, rho, N, eps
        self.sampling_direction = "reverse"

    def _rhos_alphas(self, t):
        """bridge.py:213-238."""
        one = torch.ones_like(t)
        b0, db, T = self.beta_0, self.beta_1 - self.beta_0, self.T
        if self.noise_schedule == "gmax":
            alpha_t, alpha_T = one, one
            rho_t = torch.sqrt(b0 * t + 0.5 * db * (t ** 2))
            rho_T = torch.sqrt(torch.tensor(b0 * T + 0.5 * db * (T ** 2)))
        elif self.noise_schedule == "vp":
            alpha_t = torch.exp(-0.5 * (b0 * t + 0.5 * db * (t ** 2)))
            alpha_T = torch.exp(-0.5 * torch.tensor(b0 * T + 0.5 * db * (T ** 2)))
            rho_t = torch.sqrt(self.c * (torch.exp(b0 * t + 0.5 * db * (t ** 2)) - 1))
            rho_T = torch.sqrt((self.c * (torch.exp(torch.tensor(b0 * T + 0.5 * db * (T ** 2))) - 1)))
        elif self.noise_schedule == "ve":
            alpha_t, alpha_T = one, one
            log_k2 = 2 * torch.log(torch.tensor(self.k))
            rho_t = torch.sqrt((self.c * (self.k ** (2 * t) - 1.0)) / log_k2)
            rho_T = torch.sqrt((self.c * (self.k ** (2 * T) - 1.0)) / log_k2)
        elif self.noise_schedule == "bb":
            alpha_t, alpha_T = one, one
            rho_t = torch.sqrt(torch.tensor(1) * t) * self.rho
            rho_T = one * self.rho
        else:
            raise ValueError(self.noise_schedule)
        alpha_bar_t = alpha_t / (alpha_T + self.eps)
        rho_bar_t = torch.sqrt(rho_T ** 2 - rho_t ** 2 + self.eps)
        return rho_t, rho_T, rho_bar_t, alpha_t, alpha_T, alpha_bar_t

    def sigma_t(self, t):
        rho_t, rho_T, rho_bar_t, alpha_t, _, _ = self._rhos_alphas(t)
        sigma = (alpha_t * rho_bar_t * rho_t) / (rho_T + self.eps)
        return torch.where(t == 1.0, torch.zeros_like(sigma), sigma)

    def path_param(self, t):
        """bridge.py:270-281."""
        rho_t, rho_T, rho_bar_t, alpha_t, _, alpha_bar_t = self._rhos_alphas(t)
        denom = rho_T ** 2 + self.eps
        a_t = alpha_t * rho_bar_t ** 2 / denom
        b_t = alpha_bar_t * rho_t ** 2 / denom
        sigma = (alpha_t * rho_bar_t * rho_t) / (rho_T + self.eps)
        at_T = (t == 1.0)
        return (torch.where(at_T, torch.zeros_like(a_t), a_t), torch.where(at_T, torch.ones_like(b_t), b_t),
                torch.where(at_T, torch.zeros_like(sigma), sigma))

    def sampling_param_ode_ei(self, t_curr, t_prev, batch_size, device):
        """bridge.py:308-324."""
        ones = torch.ones(batch_size, device=device)
        rho_p, rho_T, rbar_p, al_p, al_T, _ = self._rhos_alphas(t_prev * ones)
        rho_c, rho_T, rbar_c, al_c, al_T, _ = self._rhos_alphas(t_curr * ones)
        w_x = al_c * rho_c * rbar_c / (al_p * rho_p * rbar_p + self.eps)
        w_s = al_c / (rho_T ** 2 + self.eps) * (rbar_c ** 2 - rbar_p * rho_c * rbar_c / (rho_p + self.eps))
        w_y = al_c / (al_T * rho_T ** 2 + self.eps) * (rho_c ** 2 - rho_p * rho_c * rbar_c / (rbar_p + self.eps))
        return w_x, w_s, w_y

    def sampling_param_sde_ei(self, t_curr, t_prev, batch_size, device):
        """bridge.py:326-337."""
        ones = torch.ones(batch_size, device=device)
        rho_p, _, _, al_p, _, _ = self._rhos_alphas(t_prev * ones)
        rho_c, _, _, al_c, _, _ = self._rhos_alphas(t_curr * ones)
        w_x = al_c * rho_c ** 2 / (al_p * rho_p ** 2 + self.eps)
        tmp = 1 - rho_c ** 2 / (rho_p ** 2 + self.eps)
        return w_x, al_c * tmp, al_c * rho_c * torch.sqrt(tmp)


@BridgeRegistry.register("fm")
class ProbabilityPathFM(ProbabilityPath):
    """bridge.py:340-385."""

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--sigma_max", type=float, default=1.0)
        parser.add_argument("--sigma_min", type=float, default=0.01)
        parser.add_argument("--noise_schedule", type=str, default="ot")
        return parser

    def __init__(self, sigma_max=1.0, sigma_min=0.01, noise_schedule="ot", eps=1e-8, **ignored_kwargs):
        super().__init__()
        self.sigma_max, self.sigma_min, self.noise_schedule, self.eps = sigma_max, sigma_min, noise_schedule, eps
        self.sampling_direction = "forward"

    def sigma_t(self, t):
        return t * self.sigma_min + (1 - t) * self.sigma_max

    def path_param(self, t):
        return t, 1 - t, self.sigma_t(t)

    def sampling_param_ode_ei(self, t_curr, t_prev, batch_size, device):
        """bridge.py:373-385 (Euler step of OT-CFM)."""
        ones = torch.ones(batch_size, device=device)
        tp, tc = t_prev * ones, t_curr * ones
        dt = tc - tp
        sig_c, sig_p = self.sigma_t(tc), self.sigma_t(tp)
        return sig_c / (sig_p + self.eps), self.sigma_max * dt / (sig_p + self.eps), -self.sigma_min * dt / (sig_p + self.eps)
