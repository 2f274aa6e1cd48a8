"""Flow-matching / Schroedinger-bridge paths and samplers behind the reference's Bridge API
(fdbm/bridge.py:14-113, 187-385).

Host side only computes the scalar schedule (a handful of fp32 torch ops on [1]-sized CPU tensors,
in the reference's own op order so the coefficient table is bit-identical to bridge.py:308-337,
373-385); every tensor-sized operation -- prior sample, per-step update, the backbone -- runs in
libfdbm_b200 kernels.  With a fdbm_b200 backbone the whole N-step loop is one CUDA graph
(`fdbm_sampler_run`); with any other callable `model(xt, y, t)` the loop stays in Python and only
the update `x <- wx*x + ws*D + w3*{y|z}` is the fused kernel.

`pc` (bridge.py:142-166): Euler-Maruyama predictor + Langevin / annealed-Langevin corrector; all utterances of a batch
share the step's time, so every update is  x_mean = c0 x + c1 D + c2 y,  x = x_mean + c3 z  with scalar weights --
one fused kernel per update (`fdbm_bridge_update4`), the Langevin corrector's data-dependent step size being reduced on
the device (`fdbm_langevin_coef`).  The reference multiplies `[B]` weight vectors straight into `[B,1,F,T]` tensors in
`path.ode` / `path.sde` (bridge.py:283-306), which only broadcasts correctly for B = 1; here the weights are scalars, so
any batch size works and B = 1 reproduces the reference.
`ode_int` (bridge.py:115-140): the reference hands the flattened state to scipy's RK45 on the host; here the same
Dormand-Prince 5(4) scheme with scipy's step-size controller runs on device tensors (`fdbm_lincomb`, `fdbm_rk_error_norm`),
one scalar read-back per step.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import FDBM_STEP, check, current_stream, on_device, ptr
from .registry import BridgeRegistry


class Bridge:
    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--N", type=int, default=5, help="The number of steps during sampling. 5 by default.")
        parser.add_argument("--T", type=float, default=1.0, help="The total time duration of the path. 1.0 by default.")
        parser.add_argument("--sampler_type", type=str, default="ode_ei", choices=["ode_ei", "sde_ei", "ode_int", "pc"],
                            help="The sampler type to use. 'ode_ei' by default.")
        parser.add_argument("--sampling_eps", type=float, default=1e-4, help="The minimum process time for sampling.")
        return parser

    def __init__(self, path, N=5, T=1.0, sampler_type="ode_ei", sampling_eps=1e-4, noise="torch", seed=0,
                 match_torch_rng=False, **kwargs):
        self.path = BridgeRegistry.get_by_name(path)(T=T, **kwargs)
        self.N = N
        self.T = T
        self.sampler_type = sampler_type
        self.noise = noise            # "torch": z from torch.randn_like (reference RNG stream); "philox": in-kernel
        self.seed = seed
        # The reference draws z in prior_sampling even when sigma(t0) = 0 (SB: x_start = y) and discards it.  That draw
        # is a full-size HBM write per call (135 MB per 256 utterances) whose only effect is on torch's global RNG
        # stream; it is made only on request (tests that replay the reference's RNG sequence set this).
        self.match_torch_rng = match_torch_rng
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

    @on_device
    def prior_sampling(self, y: torch.Tensor) -> torch.Tensor:
        """bridge.py:45-49: x_start = b(t0) y + sigma(t0) z."""
        _, b, sig = self.path.path_param(self.start_time * torch.ones(1))
        b, sig = float(b[0]), float(sig[0])
        y = y.contiguous()
        x = torch.empty_like(y)
        z = torch.randn_like(y) if self.noise == "torch" and (sig != 0.0 or self.match_torch_rng) else None
        check(_lib.load().fdbm_prior_sample(ptr(y), ptr(z), b, sig, self.seed, self._next_offset(), y.numel(), ptr(x),
                                            current_stream()), "fdbm_prior_sample")
        return x

    def sampler(self, model, y, **kwargs):
        if self.sampler_type == "ode_ei":
            return self.ode_sampler_ei(model, y, **kwargs)
        if self.sampler_type == "sde_ei":
            return self.sde_sampler_ei(model, y, **kwargs)
        if self.sampler_type == "ode_int":
            return self.ode_sampler_int(model, y, **kwargs)
        if self.sampler_type == "pc":
            return self.pc_sampler(model, y, **kwargs)
        return None                                   # bridge.py:56-64 falls through for unknown names

    def ode_sampler_ei(self, model, y, **kwargs):
        """bridge.py:66-87."""
        return self._run(model, y, "ode_ei")

    def sde_sampler_ei(self, model, y, **kwargs):
        """bridge.py:89-113."""
        return self._run(model, y, "sde_ei")

    @on_device
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


    # ---- predictor-corrector sampler (bridge.py:142-166) ---------------------------------------------------------------
    @on_device
    def pc_sampler(self, model, y, predictor_name="reverse_diffusion", corrector_name="ald", denoise=True, snr=0.5,
                   corrector_steps=1, **kwargs):
        """bridge.py:142-166 with EulerMaruyamaPredictor / NonePredictor (util/predictors.py:39-62) and LangevinCorrector /
        AnnealedLangevinDynamics / NoneCorrector (util/correctors.py:36-95).  As in the reference the default
        predictor name 'reverse_diffusion' is not registered and raises ValueError; pass 'euler_maruyama'."""
        if predictor_name not in ("euler_maruyama", "none"):
            raise ValueError(f"Predictor with name '{predictor_name}' unknown.")
        if corrector_name not in ("langevin", "ald", "none"):
            raise ValueError(f"Corrector with name '{corrector_name}' unknown.")
        if predictor_name == "euler_maruyama" and not hasattr(self.path, "sde"):
            raise AttributeError(f"{type(self.path).__name__} object has no attribute 'sde'")
        lib = _lib.load()
        with torch.no_grad():
            y = y.contiguous()
            xt = self.prior_sampling(y)
            x_mean = xt.clone()
            B, n = xt.shape[0], xt.numel()
            timesteps = torch.linspace(self.start_time, self.end_time, self.N)
            coef = torch.empty(4, device=y.device)
            scratch = torch.empty(2 * B, dtype=torch.float64, device=y.device)
            draw = (lambda: torch.randn_like(xt)) if self.noise == "torch" else (lambda: None)
            seed = self.seed + self._next_offset()
            step_id = 0

            def update(est, z):
                nonlocal step_id
                step_id += 1
                check(lib.fdbm_bridge_update4(ptr(xt), ptr(est), ptr(y), ptr(z), ptr(coef), seed, step_id, n, ptr(x_mean),
                                              current_stream()), "fdbm_bridge_update4")

            for i in range(self.N):
                t = timesteps[i]
                stepsize = t - timesteps[i + 1] if i != self.N - 1 else timesteps[-1]
                vec_t = t * torch.ones(1)
                dev_t = (t * torch.ones(B)).to(y.device)
                # ---- corrector (correctors.py:44-81)
                if corrector_name != "none":
                    a_t, b_t, sig = (float(v[0]) for v in self.path.path_param(vec_t))
                    for _ in range(corrector_steps):
                        est = model(xt, y, dev_t).contiguous()
                        z = draw()
                        if corrector_name == "ald":
                            std = float(self._std(vec_t)[0])
                            step = (snr * std) ** 2 * 2
                            k = 1.0 / (sig ** 2 + 1e-8)
                            coef.copy_(torch.tensor([1.0 - step * k, step * k * a_t, step * k * b_t, (step * 2) ** 0.5]))
                        else:
                            check(lib.fdbm_langevin_coef(ptr(xt), ptr(est), ptr(y), ptr(z), a_t, b_t, sig, float(snr), seed, step_id + 1,
                                                         B, n // B, ptr(scratch), ptr(coef), current_stream()), "fdbm_langevin_coef")
                        update(est, z)
                # ---- predictor (predictors.py:39-62)
                if predictor_name == "euler_maruyama":
                    dt = -float(stepsize)
                    z = draw()
                    est = model(xt, y, dev_t).contiguous()
                    wx, ws, wy, gd = (float(v[0]) for v in self.path.sde_weights(vec_t))
                    coef.copy_(torch.tensor([1.0 + wx * dt, ws * dt, wy * dt, gd * (-dt) ** 0.5]))
                    update(est, z)
                else:
                    x_mean.copy_(xt)                   # NonePredictor.update_fn returns (x, x): predictors.py:54-62
            return x_mean if denoise else xt

    # ---- adaptive ODE sampler (bridge.py:115-140) ----------------------------------------------------------------------
    @on_device
    def ode_sampler_int(self, model, y, rtol=1e-5, atol=1e-5, method="RK45", max_nfev=100000, **kwargs):
        """bridge.py:115-140: integrate dx/dt = path.ode(t, x, model(x, y, t), y) from start_time to end_time with the
        explicit Runge-Kutta pair scipy.integrate.solve_ivp(method='RK45') uses (Dormand-Prince 5(4), local extrapolation,
        scipy's initial-step and step-size rules), on device tensors.  The reference integrates the COMPLEX state vector:
        scipy's error norm is the RMS of |err| / (atol + rtol max(|x|, |x_new|)) over complex elements, reproduced here."""
        if method != "RK45":
            raise NotImplementedError("fdbm_b200 implements solve_ivp's default method 'RK45'")
        from .rk45 import integrate_rk45
        with torch.no_grad():
            y = y.contiguous()
            x = self.prior_sampling(y)
            B = y.shape[0]

            def flow(t, x_state):
                vec_t = float(t) * torch.ones(1)
                est = model(x_state, y, (float(t) * torch.ones(B)).to(y.device)).contiguous()
                wx, ws, wy = (float(v[0]) for v in self.path.ode_weights(vec_t))
                return est, (wx, ws, wy)
            return integrate_rk45(flow, x, y, float(self.start_time), float(self.end_time), rtol, atol, max_nfev)


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
        self.diffusion_coeff_mode = "g"        # bridge.py:211: fixed, whatever --diffusion_coeff_mode says

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

    def auxiliary_param(self, t):
        """bridge.py:240-253: drift f and diffusion g of the reference SDE."""
        if self.noise_schedule == "ve":
            return 0.0, torch.sqrt(torch.tensor(self.c)) * self.k ** t
        if self.noise_schedule == "vp":
            lin = self.beta_0 + (self.beta_1 - self.beta_0) * t
            return -0.5 * lin, torch.sqrt(torch.tensor(self.c) * lin)
        if self.noise_schedule == "gmax":
            return 0.0, torch.sqrt(torch.as_tensor(self.beta_0 + (self.beta_1 - self.beta_0) * t))
        return 0.0, self.rho * torch.ones_like(t)

    def diffusion_coeff(self, g, t):
        """bridge.py:255-259."""
        return g if self.diffusion_coeff_mode == "g" else 0.0 * torch.ones_like(g)

    def ode_weights(self, t):
        """Scalar weights (w_x, w_s, w_y) of the probability-flow ODE, bridge.py:283-292: flow = w_x x + w_s s + w_y y."""
        rho, _, rho_bar, alpha, _, alpha_bar = self._rhos_alphas(t)
        f, g = self.auxiliary_param(t)
        w_x = f + g ** 2 * (rho_bar ** 2 - rho ** 2) / (2 * alpha ** 2 * rho ** 2 * rho_bar ** 2 + self.eps)
        w_s = -g ** 2 / (2 * alpha * rho ** 2 + self.eps)
        w_y = alpha_bar * g ** 2 / (2 * alpha ** 2 * rho_bar ** 2 + self.eps)
        return w_x, w_s, w_y

    def sde_weights(self, t):
        """Scalar weights (w_x, w_s, w_y, diffusion) of the reverse SDE, bridge.py:294-306."""
        rho, _, rho_bar, alpha, _, alpha_bar = self._rhos_alphas(t)
        f, g = self.auxiliary_param(t)
        gd = self.diffusion_coeff(g, t)
        w_x = f + ((g ** 2 + gd ** 2) * rho_bar ** 2 - (g ** 2 - gd ** 2) * rho ** 2) / (2 * alpha ** 2 * rho ** 2 * rho_bar ** 2 + self.eps)
        w_s = -(g ** 2 + gd ** 2) / (2 * alpha * rho ** 2 + self.eps)
        w_y = alpha_bar * (g ** 2 - gd ** 2) / (2 * alpha ** 2 * rho_bar ** 2 + self.eps)
        return w_x, w_s, w_y, gd

    def ode(self, t, x, s, y):
        """bridge.py:283-292 for callers that keep the reference's call (torch arithmetic; the [B] weights are broadcast over
        [B,1,F,T] correctly for any B, see the module docstring)."""
        w_x, w_s, w_y = (w.reshape(-1, 1, 1, 1) if torch.is_tensor(w) and w.dim() == 1 else w for w in self.ode_weights(t))
        return w_x * x + w_s * s + w_y * y

    def sde(self, t, x, s, y):
        """bridge.py:294-306."""
        w_x, w_s, w_y, gd = self.sde_weights(t)
        b = lambda w: w.reshape(-1, 1, 1, 1) if torch.is_tensor(w) and w.dim() == 1 else w
        return b(w_x) * x + b(w_s) * s + b(w_y) * y, gd

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

    def ode_weights(self, t):
        """bridge.py:368-371 as scalar weights: flow = ((s_min - s_max) x + s_max s - s_min y) / (sigma_t + eps)."""
        den = self.sigma_t(t) + self.eps
        return (self.sigma_min - self.sigma_max) / den, self.sigma_max / den, -self.sigma_min / den

    def ode(self, t, x, s, y):
        """bridge.py:368-371."""
        sigma_t = self.sigma_t(t)[:, None, None, None]
        return ((self.sigma_min - self.sigma_max) * x + self.sigma_max * s - self.sigma_min * y) / (sigma_t + self.eps)

    def sampling_param_ode_ei(self, t_curr, t_prev, batch_size, device):
        """bridge.py:373-385 (Euler step of OT-CFM)."""
        ones = torch.ones(batch_size, device=device)
        tp, tc = t_prev * ones, t_curr * ones
        dt = tc - tp
        sig_c, sig_p = self.sigma_t(tc), self.sigma_t(tp)
        return sig_c / (sig_p + self.eps), self.sigma_max * dt / (sig_p + self.eps), -self.sigma_min * dt / (sig_p + self.eps)
