"""ORACLE (test infrastructure, not product code): CPU restatement of the
reference sampler ``dpm_solver_pp.py`` in the configuration the live joint
path uses (``train_t2i_discrete.py:480-546``): discrete VP schedule, data
prediction (``predict_x0=True``), ``solver_type='dpm_solver'``,
``method='fast'`` singlestep orders 3/2/1 with the mask stream co-evolved
(``enable_mask_opt=True``), plus the stand-alone 2M/3M multistep updates.

Everything is float32 torch-on-CPU arithmetic in the reference's own operand
order, so on the CPU it reproduces the reference bit for bit (checked in
``tests/test_oracle_golden.py``).  Parity status: PINNED for the singlestep
path and the 2M/3M image updates; the 2M/3M *mask* stream has no reference
behaviour and is therefore "parity unpinned".

Reference lines followed (``/root/reference``):
  stable_diffusion_beta_schedule   train_t2i_discrete.py:40-44
  NoiseScheduleVP (discrete)       dpm_solver_pp.py:99-107, 121-169
  interpolate_fn                   dpm_solver_pp.py:9-52
  get_time_steps('time_uniform')   dpm_solver_pp.py:355-356
  orders for method='fast'         dpm_solver_pp.py:386-395
  model_fn (x0 conversion)         dpm_solver_pp.py:310-326
  1S / 2S / 3S updates             dpm_solver_pp.py:432-457, 524-557, 713-766
  2M / 3M updates                  dpm_solver_pp.py:602-677
  sample(method='fast')            dpm_solver_pp.py:1018-1044
  sample(method='singlestep'), two phases / ground truth   dpm_solver_pp.py:1045-1078; libs/uvit_t2i.py:486-496
  cfg_nnet                         train_t2i_discrete.py:387-439
  int2bits / bits2int              utils.py:475-518
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

from . import uvit_oracle

F32 = torch.float32


def sd_betas(n: int = 1000, lo: float = 0.00085, hi: float = 0.0120) -> torch.Tensor:
    b = torch.linspace(lo ** 0.5, hi ** 0.5, n, dtype=torch.float64) ** 2
    return b.float()


def _pwl(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear interpolation with linear extrapolation; 1-element x.
    Same segment choice and the same final formula as interpolate_fn (:9-52);
    the position of x among the knots is found by counting instead of sorting
    (identical for a stable sort: x sits before equal knots)."""
    K = xp.numel()
    idx = int((xp < x).sum())
    if idx == 0:
        lo = 0
    elif idx == K:
        lo = K - 2
    else:
        lo = idx - 1
    sx, ex, sy, ey = xp[lo], xp[lo + 1], yp[lo], yp[lo + 1]
    return sy + (x - sx) * (ey - sy) / (ex - sx)


class Schedule:
    def __init__(self, betas: Optional[torch.Tensor] = None):
        betas = sd_betas() if betas is None else betas.float()
        self.log_alpha = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        self.N = self.log_alpha.numel()
        self.t = torch.linspace(1.0 / self.N, 1.0, self.N)
        self._la_flip = torch.flip(self.log_alpha, [0])
        self._t_flip = torch.flip(self.t, [0])

    def log_mean(self, t):
        return _pwl(t, self.t, self.log_alpha)

    def alpha(self, t):
        return torch.exp(self.log_mean(t))

    def sigma(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_mean(t)))

    def lam(self, t):
        lm = self.log_mean(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))

    def inv_lam(self, lamb):
        la = -0.5 * torch.logaddexp(torch.zeros((), dtype=F32), -2.0 * lamb)
        return _pwl(la, self._la_flip, self._t_flip)


def fast_orders(steps: int, order: int = 3) -> List[int]:
    if order == 3:
        K = steps // 3 + 1
        if steps % 3 == 0:
            return [3] * (K - 2) + [2, 1]
        if steps % 3 == 1:
            return [3] * (K - 1) + [1]
        return [3] * (K - 1) + [2]
    if order == 2:
        K = steps // 2
        return [2] * K if steps % 2 == 0 else [2] * K + [1]
    raise ValueError("order must >= 2")


ModelFn = Callable[[torch.Tensor, torch.Tensor, Optional[torch.Tensor]], Tuple[torch.Tensor, Optional[torch.Tensor]]]


class Solver:
    """model(x, t_continuous (0-dim f32), mask) -> (eps, pred_mask)."""

    def __init__(self, model: ModelFn, sched: Schedule, trace: Optional[list] = None):
        self.model, self.ns, self.trace = model, sched, trace
        self.gt = False        # use_ground_truth handed to the model (dpm_solver_pp.py:310-314)
        self.mask_opt = True   # enable_mask_opt: co-evolve the mask (False: pass it through, :459-461, :592-599, :823-829)

    def _x0(self, x, t, mask):
        a, s = self.ns.alpha(t), self.ns.sigma(t)
        eps, pm = self.model(x, t, mask, gt=True) if self.gt else self.model(x, t, mask)
        x0 = (x - s * eps) / a
        if self.trace is not None:
            self.trace.append(dict(t=float(t), x_in=x.clone(), m_in=None if mask is None else mask.clone(),
                                   eps=eps.clone(), pm=None if pm is None else pm.clone()))
        return x0, pm

    def first(self, x, s, t, mask):
        ns = self.ns
        h = ns.lam(t) - ns.lam(s)
        sig_s, sig_t, a_t = ns.sigma(s), ns.sigma(t), torch.exp(ns.log_mean(t))
        phi_1 = (torch.exp(-h) - 1.0) / (-1.0)
        X, P = self._x0(x, s, mask)
        x_t = (sig_t / sig_s) * x + (a_t * phi_1) * X
        if not self.mask_opt:
            return x_t, P, P
        m_t = None if mask is None else (sig_t / sig_s) * mask + (a_t * phi_1) * P
        return x_t, P, m_t

    def second(self, x, s, t, r1, mask):
        ns = self.ns
        lam_s, lam_t = ns.lam(s), ns.lam(t)
        h = lam_t - lam_s
        s1 = ns.inv_lam(lam_s + r1 * h)
        sig_s, sig_s1, sig_t = ns.sigma(s), ns.sigma(s1), ns.sigma(t)
        a_s1, a_t = torch.exp(ns.log_mean(s1)), torch.exp(ns.log_mean(t))
        phi_11 = torch.expm1(-r1 * h)
        phi_1 = torch.expm1(-h)
        X, P = self._x0(x, s, mask)
        x_s1 = (sig_s1 / sig_s) * x - (a_s1 * phi_11) * X
        m_s1 = None if mask is None else (sig_s1 / sig_s) * mask + (a_s1 * phi_11) * P  # sign quirk (:536-539)
        if not self.mask_opt:
            m_s1 = mask
        X1, P1 = self._x0(x_s1, s1, m_s1)
        x_t = (sig_t / sig_s) * x - (a_t * phi_1) * X - (0.5 / r1) * (a_t * phi_1) * (X1 - X)
        if not self.mask_opt:
            return x_t, P, P
        m_t = None
        if mask is not None:
            m_t = (sig_t / sig_s) * mask - (a_t * phi_1) * P - (0.5 / r1) * (a_t * phi_1) * (P1 - P)
        return x_t, P, m_t

    def third(self, x, s, t, r1, r2, mask):
        ns = self.ns
        lam_s, lam_t = ns.lam(s), ns.lam(t)
        h = lam_t - lam_s
        s1 = ns.inv_lam(lam_s + r1 * h)
        s2 = ns.inv_lam(lam_s + r2 * h)
        sig_s, sig_s1, sig_s2, sig_t = ns.sigma(s), ns.sigma(s1), ns.sigma(s2), ns.sigma(t)
        a_s1, a_s2, a_t = torch.exp(ns.log_mean(s1)), torch.exp(ns.log_mean(s2)), torch.exp(ns.log_mean(t))
        phi_11 = torch.expm1(-r1 * h)
        phi_12 = torch.expm1(-r2 * h)
        phi_1 = torch.expm1(-h)
        phi_22 = torch.expm1(-r2 * h) / (r2 * h) + 1.0
        phi_2 = phi_1 / h + 1.0
        X, P = self._x0(x, s, mask)
        x_s1 = (sig_s1 / sig_s) * x - (a_s1 * phi_11) * X
        m_s1 = None if mask is None else (sig_s1 / sig_s) * mask + (a_s1 * phi_11) * P  # sign quirk (:730-733)
        if not self.mask_opt:
            m_s1 = mask
        X1, P1 = self._x0(x_s1, s1, m_s1)
        x_s2 = (sig_s2 / sig_s) * x - (a_s2 * phi_12) * X + r2 / r1 * (a_s2 * phi_22) * (X1 - X)
        m_s2 = None
        if mask is not None:
            m_s2 = (sig_s2 / sig_s) * mask - (a_s2 * phi_12) * P + r2 / r1 * (a_s2 * phi_22) * (P1 - P)
        if not self.mask_opt:
            m_s2 = mask
        X2, P2 = self._x0(x_s2, s2, m_s2)
        x_t = (sig_t / sig_s) * x - (a_t * phi_1) * X + (1.0 / r2) * (a_t * phi_2) * (X2 - X)
        if not self.mask_opt:
            return x_t, P, P
        m_t = None
        if mask is not None:
            m_t = (sig_t / sig_s) * mask - (a_t * phi_1) * P + (1.0 / r2) * (a_t * phi_2) * (P2 - P)
        return x_t, P, m_t

    def sample_fast(self, x, mask, steps, order=3, eps=1e-3, T=1.0):
        ns = self.ns
        orders = fast_orders(steps, order)
        ts = torch.linspace(T, eps, steps + 1)
        i = 0
        pred_mask, mask_t = mask, mask
        for o in orders:
            s, t = ts[i], ts[i + o]
            h = ns.lam(ts[i + o]) - ns.lam(ts[i])
            r1 = None if o <= 1 else (ns.lam(ts[i + 1]) - ns.lam(ts[i])) / h
            r2 = None if o <= 2 else (ns.lam(ts[i + 2]) - ns.lam(ts[i])) / h
            if o == 1:
                x, pred_mask, mask_t = self.first(x, s, t, mask_t)
            elif o == 2:
                x, pred_mask, mask_t = self.second(x, s, t, r1, mask_t)
            else:
                x, pred_mask, mask_t = self.third(x, s, t, r1, r2, mask_t)
            i += o
        return x, pred_mask

    def sample_singlestep(self, x, mask, steps, order=3, eps=1e-3, T=1.0, two_phases=False):
        """method='singlestep' (dpm_solver_pp.py:1045-1078): steps // order updates of one fixed order with the default
        r1 / r2 on a uniform time grid; with ``two_phases`` the same grid is walked a second time from the phase-one image with
        the phase-one mask held fixed and handed to the network as ground truth (enable_mask_opt=False,
        use_ground_truth=True); the returned pred_mask is phase one's."""
        n = steps // order
        ts = torch.linspace(T, eps, n + 1)

        def walk(x, mask_t):
            pred_mask = mask_t
            for i in range(n):
                s, t = ts[i], ts[i + 1]
                if order == 1:
                    x, pred_mask, mask_t = self.first(x, s, t, mask_t)
                elif order == 2:
                    x, pred_mask, mask_t = self.second(x, s, t, 0.5, mask_t)
                else:
                    x, pred_mask, mask_t = self.third(x, s, t, 1.0 / 3.0, 2.0 / 3.0, mask_t)
            return x, pred_mask, mask_t

        x, pred_mask, mask_t = walk(x, mask)
        if two_phases:
            keep = (self.gt, self.mask_opt)
            self.gt, self.mask_opt = True, False
            x, _, _ = walk(x, mask_t)
            self.gt, self.mask_opt = keep
        return x, pred_mask

    # --- multistep pure updates (dpm_solver_pp.py:602-677), data prediction, 'dpm_solver' ---
    def multistep_second(self, x, X_list, t_list, t):
        ns = self.ns
        X1, X0 = X_list
        t1, t0 = t_list
        lam1, lam0, lam_t = ns.lam(t1), ns.lam(t0), ns.lam(t)
        sig0, sig_t, a_t = ns.sigma(t0), ns.sigma(t), torch.exp(ns.log_mean(t))
        h_0 = lam0 - lam1
        h = lam_t - lam0
        r0 = h_0 / h
        D1_0 = (1.0 / r0) * (X0 - X1)
        return (sig_t / sig0) * x - (a_t * (torch.exp(-h) - 1.0)) * X0 - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * D1_0

    def multistep_third(self, x, X_list, t_list, t):
        ns = self.ns
        X2, X1, X0 = X_list
        t2, t1, t0 = t_list
        lam2, lam1, lam0, lam_t = ns.lam(t2), ns.lam(t1), ns.lam(t0), ns.lam(t)
        sig0, sig_t, a_t = ns.sigma(t0), ns.sigma(t), torch.exp(ns.log_mean(t))
        h_1 = lam1 - lam2
        h_0 = lam0 - lam1
        h = lam_t - lam0
        r0, r1 = h_0 / h, h_1 / h
        D1_0 = (1.0 / r0) * (X0 - X1)
        D1_1 = (1.0 / r1) * (X1 - X2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
        return ((sig_t / sig0) * x - (a_t * (torch.exp(-h) - 1.0)) * X0
                + (a_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1
                - (a_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5)) * D2)


def cfg_model(sd, cfg, context, empty_context, scale, dtype=F32, n_time=1000) -> ModelFn:
    """train_t2i_discrete.py:387-439 + :506-516: two forwards (cond / empty context),
    CFG on both eps and the mask prediction, model time = 1000 * t."""

    def fn(x, t_cont, mask, gt=False):
        B = x.shape[0]
        t = (torch.ones(B, device=x.device) * t_cont) * n_time
        ec = empty_context.unsqueeze(0).expand(B, -1, -1)
        if mask is None:
            c = uvit_oracle.uvit_forward(sd, cfg, x, t, context, None, dtype).float()
            u = uvit_oracle.uvit_forward(sd, cfg, x, t, ec, None, dtype).float()
            return c + scale * (c - u), None
        c, pc = uvit_oracle.uvit_forward(sd, cfg, x, t, context, mask, dtype, use_ground_truth=gt)
        u, pu = uvit_oracle.uvit_forward(sd, cfg, x, t, ec, mask, dtype, use_ground_truth=gt)
        c, pc, u, pu = c.float(), pc.float(), u.float(), pu.float()
        pm = pc + scale * (pc - pu)
        return c + scale * (c - u), pm

    return fn


def joint_sample(sd, cfg, z_init, mask_init, context, empty_context, scale, steps,
                 order=3, eps=1e-3, T=1.0, dtype=F32, trace=None):
    solver = Solver(cfg_model(sd, cfg, context, empty_context, scale, dtype), Schedule(), trace)
    return solver.sample_fast(z_init, mask_init, steps, order, eps, T)


# --- analog-bit codec (utils.py:475-518) ---
def int2bits(x: torch.Tensor, n: int = 8) -> torch.Tensor:
    x = x.to(torch.int32)
    planes = [torch.bitwise_right_shift(x, n - 1 - i) for i in range(n)]  # MSB first
    return torch.remainder(torch.cat(planes, dim=1), 2)


def bits2int(bits: torch.Tensor, n: int = 8) -> torch.Tensor:
    b = bits.to(torch.int32)
    y = torch.zeros(b.shape[0], 1, b.shape[2], b.shape[3])
    for i in range(n):
        y[:, 0] += b[:, i] * (2 ** (n - 1 - i))
    return y
