"""Drop-in ``NoiseScheduleVP`` / ``DPM_Solver`` for the joint image+mask sampler.

API mirrors the reference ``dpm_solver_pp.py`` (ctor ``:55``/``:291-308``, ``sample`` ``:927-930``), but
the work is split differently:

* all solver scalars (alpha/sigma/lambda, r1/r2, phi coefficients) are data independent, so the HOST
  builds a flat *plan* once (``build_plan``): one 16-float record per network evaluation.  This
  removes the reference's ~400 ``interpolate_fn`` sort/gather launches per sample batch;
* per evaluation the device runs ONE fused kernel (``pdm_cfg_update``): CFG combine + eps->x0 + the
  linear singlestep update for the image and the mask stream;
* when the model is a ``CFGModel`` (our ``UViT`` + classifier-free guidance) the whole loop runs
  inside ``pdm_sample`` as a CUDA graph; any other callable ``model_fn`` is driven from Python with
  the same plan and kernel (reference callback semantics, ``dpm_solver_pp.py:310-328``).

Scalars are computed with float32 torch CPU ops in the reference's operand order, so on identical
inputs the plan is bit-identical to the numbers the reference derives on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

PLAN_STRIDE = _lib.PLAN_STRIDE


def _pwl(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear interpolation of ``x`` (any shape) on knots ``xp`` (ascending) with linear
    extrapolation beyond the end knots; value-compatible with ``interpolate_fn`` (dpm_solver_pp.py:9-52):
    the segment is the one whose right knot is the first knot >= x."""
    K = xp.numel()
    flat = x.reshape(-1)
    idx = torch.searchsorted(xp, flat, right=False)           # number of knots strictly below x
    lo = torch.clamp(idx - 1, 0, K - 2)
    sx, ex, sy, ey = xp[lo], xp[lo + 1], yp[lo], yp[lo + 1]
    return (sy + (flat - sx) * (ey - sy) / (ex - sx)).reshape(x.shape)


class NoiseScheduleVP:
    """Forward VP SDE wrapper; 'discrete' (used by the t2i path) and 'linear' schedules."""

    def __init__(self, schedule="discrete", beta_0=1e-4, beta_1=2e-2, total_N=1000, betas=None, alphas_cumprod=None):
        if schedule not in ("linear", "discrete", "cosine"):
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be 'linear' or 'cosine'".format(schedule))
        if schedule == "cosine":
            raise NotImplementedError("cosine schedule is not used by the t2i sampling path")
        self.schedule = schedule
        self.total_N = total_N
        self.beta_0 = beta_0 * 1000.0
        self.beta_1 = beta_1 * 1000.0
        self.T = 1.0
        if schedule == "discrete":
            if betas is not None:
                betas = torch.as_tensor(betas).detach().float().cpu()
                log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
            else:
                assert alphas_cumprod is not None
                log_alphas = 0.5 * torch.log(torch.as_tensor(alphas_cumprod).detach().float().cpu())
            self.total_N = len(log_alphas)
            self.t_discrete = torch.linspace(1.0 / self.total_N, 1.0, self.total_N).reshape((1, -1))
            self.log_alpha_discrete = log_alphas.reshape((1, -1))
            self._t = self.t_discrete.reshape(-1)
            self._la = self.log_alpha_discrete.reshape(-1)
            self._la_rev = torch.flip(self._la, [0])
            self._t_rev = torch.flip(self._t, [0])

    @staticmethod
    def _host(t) -> torch.Tensor:
        return torch.as_tensor(t, dtype=torch.float32).detach().cpu()

    def marginal_log_mean_coeff(self, t):
        t = self._host(t)
        if self.schedule == "linear":
            return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return _pwl(t, self._t, self._la)

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))

    def inverse_lambda(self, lamb):
        lamb = self._host(lamb)
        if self.schedule == "linear":
            tmp = 2.0 * (self.beta_1 - self.beta_0) * torch.logaddexp(-2.0 * lamb, torch.zeros((1,)))
            delta = self.beta_0 ** 2 + tmp
            return tmp / (torch.sqrt(delta) + self.beta_0) / (self.beta_1 - self.beta_0)
        la = -0.5 * torch.logaddexp(torch.zeros((1,)), -2.0 * lamb)
        return _pwl(la, self._la_rev, self._t_rev).reshape(lamb.shape)


def fast_orders(steps: int, order: int) -> List[int]:
    """Order list of DPM-Solver-fast for ``steps`` function evaluations (dpm_solver_pp.py:386-403)."""
    if order == 3:
        K = steps // 3 + 1
        rem = steps % 3
        tail = {0: [2, 1], 1: [1], 2: [2]}[rem]
        return [3] * (K - len(tail)) + tail
    if order == 2:
        return [2] * (steps // 2) + ([1] if steps % 2 else [])
    raise ValueError("order must >= 2")


def _time_steps(ns: NoiseScheduleVP, skip_type: str, t_T: float, t_0: float, N: int) -> torch.Tensor:
    if skip_type == "time_uniform":
        return torch.linspace(t_T, t_0, N + 1)
    if skip_type == "logSNR":
        lam_T = ns.marginal_lambda(torch.tensor(t_T))
        lam_0 = ns.marginal_lambda(torch.tensor(t_0))
        return ns.inverse_lambda(torch.linspace(lam_T.item(), lam_0.item(), N + 1))
    if skip_type == "t2":
        return torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2)
    raise ValueError("Unsupported skip_type {}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'".format(skip_type))


def _rec(t_eval, ns, A, B_img, C_img, B_msk, C_msk, stage, last, n_time) -> List[float]:
    has_c = 0.0 if C_img is None else 1.0
    f = lambda v: 0.0 if v is None else float(v)  # noqa: E731  (float32 tensor -> python float is exact)
    rec = [float(t_eval * n_time), float(ns.marginal_alpha(t_eval)), float(ns.marginal_std(t_eval)), f(A), f(B_img),
           f(C_img), f(B_msk), f(C_msk), float(stage), has_c, 1.0 if last else 0.0]
    return rec + [0.0] * (PLAN_STRIDE - len(rec))


def _step_records(ns, s, t, order, r1, r2, mask_opt: bool, n_time: float) -> List[List[float]]:
    """Plan records of ONE singlestep update s -> t (data prediction, solver_type='dpm_solver').
    Formulas: dpm_solver_pp.py:432-457 (1S), :511-557 (2S), :700-766 (3S); SURVEY App. A.3.
    ``mask_opt=False`` reproduces the reference's pass-through: intermediate mask = mask_token, and the
    mask handed to the next step is the prediction itself (``return x_t, pred_mask, pred_mask``)."""
    lam_s, lam_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
    h = lam_t - lam_s
    sig_s, sig_t = ns.marginal_std(s), ns.marginal_std(t)
    a_t = torch.exp(ns.marginal_log_mean_coeff(t))
    one, zero = torch.tensor(1.0), torch.tensor(0.0)
    recs = []
    if order == 1:
        phi_1 = (torch.exp(-h) - 1.0) / (-1.0)
        B = a_t * phi_1
        if mask_opt:
            recs.append(_rec(s, ns, sig_t / sig_s, B, None, B, None, 0, True, n_time))
        else:
            recs.append(_rec_split(s, ns, sig_t / sig_s, B, None, zero, one, None, 0, True, n_time))
        return recs
    if order == 2:
        r1 = 0.5 if r1 is None else r1
        s1 = ns.inverse_lambda(lam_s + r1 * h)
        sig_s1 = ns.marginal_std(s1)
        a_s1 = torch.exp(ns.marginal_log_mean_coeff(s1))
        phi_11 = torch.expm1(-r1 * h)
        phi_1 = torch.expm1(-h)
        B0 = a_s1 * phi_11
        Bt = a_t * phi_1
        Ct = (0.5 / r1) * (a_t * phi_1)
        if mask_opt:
            recs.append(_rec(s, ns, sig_s1 / sig_s, -B0, None, B0, None, 0, False, n_time))   # '+' quirk on the mask (:536-539)
            recs.append(_rec(s1, ns, sig_t / sig_s, -Bt, -Ct, -Bt, -Ct, 1, True, n_time))
        else:
            recs.append(_rec_split(s, ns, sig_s1 / sig_s, -B0, None, one, zero, None, 0, False, n_time))
            recs.append(_rec_split(s1, ns, sig_t / sig_s, -Bt, -Ct, zero, one, zero, 1, True, n_time))
        return recs
    if order == 3:
        r1 = 1.0 / 3.0 if r1 is None else r1
        r2 = 2.0 / 3.0 if r2 is None else r2
        s1 = ns.inverse_lambda(lam_s + r1 * h)
        s2 = ns.inverse_lambda(lam_s + r2 * h)
        sig_s1, sig_s2 = ns.marginal_std(s1), ns.marginal_std(s2)
        a_s1 = torch.exp(ns.marginal_log_mean_coeff(s1))
        a_s2 = torch.exp(ns.marginal_log_mean_coeff(s2))
        phi_11 = torch.expm1(-r1 * h)
        phi_12 = torch.expm1(-r2 * h)
        phi_1 = torch.expm1(-h)
        phi_22 = torch.expm1(-r2 * h) / (r2 * h) + 1.0
        phi_2 = phi_1 / h + 1.0
        B0 = a_s1 * phi_11
        B1 = a_s2 * phi_12
        C1 = r2 / r1 * (a_s2 * phi_22)
        Bt = a_t * phi_1
        Ct = (1.0 / r2) * (a_t * phi_2)
        if mask_opt:
            recs.append(_rec(s, ns, sig_s1 / sig_s, -B0, None, B0, None, 0, False, n_time))   # '+' quirk (:730-733)
            recs.append(_rec(s1, ns, sig_s2 / sig_s, -B1, C1, -B1, C1, 1, False, n_time))
            recs.append(_rec(s2, ns, sig_t / sig_s, -Bt, Ct, -Bt, Ct, 2, True, n_time))
        else:
            recs.append(_rec_split(s, ns, sig_s1 / sig_s, -B0, None, one, zero, None, 0, False, n_time))
            recs.append(_rec_split(s1, ns, sig_s2 / sig_s, -B1, C1, one, zero, zero, 1, False, n_time))
            recs.append(_rec_split(s2, ns, sig_t / sig_s, -Bt, Ct, zero, one, zero, 2, True, n_time))
        return recs
    raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))


def _rec_split(t_eval, ns, A, B_img, C_img, A_msk, B_msk, C_msk, stage, last, n_time) -> List[float]:
    """Record whose mask stream uses its own A (pass-through encodings); stored in the reserved slot 11
    with flag slot 12 = 1."""
    rec = _rec(t_eval, ns, A, B_img, C_img, B_msk, C_msk, stage, last, n_time)
    rec[11] = float(A_msk)
    rec[12] = 1.0
    return rec


def build_plan(ns: NoiseScheduleVP, steps: int, order: int = 3, eps: float = 1e-3, T: Optional[float] = None,
               skip_type: str = "time_uniform", method: str = "fast", mask_opt: bool = True,
               n_time: float = 1000.0) -> np.ndarray:
    """Flat per-evaluation coefficient table for ``method`` in {'fast', 'singlestep'}
    (dpm_solver_pp.py:1018-1078).  Shape [n_evals, PLAN_STRIDE], float32."""
    t_0 = eps
    t_T = ns.T if T is None else T
    recs: List[List[float]] = []
    if method == "fast":
        orders = fast_orders(steps, order)
        ts = _time_steps(ns, skip_type, t_T, t_0, steps)
        i = 0
        for o in orders:
            h = ns.marginal_lambda(ts[i + o]) - ns.marginal_lambda(ts[i])
            r1 = None if o <= 1 else (ns.marginal_lambda(ts[i + 1]) - ns.marginal_lambda(ts[i])) / h
            r2 = None if o <= 2 else (ns.marginal_lambda(ts[i + 2]) - ns.marginal_lambda(ts[i])) / h
            recs += _step_records(ns, ts[i], ts[i + o], o, r1, r2, mask_opt, n_time)
            i += o
    elif method == "singlestep":
        n_steps = steps // order
        ts = _time_steps(ns, skip_type, t_T, t_0, n_steps)
        for i in range(n_steps):
            recs += _step_records(ns, ts[i], ts[i + 1], order, None, None, mask_opt, n_time)
    else:
        raise ValueError(f"build_plan: unsupported method {method!r}")
    return np.asarray(recs, dtype=np.float32).reshape(-1, PLAN_STRIDE)


def build_plan_c(betas, steps: int, order: int = 3, eps: float = 1e-3, T: float = 1.0, skip_type: str = "time_uniform",
                 method: str = "fast", mask_opt: bool = True, n_time: float = 1000.0) -> np.ndarray:
    """The same table from the C ABI's host planner (``pdm_solver_plan``, csrc/plan.cu) -- what a non-Python consumer of
    libpdm.so calls.  Agrees with ``build_plan`` / ``build_multistep_plan`` to a few float32 ulp (libm vs SLEEF rounding)."""
    L = _lib.lib()
    b = np.ascontiguousarray(torch.as_tensor(betas).detach().float().cpu().numpy(), dtype=np.float32)
    n = C.c_int32(0)
    args = (b.ctypes.data_as(C.POINTER(C.c_float)), b.size, steps, order, _lib.METHOD_CODES[method], _lib.SKIP_CODES[skip_type],
            float(eps), float(T), 1 if mask_opt else 0, float(n_time))
    _lib.check(L.pdm_solver_plan(*args, None, 0, C.byref(n)))
    out = np.zeros((n.value, PLAN_STRIDE), dtype=np.float32)
    _lib.check(L.pdm_solver_plan(*args, out.ctypes.data_as(C.POINTER(C.c_float)), n.value, C.byref(n)))
    return out


class DPM_Solver:
    def __init__(self, model_fn, noise_schedule, predict_x0=False, thresholding=False, max_val=1.0, n_time=1000.0):
        """``model_fn(x, t_continuous, panoptic=, mask_token=, use_ground_truth=, enable_panoptic=) -> (noise, pred_mask)``
        (dpm_solver_pp.py:291-308).  Only the data-prediction mode the t2i path uses is implemented."""
        if not predict_x0:
            raise NotImplementedError("the joint sampling path runs DPM-Solver++ (predict_x0=True)")
        if thresholding:
            raise NotImplementedError("thresholding is off on the t2i path (train_t2i_discrete.py:518)")
        self.model = model_fn
        self.noise_schedule = noise_schedule
        self.predict_x0 = predict_x0
        self.thresholding = thresholding
        self.max_val = max_val
        self.n_time = n_time
        self.use_graph = True

    def get_time_steps(self, skip_type, t_T, t_0, N, device=None):
        return _time_steps(self.noise_schedule, skip_type, t_T, t_0, N)

    def get_time_steps_for_dpm_solver_fast(self, skip_type, t_T, t_0, steps, order, device=None):
        orders = fast_orders(steps, order)
        K = steps // 3 + 1 if order == 3 else steps // 2
        return orders, _time_steps(self.noise_schedule, skip_type, t_T, t_0, K)

    # -------------------------------------------------------------------------------------------
    def sample(self, x, steps=10, eps=1e-4, T=None, order=3, panoptic=None, skip_type="time_uniform", denoise=False,
               method="fast", solver_type="dpm_solver", atol=0.0078, rtol=0.05, mask_token=None, use_twophases=False,
               use_ground_truth=False, enable_panoptic=False, enable_mask_opt=False):
        """Returns ``(x_0, pred_mask)`` like the reference (dpm_solver_pp.py:1044)."""
        if solver_type != "dpm_solver":
            raise NotImplementedError("solver_type 'taylor' is not on the sampling hot path")
        if use_twophases and method != "singlestep":
            use_twophases = False  # the reference only looks at the flag in the singlestep driver (dpm_solver_pp.py:1071)
        if denoise:
            raise NotImplementedError("denoise=True is unreachable in the reference for these methods")
        if method == "multistep":
            return self._sample_multistep(x, steps, eps, T, order, skip_type, mask_token)
        if method not in ("fast", "singlestep"):
            raise NotImplementedError(f"method {method!r} is not on the sampling hot path")
        if method == "singlestep" and order not in (1, 2, 3):
            raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))
        if not x.is_cuda:
            raise RuntimeError("DPM_Solver (libpdm) has no CPU path: x must be a CUDA tensor")
        plan = build_plan(self.noise_schedule, steps, order, eps, T, skip_type, method,
                          mask_opt=bool(enable_mask_opt) or mask_token is None, n_time=self.n_time)
        fast = getattr(self.model, "_pdm_fast_path", False) and not use_ground_truth and not use_twophases
        if fast:
            return self.model.run_plan(x, mask_token, plan, use_graph=self.use_graph)
        x, pred_mask, mask_t = self._sample_callback(x, mask_token, plan, enable_panoptic, use_ground_truth)
        if use_twophases:
            # phase two (dpm_solver_pp.py:1071-1075): the same time grid again, starting from the phase-one image, with the
            # phase-one mask held fixed and fed to the network as ground truth (the reference runs this pass whether or not
            # there is a mask stream); the returned pred_mask is phase one's
            plan2 = build_plan(self.noise_schedule, steps, order, eps, T, skip_type, method, mask_opt=False, n_time=self.n_time)
            x, _, _ = self._sample_callback(x, mask_t, plan2, True, True)
        return x, pred_mask

    # ---- generic model_fn: Python loop, ONE fused K12 kernel per evaluation ----------------------------
    @torch.no_grad()
    def _sample_callback(self, x, mask_token, plan, enable_panoptic, use_ground_truth=False):
        """Reference callback semantics (dpm_solver_pp.py:310-328) for an arbitrary ``model_fn``.  A model that offers
        ``eval_pair`` (our CFGModel) hands back the un-combined cond / uncond outputs so that the guidance combine stays
        inside the update kernel here as well."""
        L = _lib.lib()
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        xbase = x.to(**f32).contiguous().clone()
        xin = torch.empty_like(xbase)
        X0 = torch.empty_like(xbase)
        has_mask = mask_token is not None
        if has_mask:
            mbase = mask_token.to(**f32).contiguous().clone()
            min_, P0 = torch.empty_like(mbase), torch.empty_like(mbase)
        else:
            mbase = min_ = P0 = None
        B = x.shape[0]
        pred_mask = mask_token
        pair = getattr(self.model, "eval_pair", None)
        with torch.cuda.device(dev):
            for rec in plan:
                stage, last = int(rec[8]), rec[10] != 0
                cur_x = xbase if stage == 0 else xin
                cur_m = (mbase if stage == 0 else min_) if has_mask else None
                t_cont = torch.full((B,), float(rec[0]) / self.n_time, **f32)
                if pair is not None:
                    ec, eu, pc, pu, scale = pair(cur_x, t_cont, mask_token=cur_m, use_ground_truth=use_ground_truth)
                else:
                    noise, pm = self.model(cur_x, t_cont, panoptic=pred_mask, mask_token=cur_m,
                                           use_ground_truth=use_ground_truth, enable_panoptic=enable_panoptic)
                    ec, eu, scale = noise.to(**f32).contiguous(), None, 0.0
                    pc, pu = (pm.to(**f32).contiguous() if (has_mask and pm is not None) else None), None
                coef = np.ascontiguousarray(rec, dtype=np.float32)
                m_out = (mbase if last else min_) if has_mask else None
                _lib.check(L.pdm_cfg_update(
                    _lib.ptr(ec), _lib.ptr(eu), _lib.ptr(pc), _lib.ptr(pu), _lib.ptr(cur_x), _lib.ptr(xbase), _lib.ptr(X0),
                    _lib.ptr(xbase if last else xin), _lib.ptr(mbase), _lib.ptr(P0), _lib.ptr(m_out),
                    coef.ctypes.data_as(C.POINTER(C.c_float)), float(scale), xbase.numel(),
                    mbase.numel() if has_mask else 0, _lib.current_stream()))
                if stage == 0 and has_mask:
                    pred_mask = P0
        return xbase, (P0 if has_mask else None), (mbase if has_mask else None)

    # ---- multistep 2M / 3M (dpm_solver_pp.py:602-677, driver :995-1017 repaired) ----------
    @torch.no_grad()
    def _sample_multistep(self, x, steps, eps, T, order, skip_type, mask_token):
        from .multistep import sample_multistep
        return sample_multistep(self, x, steps, eps, T, order, skip_type, mask_token)
