"""DPM-Solver++ multistep (2M / 3M) sampling, data prediction, ``solver_type='dpm_solver'``.

The reference's driver for ``method='multistep'`` is broken (``dpm_solver_pp.py:995-1017``: ``timesteps`` is read
before assignment and ``model_fn`` returns a tuple), but its pure update functions
(``dpm_multistep_second_update`` ``:602-642``, ``dpm_multistep_third_update`` ``:645-677``, order-1 warm-up through
``dpm_solver_first_update``) work.  This module repairs the driver: warm-up with orders 1..order-1, then order
``order`` -- every model evaluation is followed by exactly one update, so each step is ONE fused device kernel
(``pdm_multistep_update``: CFG combine + eps->x0 + the 2M/3M linear update).

Mask stream: the reference defines none for this method; here the mask state gets the same update with the mask
prediction as its data prediction (*parity unpinned*).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np
import torch

from . import _lib

STRIDE = _lib.PLAN_STRIDE


def multistep_record(ns, t_hist, t, n_time: float = 1000.0) -> List[float]:
    """Coefficients of ONE multistep update from t_hist[-1] to t using len(t_hist) cached predictions (order =
    len(t_hist)), in the reference's float32 operand order (dpm_solver_pp.py:432-446, 606-628, 649-669)."""
    o = len(t_hist)
    p0 = t_hist[-1]
    lam0, lam_t = ns.marginal_lambda(p0), ns.marginal_lambda(t)
    sig0, sig_t = ns.marginal_std(p0), ns.marginal_std(t)
    a_t = torch.exp(ns.marginal_log_mean_coeff(t))
    h = lam_t - lam0
    rec = [0.0] * STRIDE
    rec[0] = float(p0 * n_time)
    rec[1], rec[2] = float(ns.marginal_alpha(p0)), float(ns.marginal_std(p0))
    rec[3] = float(sig_t / sig0)
    rec[11] = float(o)
    rec[15] = 1.0  # record kind: multistep (include/pdm.h)
    if o == 1:
        phi_1 = (torch.exp(-h) - 1.0) / (-1.0)
        rec[4] = float(a_t * phi_1)
        return rec
    lam1 = ns.marginal_lambda(t_hist[-2])
    h_0 = lam0 - lam1
    r0 = h_0 / h
    B = a_t * (torch.exp(-h) - 1.0)
    rec[4], rec[7], rec[12] = float(B), float(1.0 / r0), float(0.5 * B)
    if o == 3:
        lam2 = ns.marginal_lambda(t_hist[-3])
        h_1 = lam1 - lam2
        r1 = h_1 / h
        rec[5] = float(a_t * ((torch.exp(-h) - 1.0) / h + 1.0))
        rec[6] = float(a_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5))
        rec[8] = float(1.0 / r1)
        rec[9] = float(r0 / (r0 + r1))
        rec[10] = float(1.0 / (r0 + r1))
    return rec


def build_multistep_plan(ns, steps: int, order: int, eps: float, T, skip_type: str, n_time: float = 1000.0) -> np.ndarray:
    """One record per model evaluation k (at ts[k]): the update ts[k] -> ts[k+1] of order min(k+1, order)
    (warm-up with lower orders exactly like the reference driver, dpm_solver_pp.py:1004-1008)."""
    from .dpm_solver_pp import _time_steps
    assert steps >= order
    t_T = ns.T if T is None else T
    ts = _time_steps(ns, skip_type, t_T, eps, steps)
    recs = []
    for k in range(steps):
        o = min(k + 1, order)
        recs.append(multistep_record(ns, [ts[i] for i in range(k - o + 1, k + 1)], ts[k + 1], n_time))
    return np.asarray(recs, dtype=np.float32).reshape(-1, STRIDE)


@torch.no_grad()
def sample_multistep(solver, x, steps, eps, T, order, skip_type, mask_token):
    if order not in (1, 2, 3):
        raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))
    if not x.is_cuda:
        raise RuntimeError("DPM_Solver (libpdm) has no CPU path: x must be a CUDA tensor")
    plan = build_multistep_plan(solver.noise_schedule, steps, order, eps, T, skip_type, solver.n_time)
    if getattr(solver.model, "_pdm_fast_path", False):
        # our UViT + guidance: the whole 2M / 3M loop runs on the device (pdm_sample: one forward + ONE fused update kernel
        # per step, history ring in the engine workspace, captured into a CUDA graph)
        return solver.model.run_plan(x, mask_token, plan, use_graph=solver.use_graph)
    L = _lib.lib()
    dev = x.device
    f32 = dict(device=dev, dtype=torch.float32)
    cur = x.to(**f32).contiguous().clone()
    nxt = torch.empty_like(cur)
    hist = [torch.empty_like(cur) for _ in range(3)]          # ring of data predictions
    has_mask = mask_token is not None
    if has_mask:
        m_cur = mask_token.to(**f32).contiguous().clone()
        m_nxt = torch.empty_like(m_cur)
        phist = [torch.empty_like(m_cur) for _ in range(3)]
    B = x.shape[0]
    model = solver.model
    pair = getattr(model, "eval_pair", None)
    with torch.cuda.device(dev):
        for k, rec in enumerate(plan):
            t_model = float(rec[0])
            X0, X1, X2 = hist[k % 3], hist[(k - 1) % 3], hist[(k - 2) % 3]
            t_cont = torch.full((B,), t_model / solver.n_time, **f32)
            if pair is not None:
                ec, eu, pc, pu, scale = pair(cur, t_cont, mask_token=m_cur if has_mask else None)
            else:
                out = model(cur, t_cont, panoptic=None, mask_token=m_cur if has_mask else None)
                noise, pm = out if isinstance(out, tuple) else (out, None)
                ec, eu, scale = noise.to(**f32).contiguous(), None, 0.0
                pc, pu = (pm.to(**f32).contiguous() if (has_mask and pm is not None) else None), None
            coef = np.ascontiguousarray(rec, dtype=np.float32)
            _lib.check(L.pdm_multistep_update(
                _lib.ptr(ec), _lib.ptr(eu), _lib.ptr(pc), _lib.ptr(pu), _lib.ptr(cur), _lib.ptr(X1), _lib.ptr(X2),
                _lib.ptr(X0), _lib.ptr(nxt),
                _lib.ptr(m_cur) if has_mask else None, _lib.ptr(phist[(k - 1) % 3]) if has_mask else None,
                _lib.ptr(phist[(k - 2) % 3]) if has_mask else None, _lib.ptr(phist[k % 3]) if has_mask else None,
                _lib.ptr(m_nxt) if has_mask else None, coef.ctypes.data_as(C.POINTER(C.c_float)), float(scale),
                cur.numel(), m_cur.numel() if has_mask else 0, _lib.current_stream()))
            cur, nxt = nxt, cur
            if has_mask:
                m_cur, m_nxt = m_nxt, m_cur
    pred_mask = phist[(len(plan) - 1) % 3] if has_mask else None
    return cur, pred_mask
