"""Classifier-free-guidance wiring of the joint sampler -- the working equivalent of the live
reference path ``train_t2i_discrete.py:387-439`` (``cfg_nnet``) + ``:480-546`` (``dpm_solver_sample``).

``CFGModel`` is both a reference-style ``model_fn`` (callable from any ``DPM_Solver``) and the
object the fast path recognises: ``DPM_Solver.sample`` hands it the whole plan and the loop runs on
the device (``pdm_sample``: cond+uncond batched as one 2B forward, K12 fused update, CUDA graph).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .dpm_solver_pp import DPM_Solver, NoiseScheduleVP


def stable_diffusion_beta_schedule(linear_start=0.00085, linear_end=0.0120, n_timestep=1000):
    """train_t2i_discrete.py:40-44 (float64 linspace of sqrt-betas, squared)."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


class CFGModel:
    """model_fn(x, t_continuous, panoptic=None, mask_token=None, ...) -> (eps, pred_mask) with guidance on
    both outputs: eps = c + s (c - u), pred_mask = pm_c + s (pm_c - pm_u)."""

    _pdm_fast_path = True

    def __init__(self, nnet, context: torch.Tensor, empty_context: Optional[torch.Tensor], scale: float,
                 cfg: bool = True, n_time: int = 1000):
        self.nnet = nnet
        self.context = context
        self.empty_context = empty_context if cfg else None
        self.scale = float(scale)
        self.n_time = n_time

    # ---- reference-style callback (one evaluation) ----
    @torch.no_grad()
    def __call__(self, x, t_continuous, panoptic=None, mask_token=None, use_ground_truth=False, enable_panoptic=False):
        t = t_continuous * self.n_time
        B = x.shape[0]
        gt = bool(use_ground_truth)
        if self.empty_context is None:
            out = self.nnet(x, t, self.context, mask_token=mask_token, use_ground_truth=gt)
            return out if mask_token is not None else (out, None)
        # cond + uncond as one 2B batch (bit-identical to two B forwards: no cross-sample op)
        ctx2 = torch.cat([self.context, self.empty_context.unsqueeze(0).expand(B, -1, -1)], dim=0)
        x2 = torch.cat([x, x], dim=0)
        t2 = torch.cat([t, t], dim=0) if torch.is_tensor(t) and t.dim() > 0 else t
        if mask_token is None:
            out = self.nnet(x2, t2, ctx2)
            c, u = out[:B], out[B:]
            return c + self.scale * (c - u), None
        noise, y = self.nnet(x2, t2, ctx2, mask_token=torch.cat([mask_token, mask_token], dim=0), use_ground_truth=gt)
        c, u, pc, pu = noise[:B], noise[B:], y[:B], y[B:]
        return c + self.scale * (c - u), pc + self.scale * (pc - pu)

    # ---- un-combined evaluation for the solver's own callback loop: the combine stays in the fused update kernel ----
    @torch.no_grad()
    def eval_pair(self, x, t_continuous, mask_token=None, use_ground_truth=False):
        """-> (eps_c, eps_u, pm_c, pm_u, scale): cond / uncond halves of ONE 2B-row forward (eps_u / pm_u None without
        guidance); contiguous views, no arithmetic."""
        t = t_continuous * self.n_time
        B = x.shape[0]
        gt = bool(use_ground_truth)
        if self.empty_context is None:
            out = self.nnet(x, t, self.context, mask_token=mask_token, use_ground_truth=gt)
            noise, y = out if mask_token is not None else (out, None)
            return noise, None, y, None, 0.0
        ctx2 = torch.cat([self.context, self.empty_context.unsqueeze(0).expand(B, -1, -1)], dim=0)
        x2 = torch.cat([x, x], dim=0)
        t2 = torch.cat([t, t], dim=0) if torch.is_tensor(t) and t.dim() > 0 else t
        if mask_token is None:
            noise = self.nnet(x2, t2, ctx2)
            return noise[:B], noise[B:], None, None, self.scale
        noise, y = self.nnet(x2, t2, ctx2, mask_token=torch.cat([mask_token, mask_token], dim=0), use_ground_truth=gt)
        return noise[:B], noise[B:], y[:B], y[B:], self.scale

    # ---- fast path: whole loop on the device ----
    @torch.no_grad()
    def run_plan(self, x, mask_token, plan: np.ndarray, use_graph: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        if not x.is_cuda:
            raise RuntimeError("libpdm has no CPU path: x must be a CUDA tensor")
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        h = self.nnet.engine()
        z = x.to(**f32).contiguous()
        m = None if mask_token is None else mask_token.to(**f32).contiguous()
        ctx = self.context.to(**f32).contiguous()
        ec = None if self.empty_context is None else self.empty_context.to(**f32).contiguous()
        out_z = torch.empty_like(z)
        out_pm = None if m is None else torch.empty_like(m)
        plan = np.ascontiguousarray(plan, dtype=np.float32)
        if z.shape[0] == 0:  # empty shard (e.g. a rank past the end of the prompt list): nothing to launch
            return out_z, out_pm
        if plan.shape[0] == 0:  # steps < order with method='singlestep': the reference's loop body never runs
            return z.clone(), (None if m is None else m.clone())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pdm_sample(
                h, plan.ctypes.data_as(C.POINTER(C.c_float)), plan.shape[0], _lib.ptr(z), _lib.ptr(m), _lib.ptr(ctx),
                _lib.ptr(ec), self.scale, _lib.ptr(out_z), _lib.ptr(out_pm), z.shape[0], self.nnet.prec_code(),
                1 if use_graph else 0, _lib.current_stream()))
        return out_z, out_pm


class JointSampler:
    """``dpm_solver_sample`` of the live path: noise init, mask init, DPM-Solver-fast order 3, CFG."""

    def __init__(self, nnet, z_shape=(4, 32, 32), mask_channels: int = 8, scale: float = 1.0, cfg: bool = True,
                 sample_steps: int = 50, betas=None, method: str = "fast"):
        self.nnet = nnet
        self.z_shape = tuple(z_shape)
        self.mask_channels = mask_channels
        self.scale, self.cfg, self.sample_steps = scale, cfg, sample_steps
        self.method = method  # 'fast' (singlestep 3,...,3,2: the live path, train_t2i_discrete.py:516) | 'multistep' (3M)
        betas = stable_diffusion_beta_schedule() if betas is None else betas
        self.N = len(betas)
        self.noise_schedule = NoiseScheduleVP(schedule="discrete", betas=torch.tensor(betas).float())

    @torch.no_grad()
    def sample(self, context: torch.Tensor, empty_context: Optional[torch.Tensor], z_init: Optional[torch.Tensor] = None,
               mask_init: Optional[torch.Tensor] = None, use_panoptic: bool = True, generator=None, steps=None,
               use_graph: bool = True):
        dev = context.device
        B = context.shape[0]
        if z_init is None:
            z_init = torch.randn(B, *self.z_shape, device=dev, generator=generator)
        if use_panoptic and mask_init is None:
            mask_init = torch.randn(B, self.mask_channels, *self.z_shape[1:], device=dev, generator=generator)
        model = CFGModel(self.nnet, context, empty_context, self.scale, self.cfg, self.N)
        solver = DPM_Solver(model, self.noise_schedule, predict_x0=True, thresholding=False, n_time=float(self.N))
        solver.use_graph = use_graph
        steps = self.sample_steps if steps is None else steps
        if use_panoptic:
            return solver.sample(z_init, steps=steps, eps=1.0 / self.N, T=1.0, order=3, mask_token=mask_init,
                                 enable_mask_opt=True, enable_panoptic=True, method=self.method)
        z, _ = solver.sample(z_init, steps=steps, eps=1.0 / self.N, T=1.0, order=3, method=self.method)
        return z, None
