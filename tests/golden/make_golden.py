"""Generate the golden fixtures by EXECUTING the real reference (read-only,
``/root/reference``) on the CPU.  Run in the dev container only:

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  The reference cannot travel to the GPU box, so
the fixtures (and this script) are what is committed.  Harness recipe:
SURVEY.md App. B (panopticapi stub; live wiring = train_t2i_discrete.py:387-439,
:480-546).
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PDM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    for n in ("panopticapi", "panopticapi.utils"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["panopticapi.utils"].IdGenerator = object
    from libs.uvit_t2i import UViT, timestep_embedding  # noqa
    import dpm_solver_pp  # noqa
    return UViT, dpm_solver_pp


TINY = dict(img_size=8, patch_size=2, in_chans=4, embed_dim=64, depth=2, num_heads=1, mlp_ratio=4,
            qkv_bias=False, mlp_time_embed=False, clip_dim=32, num_clip_token=5,
            enable_panoptic=True, use_ground_truth=False, num_panoptic_class=8)


def build(UViT, separate, seed=1234):
    torch.manual_seed(seed)
    net = UViT(separate=separate, **TINY).eval()
    with torch.no_grad():
        for k, p in net.named_parameters():
            # make every weight participate: zero-initialised bridges / biases get small noise
            if k.startswith("zero_convs") or k.endswith(".bias"):
                p.copy_(torch.randn_like(p) * 0.02)
            if "norm" in k and k.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
    return net


def sd_betas():
    return (torch.linspace(0.00085 ** 0.5, 0.0120 ** 0.5, 1000, dtype=torch.float64) ** 2).float()


def main():
    UViT, dpm = import_reference()
    out = {}

    # --- KAT from the interpolate_fn docstring (dpm_solver_pp.py:21-24) ---
    a = dpm.interpolate_fn(torch.tensor([[0.5]]), torch.tensor([[0.0, 1.0]]), torch.tensor([[0.0, 2.0]]))
    b = dpm.interpolate_fn(torch.tensor([[-10.0]]), torch.tensor([[0.0, 1.0]]), torch.tensor([[0.0, 2.0]]))
    assert float(a) == 1.0 and float(b) == -20.0

    # --- schedule scalars ---
    ns = dpm.NoiseScheduleVP("discrete", betas=sd_betas())
    tq = torch.tensor([1.0, 0.98002, 0.5, 0.3337, 0.0410, 0.001, 0.0005, 1.2], dtype=torch.float32)
    sched = dict(t=tq.numpy(), log_alpha=ns.marginal_log_mean_coeff(tq).numpy(),
                 sigma=ns.marginal_std(tq).numpy(), lam=ns.marginal_lambda(tq).numpy())
    sched["inv_lam"] = ns.inverse_lambda(torch.from_numpy(sched["lam"])).numpy()
    np.savez(os.path.join(HERE, "schedule.npz"), **sched)

    B = 2
    for name, separate in (("single", False), ("two", True)):
        net = build(UViT, separate)
        g = torch.Generator().manual_seed(4321)
        x = torch.randn(B, 4, 8, 8, generator=g)
        m = torch.randn(B, 8, 8, 8, generator=g)
        ctx = torch.randn(B, 5, 32, generator=g)
        empty = torch.randn(5, 32, generator=g)
        t = torch.tensor([999.0, 123.456])
        with torch.no_grad():
            noise, y = net(x, t, ctx, mask_token=m)
            noise_nomask = net(x, t, ctx)

        scale = 2.0

        def model_fn(xx, t_cont, panoptic=None, mask_token=None, use_ground_truth=False, enable_panoptic=False):
            tt = t_cont * 1000
            ec = empty.unsqueeze(0).expand(xx.shape[0], -1, -1)
            c, pc = net(xx, tt, context=ctx, mask_token=mask_token)
            u, pu = net(xx, tt, context=ec, mask_token=mask_token)
            pm = pc + scale * (pc - pu)
            return c + scale * (c - u), pm

        res = {}
        for steps in (20, 7, 9):  # orders [3]*6+[2]; [3,3,1]; [3,3,2,1]
            solver = dpm.DPM_Solver(model_fn, ns, predict_x0=True, thresholding=False)
            with torch.no_grad():
                z, pm = solver.sample(x.clone(), steps=steps, eps=1e-3, T=1.0, order=3, mask_token=m.clone(),
                                      enable_mask_opt=True, enable_panoptic=True)
            res[f"z{steps}"] = z.numpy()
            res[f"pm{steps}"] = pm.numpy()

        # --- ground-truth evaluation (libs/uvit_t2i.py:486-496) and the two-phase singlestep driver (dpm_solver_pp.py:1045-1078)
        with torch.no_grad():
            noise_gt, y_gt = net(x, t, ctx, mask_token=m, use_ground_truth=True)
        assert torch.equal(y_gt, m)
        res["noise_gt"] = noise_gt.numpy()

        def model_fn_gt(xx, t_cont, panoptic=None, mask_token=None, use_ground_truth=False, enable_panoptic=False):
            tt = t_cont * 1000
            ec = empty.unsqueeze(0).expand(xx.shape[0], -1, -1)
            c, pc = net(xx, tt, context=ctx, mask_token=mask_token, use_ground_truth=use_ground_truth)
            u, pu = net(xx, tt, context=ec, mask_token=mask_token, use_ground_truth=use_ground_truth)
            return c + scale * (c - u), pc + scale * (pc - pu)

        for order, steps in ((3, 9), (2, 8), (1, 4)):
            solver = dpm.DPM_Solver(model_fn_gt, ns, predict_x0=True, thresholding=False)
            with torch.no_grad():
                z, pm = solver.sample(x.clone(), steps=steps, eps=1e-3, T=1.0, order=order, method="singlestep",
                                      mask_token=m.clone(), enable_mask_opt=True, enable_panoptic=True, use_twophases=True)
                z1, pm1 = solver.sample(x.clone(), steps=steps, eps=1e-3, T=1.0, order=order, method="singlestep",
                                        mask_token=m.clone(), enable_mask_opt=True, enable_panoptic=True)
            res[f"z_tp{order}"], res[f"pm_tp{order}"] = z.numpy(), pm.numpy()
            res[f"z_ss{order}"], res[f"pm_ss{order}"] = z1.numpy(), pm1.numpy()

        sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
        np.savez_compressed(
            os.path.join(HERE, f"tiny_{name}.npz"),
            x=x.numpy(), m=m.numpy(), ctx=ctx.numpy(), empty=empty.numpy(), t=t.numpy(),
            noise=noise.numpy(), y=y.numpy(), noise_nomask=noise_nomask.numpy(), scale=np.float32(scale),
            **res, **{"sd/" + k: v for k, v in sd.items()})

    # --- multistep pure updates (dpm_solver_pp.py:602-677) ---
    solver = dpm.DPM_Solver(lambda *a, **k: None, ns, predict_x0=True)
    g = torch.Generator().manual_seed(7)
    xs = torch.randn(2, 4, 8, 8, generator=g)
    X2, X1, X0 = (torch.randn(2, 4, 8, 8, generator=g) for _ in range(3))
    tt = [torch.full((2,), v) for v in (0.9, 0.8, 0.7, 0.6)]
    m2 = solver.dpm_multistep_second_update(xs, [X1, X0], [tt[1], tt[2]], tt[3], solver_type="dpm_solver")
    m3 = solver.dpm_multistep_third_update(xs, [X2, X1, X0], [tt[0], tt[1], tt[2]], tt[3], solver_type="dpm_solver")
    np.savez(os.path.join(HERE, "multistep.npz"), x=xs.numpy(), X2=X2.numpy(), X1=X1.numpy(), X0=X0.numpy(),
             t=np.array([0.9, 0.8, 0.7, 0.6], dtype=np.float32), m2=m2.numpy(), m3=m3.numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
