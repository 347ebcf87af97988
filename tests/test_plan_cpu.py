"""Host planner (panopticdiffusionmodels_b200.dpm_solver_pp.build_plan) against the oracle: the plan,
executed by a plain-torch emulation of the K12 kernel arithmetic, must reproduce the oracle's joint
sample bit for bit (the oracle itself is pinned to the reference in test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import TINY, load_golden
from oracle import dpm_oracle
from panopticdiffusionmodels_b200 import dpm_solver_pp as P


def emulate_plan(plan, model, x, m):
    """Same dataflow as pdm_sample / update_kernel (csrc/elementwise.cu), CFG already inside `model`."""
    f = torch.float32
    xbase, mbase = x.clone(), (None if m is None else m.clone())
    xin = X0 = min_ = P0 = None
    for rec in plan:
        rec = [torch.tensor(float(v), dtype=f) for v in rec]
        stage, has_c, last = int(rec[8]), rec[9] != 0, rec[10] != 0
        cur_x = xbase if stage == 0 else xin
        cur_m = mbase if (stage == 0 or m is None) else min_
        eps, pm = model(cur_x, rec[0] / 1000.0, cur_m)
        X = (cur_x - rec[2] * eps) / rec[1]
        if stage == 0:
            X0, P0 = X, pm
        out = rec[3] * xbase + rec[4] * X0
        if has_c:
            out = out + rec[5] * (X - X0)
        if m is not None:
            mo = rec[3] * mbase + rec[6] * P0
            if has_c:
                mo = mo + rec[7] * (pm - P0)
        if last:
            xbase = out
            if m is not None:
                mbase = mo
        else:
            xin = out
            if m is not None:
                min_ = mo
    return xbase, P0


def test_schedule_matches_oracle():
    ns = P.NoiseScheduleVP("discrete", betas=dpm_oracle.sd_betas())
    s = dpm_oracle.Schedule()
    ts = torch.linspace(1.0, 1e-3, 51)
    for t in ts:
        assert float(ns.marginal_lambda(t)) == float(s.lam(t))
        assert float(ns.marginal_std(t)) == float(s.sigma(t))
        assert float(ns.inverse_lambda(s.lam(t))) == float(s.inv_lam(s.lam(t)))
    g, _ = load_golden("schedule.npz")
    assert torch.equal(ns.marginal_log_mean_coeff(g["t"]), g["log_alpha"])
    assert torch.equal(ns.inverse_lambda(g["lam"]), g["inv_lam"])


def test_orders():
    for steps in range(3, 60):
        assert P.fast_orders(steps, 3) == dpm_oracle.fast_orders(steps, 3)
        assert sum(P.fast_orders(steps, 3)) == steps
    assert P.fast_orders(50, 3) == [3] * 16 + [2]
    assert P.fast_orders(20, 3) == [3] * 6 + [2]
    assert P.fast_orders(5, 2) == [2, 2, 1]


def test_plan_shape_and_times():
    ns = P.NoiseScheduleVP("discrete", betas=dpm_oracle.sd_betas())
    plan = P.build_plan(ns, 50, 3, eps=1e-3, T=1.0)
    assert plan.shape == (50, P.PLAN_STRIDE) and plan.dtype == np.float32
    assert plan[0, 0] == 1000.0 and abs(plan[1, 0] - 980.02) < 1e-2 and abs(plan[-1, 0] - 20.98) < 2e-2
    assert list(plan[:6, 8]) == [0, 1, 2, 0, 1, 2] and list(plan[-2:, 8]) == [0, 1]
    assert list(plan[:3, 10]) == [0, 0, 1] and plan[-1, 10] == 1
    # sign quirk: at stage 0 the mask coefficient has the opposite sign of the image one
    assert plan[0, 4] == -plan[0, 6] and plan[0, 6] < 0
    assert plan[1, 4] == plan[1, 6] and plan[2, 5] == plan[2, 7]


@pytest.mark.parametrize("name,separate", [("single", False), ("two", True)])
@pytest.mark.parametrize("steps", [20, 7, 9])
def test_plan_reproduces_oracle_bit_exact(name, separate, steps):
    g, sd = load_golden(f"tiny_{name}.npz")
    cfg = dict(TINY, separate=separate)
    model = dpm_oracle.cfg_model(sd, cfg, g["ctx"], g["empty"], float(g["scale"]))
    z_ref, pm_ref = dpm_oracle.Solver(model, dpm_oracle.Schedule()).sample_fast(g["x"], g["m"], steps)
    ns = P.NoiseScheduleVP("discrete", betas=dpm_oracle.sd_betas())
    plan = P.build_plan(ns, steps, 3, eps=1e-3, T=1.0)
    assert plan.shape[0] == steps
    z, pm = emulate_plan(plan, model, g["x"], g["m"])
    assert torch.equal(z, z_ref)
    assert torch.equal(pm, pm_ref)


def test_plan_image_only():
    g, sd = load_golden("tiny_single.npz")
    model = dpm_oracle.cfg_model(sd, dict(TINY, separate=False), g["ctx"], g["empty"], 2.0)
    z_ref, _ = dpm_oracle.Solver(model, dpm_oracle.Schedule()).sample_fast(g["x"], None, 8)
    ns = P.NoiseScheduleVP("discrete", betas=dpm_oracle.sd_betas())
    z, _ = emulate_plan(P.build_plan(ns, 8, 3, eps=1e-3, T=1.0), model, g["x"], None)
    assert torch.equal(z, z_ref)


def test_multistep_plan_matches_reference_updates():
    """multistep_record + the kernel's operation order (emulated with torch float32) == the reference's
    dpm_multistep_second/third_update outputs stored in tests/golden/multistep.npz (bit exact)."""
    from panopticdiffusionmodels_b200.multistep import build_multistep_plan, multistep_record
    g, _ = load_golden("multistep.npz")
    ns = P.NoiseScheduleVP("discrete", betas=dpm_oracle.sd_betas())
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)
    t = [g["t"][i] for i in range(4)]
    x, X2, X1, X0 = g["x"], g["X2"], g["X1"], g["X0"]
    rec = multistep_record(ns, t[:3], t[3])
    assert rec[11] == 3
    A, B, C1, C2, ir0, ir1, q, ir01 = (f(rec[i]) for i in (3, 4, 5, 6, 7, 8, 9, 10))
    D10 = ir0 * (X0 - X1)
    D11 = ir1 * (X1 - X2)
    d = D10 - D11
    out3 = A * x - B * X0 + C1 * (D10 + q * d) - C2 * (ir01 * d)
    assert torch.equal(out3, g["m3"])
    rec = multistep_record(ns, t[1:3], t[3])
    assert rec[11] == 2
    A, B, ir0, hB = (f(rec[i]) for i in (3, 4, 7, 12))
    out2 = A * x - B * X0 - hB * (ir0 * (X0 - X1))
    assert torch.equal(out2, g["m2"])
    plan = build_multistep_plan(ns, 10, 3, eps=1e-3, T=1.0, skip_type="time_uniform")
    assert plan.shape == (10, P.PLAN_STRIDE) and list(plan[:4, 11]) == [1, 2, 3, 3] and plan[0, 0] == 1000.0


@pytest.mark.parametrize("method,order,steps", [("fast", 3, 50), ("fast", 3, 20), ("fast", 3, 7), ("fast", 3, 9), ("fast", 2, 11),
                                                ("singlestep", 3, 9), ("singlestep", 2, 8), ("singlestep", 1, 4),
                                                ("multistep", 3, 20), ("multistep", 2, 10), ("multistep", 1, 5)])
@pytest.mark.parametrize("mask_opt", [True, False])
@pytest.mark.parametrize("skip_type", ["time_uniform", "logSNR", "t2"])
def test_c_planner_matches_host_planner(method, order, steps, mask_opt, skip_type):
    """pdm_solver_plan (C ABI host planner, csrc/plan.cu) against the Python host planner, whose float32 torch ops are what the
    reference evaluates.  Structure (stages, flags, orders, record kinds) is identical; every coefficient agrees to <= 4e-6
    relative (+ 2e-7 absolute; 3e-5 / 3e-4 for the ill-conditioned difference-term coefficients): the residue is libm-vs-SLEEF rounding of exp / log / log1p / expm1 and torch.linspace's
    per-SIMD-chunk rounding, amplified where a coefficient is a difference of nearby numbers."""
    from panopticdiffusionmodels_b200 import dpm_solver_pp as P
    from panopticdiffusionmodels_b200.multistep import build_multistep_plan
    from panopticdiffusionmodels_b200.sampling import stable_diffusion_beta_schedule
    betas = torch.tensor(stable_diffusion_beta_schedule()).float()
    ns = P.NoiseScheduleVP("discrete", betas=betas)
    if method == "multistep":
        if not mask_opt:
            pytest.skip("multistep has no pass-through variant")
        want = build_multistep_plan(ns, steps, order, 1e-3, 1.0, skip_type)
    else:
        want = P.build_plan(ns, steps, order, 1e-3, 1.0, skip_type, method, mask_opt=mask_opt)
    got = P.build_plan_c(betas, steps, order, 1e-3, 1.0, skip_type, method, mask_opt=mask_opt)
    assert got.shape == want.shape and got.dtype == np.float32
    flags = [8, 9, 10, 11, 12, 15] if method != "multistep" else [11, 15]
    assert np.array_equal(got[:, flags], want[:, flags])
    assert np.array_equal(got[:, 13:15], want[:, 13:15])
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    tol = 4e-6 * np.abs(want.astype(np.float64)) + 2e-7
    if method != "multistep":
        # the difference-term coefficients carry phi_22 = expm1(-r2 h) / (r2 h) + 1 and phi_2 = expm1(-h) / h + 1, which cancel
        # to O(h): an ulp of expm1 is amplified by ~2 / h (small steps of the 'logSNR' / 't2' grids)
        tol[:, [5, 7]] = 3e-5 * np.abs(want[:, [5, 7]].astype(np.float64)) + 2e-7
    if method == "multistep":
        # the 3M coefficients (e^-h - 1) / h + 1 and (e^-h - 1 + h) / h^2 - 1/2 cancel to O(h): a float32 ulp of e^-h is
        # amplified by 1 / h .. 1 / h^2 -- in the reference's own arithmetic as much as here
        tol[:, 5:7] = 3e-4 * np.abs(want[:, 5:7].astype(np.float64)) + 2e-7
    assert (err <= tol).all(), (float((err / (np.abs(want) + 1e-30)).max()), np.argwhere(err > tol)[:5])


def test_c_planner_errors():
    from panopticdiffusionmodels_b200 import dpm_solver_pp as P
    from panopticdiffusionmodels_b200.sampling import stable_diffusion_beta_schedule
    betas = torch.tensor(stable_diffusion_beta_schedule()).float()
    with pytest.raises(RuntimeError, match="order"):
        P.build_plan_c(betas, 10, 4, method="singlestep")
    with pytest.raises(RuntimeError, match="steps >= order"):
        P.build_plan_c(betas, 2, 3, method="multistep")
