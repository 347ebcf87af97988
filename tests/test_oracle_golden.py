"""Pins the oracle (oracle/*.py) to the reference: golden fixtures made by
executing /root/reference (tests/golden/make_golden.py)."""
import numpy as np
import os
import sys

import pytest
import torch

from conftest import TINY, load_golden
from oracle import dpm_oracle, uvit_oracle


def test_interp_docstring_kat():
    # dpm_solver_pp.py:21-24
    xp, yp = torch.tensor([0.0, 1.0]), torch.tensor([0.0, 2.0])
    assert float(dpm_oracle._pwl(torch.tensor(0.5), xp, yp)) == 1.0
    assert float(dpm_oracle._pwl(torch.tensor(-10.0), xp, yp)) == -20.0


def test_schedule_bit_exact():
    g, _ = load_golden("schedule.npz")
    s = dpm_oracle.Schedule()
    for i, t in enumerate(g["t"]):
        assert float(s.log_mean(t)) == float(g["log_alpha"][i])
        assert float(s.sigma(t)) == float(g["sigma"][i])
        assert float(s.lam(t)) == float(g["lam"][i])
        assert float(s.inv_lam(g["lam"][i])) == float(g["inv_lam"][i])


@pytest.mark.parametrize("name,separate", [("single", False), ("two", True)])
def test_forward_matches_reference(name, separate):
    g, sd = load_golden(f"tiny_{name}.npz")
    cfg = dict(TINY, separate=separate)
    noise, y = uvit_oracle.uvit_forward(sd, cfg, g["x"], g["t"], g["ctx"], g["m"])
    assert torch.allclose(noise, g["noise"], rtol=0, atol=2e-6)
    assert torch.allclose(y, g["y"], rtol=0, atol=2e-6)
    n2 = uvit_oracle.uvit_forward(sd, cfg, g["x"], g["t"], g["ctx"], None)
    assert torch.allclose(n2, g["noise_nomask"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name,separate", [("single", False), ("two", True)])
@pytest.mark.parametrize("steps", [20, 7, 9])
def test_joint_sample_matches_reference(name, separate, steps):
    g, sd = load_golden(f"tiny_{name}.npz")
    cfg = dict(TINY, separate=separate)
    z, pm = dpm_oracle.joint_sample(sd, cfg, g["x"], g["m"], g["ctx"], g["empty"], float(g["scale"]), steps)
    ref_z, ref_pm = g[f"z{steps}"], g[f"pm{steps}"]
    assert (z - ref_z).abs().max() <= 2e-4 * ref_z.abs().max()
    assert (pm - ref_pm).abs().max() <= 2e-4


@pytest.mark.parametrize("name,separate", [("single", False), ("two", True)])
def test_ground_truth_forward_matches_reference(name, separate):
    """use_ground_truth=True evaluation (libs/uvit_t2i.py:486-496): fixture = the real reference's output."""
    g, sd = load_golden(f"tiny_{name}.npz")
    cfg = dict(TINY, separate=separate)
    noise, y = uvit_oracle.uvit_forward(sd, cfg, g["x"], g["t"], g["ctx"], g["m"], use_ground_truth=True)
    assert torch.allclose(noise, g["noise_gt"], rtol=0, atol=2e-6)
    assert torch.equal(y, g["m"])


@pytest.mark.parametrize("name,separate", [("single", False), ("two", True)])
@pytest.mark.parametrize("order,steps", [(3, 9), (2, 8), (1, 4)])
def test_singlestep_and_two_phase_match_reference(name, separate, order, steps):
    """method='singlestep' with and without use_twophases (dpm_solver_pp.py:1045-1078), fixtures from the reference."""
    g, sd = load_golden(f"tiny_{name}.npz")
    cfg = dict(TINY, separate=separate)
    model = dpm_oracle.cfg_model(sd, cfg, g["ctx"], g["empty"], float(g["scale"]))
    for two, tag in ((False, "ss"), (True, "tp")):
        z, pm = dpm_oracle.Solver(model, dpm_oracle.Schedule()).sample_singlestep(g["x"], g["m"], steps, order, two_phases=two)
        ref_z, ref_pm = g[f"z_{tag}{order}"], g[f"pm_{tag}{order}"]
        assert (z - ref_z).abs().max() <= 2e-4 * ref_z.abs().max(), (tag, order)
        assert (pm - ref_pm).abs().max() <= 2e-4, (tag, order)


def test_config1_joint_sample_matches_reference():
    """BASELINE config 1 (mscoco_uvit_small as shipped, random-init, batch 4, DPM-Solver++ 20 NFE, CFG 2.0, fp32 on CPU):
    the oracle against the REAL reference's output (tests/golden/make_config1.py).  ~1 minute of CPU."""
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_config1 import build_ours, inputs
    net, kw = build_ours()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x, m, ctx, empty = inputs()
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_small_two.npz"))
    torch.set_num_threads(os.cpu_count())
    z, pm = dpm_oracle.joint_sample(sd, kw, x, m, ctx, empty, float(ref["scale"]), int(ref["steps"]))
    ref_z, ref_pm = torch.from_numpy(ref["z"]), torch.from_numpy(ref["pm"])
    assert (z - ref_z).abs().max() <= 5e-4 * ref_z.abs().max()
    assert (pm - ref_pm).abs().max() <= 5e-4


def test_multistep_bit_exact():
    g, _ = load_golden("multistep.npz")
    s = dpm_oracle.Solver(None, dpm_oracle.Schedule())
    t = [g["t"][i] for i in range(4)]
    m2 = s.multistep_second(g["x"], [g["X1"], g["X0"]], [t[1], t[2]], t[3])
    m3 = s.multistep_third(g["x"], [g["X2"], g["X1"], g["X0"]], [t[0], t[1], t[2]], t[3])
    assert torch.equal(m2, g["m2"])
    assert torch.equal(m3, g["m3"])


def test_bits_roundtrip():
    ids = torch.arange(256).reshape(1, 1, 16, 16)
    bits = dpm_oracle.int2bits(ids)
    assert bits.shape == (1, 8, 16, 16)
    assert int(bits[0, 0, 15, 15]) == 1 and int(bits[0, 7, 0, 1]) == 1  # MSB first
    assert torch.equal(dpm_oracle.bits2int((bits * 2.0 - 1.0) > 0).long(), ids)


def test_codec_matches_reference():
    """oracle int2bits / bits2int against the outputs of the reference's utils.int2bits / utils.bits2int (utils.py:475-518;
    tests/golden/make_round2.py codec): all 256 ids, random bit patterns, sign-thresholded analog bits incl. exact zeros."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "codec.npz"))
    ids = torch.from_numpy(z["ids"])
    assert torch.equal(dpm_oracle.int2bits(ids).to(torch.int32), torch.from_numpy(z["bits"]).to(torch.int32))
    assert torch.equal(dpm_oracle.bits2int(torch.from_numpy(z["pattern"])), torch.from_numpy(z["labels"]))
    assert torch.equal(dpm_oracle.bits2int(torch.from_numpy(z["analog"]) > 0), torch.from_numpy(z["labels_analog"]))


@pytest.mark.parametrize("tag", ["mid", "large", "small_512"])
def test_full_depth_forward_matches_reference(tag):
    """BASELINE configs 3 / 4 / 5 at full depth and real geometry (batch 1): the oracle against the REAL reference forward
    (tests/golden/make_round2.py forwards).  A few seconds of CPU each."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_round2 import build_model, fwd_inputs
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "configs_fwd.npz"))
    net, kw = build_model(tag)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x, m, ctx, t = fwd_inputs(kw)
    torch.set_num_threads(os.cpu_count())
    noise, y = uvit_oracle.uvit_forward(sd, kw, x, t, ctx, m)
    rn, ry = torch.from_numpy(ref[f"{tag}/noise"]), torch.from_numpy(ref[f"{tag}/y"])
    assert (noise - rn).abs().max() <= 2e-5 * rn.abs().max(), float((noise - rn).abs().max() / rn.abs().max())
    assert (y - ry).abs().max() <= 2e-5 * max(1.0, float(ry.abs().max()))


def test_vae_decode_matches_reference():
    """oracle/vae_oracle.py against the REAL reference's FrozenAutoencoderKL.decode (tests/golden/make_vae.py): SD autoencoder
    layout, random weights rebuilt from the seed, 2 latents of 16 x 16 -> 128 x 128 px."""
    from oracle import vae_oracle
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_vae import SCALE, build_vae, latents
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae_decode.npz"))
    vae = build_vae()
    sd = {k: v.detach().clone() for k, v in vae.state_dict().items()}
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = vae_oracle.vae_decode(sd, vae.ddconfig, latents(), SCALE)
    want = torch.from_numpy(ref["out"])
    assert out.shape == want.shape
    assert (out - want).abs().max() <= 2e-5 * want.abs().max()
