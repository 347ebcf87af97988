"""Edge cases of the hot path on the B200 (through the C ABI): ragged and single-sample batches, an empty batch, workspace
growth and reuse, non-default streams, bit reproducibility, shared timestep scalars, and the error behaviour of the C entry
points (status code + ``pdm_last_error`` instead of a crash).  The reference has no tests of its own; these
follow the shapes its call sites can produce: ``amortize`` hands the sampler a ragged last mini-batch (``utils.py:452-455``),
the per-rank share of a short prompt list can be empty (``sample_t2i_discrete.py:70``), and ``model_fn`` passes one scalar
time for the whole batch (``dpm_solver_pp.py:310-328``).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import TINY

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    DEV = torch.device("cuda:0")


def maxrel(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def cosine(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm()))


def make_net(separate, precision, cfg=TINY, seed=0):
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    torch.manual_seed(seed)
    net = UViT(separate=separate, **cfg)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.startswith("zero_convs") or k.endswith(".bias"):
                p.copy_(torch.randn_like(p) * 0.02)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(DEV).eval()
    net.precision = precision
    return net, sd


def inputs(B, cfg=TINY, seed=7):
    g = torch.Generator().manual_seed(seed)
    s = cfg["img_size"]
    x = torch.randn(B, cfg["in_chans"], s, s, generator=g)
    m = torch.randn(B, cfg["num_panoptic_class"], s, s, generator=g)
    ctx = torch.randn(B, cfg["num_clip_token"], cfg["clip_dim"], generator=g)
    t = torch.rand(B, generator=g) * 999 + 1
    return x, m, ctx, t


@pytest.mark.parametrize("B", [1, 3, 5, 17])
@pytest.mark.parametrize("separate", [False, True])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ragged_batches_vs_oracle(B, separate, precision):
    """Row counts that are not a multiple of any tile (B * L = 21 ... 629 rows) against the oracle, every sample checked on
    its own so that a fault in the last partial tile cannot hide behind the batch maximum."""
    from oracle import uvit_oracle
    net, sd = make_net(separate, precision)
    x, m, ctx, t = inputs(B)
    kw = dict(TINY, separate=separate)
    ref_n, ref_y = uvit_oracle.uvit_forward(sd, kw, x, t, ctx, m)
    noise, y = net(x.to(DEV), t.to(DEV), ctx.to(DEV), mask_token=m.to(DEV))
    for b in range(B):
        if precision == "fp32":
            assert maxrel(noise[b], ref_n[b]) <= 1e-3 and maxrel(y[b], ref_y[b]) <= 1e-3, b
        else:
            assert cosine(noise[b], ref_n[b]) >= 0.999 and cosine(y[b], ref_y[b]) >= 0.998, b


@pytest.mark.parametrize("separate", [False, True])
def test_image_only_call_ragged(separate):
    """``nnet(x, t, ctx)`` without a mask_token (libs/uvit_t2i.py:408-410) on a ragged batch."""
    from oracle import uvit_oracle
    net, sd = make_net(separate, "fp32")
    x, _, ctx, t = inputs(3)
    ref = uvit_oracle.uvit_forward(sd, dict(TINY, separate=separate), x, t, ctx, None)
    ref = ref[0] if isinstance(ref, tuple) else ref
    out = net(x.to(DEV), t.to(DEV), ctx.to(DEV))
    assert maxrel(out, ref) <= 1e-3


def test_empty_batch_is_a_no_op():
    """B = 0 flows through the reference's torch ops as empty tensors; here it launches nothing and returns empty outputs."""
    from panopticdiffusionmodels_b200.sampling import JointSampler
    net, _ = make_net(False, "bf16")
    x, m, ctx, t = inputs(0)
    noise, y = net(x.to(DEV), t.to(DEV), ctx.to(DEV), mask_token=m.to(DEV))
    assert tuple(noise.shape) == (0, 4, 8, 8) and tuple(y.shape) == (0, 8, 8, 8)
    sampler = JointSampler(net, z_shape=(4, 8, 8), scale=2.0, sample_steps=8)
    z, pm = sampler.sample(ctx.to(DEV), torch.randn(5, 32, device=DEV))
    assert tuple(z.shape) == (0, 4, 8, 8) and tuple(pm.shape) == (0, 8, 8, 8)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_workspace_growth_and_reuse_is_bit_stable(precision):
    """The engine's private workspace grows with the largest batch seen; a small batch must give the same bits before and
    after it grew, and a batch must not depend on what ran before it."""
    net, _ = make_net(True, precision)
    xs, ms, cs, ts = (v.to(DEV) for v in inputs(2, seed=1))
    xl, ml, cl, tl = (v.to(DEV) for v in inputs(9, seed=2))
    n0, y0 = net(xs, ts, cs, mask_token=ms)
    n0, y0 = n0.clone(), y0.clone()
    nl, yl = net(xl, tl, cl, mask_token=ml)
    nl, yl = nl.clone(), yl.clone()
    n1, y1 = net(xs, ts, cs, mask_token=ms)
    assert torch.equal(n0, n1) and torch.equal(y0, y1)
    nl2, yl2 = net(xl, tl, cl, mask_token=ml)
    assert torch.equal(nl, nl2) and torch.equal(yl, yl2)
    # the first two samples of the large batch alone: same rows, same bits (no cross-sample op)
    n2, y2 = net(xl[:2].contiguous(), tl[:2].contiguous(), cl[:2].contiguous(), mask_token=ml[:2].contiguous())
    assert torch.equal(n2, nl[:2]) and torch.equal(y2, yl[:2])


def test_precision_switch_on_one_handle():
    """fp32 and bf16 forwards alternate on one engine handle (two workspaces, one parameter set)."""
    net, _ = make_net(False, "fp32")
    x, m, ctx, t = (v.to(DEV) for v in inputs(3))
    a32, b32 = (v.clone() for v in net(x, t, ctx, mask_token=m))
    net.precision = "bf16"
    a16, b16 = (v.clone() for v in net(x, t, ctx, mask_token=m))
    net.precision = "fp32"
    c32, d32 = net(x, t, ctx, mask_token=m)
    assert torch.equal(a32, c32) and torch.equal(b32, d32)
    assert cosine(a16, a32) >= 0.999 and cosine(b16, b32) >= 0.998 and not torch.equal(a16, a32)


def test_side_stream_matches_default_stream():
    """All work is enqueued on the caller's current stream (no hidden default-stream launches, no hidden syncs)."""
    net, _ = make_net(True, "bf16")
    x, m, ctx, t = (v.to(DEV) for v in inputs(4))
    n0, y0 = (v.clone() for v in net(x, t, ctx, mask_token=m))
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side):
        n1, y1 = net(x, t, ctx, mask_token=m)
    side.synchronize()
    assert torch.equal(n0, n1) and torch.equal(y0, y1)


def test_scalar_timestep_broadcasts():
    """``model_fn`` hands the network one time for the whole batch; a 1-element ``timesteps`` broadcasts like the reference's
    ``timesteps.expand`` would (train_t2i_discrete.py:506-516)."""
    net, _ = make_net(False, "fp32")
    x, m, ctx, _ = (v.to(DEV) for v in inputs(3))
    a, b = (v.clone() for v in net(x, torch.tensor([417.25], device=DEV), ctx, mask_token=m))
    c, d = net(x, torch.full((3,), 417.25, device=DEV), ctx, mask_token=m)
    assert torch.equal(a, c) and torch.equal(b, d)


@pytest.mark.parametrize("use_graph", [True, False])
def test_whole_sample_is_bit_reproducible(use_graph):
    """Two runs of the device loop from the same inputs give the same bits (no atomics anywhere on the path), with and
    without the CUDA graph, and the graph replay equals the eager enqueue."""
    from panopticdiffusionmodels_b200.sampling import JointSampler
    net, _ = make_net(True, "bf16")
    g = torch.Generator().manual_seed(3)
    B = 3
    ctx = torch.randn(B, 5, 32, generator=g).to(DEV)
    ec = torch.randn(5, 32, generator=g).to(DEV)
    z0 = torch.randn(B, 4, 8, 8, generator=g).to(DEV)
    m0 = torch.randn(B, 8, 8, 8, generator=g).to(DEV)
    sampler = JointSampler(net, z_shape=(4, 8, 8), scale=2.0, sample_steps=11)
    z1, p1 = (v.clone() for v in sampler.sample(ctx, ec, z0, m0, use_graph=use_graph))
    z2, p2 = sampler.sample(ctx, ec, z0, m0, use_graph=use_graph)
    assert torch.equal(z1, z2) and torch.equal(p1, p2)
    z3, p3 = sampler.sample(ctx, ec, z0, m0, use_graph=not use_graph)
    assert torch.equal(z1, z3) and torch.equal(p1, p3)
    assert torch.isfinite(z1).all() and torch.isfinite(p1).all()


def test_changed_batch_size_between_samples():
    """A graph captured for one batch size must not be replayed for another (the last mini-batch of ``amortize`` is ragged)."""
    from panopticdiffusionmodels_b200.sampling import JointSampler
    net, _ = make_net(False, "bf16")
    sampler = JointSampler(net, z_shape=(4, 8, 8), scale=2.0, sample_steps=8)
    g = torch.Generator().manual_seed(5)
    ec = torch.randn(5, 32, generator=g).to(DEV)
    outs = {}
    for B in (4, 2, 4, 1):
        gg = torch.Generator().manual_seed(100 + B)
        ctx = torch.randn(B, 5, 32, generator=gg).to(DEV)
        z0 = torch.randn(B, 4, 8, 8, generator=gg).to(DEV)
        m0 = torch.randn(B, 8, 8, 8, generator=gg).to(DEV)
        z, pm = (v.clone() for v in sampler.sample(ctx, ec, z0, m0))
        assert torch.isfinite(z).all()
        if B in outs:
            assert torch.equal(outs[B][0], z) and torch.equal(outs[B][1], pm)
        outs[B] = (z, pm)
    # sample 0 of the batch-4 run == the same sample run alone (samples are independent)
    gg = torch.Generator().manual_seed(104)
    ctx = torch.randn(4, 5, 32, generator=gg).to(DEV)
    z0 = torch.randn(4, 4, 8, 8, generator=gg).to(DEV)
    m0 = torch.randn(4, 8, 8, 8, generator=gg).to(DEV)
    z, pm = sampler.sample(ctx[:1].contiguous(), ec, z0[:1].contiguous(), m0[:1].contiguous())
    assert torch.equal(z, outs[4][0][:1]) and torch.equal(pm, outs[4][1][:1])


# ---- error behaviour of the C entry points -----------------------------------------------------------------------------
def _err():
    from panopticdiffusionmodels_b200 import _lib
    return _lib.lib().pdm_last_error().decode()


def test_c_abi_rejects_bad_arguments_with_a_message():
    from panopticdiffusionmodels_b200 import _lib
    L = _lib.lib()
    net, _ = make_net(False, "bf16")
    h = net.engine()
    stream = _lib.current_stream()
    w = torch.zeros(64, 64, device=DEV)
    shape = (C.c_int64 * 2)(64, 64)
    # unknown state_dict key; wrong shape for a known key
    assert L.pdm_set_param(h, b"blocks.0.not_a_key", w.data_ptr(), shape, 2, stream) != 0 and "not_a_key" in _err()
    assert L.pdm_set_param(h, b"in_blocks.0.attn.proj.weight", w.data_ptr(), (C.c_int64 * 2)(64, 32), 2, stream) != 0
    assert "size mismatch" in _err()
    # null pointers, bad precision, bad batch
    x, m, ctx, t = (v.to(DEV) for v in inputs(2))
    out = torch.empty_like(x)
    assert L.pdm_nnet_forward(h, None, _lib.ptr(t), _lib.ptr(ctx), None, _lib.ptr(out), None, 2, 0, stream) != 0
    assert L.pdm_nnet_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(ctx), None, _lib.ptr(out), None, 2, 7, stream) != 0
    assert "precision" in _err()
    assert L.pdm_nnet_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(ctx), None, _lib.ptr(out), None, 0, 0, stream) != 0
    # a mask without an output buffer for it
    assert L.pdm_nnet_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(ctx), _lib.ptr(m), _lib.ptr(out), None, 2, 0, stream) != 0
    assert "out_mask" in _err()
    # unknown forward flag
    assert L.pdm_nnet_forward_ex(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(ctx), None, _lib.ptr(out), None, 2, 0, 0x40, stream) != 0
    # the handle is still usable after every rejected call
    torch.cuda.synchronize()
    good = net(x, t, ctx, mask_token=m)
    assert torch.isfinite(good[0]).all()


def test_c_abi_rejects_malformed_plans():
    from panopticdiffusionmodels_b200 import _lib
    from panopticdiffusionmodels_b200.dpm_solver_pp import NoiseScheduleVP, build_plan
    from panopticdiffusionmodels_b200.sampling import stable_diffusion_beta_schedule
    L = _lib.lib()
    net, _ = make_net(False, "bf16")
    h = net.engine()
    ns = NoiseScheduleVP(schedule="discrete", betas=torch.tensor(stable_diffusion_beta_schedule()).float())
    plan = np.ascontiguousarray(build_plan(ns, steps=8, eps=1e-3, T=1.0, order=3, mask_opt=True), dtype=np.float32)
    g = torch.Generator().manual_seed(0)
    ctx = torch.randn(2, 5, 32, generator=g).to(DEV)
    ec = torch.randn(5, 32, generator=g).to(DEV)
    z0 = torch.randn(2, 4, 8, 8, generator=g).to(DEV)
    m0 = torch.randn(2, 8, 8, 8, generator=g).to(DEV)
    oz, om = torch.empty_like(z0), torch.empty_like(m0)

    def run(p, n):
        return L.pdm_sample(h, p.ctypes.data_as(C.POINTER(C.c_float)), n, _lib.ptr(z0), _lib.ptr(m0), _lib.ptr(ctx), _lib.ptr(ec),
                            2.0, _lib.ptr(oz), _lib.ptr(om), 2, 0, 0, _lib.current_stream())

    assert run(plan, plan.shape[0]) == 0
    torch.cuda.synchronize()
    ref = oz.clone()
    assert run(plan, 0) != 0                                   # no evaluations
    bad = plan.copy(); bad[0, 8] = 2.0                         # the first record of a step must be stage 0
    assert run(bad, bad.shape[0]) != 0 and "stage" in _err()
    bad = plan.copy(); bad[-1, 10] = 0.0                       # the plan must end on a step boundary
    assert run(bad, bad.shape[0]) != 0
    bad = plan.copy(); bad[1, 0] = float("nan")                # non-finite time
    assert run(bad, bad.shape[0]) != 0
    assert run(plan, plan.shape[0]) == 0                       # and the handle still works
    torch.cuda.synchronize()
    assert torch.equal(ref, oz)


@pytest.mark.parametrize("separate", [False, True])
def test_parameter_update_after_first_use(separate):
    """Weights changed in place after an engine (and a captured CUDA graph) exists: the next call re-uploads them, rebuilds the
    derived buffers IN PLACE (bf16 copies, LayerNorm-folded weights, the concatenated fc2 | zero-conv weight of a two-stream
    layer) and the replayed graph computes with the new values -- bit-equal to a fresh module built from the same state."""
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    from panopticdiffusionmodels_b200.sampling import JointSampler
    net, _ = make_net(separate, "bf16")
    g = torch.Generator().manual_seed(11)
    B = 2
    ctx = torch.randn(B, 5, 32, generator=g).to(DEV)
    ec = torch.randn(5, 32, generator=g).to(DEV)
    z0 = torch.randn(B, 4, 8, 8, generator=g).to(DEV)
    m0 = torch.randn(B, 8, 8, 8, generator=g).to(DEV)
    sampler = JointSampler(net, z_shape=(4, 8, 8), scale=2.0, sample_steps=8)
    z_old, p_old = (v.clone() for v in sampler.sample(ctx, ec, z0, m0))
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("mlp.fc2.weight") or k.endswith("norm2.weight") or k.startswith("zero_convs") or k == "pos_embed":
                p.mul_(1.25)
    z_new, p_new = (v.clone() for v in sampler.sample(ctx, ec, z0, m0))
    assert not torch.equal(z_old, z_new)
    fresh = UViT(separate=separate, **TINY)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, strict=True)
    fresh = fresh.to(DEV).eval()
    fresh.precision = "bf16"
    z_ref, p_ref = JointSampler(fresh, z_shape=(4, 8, 8), scale=2.0, sample_steps=8).sample(ctx, ec, z0, m0)
    assert torch.equal(z_new, z_ref) and torch.equal(p_new, p_ref)
