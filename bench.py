#!/usr/bin/env python
"""bench.py -- joint image+mask samples/sec (DPM-Solver++ 50 NFE, CFG) on N B200s.

A "step" is one full pass of the hot path over one synthetic batch: a complete 50-NFE classifier-free-
guided joint sample of `--batch` images+masks per GPU (= 100 U-ViT forwards on 2B rows + 50 fused
CFG/solver updates), VAE / CLIP excluded (they run once per sample outside the loop).

  python bench.py --gpus 1 --steps 3 --warmup 3                  # our arm (libpdm, bf16)
  python bench.py --impl reference --steps 1 --warmup 1          # CPU arm: oracle port on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              # weak scaling: batch per GPU fixed

One JSON line on stdout (rank 0).  `value` is BASELINE config 2 (mscoco_uvit_small as shipped, batch 256 per GPU); the
`configs` object of the same line carries short runs (3 warm-ups + 3 timed steps) of the other BASELINE configs --
large (U-ViT-L/2, batch 128 per GPU: the north-star headline), mid (global batch 512 sharded over the GPUs: strong scaling)
and small_512 (batch 32 per GPU, attention-bound) -- each with its own samples/s, roofline (GEMM and attention) and clocks.
Baselines beside it, all outside the timed regions: `cpu_baseline` (oracle port on the host cores, rank 0, N = 1) and
`torch_eager_b200` (the reference's network in eager PyTorch on this same GPU, fp32 and fp16 autocast).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "joint image+mask samples/sec (DPM-Solver++ 50 steps)"
CONFIG_NAMES = {"small": "mscoco_uvit_small", "mid": "mscoco_uvit_mid", "large": "mscoco_uvit_large",
                "small_512": "mscoco_uvit_small_512"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=list(CONFIG_NAMES),
                    help="headline workload (default: small = BASELINE config 2, plus short runs of the others)")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (default: BASELINE config)")
    ap.add_argument("--nfe", type=int, default=50)
    ap.add_argument("--scale", type=float, default=2.0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--single-stream", action="store_true", help="override separate=True of the shipped small config")
    ap.add_argument("--method", default="fast", choices=["fast", "multistep"],
                    help="solver driver: 'fast' (singlestep orders 3..3,2 -- the live path) or 'multistep' (DPM-Solver++ 3M)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-this-GPU baseline leg")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the short runs of the other BASELINE configs (large / mid / small_512) in the `configs` object")
    return ap.parse_args()


def default_batch(config: str, gpus: int) -> int:
    # BASELINE.json configs: small 256/GPU; mid 512 global; large 1024 global over 8; small_512 256 global over 8
    return {"small": 256, "mid": max(64, 512 // max(gpus, 2)), "large": 128, "small_512": 32}[config]


def nnet_kwargs(config: str, single_stream: bool = False):
    from panopticdiffusionmodels_b200 import configs
    cfg = configs.get_config(CONFIG_NAMES[config])
    kw = dict(cfg.nnet)
    kw.pop("name")
    if config == "mid":
        kw["enable_panoptic"] = True  # BASELINE config 3 is the joint model
    if single_stream:
        kw["separate"] = False
    return cfg, kw


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def gemm_traffic(config: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` capture OF THIS CONFIG (profiles/*_gemm_traffic_<config>.json, written by tools/summarise_profiles.py;
    the round-1 captures, all of the small config, are named *_gemm_traffic.json); None when no capture of the config is
    committed."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_gemm_traffic_{config}.json")))
    if not cands and config == "small":
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json")))
    if not cands:
        return None
    return json.load(open(cands[-1])).get("dram_bytes_per_launch_mean")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit())}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_eval_rate(kw, nfe, scale, steps=1, warmup=1, batch=4):
    """Oracle port (oracle/*.py = CPU restatement of the reference path) on the host cores.
    Bounded sample: `steps` CFG model evaluations (2 forwards each) at batch `batch` (4 = BASELINE config 1); a full sample
    costs `nfe` such evaluations, so samples/s = batch / (nfe * t_eval)."""
    import torch
    from oracle import dpm_oracle
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1234)
    net = UViT(**kw)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    S = kw["img_size"]
    g = torch.Generator().manual_seed(1234)
    x, m = torch.randn(batch, 4, S, S, generator=g), torch.randn(batch, 8, S, S, generator=g)
    ctx, empty = torch.randn(batch, 77, 768, generator=g), torch.randn(77, 768, generator=g)
    model = dpm_oracle.cfg_model(sd, kw, ctx, empty, scale)
    t = torch.tensor(0.5)
    with torch.no_grad():
        for _ in range(warmup):
            model(x, t, m)
        t0 = time.perf_counter()
        for _ in range(steps):
            model(x, t, m)
        dt = (time.perf_counter() - t0) / steps
    return dict(value=batch / (nfe * dt), unit="samples/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} CFG model evaluation(s) (2 U-ViT forwards each) at batch {batch}, fp32, extrapolated x{nfe} evals/sample",
                s_per_eval=dt)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config = a.config or "small"
    cfg, kw = nnet_kwargs(config, a.single_stream)
    B = a.batch or default_batch(config, a.gpus)
    cb = cpu_eval_rate(kw, a.nfe, a.scale, steps=max(1, a.steps), warmup=max(1, min(a.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["s_per_eval"] * a.nfe * 1e3 * (B / 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload(config, kw, B, a), "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload(config, kw, B, a, scaling="weak"):
    topo = "two-stream (separate=True, as shipped)" if kw.get("separate") else "single-stream"
    solver = "DPM-Solver++ fast order 3" if a.method == "fast" else "DPM-Solver++ multistep 3M"
    return {"workload": f"{CONFIG_NAMES[config]} {topo} U-ViT D={kw['embed_dim']} depth={kw['depth']}, "
                        f"{kw['img_size']}x{kw['img_size']}x4 latent + 8-bit mask, batch {B}/GPU, {solver}, "
                        f"{a.nfe} NFE, CFG scale {a.scale}", "batch_per_gpu": B, "nfe": a.nfe, "cfg_scale": a.scale,
            "precision": a.precision, "l2": "working set (GBs of activations) larger than L2; no flush needed",
            "weights": "random-init (reference init, seed 1234; zero-conv bridges randomised)"}


# ------------------------------------------------------------------------------------------------ eager-PyTorch-on-GPU leg
def _reference_uvit():
    """The UNMODIFIED reference network class, if its three source files were staged under the git-ignored baseline/_ref/
    (by __graft_entry__.build() in the dev container; they travel to the GPU box with the snapshot)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "libs", "uvit_t2i.py")):
        return None
    try:
        sys.path.insert(0, ref)
        from libs.uvit_t2i import UViT as RefUViT  # noqa: E402
        return RefUViT
    except Exception:
        return None
    finally:
        if sys.path and sys.path[0] == ref:
            sys.path.pop(0)


def eager_torch_rate(kw, nfe, scale, dev, batch=32, evals=3):
    """The like-for-like "beat this" number: the reference's network evaluated by eager PyTorch
    on THIS GPU -- fp32 (allow_tf32 = False: the parity setting) and under fp16 autocast (the authors' launch mode,
    run_commands.sh:37-38) -- as cfg_nnet runs it (two forwards per model evaluation, train_t2i_discrete.py:387-439), on a
    bounded sample: `evals` CFG evaluations at batch `batch`, CUDA-event timed, extrapolated x nfe.  The network is the
    reference's own `libs/uvit_t2i.UViT` (nn.Linear / nn.LayerNorm / SDPA) when baseline/_ref/ holds it (`kind:
    "reference"`), else the oracle port running on the device (`kind: "port"`: plain unfused torch ops, slower than the
    reference would be).  Baseline only, outside every timed region, never on the product path."""
    import torch
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    torch.manual_seed(1234)
    net = UViT(**kw)
    sd = {k: v.detach().clone().to(dev) for k, v in net.state_dict().items()}
    del net
    S = kw["img_size"]
    g = torch.Generator().manual_seed(1234)
    x, m = torch.randn(batch, 4, S, S, generator=g).to(dev), torch.randn(batch, 8, S, S, generator=g).to(dev)
    ctx, empty = torch.randn(batch, 77, 768, generator=g).to(dev), torch.randn(77, 768, generator=g).to(dev)
    t = torch.tensor(0.5, device=dev)
    RefUViT = _reference_uvit()
    if RefUViT is not None:
        ref = RefUViT(**{k: v for k, v in kw.items() if k != "patch_factor"}).to(dev).eval()
        ref.load_state_dict(sd, strict=True)
        ec = empty.unsqueeze(0).expand(batch, -1, -1)

        def model(xx, tc, mm):
            tt = torch.ones(batch, device=dev) * tc * 1000
            c, pc = ref(xx, tt, context=ctx, mask_token=mm)
            u, pu = ref(xx, tt, context=ec, mask_token=mm)
            return c + scale * (c - u), pc + scale * (pc - pu)
        kind = "reference"
        what = "the reference's libs/uvit_t2i.UViT (staged in baseline/_ref)"
    else:
        from oracle import dpm_oracle
        model = dpm_oracle.cfg_model(sd, kw, ctx, empty, scale)
        kind = "port"
        what = "eager torch ops of the oracle port on the device"
    out = {}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        for name, ctxmgr in (("fp32", torch.autocast("cuda", enabled=False)), ("fp16_autocast", torch.autocast("cuda", dtype=torch.float16))):
            with torch.no_grad(), ctxmgr:
                model(x, t, m)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(evals):
                    model(x, t, m)
                e1.record()
                torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) / 1e3 / evals
            out[name] = {"samples_per_s": round(batch / (nfe * dt), 3), "ms_per_cfg_eval": round(dt * 1e3, 2)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    out.update(batch=batch, kind=kind, sample=f"{evals} CFG model evaluations (2 U-ViT forwards each) at batch {batch}, "
               f"extrapolated x{nfe} evals/sample; {what}")
    del sd, model
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ our arm
def vae_decode_rate(dev, batch=32, latent=32, iters=3):
    """SURVEY 8(f) row 3, the step after the loop: latents -> 256 px images through `pdm_vae_decode` (SD decoder layout, random
    weights, batch 32).  Outside every timed region of the headline metric, which excludes the VAE like the reference's
    throughput does; reported so that the row has a driver-run number."""
    import torch
    from panopticdiffusionmodels_b200.libs.autoencoder import get_model
    torch.manual_seed(0)
    vae = get_model(None, 0.23010).to(dev)
    z = torch.randn(batch, 4, latent, latent, device=dev)
    vae.decode(z, max_batch=batch)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        img = vae.decode(z, max_batch=batch)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    ok = bool(torch.isfinite(img).all())
    del vae, z, img
    torch.cuda.empty_cache()
    return {"images_per_s": round(batch / ms * 1e3, 1), "ms_per_batch": round(ms, 2), "batch": batch, "image_px": latent * 8,
            "finite": ok, "note": "pdm_vae_decode (csrc/vae.cu), random weights; not part of `value`"}


def gemm_attn_flops(kw, B, nfe):
    """Algorithmic FLOPs of one step on one GPU, split into the GEMM kernels and the attention kernel."""
    D, depth, P = kw["embed_dim"], kw["depth"], (kw["img_size"] // 2) ** 2
    two = bool(kw.get("separate"))
    Ls = [78 + P, 78 + 2 * P] if two else [78 + 2 * P]
    gemm = sum((depth + 1) * 24 * Lx * D * D + (depth // 2) * 4 * Lx * D * D for Lx in Ls)
    if two:
        gemm += (depth + 1) * 2 * (78 + P) * D * D
    attn = sum((depth + 1) * 4 * Lx * Lx * D for Lx in Ls)
    return gemm * 2 * B * nfe, attn * 2 * B * nfe


def measure(a, config, B, steps, warmup, rank, world, dev, with_e2e, with_profile, scaling="weak"):
    """One workload on this process group: `warmup` untimed steps, `steps` timed ones (barrier + synchronize on both sides,
    CUDA events, max over ranks), optionally the end-to-end variant and one extra eager step with events around every
    launch (rank 0) for the per-kernel table."""
    import torch
    import torch.distributed as dist
    from panopticdiffusionmodels_b200 import _lib
    from panopticdiffusionmodels_b200.distributed import gather_samples
    from panopticdiffusionmodels_b200.flops import flops_per_forward
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    from panopticdiffusionmodels_b200.sampling import JointSampler

    cfg, kw = nnet_kwargs(config, a.single_stream)
    S = kw["img_size"]
    torch.manual_seed(1234)
    net = UViT(**kw)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.startswith("zero_convs"):
                torch.nn.init.trunc_normal_(p, std=0.02)
    net = net.to(dev).eval()
    net.precision = a.precision
    js = JointSampler(net, z_shape=(4, S, S), mask_channels=8, scale=a.scale, cfg=True, sample_steps=a.nfe, method=a.method)

    g = torch.Generator().manual_seed(1234 + rank)
    pin = lambda *s: torch.randn(*s, generator=g).pin_memory()  # noqa: E731
    h_ctx, h_empty, h_z, h_m = pin(B, 77, 768), pin(77, 768), pin(B, 4, S, S), pin(B, 8, S, S)
    d_ctx, d_empty, d_z, d_m = (t.to(dev) for t in (h_ctx, h_empty, h_z, h_m))
    out_host_z = torch.empty(B, 4, S, S).pin_memory()
    out_host_m = torch.empty(B, 8, S, S).pin_memory()

    def step_resident():
        z, pm = js.sample(d_ctx, d_empty, z_init=d_z, mask_init=d_m)
        return gather_samples(z, pm)  # the path's one exchange: NCCL all-gather of finished latents + mask predictions

    def step_e2e():
        c, e = h_ctx.to(dev, non_blocking=True), h_empty.to(dev, non_blocking=True)
        z0, m0 = h_z.to(dev, non_blocking=True), h_m.to(dev, non_blocking=True)
        z, pm = js.sample(c, e, z_init=z0, mask_init=m0)
        gz, gm = gather_samples(z, pm)
        out_host_z.copy_(gz[rank * B:(rank + 1) * B], non_blocking=True)
        out_host_m.copy_(gm[rank * B:(rank + 1) * B], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_count() - l0

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(dev.index) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(step_resident, steps)
    clocks = sampler.summary() if sampler else None
    ms_e2e = None
    if with_e2e:
        step_e2e()
        ms_e2e, _ = timed(step_e2e, steps)

    samples = world * B * steps
    F = flops_per_forward(dict(kw, clip_dim=768), with_mask=True)   # per sample per forward
    flops_step = 2 * a.nfe * F * B                                   # per GPU per step (cond + uncond)
    pk = peaks()
    res = {
        "value": round(samples / (ms / 1e3), 3), "ms_per_step": round(ms / steps, 2), "steps": steps, "warmup": warmup,
        "scaling": scaling, "config": workload(config, kw, B, a, scaling),
        "ms_per_nnet_step": round(ms / steps / (2 * a.nfe), 3), "ms_per_cfg_eval": round(ms / steps / a.nfe, 3),
        "model_tflops_per_gpu": round(flops_step / (ms / steps / 1e3) / 1e12, 1),
        "frac_of_bf16_peak": round(flops_step / (ms / steps / 1e3) / 1e12 / pk["tf_sust"], 4),
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if ms_e2e is not None:
        h2d = (h_ctx.numel() + h_empty.numel() + h_z.numel() + h_m.numel()) * 4
        d2h = (out_host_z.numel() + out_host_m.numel()) * 4
        res["e2e"] = {"value": round(samples / (ms_e2e / 1e3), 3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h}

    # ---- per-kernel timing of one extra eager step with CUDA events around every launch (rank 0) ----
    if rank == 0 and with_profile:
        import ctypes as C
        L = _lib.lib()
        h = net.engine()
        _lib.check(L.pdm_set_profiling(h, 1))
        js.sample(d_ctx, d_empty, z_init=d_z, mask_init=d_m)
        buf_ms = (C.c_float * 64)()
        names = C.create_string_buffer(4096)
        cnt = C.c_int32(0)
        _lib.check(L.pdm_get_profile(h, buf_ms, 64, names, 4096, C.byref(cnt)))
        _lib.check(L.pdm_set_profiling(h, 0))
        kern = {}
        for i, nm in enumerate(names.value.decode().strip().split("\n")[:cnt.value]):
            n, c = nm.rsplit(":", 1)
            kern[n] = {"launches": int(c), "total_ms": round(buf_ms[i], 3), "avg_ms": round(buf_ms[i] / int(c), 4)}
        tot = sum(v["total_ms"] for v in kern.values())
        for v in kern.values():
            v["share"] = round(v["total_ms"] / tot, 4)
        gem = {k: v for k, v in kern.items() if k.startswith("gemm_")}
        gemm_flops, attn_flops = gemm_attn_flops(kw, B, a.nfe)
        g_ms = sum(v["total_ms"] for v in gem.values())
        g_n = sum(v["launches"] for v in gem.values())
        ach = gemm_flops / (g_ms / 1e3) / 1e12
        res["roofline"] = {
            "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM: qkv/proj/fc1/fc2/skip/zero-conv)", "bound": "tensor",
            "achieved": round(ach, 1), "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_sust"], 4),
            "peak_src": pk["src"] + " (sustained bf16)", "launches": g_n, "avg_launch_ms": round(g_ms / g_n, 4),
            "flops_per_launch": gemm_flops / g_n, "share_of_step": round(g_ms / tot, 4), "traffic": gemm_traffic(config)}
        if "attention" in kern:
            at = kern["attention"]
            ach_a = attn_flops / (at["total_ms"] / 1e3) / 1e12
            res["roofline_attention"] = {
                "kernel": "attention_tc3_kernel (tcgen05 flash attention, head dim 64)", "bound": "tensor",
                "achieved": round(ach_a, 1), "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": round(ach_a / pk["tf_sust"], 4),
                "peak_src": pk["src"] + " (sustained bf16)", "launches": at["launches"], "avg_launch_ms": at["avg_ms"],
                "flops_per_launch": attn_flops / at["launches"], "share_of_step": at["share"], "traffic": None,
                "note": "exp2 on the MUFU pipe (16/clk/SM) bounds head-dim-64 attention at <= 50 % of the tensor peak"}
        res["kernels"] = kern
    del js, net
    torch.cuda.empty_cache()
    return res, kw


def run_ours(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # keep stdout to the ONE JSON line: NCCL writes its version banner to the C-level stdout whenever NCCL_DEBUG >= VERSION
    # (WARN included), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved one
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    config = a.config or "small"
    B = a.batch or default_batch(config, a.gpus)
    head, kw = measure(a, config, B, a.steps, max(a.warmup, 3), rank, world, dev, with_e2e=True,
                       with_profile=not a.no_kernel_profile)

    # ---- the other BASELINE configs, short runs in the same line (default invocation only) ----
    extra = {}
    if a.config is None and not a.no_extra_configs and a.method == "fast" and not a.batch:
        for name, scaling in (("large", "weak"), ("mid", "strong"), ("small_512", "weak")):
            Bx = default_batch(name, a.gpus) if name != "mid" else max(64, 512 // world)
            r, _ = measure(a, name, Bx, 3, 3, rank, world, dev, with_e2e=False, with_profile=not a.no_kernel_profile,
                           scaling=scaling)
            r.pop("kernels", None)
            extra[name] = {"samples_per_s": r.pop("value"), "unit": "samples/s", "global_batch": Bx * world, **r}

    if rank == 0:
        # the CPU and eager-GPU baselines are rank-0, N = 1 measurements (under torchrun the host cores are shared by all ranks
        # and OMP_NUM_THREADS is pinned to 1: the CPU number would be meaningless), outside every timed region
        cb = None if (a.no_cpu_baseline or world > 1) else cpu_eval_rate(kw, a.nfe, a.scale, steps=1, warmup=1)
        eager = None if (a.no_eager_baseline or world > 1) else eager_torch_rate(kw, a.nfe, a.scale, dev)
        line = {
            "metric": METRIC, "value": head.pop("value"), "unit": "samples/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": head.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        }
        head.pop("steps"); head.pop("warmup"); head.pop("scaling")
        line.update(head)
        line["cpu_baseline"] = None if cb is None else {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["torch_eager_b200"] = eager
        if extra:
            line["configs"] = extra
            line["vae_decode"] = vae_decode_rate(dev) if world == 1 else None
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    os.close(json_fd)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
