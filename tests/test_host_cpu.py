"""CPU-side checks of the host layer: the C-ABI library loads and exports every symbol include/pdm.h
declares, the drop-in module has the reference state_dict layout, and the product refuses to run
without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, TINY, load_golden


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pdm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from panopticdiffusionmodels_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libpdm.so must be built (python -m panopticdiffusionmodels_b200.build)"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/pdm.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib().pdm_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("name,separate", [("single", False), ("two", True)])
def test_state_dict_layout_matches_reference(name, separate):
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    _, sd = load_golden(f"tiny_{name}.npz")
    net = UViT(separate=separate, **TINY)
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys()) or set(mine.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(mine[k].shape) == tuple(v.shape), k
    net.load_state_dict(sd, strict=True)


def test_reference_init_statistics():
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    torch.manual_seed(0)
    net = UViT(separate=True, **TINY)
    sd = net.state_dict()
    assert float(sd["zero_convs.1.conv.weight"].abs().max()) == 0.0          # ControlNet-style zero bridges
    assert float(sd["in_blocks.0.norm1.weight"].min()) == 1.0
    assert float(sd["in_blocks.0.mlp.fc1.bias"].abs().max()) == 0.0
    w = sd["in_blocks.0.attn.qkv.weight"]
    assert 0.015 < float(w.std()) < 0.025 and float(w.abs().max()) <= 2.0


def test_config_files_load_and_construct():
    from panopticdiffusionmodels_b200 import configs, utils
    for name, D, depth, sep in (("mscoco_uvit_small", 512, 12, True), ("mscoco_uvit_mid", 768, 16, False),
                                ("mscoco_uvit_large", 1024, 20, False), ("mscoco_uvit_small_512", 512, 12, False)):
        cfg = configs.get_config(name)
        assert cfg.nnet.embed_dim == D and cfg.nnet.depth == depth and cfg.sample.sample_steps == 50
        assert tuple(cfg.z_shape)[0] == 4
    cfg = configs.get_config("mscoco_uvit_small")
    assert cfg.nnet.separate is True and cfg.nnet.patch_factor == 2          # accepted and ignored
    with torch.device("meta"):
        net = utils.get_nnet(**cfg.nnet)
    n = sum(p.numel() for p in net.parameters())
    assert abs(n - 95.81e6) < 0.05e6


def test_no_cpu_fallback():
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    from panopticdiffusionmodels_b200 import dpm_solver_pp as P
    net = UViT(separate=False, **TINY)
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        net(x, torch.zeros(1), torch.zeros(1, 5, 32))
    ns = P.NoiseScheduleVP("discrete", betas=torch.linspace(1e-4, 2e-2, 1000))
    with pytest.raises(RuntimeError):
        P.DPM_Solver(lambda *a, **k: None, ns, predict_x0=True).sample(x, steps=6, eps=1e-3, T=1.0)


def test_c_abi_null_arguments_return_a_status_not_a_crash():
    """Every entry point reports bad arguments through its int status + pdm_last_error() before touching the device
    (include/pdm.h: 0 = OK); none of these calls computes anything, so they are safe without a GPU."""
    import ctypes as C
    from panopticdiffusionmodels_b200 import _lib
    L = _lib.lib()
    assert L.pdm_abi_version() == _lib.ABI_VERSION
    h = C.c_void_p()
    assert L.pdm_create(None, C.byref(h)) != 0 and "null" in L.pdm_last_error().decode()
    assert L.pdm_set_param(None, b"pos_embed", None, None, 0, None) != 0
    assert L.pdm_finalize_params(None, None) != 0 and "null" in L.pdm_last_error().decode()
    n = C.c_size_t()
    assert L.pdm_workspace_bytes(None, 4, 0, C.byref(n)) != 0
    assert L.pdm_nnet_forward(None, None, None, None, None, None, None, 1, 0, None) != 0
    assert L.pdm_sample(None, None, 0, None, None, None, None, 1.0, None, None, 1, 0, 1, None) != 0
    assert L.pdm_vae_create(None, C.byref(h)) != 0
    assert L.pdm_vae_decode(None, None, None, 1, 32, None) != 0
    assert L.pdm_destroy(None) == 0  # destroying nothing is fine


def test_rejected_options():
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    with pytest.raises(NotImplementedError):
        UViT(**dict(TINY, mlp_time_embed=True))
    with pytest.raises(TypeError):
        UViT(**dict(TINY, bogus_kwarg=1))


def test_config_table_matches_reference_files():
    """configs.get_config == the reference's configs/mscoco_uvit_*.py key by key (fixture: the flattened tables of the four
    reference files, written by executing them through the ml_collections shim), except the machine-specific paths."""
    import json
    from panopticdiffusionmodels_b200 import configs

    def flat(c, pre=""):
        out = {}
        for k, v in c.items():
            if hasattr(v, "items"):
                out.update(flat(v, pre + k + "."))
            else:
                out[pre + k] = list(v) if isinstance(v, tuple) else v
        return out

    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_configs.json")))
    local_paths = {"dataset.path", "sample.path", "pretrained"}
    for name in configs.NAMES:
        ours, want = flat(configs.get_config(name)), ref[name]
        assert set(ours) == set(want), name
        for k in want:
            if k not in local_paths:
                assert ours[k] == want[k], (name, k, ours[k], want[k])


def test_module_copies_do_not_share_the_engine_handle():
    """copy.deepcopy(nnet) (e.g. an EMA twin), pickling and torch.save of the
    whole module must work after an engine exists, and a copy must never alias the C handle (double pdm_destroy)."""
    import copy
    import ctypes as C
    import pickle
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    net = UViT(separate=False, **TINY)
    net._handle = C.c_void_p(0x1234)          # stand-in for a live engine (no GPU here)
    net._handle_device, net._fingerprint = "cuda:0", ("x",)
    try:
        for dup in (copy.deepcopy(net), copy.copy(net), pickle.loads(pickle.dumps(net))):
            assert dup._handle is None and dup._handle_device is None and dup._fingerprint is None
            assert set(dup.state_dict()) == set(net.state_dict())
    finally:
        net._handle = None                    # do not hand the fake pointer to pdm_destroy


def test_vae_state_dict_layout_matches_reference():
    """libs.autoencoder.FrozenAutoencoderKL holds exactly the reference module's parameters (keys and shapes recorded from the
    reference's own state_dict by tests/golden/make_vae.py, whose get_model also loaded ours with its no-missing /
    no-unexpected assertion) and refuses to run without a CUDA device."""
    import json
    from panopticdiffusionmodels_b200.libs.autoencoder import get_model
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "vae_keys.json")))
    vae = get_model(None, 0.23010)
    mine = {k: list(v.shape) for k, v in vae.state_dict().items()}
    assert mine == want
    with pytest.raises(RuntimeError, match="no CPU path"):
        vae.decode(torch.zeros(1, 4, 16, 16))
    with pytest.raises(NotImplementedError):
        vae.encode(torch.zeros(1, 3, 128, 128))
