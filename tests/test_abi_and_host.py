"""CPU-side checks: the C-ABI library loads and exports every symbol include/spinnerf_b200.h declares,
host logic of the drop-in module, flat parameter storage, the config shim.  No compute calls."""
import ctypes
import importlib
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import nerf_oracle as O

spn = importlib.import_module("spin-nerf_b200")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "spinnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 25
    l = ctypes.CDLL(spn.LIB_PATH)
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/spinnerf_b200.h but not exported"
    assert set(spn._lib.EXPORTS) == set(syms), set(spn._lib.EXPORTS) ^ set(syms)
    assert l.spn_version() == 100


def test_param_offsets_match_module_layout():
    off = spn._lib.param_offsets()
    assert off[-1] == spn.MLP_NPARAMS == 595844
    sizes = [int(np.prod(s)) for _, s in O.PARAM_SHAPES]
    assert off[:-1] == list(np.cumsum([0] + sizes[:-1]))


def test_nerf_module_is_reference_shaped_and_flat():
    net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    sd = net.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, s) for k, s in O.PARAM_SHAPES]
    flat = net.flat_params()
    assert flat.numel() == 595844 and net._aliased()
    p = O.init_params(5)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})     # copies in place: alias survives
    assert net._aliased()
    np.testing.assert_array_equal(net.flat_params().numpy(), np.concatenate([p[k].reshape(-1) for k, _ in O.PARAM_SHAPES]))
    net.double(); net.float()                                                       # breaks the alias ...
    assert net.flat_params().numel() == 595844 and net._aliased()                  # ... flat_params() re-establishes it
    with pytest.raises(RuntimeError):
        net(torch.zeros(3, 6))                                                      # CPU tensor: loud, no fallback


def test_seeded_init_is_the_checkers_weight_stream():
    """bench.py / tools derive their synthetic weights inside the product package (no oracle import on the product
    path); the tests' checker derives the same ones from the same seed."""
    net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).seeded_init_(7)
    ref = O.init_params(7)
    sd = net.state_dict()
    assert all(np.array_equal(sd[k].numpy(), ref[k]) for k in ref)
    np.testing.assert_array_equal(net.flat_params().numpy(), np.concatenate([ref[k].reshape(-1) for k, _ in O.PARAM_SHAPES]))


def test_nerf_rgb_state_dict_has_no_alpha_linear():
    a = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m = spn.NeRF_RGB(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, alpha_model=a)
    keys = [k for k in m.state_dict() if not k.startswith("alpha_model.")]
    assert not any("alpha_linear" in k for k in keys) and len(keys) == 22
    m.load_state_dict(m.state_dict())
    assert len(m._flat_params()) == 24


def test_dropin_module_exports_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "spin-nerf_b200", "dropin"))
    try:
        sys.modules.pop("run_nerf_helpers", None)
        h = importlib.import_module("run_nerf_helpers")
    finally:
        sys.path.pop(0)
    for name in ["get_embedder", "NeRF", "NeRF_RGB", "raw2outputs", "sample_pdf", "get_rays", "get_rays_np",
                 "get_rays_by_coord_np", "ndc_rays", "sample_sigma", "visualize_sigma", "img2mse", "img2l1", "mse2psnr",
                 "to8b", "torch", "torchvision", "cv2", "np", "nn", "F", "plt", "searchsorted", "device", "Embedder"]:
        assert hasattr(h, name), name
    fn, dim = h.get_embedder(10, 0); assert dim == 63
    fn4, dim4 = h.get_embedder(4, 0); assert dim4 == 27
    ident, d3 = h.get_embedder(10, -1); assert d3 == 3 and isinstance(ident, torch.nn.Identity)
    x = torch.randn(5, 3)
    assert fn(x) is x                      # lazy embedder hands the point through; NeRF.forward encodes in-kernel
    c2w = np.eye(4, dtype=np.float32)[:3]
    ro, rd = h.get_rays_np(6, 8, 7.0, c2w); o2, d2 = O.get_rays(6, 8, 7.0, c2w)
    np.testing.assert_allclose(rd, d2, rtol=1e-6); np.testing.assert_allclose(ro, o2)
    co = np.array([[1., 2.], [3., 4.]], np.float32)
    _, dcoord = h.get_rays_by_coord_np(6, 8, 7.0, c2w, co)
    np.testing.assert_allclose(dcoord, rd[[2, 4], [1, 3]], rtol=1e-6)
    assert float(h.mse2psnr(torch.tensor(0.01))) == pytest.approx(20.0, abs=1e-4)
    assert h.to8b(np.array([0.5, 2.0, -1.0])).tolist() == [127, 255, 0]


def test_configargparse_shim_reads_reference_config(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "spin-nerf_b200", "compat"))
    try:
        sys.modules.pop("configargparse", None)
        cap = importlib.import_module("configargparse")
    finally:
        sys.path.pop(0)
    cfg = tmp_path / "c.txt"
    cfg.write_text("expname = statue\nN_rand = 1024\nuse_viewdirs = True\nlrate = 0.03\nno_ndc = True\n# c\nunknown = 1\n")
    p = cap.ArgumentParser()
    p.add_argument('--config', is_config_file=True)
    p.add_argument('--expname', type=str); p.add_argument('--N_rand', type=int, default=32)
    p.add_argument('--use_viewdirs', action='store_true'); p.add_argument('--no_ndc', action='store_true')
    p.add_argument('--lrate', type=float, default=5e-4); p.add_argument('--lindisp', action='store_true')
    a = p.parse_args(['--config', str(cfg), '--N_rand', '4096'])
    assert (a.expname, a.N_rand, a.use_viewdirs, a.no_ndc, a.lrate, a.lindisp) == ("statue", 4096, True, True, 0.03, False)


def test_oracle_backward_matches_finite_differences():
    rng = np.random.default_rng(0)
    raw = rng.standard_normal((3, 9, 4)); z = np.sort(rng.uniform(1, 4, (3, 9)), -1); rd = rng.standard_normal((3, 3))
    gr, gd, ga, gw, gz = (rng.standard_normal(s) for s in ((3, 3), (3,), (3,), (3, 9), (3,)))
    def loss(r):
        o = O.raw2outputs(r.astype(np.float32), z, rd, None, True)
        return float((o[0] * gr).sum() + (o[1] * gd).sum() + (o[2] * ga).sum() + (o[3] * gw).sum() + (o[4] * gz).sum())
    ana = O.raw2outputs_backward(raw, z, rd, gr, gd, ga, gw, gz, None, True, False)
    num = np.zeros_like(raw)
    for idx in np.ndindex(raw.shape):
        d = np.zeros_like(raw); d[idx] = 1e-3
        num[idx] = (loss(raw + d) - loss(raw - d)) / 2e-3
    np.testing.assert_allclose(ana, num, rtol=5e-2, atol=5e-3)
