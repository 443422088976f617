"""create_nerf / run_network / batchify of the render API (SURVEY.md section 8 a4 + the model set-up of a1-a3) against the
unmodified reference's functions (imported through oracle/ref_loader.py where the reference checkout exists)."""
import argparse
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader

spn = importlib.import_module("spin-nerf_b200")
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


def _args(tmp, **over):
    a = argparse.Namespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=64, N_samples=64,
                           netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, alpha_model_path=None,
                           no_coarse=False, netchunk=4096, lrate=5e-4, basedir=str(tmp), expname="exp", ft_path=None,
                           no_reload=False, perturb=1.0, white_bkgd=True, raw_noise_std=1.0, dataset_type="llff", no_ndc=True,
                           lindisp=True, sigma_loss=False)
    a.__dict__.update(over)
    os.makedirs(os.path.join(a.basedir, a.expname), exist_ok=True)
    return a


def test_run_network_hands_the_mlp_points_and_viewdirs():
    """With the lazy embedders the network receives the raw [pt, viewdir] rows (what the fused kernel encodes); with
    materialising embedders of the reference's layout it receives 90 columns whose raw part NeRF._as_points extracts."""
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(5, 7, 3, generator=g); dirs = torch.randn(5, 3, generator=g)
    seen = []
    fn = lambda x: (seen.append(x), x[:, :4] * 2)[1]
    e, d = spn.get_embedder(10, 0); ev, dv = spn.get_embedder(4, 0)
    assert (d, dv) == (63, 27)
    out = spn.run_network(pts, dirs, fn, e, ev, netchunk=16)
    assert out.shape == (5, 7, 4) and [s.shape[0] for s in seen] == [16, 16, 3]        # batchify slices
    x = torch.cat(seen, 0)
    assert torch.equal(x[:, :3], pts.reshape(-1, 3)) and torch.equal(x[:, 3:], dirs[:, None].expand(5, 7, 3).reshape(-1, 3))
    net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    fat = torch.cat([x[:, :3], torch.zeros(35, 60), x[:, 3:], torch.zeros(35, 24)], -1)   # [gamma(pt) | gamma(dir)] layout
    assert torch.equal(net._as_points(fat), x) and net._as_points(x) is x
    assert spn.batchify(fn, None) is fn


@needs_ref
def test_run_network_matches_the_references_plumbing():
    _, R = ref_loader.load()
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(4, 6, 3, generator=g); dirs = torch.randn(4, 3, generator=g)
    ident = lambda x: x
    fn = lambda x: x.sum(-1, keepdim=True).expand(-1, 4).contiguous()
    ours = spn.run_network(pts, dirs, fn, ident, ident, netchunk=5)
    ref = R.run_network(pts, dirs, fn, ident, ident, netchunk=5)
    assert torch.equal(ours, ref)
    assert torch.equal(spn.batchify(fn, 7)(pts.reshape(-1, 3)), R.batchify(fn, 7)(pts.reshape(-1, 3)))


@needs_ref
@pytest.mark.parametrize("variant", ["plain", "ndc", "alpha_model", "no_coarse"])
def test_create_nerf_builds_what_the_reference_builds(tmp_path, variant):
    H, R = ref_loader.load()
    over = {}
    if variant == "ndc":
        over.update(no_ndc=False)
    if variant in ("alpha_model", "no_coarse"):
        provider = H.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        path = str(tmp_path / "provider.tar")
        torch.save({"network_fine_state_dict": provider.state_dict()}, path)
        over.update(alpha_model_path=path, no_coarse=(variant == "no_coarse"))
    a_ref, a_our = _args(tmp_path / "ref", **over), _args(tmp_path / "our", **over)
    saved_dev = R.device
    R.device = torch.device("cpu")
    try:
        torch.manual_seed(0)
        tr_r, te_r, start_r, gv_r, opt_r = R.create_nerf(a_ref)
    finally:
        R.device = saved_dev
    tr_o, te_o, start_o, gv_o, opt_o = spn.create_nerf(a_our, device=torch.device("cpu"))
    assert start_o == start_r == 0
    assert set(tr_o) == set(tr_r) and set(te_o) == set(te_r)
    for k in tr_r:
        if k not in ("network_query_fn", "network_fn", "network_fine"):
            assert tr_o[k] == tr_r[k] and te_o[k] == te_r[k], k
    assert te_o["perturb"] is False and te_o["raw_noise_std"] == 0.
    assert [tuple(p.shape) for p in gv_o] == [tuple(p.shape) for p in gv_r]            # trainable tensors, in order
    assert opt_o.defaults["lr"] == opt_r.defaults["lr"] and opt_o.defaults["betas"] == opt_r.defaults["betas"]
    for k in ("network_fn", "network_fine"):
        m_o, m_r = tr_o[k], tr_r[k]
        assert (m_o is None) == (m_r is None)
        if m_r is not None:
            assert type(m_o).__name__ == type(m_r).__name__
            assert [(n, tuple(v.shape)) for n, v in m_o.state_dict().items()] == [(n, tuple(v.shape)) for n, v in m_r.state_dict().items()]
    if variant in ("alpha_model", "no_coarse"):                                        # the provider's weights were loaded
        got = tr_o["network_fine"].alpha_model.state_dict()
        for n, v in provider.state_dict().items():
            assert torch.equal(got[n], v)


@needs_ref
def test_create_nerf_resumes_from_a_reference_checkpoint(tmp_path):
    """Checkpoints interchange (run_nerf.py:443-461, 1626-1636): a .tar written from the reference's modules and optimizer
    is picked up as the newest checkpoint of basedir/expname."""
    H, R = ref_loader.load()
    a = _args(tmp_path)
    torch.manual_seed(3)
    nets = [H.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True) for _ in range(2)]
    params = list(nets[0].parameters()) + list(nets[1].parameters())
    opt = torch.optim.Adam(params=params, lr=a.lrate, betas=(0.9, 0.999))
    for p in params:
        p.grad = torch.ones_like(p) * 1e-3
    opt.step()
    logdir = os.path.join(a.basedir, a.expname)
    torch.save({"global_step": 1}, os.path.join(logdir, "000001.tar"))                 # an older one: must not be chosen
    torch.save({"global_step": 1234, "network_fn_state_dict": nets[0].state_dict(), "network_fine_state_dict": nets[1].state_dict(),
                "optimizer_state_dict": opt.state_dict()}, os.path.join(logdir, "001234.tar"))
    tr, te, start, gv, opt2 = spn.create_nerf(a, device=torch.device("cpu"))
    assert start == 1234
    for k, net in (("network_fn", nets[0]), ("network_fine", nets[1])):
        for n, v in net.state_dict().items():
            assert torch.equal(tr[k].state_dict()[n], v)
    st = opt2.state_dict()["state"]
    assert len(st) == 48 and all(int(s["step"]) == 1 for s in st.values())
    a.no_reload = True
    assert spn.create_nerf(a, device=torch.device("cpu"))[2] == 0


@needs_ref
def test_trainer_checkpoints_interchange_with_the_reference(tmp_path):
    """Trainer.checkpoint / load_checkpoint use the reference trainer's file layout (run_nerf.py:1626-1636): a checkpoint of
    reference modules + torch.optim.Adam resumes our fused trainer (flat moments), and ours resumes the reference's
    optimizer — weights, both Adam moments and the step count survive the round trip in both directions."""
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
    H, R = ref_loader.load()
    torch.manual_seed(5)
    ref_nets = [H.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True) for _ in range(2)]
    params = list(ref_nets[0].parameters()) + list(ref_nets[1].parameters())
    opt = torch.optim.Adam(params=params, lr=5e-4, betas=(0.9, 0.999))
    for k in range(3):                                              # three reference optimizer steps with synthetic gradients
        for j, p in enumerate(params):
            p.grad = torch.full_like(p, 1e-3 * (k + 1) * ((j % 5) - 2))
        opt.step()
        for group in opt.param_groups:                              # run_nerf.py:1616-1622 with global_step = k
            group["lr"] = 5e-4 * (0.1 ** (k / (250 * 1000)))
    path = str(tmp_path / "000002.tar")
    torch.save({"global_step": 2, "network_fn_state_dict": ref_nets[0].state_dict(),
                "network_fine_state_dict": ref_nets[1].state_dict(), "optimizer_state_dict": opt.state_dict()}, path)

    mk = lambda: spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    tr = trainer_mod.Trainer(mk(), mk()).load_checkpoint(torch.load(path, weights_only=False))
    assert tr.global_step == 3
    flat = lambda ts: torch.cat([t.reshape(-1) for t in ts])
    for net, ref, m, v, lo in ((tr.net_c, ref_nets[0], tr.m[0], tr.v[0], 0), (tr.net_f, ref_nets[1], tr.m[1], tr.v[1], 24)):
        assert torch.equal(net.flat_params(), flat(ref.parameters()))
        assert torch.equal(m, flat([opt.state[p]["exp_avg"] for p in params[lo:lo + 24]]))
        assert torch.equal(v, flat([opt.state[p]["exp_avg_sq"] for p in params[lo:lo + 24]]))

    # ... and back: the reference's create_nerf-style reload of OUR checkpoint
    out = str(tmp_path / "000003_ours.tar")
    torch.save(tr.checkpoint(), out)
    ck = torch.load(out, weights_only=False)
    nets2 = [H.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True) for _ in range(2)]
    params2 = list(nets2[0].parameters()) + list(nets2[1].parameters())
    opt2 = torch.optim.Adam(params=params2, lr=5e-4, betas=(0.9, 0.999))
    opt2.load_state_dict(ck["optimizer_state_dict"])                # run_nerf.py:456
    nets2[0].load_state_dict(ck["network_fn_state_dict"]); nets2[1].load_state_dict(ck["network_fine_state_dict"])
    assert ck["global_step"] == 2 and opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
    for p, q in zip(params, params2):
        assert torch.equal(p, q)
        assert torch.equal(opt.state[p]["exp_avg"], opt2.state[q]["exp_avg"])
        assert torch.equal(opt.state[p]["exp_avg_sq"], opt2.state[q]["exp_avg_sq"])
        assert float(opt2.state[q]["step"]) == 3.0
    for j, (p, q) in enumerate(zip(params, params2)):               # and the next optimizer step agrees
        p.grad = torch.full_like(p, 2e-3); q.grad = torch.full_like(q, 2e-3)
    opt.step(); opt2.step()
    assert all(torch.equal(p, q) for p, q in zip(params, params2))


def test_need_alpha_without_fine_pass_fails_like_the_reference():
    """the reference raises NameError (alpha0 undefined) for need_alpha=True with N_importance == 0;
    ours fails the same way before touching the GPU instead of returning an undefined alpha0."""
    net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    with pytest.raises(NameError):
        spn.render_rays(torch.zeros(4, 11), net, None, 64, N_importance=0, need_alpha=True)
    with pytest.raises(NotImplementedError):
        spn.render_rays(torch.zeros(4, 11), net, None, 64, N_importance=0, sigma_loss=object())


@needs_ref
def test_dropin_host_helpers_equal_the_references():
    """The drop-in module's host-side helpers (numpy ray generation used to precompute the training rays, the loss lambdas)
    against the reference module's, bit for bit."""
    import sys
    H, _ = ref_loader.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "spin-nerf_b200", "dropin"))
    try:
        sys.modules.pop("run_nerf_helpers", None)
        D = importlib.import_module("run_nerf_helpers")
    finally:
        sys.path.pop(0); sys.modules.pop("run_nerf_helpers", None)
    assert D.__file__.endswith(os.path.join("dropin", "run_nerf_helpers.py"))
    rng = np.random.default_rng(0)
    c2w = rng.standard_normal((3, 4)).astype(np.float32)
    for a, b in zip(D.get_rays_np(13, 17, 15.5, c2w), H.get_rays_np(13, 17, 15.5, c2w)):
        np.testing.assert_array_equal(a, b)
    coords = rng.uniform(0, 16, (40, 2))
    for a, b in zip(D.get_rays_by_coord_np(13, 17, 15.5, c2w, coords), H.get_rays_by_coord_np(13, 17, 15.5, c2w, coords)):
        np.testing.assert_array_equal(a, b)
    x, y = torch.rand(7, 3), torch.rand(7, 3)
    assert torch.equal(D.img2mse(x, y), H.img2mse(x, y)) and torch.equal(D.img2l1(x, y), H.img2l1(x, y))
    assert torch.equal(D.mse2psnr(D.img2mse(x, y)), H.mse2psnr(H.img2mse(x, y)))
    img = rng.uniform(-0.2, 1.2, (5, 6, 3))
    np.testing.assert_array_equal(D.to8b(img), H.to8b(img))
    e, d = D.get_embedder(10, 0)
    assert d == H.get_embedder(10, 0)[1] == 63 and D.get_embedder(4, 0)[1] == H.get_embedder(4, 0)[1] == 27
    ident, d3 = D.get_embedder(10, -1)
    assert d3 == 3 and isinstance(ident, torch.nn.Identity)
