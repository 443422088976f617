"""Parity of the CUDA path (called through the C ABI) against the CPU oracle and the golden
fixtures produced by the unmodified reference.  Needs a B200:  pytest -m gpu

Tolerances (fp32 mode): weights/rgb/acc abs 1e-5..2e-4 through the MLP,
depth/z rel 1e-5, disp rel 1e-4, sample_pdf indices bit-exact given identical cdf and u.
bf16 (tcgen05) mode: raw abs 2e-2 * max|raw| (bf16 operands, fp32 accumulation over 9 layers).
"""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

spn = importlib.import_module("spin-nerf_b200")
ops = spn.ops
DEV = "cuda"


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


def N(t):
    return t.detach().cpu().numpy()


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(N(a) if torch.is_tensor(a) else a, b, rtol=rtol, atol=atol)


def close_mostly(a, b, rtol, atol, max_frac=0.01, hard=5e-2):
    a = np.asarray(N(a) if torch.is_tensor(a) else a, np.float64); b = np.asarray(b, np.float64)
    err = np.abs(a - b); tol = atol + rtol * np.abs(b)
    assert np.mean(err > tol) <= max_frac, (np.mean(err > tol), err.max())
    assert err.max() <= hard * max(1.0, np.abs(b).max()), err.max()


def make_net(seed, precision, bias_sigma=0.0):
    p = O.init_params(seed)
    p["alpha_linear.bias"] = p["alpha_linear.bias"] + np.float32(bias_sigma)
    net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    net = net.to(DEV)
    net.precision = precision
    return net, p


def test_library_is_the_native_one():
    assert spn._lib.lib().spn_version() == 100
    sms = __import__("ctypes").c_int()
    spn._lib.check(spn._lib.lib().spn_device_info(sms, None, None))
    assert sms.value >= 100


# ---- small ops --------------------------------------------------------------------------------
def test_embed():
    g = load_golden("embed")
    close(ops.embed(T(g["x"]), 10), g["e10"], rtol=0, atol=2e-6)
    close(ops.embed(T(g["x"]), 4), g["e4"], rtol=0, atol=2e-6)


def test_rays_and_ndc():
    g = load_golden("rays")
    ro, rd = ops.get_rays(12, 16, 14.4, T(g["c2w"]))
    close(ro, g["ro"]); close(rd, g["rd"])
    no, nd = ops.ndc_rays(12, 16, 14.4, 1.0, ro, rd)
    close(no, g["ndc_o"], rtol=1e-5, atol=1e-6); close(nd, g["ndc_d"], rtol=1e-5, atol=1e-6)
    # patch window == python slicing of the full grid (run_nerf.py:120-123)
    po, pd = ops.get_rays(12, 16, 14.4, T(g["c2w"]), patch=(3, 5, 4, 7))
    close(pd, g["rd"][3:7, 5:12])
    rb = ops.build_ray_batch(ro, rd, 0.0, 1.0, ndc=True, H=12, W=16, focal=14.4)
    close(rb, O.make_ray_batch(g["ro"], g["rd"], 0.0, 1.0, ndc=True, H=12, W=16, focal=14.4), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("lindisp", [False, True])
@pytest.mark.parametrize("perturb", [False, True])
def test_sample_z(lindisp, perturb):
    rng = np.random.default_rng(3)
    n, S = 77, 64
    rays = np.zeros((n, 11), np.float32)
    rays[:, 6] = rng.uniform(0.5, 2.0, n); rays[:, 7] = rays[:, 6] + rng.uniform(1, 8, n)
    t_rand = rng.uniform(0, 1, (n, S)).astype(np.float32) if perturb else None
    z = ops.sample_z(T(rays), S, lindisp, None if t_rand is None else T(t_rand))
    ref = O.sample_z(rays[:, 6], rays[:, 7], S, lindisp, t_rand)
    assert np.array_equal(N(z), ref) or np.abs(N(z) - ref).max() <= 2e-6 * np.abs(ref).max()
    assert (np.diff(N(z), axis=1) >= 0).all()


@pytest.mark.parametrize("tag,white,detach", [("plain", False, False), ("white", True, False),
                                              ("noise_detach", True, True)])
def test_raw2outputs_forward_backward(tag, white, detach):
    g = {k.split("__", 1)[1]: v for k, v in load_golden("raw2outputs").items() if k.startswith(tag + "__")}
    raw = T(g["raw"]).requires_grad_(True)
    noise = T(g["noise"]) if g["noise"].any() else None
    outs = spn.ops._Raw2Outputs.apply(raw, T(g["z"]), T(g["rd"]), noise, white, True, detach)
    rgb, disp, acc, w, depth, alpha = outs
    close(w, g["w"], atol=2e-6); close(alpha, g["alpha"], atol=2e-6)
    close(rgb, g["rgb"], atol=3e-6); close(acc, g["acc"], atol=3e-6)
    close(depth, g["depth"], rtol=1e-5); close(disp, g["disp"], rtol=1e-4)
    loss = sum((o * T(g[k])).sum() for o, k in zip((rgb, disp, acc, w, depth), ("g_rgb", "g_disp", "g_acc", "g_w", "g_depth")))
    loss.backward()
    close(raw.grad, g["d_raw"], rtol=2e-3, atol=2e-4 * np.abs(g["d_raw"]).max())


def test_raw2outputs_edge_cases():
    # ragged sample counts (not a multiple of 32), single sample, empty batch, zero density
    rng = np.random.default_rng(5)
    for S in (1, 7, 33, 100, 200):
        raw = (rng.standard_normal((9, S, 4))).astype(np.float32)
        z = np.sort(rng.uniform(1, 5, (9, S)).astype(np.float32), -1)
        rd = rng.standard_normal((9, 3)).astype(np.float32)
        got = ops.raw2outputs(T(raw), T(z), T(rd), need_alpha=True)
        ref = O.raw2outputs(raw, z, rd, need_alpha=True)
        for a, b in zip(got, ref):
            close(a, b, rtol=2e-5, atol=3e-6)
    e = ops.raw2outputs(torch.empty(0, 64, 4, device=DEV), torch.empty(0, 64, device=DEV), torch.empty(0, 3, device=DEV))
    assert e[0].shape == (0, 3) and e[3].shape == (0, 64)
    raw = np.zeros((2, 8, 4), np.float32); raw[..., 3] = -1.0       # sigma<=0 everywhere -> acc=0, disp NaN like torch
    out = ops.raw2outputs(T(raw), T(np.tile(np.arange(1, 9, dtype=np.float32), (2, 1))), T(np.ones((2, 3), np.float32)))
    assert float(out[2].abs().max()) == 0.0 and torch.isnan(out[1]).all()
    with pytest.raises(RuntimeError):
        ops.raw2outputs(torch.zeros(2, 300, 4, device=DEV, requires_grad=True), torch.zeros(2, 300, device=DEV),
                        torch.ones(2, 3, device=DEV))[0].sum().backward()       # S > 256 unsupported in bwd


def test_sample_pdf_indices_bit_exact():
    g = load_golden("sample_pdf")
    for u, ref_s in ((g["u_det"], g["det"]), (g["u_sto"], g["sto"])):
        s, inds, cdf = ops.sample_pdf(T(g["bins"]), T(g["weights"]), 64, u=T(u), return_inds=True, return_cdf=True)
        assert np.abs(N(cdf) - g["cdf"]).max() <= 4e-7          # fp32 row total may round 1 ulp differently
        # bit-exact GIVEN identical cdf and u: the oracle search/lerp on the GPU's own cdf
        s_o, i_o = O.sample_pdf_from_cdf(g["bins"], N(cdf), u)
        assert np.array_equal(N(inds), i_o)
        assert np.array_equal(N(s), s_o)
        # and against the reference's samples: equal except where a ~1e-5 cdf step amplifies 1 ulp
        err = np.abs(N(s) - ref_s)
        assert np.mean(err > 1e-5) < 0.01 and err.max() < 2e-2
    # weights whose fp32 sums are exact in any order -> cdf identical to the reference -> indices identical
    rng = np.random.default_rng(9)
    w = (rng.integers(0, 64, (50, 62)) / 64.0).astype(np.float32) - np.float32(1e-5)
    bins = np.sort(rng.uniform(1, 8, (50, 63)).astype(np.float32), -1)
    uu = rng.uniform(0, 1, (50, 64)).astype(np.float32)
    s, inds, cdf = ops.sample_pdf(T(bins), T(w), 64, u=T(uu), return_inds=True, return_cdf=True)
    ref_cdf = O.pdf_to_cdf(w)
    s_o, i_o = O.sample_pdf_from_cdf(bins, ref_cdf, uu)
    same = np.all(N(cdf) == ref_cdf, axis=1)
    assert same.mean() > 0.5 and np.array_equal(N(inds)[same], i_o[same])
    s_own, i_own = O.sample_pdf_from_cdf(bins, N(cdf), uu)          # every row: exact given the GPU's own cdf
    assert np.array_equal(N(inds), i_own) and np.array_equal(N(s), s_own)
    # torch.searchsorted(right=True) KAT (cdf [0,.2,.2,.7,1], u [0,.2,.7,.99,1,1.1] -> [1,3,4,4,5,5]),
    # unconditionally: the KAT cdf goes straight through the library's search (numpy side='right' = torch right=True)
    g2 = load_golden("searchsorted_kat")
    got = ops.searchsorted(T(g2["cdf"].astype(np.float32)), T(g2["u"].astype(np.float32).reshape(1, -1)), side='right')
    assert N(got).reshape(-1).tolist() == np.asarray(g2["inds"]).reshape(-1).tolist()
    # and through sample_pdf's own search when its cdf reproduces the KAT's bit for bit
    cdf_k = g2["cdf"][0]
    wk = np.diff(cdf_k).astype(np.float32)        # build weights that reproduce the KAT cdf
    _, inds_k, cdf_g = ops.sample_pdf(T(np.zeros((1, 5), np.float32)), T(wk[None] - np.float32(1e-5)), 6, u=T(g2["u"]),
                                      return_inds=True, return_cdf=True)
    i_k = O.sample_pdf_from_cdf(np.zeros((1, 5), np.float32), N(cdf_g), np.asarray(g2["u"], np.float32).reshape(1, -1))[1]
    assert np.array_equal(N(inds_k), i_k)
    if np.array_equal(N(cdf_g)[0], cdf_k):
        assert N(inds_k).tolist() == g2["inds"].tolist()


def test_merge_and_resample():
    rng = np.random.default_rng(11)
    a = np.sort(rng.uniform(1, 8, (33, 64)).astype(np.float32), -1)
    b = rng.uniform(1, 8, (33, 64)).astype(np.float32)                 # unsorted, like stochastic z_samples
    close(ops.merge_sorted(T(a), T(b)), O.merge_sorted(a, b), rtol=0, atol=0)
    close(ops.merge_sorted(T(a[:, :5]), T(b[:, :3])), O.merge_sorted(a[:, :5], b[:, :3]), rtol=0, atol=0)
    w = (rng.uniform(0, 1, (33, 64)) ** 3).astype(np.float32)
    u = rng.uniform(0, 1, (33, 64)).astype(np.float32)
    zm, zstd, zs, inds = ops.resample(T(a), T(w), 64, T(u), want_samples=True, want_inds=True)
    mid = np.float32(0.5) * (a[:, 1:] + a[:, :-1])
    s_o, i_o = O.sample_pdf(mid, w[:, 1:-1], 64, det=False, u=u)
    assert (N(inds) == i_o).mean() > 0.999                     # vs the oracle's own cdf (may differ by an ulp per row) ...
    # ... and EXACT given the GPU's cdf: sample_pdf's kernel on the same mids / weights returns the cdf the fused kernel builds
    _, i_g, cdf_g = ops.sample_pdf(T(mid), T(w[:, 1:-1].copy()), 64, u=T(u), return_inds=True, return_cdf=True)
    s_x, i_x = O.sample_pdf_from_cdf(mid, N(cdf_g), u)
    assert np.array_equal(N(inds), i_x) and np.array_equal(N(zs), s_x)
    close_mostly(zs, s_o, rtol=1e-5, atol=1e-5)
    assert np.array_equal(N(zm), np.sort(np.concatenate([a, N(zs)], -1), -1))       # sortedness + multiset
    close(zstd, N(zs).std(-1), rtol=1e-4, atol=1e-5)


# ---- MLP ----------------------------------------------------------------------------------------
def _mlp_case():
    g = load_golden("mlp")
    x90 = g["x90"]
    x6 = np.concatenate([x90[:, 0:3], x90[:, 63:66]], -1)
    return g, x90, x6


def test_mlp_fp32_forward_backward():
    g, x90, x6 = _mlp_case()
    net, p = make_net(11, spn.PREC_FP32)
    raw = net(T(x6))
    close(raw, g["raw"], rtol=1e-4, atol=3e-5)
    close(net(T(x90)), g["raw"], rtol=1e-4, atol=3e-5)        # pre-embedded input (non-lazy embedder) path
    raw.backward(T(g["draw"]))
    for k, v in net.named_parameters():
        gv = N(v.grad)
        close(gv.reshape(-1)[::97], g["g_sub__" + k], rtol=2e-3, atol=2e-4)
        assert abs(np.abs(gv).sum(dtype=np.float64) - g["g_abs__" + k]) <= 1e-3 * g["g_abs__" + k] + 1e-4


def test_mlp_ragged_sizes_fp32():
    net, p = make_net(11, spn.PREC_FP32)
    rng = np.random.default_rng(2)
    for m in (1, 127, 129, 1000):
        x6 = np.concatenate([rng.standard_normal((m, 3)) * 2, O.embed(rng.standard_normal((m, 3)), 0)], -1).astype(np.float32)
        x6[:, 3:] /= np.linalg.norm(x6[:, 3:], axis=1, keepdims=True)
        ref = O.mlp_forward(p, np.concatenate([O.embed(x6[:, :3], 10), O.embed(x6[:, 3:], 4)], -1))
        with torch.no_grad():
            close(net(T(x6)), ref, rtol=1e-4, atol=3e-5)
    with torch.no_grad():
        assert net(torch.empty(0, 6, device=DEV)).shape == (0, 4)


# ---- render end to end ----------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["det_lindisp_white", "det_ndc", "sto_lindisp_white", "coarse_only"])
def test_render_matches_reference_fp32(tag):
    g = load_golden("render")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    netc, _ = make_net(11, spn.PREC_FP32, 1.0)
    netf, _ = make_net(12, spn.PREC_FP32, 1.0)
    ndc = tag == "det_ndc"
    near, far = (0.0, 1.0) if ndc else (1.2, 8.0)
    n_imp = 0 if tag == "coarse_only" else 64
    sto = tag.startswith("sto")
    kw = dict(network_query_fn=None, network_fn=netc, network_fine=netf if n_imp else None, N_samples=64,
              N_importance=n_imp, lindisp="lindisp" in tag, white_bkgd="white" in tag, perturb=1. if sto else 0.,
              raw_noise_std=1. if sto else 0., pytest=sto)
    with torch.no_grad():
        rgb, disp, acc, depth, ex = spn.render(H, W, f, chunk=32768, rays=T(g["rays"]), retraw=True, use_viewdirs=True,
                                               ndc=ndc, near=near, far=far, need_alpha=bool(n_imp), **kw)
    G = lambda k: g[f"{tag}__{k}"]
    cm = (lambda a, b, rtol=1e-5, atol=1e-6: close(a, b, rtol, atol)) if n_imp == 0 else close_mostly
    cm(ex["z_vals"], G("z_vals"), rtol=2e-5, atol=1e-5)
    cm(ex["raw"], G("raw"), rtol=1e-3, atol=1e-3)
    cm(ex["weights"], G("weights"), rtol=0, atol=2e-4)
    cm(rgb, G("rgb"), rtol=0, atol=2e-4); cm(acc, G("acc"), rtol=0, atol=2e-4)
    cm(depth, G("depth"), rtol=2e-4, atol=2e-4); cm(disp, G("disp"), rtol=5e-4, atol=0)
    if n_imp:
        close(ex["rgb0"], G("rgb0"), atol=2e-4); close(ex["disp0"], G("disp0"), rtol=5e-4)
        cm(ex["z_std"], G("z_std"), rtol=1e-3, atol=1e-4)
        cm(ex["alpha"], G("alpha"), rtol=0, atol=2e-4); close(ex["alpha0"], G("alpha0"), atol=2e-4)
        mse = float(((rgb - T(G("rgb"))) ** 2).mean())
        assert -10 * np.log10(max(mse, 1e-20)) > 60.0          # PSNR of our render vs the reference's render


def test_train_step_matches_reference_fp32():
    """loss, gradients and two Adam steps through render() (golden: tests/golden/make_golden.py train_step)."""
    g = load_golden("render"); tr = load_golden("train_step")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    netc, _ = make_net(11, spn.PREC_FP32, 1.0)
    netf, _ = make_net(12, spn.PREC_FP32, 1.0)
    opt = torch.optim.Adam(list(netc.parameters()) + list(netf.parameters()), lr=5e-4, betas=(0.9, 0.999))
    target, tdisp = T(tr["target"]), T(tr["tdisp"])
    mse = lambda a, b: torch.mean((a - b) ** 2)
    for it in range(2):
        rgb, disp, acc, depth, ex = spn.render(H, W, f, chunk=32768, rays=T(g["rays"]), retraw=True, use_viewdirs=True,
                                               network_query_fn=None, network_fn=netc, network_fine=netf, N_samples=64,
                                               N_importance=64, ndc=False, lindisp=True, white_bkgd=True, perturb=0.,
                                               raw_noise_std=0., near=1.2, far=8.0)
        opt.zero_grad()
        loss = mse(rgb, target) + mse(ex["rgb0"], target) + mse(disp, tdisp) + mse(ex["disp0"], tdisp)
        loss.backward()
        assert abs(loss.item() - float(tr[f"loss{it}"])) <= 2e-4 * abs(float(tr[f"loss{it}"])), (it, loss.item())
        if it == 0:
            for nm, net in (("c", netc), ("f", netf)):
                for k, v in net.named_parameters():
                    gv = N(v.grad)
                    ref_abs = float(tr[f"g_abs__{nm}__{k}"])
                    assert abs(np.abs(gv).sum(dtype=np.float64) - ref_abs) <= 5e-3 * ref_abs + 1e-6, (nm, k)
                    close_mostly(gv.reshape(-1)[::997], tr[f"g_sub__{nm}__{k}"], rtol=1e-2, atol=1e-3 * np.abs(gv).max() + 1e-9,
                                 max_frac=0.03, hard=1.0)
        opt.step()
    for nm, net in (("c", netc), ("f", netf)):
        for k, v in net.named_parameters():
            close_mostly(N(v).reshape(-1)[::997], tr[f"p_sub__{nm}__{k}"], rtol=0, atol=2e-4, max_frac=0.03, hard=1.0)


# ---- tcgen05 (bf16) path ----------------------------------------------------------------------------
def bf16r(a):
    """round-to-nearest-even to bfloat16, returned as float32 (what cvt.rn.bf16x2.f32 does)."""
    return N(torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).float())


def mlp_forward_bf16(p, x6):
    """The kernel's arithmetic restated on the CPU: bf16 operands (weights, encodings, activations), fp32
    accumulation, fp32 bias/ReLU, sigma from the fp32 h7, rgb from the fp32 hv."""
    W = lambda k: bf16r(p[k])
    xp = bf16r(O.embed(x6[:, :3], 10)); xd = bf16r(O.embed(x6[:, 3:], 4))
    acts = {"xp": xp, "xd": xd}
    h = xp
    for i in range(8):
        w = W(f"pts_linears.{i}.weight")
        pre = (h @ w[:, 63:].T + xp @ w[:, :63].T) if i == 5 else h @ w.T
        hf = np.maximum(pre + p[f"pts_linears.{i}.bias"], 0).astype(np.float32)
        h = bf16r(hf); acts[f"h{i}"] = h
    alpha = hf @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
    feat = bf16r(h @ W("feature_linear.weight").T + p["feature_linear.bias"]); acts["feat"] = feat
    wv = W("views_linears.0.weight")
    hv = np.maximum(feat @ wv[:, :256].T + xd @ wv[:, 256:].T + p["views_linears.0.bias"], 0).astype(np.float32)
    acts["hv"] = bf16r(hv)
    rgb = hv @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
    return np.concatenate([rgb, alpha], -1).astype(np.float32), acts


@pytest.mark.parametrize("N_,K", [(256, 64), (256, 256), (128, 128), (128, 64)])
def test_tc_selftest_gemm(N_, K):
    rng = np.random.default_rng(N_ + K)
    A = rng.standard_normal((128, K)).astype(np.float32); B = rng.standard_normal((N_, K)).astype(np.float32)
    D = torch.full((128, N_), float("nan"), device=DEV)
    tA, tB = T(A), T(B)            # keep the device buffers alive across the call
    spn._lib.check(spn._lib.lib().spn_tc_selftest_gemm(spn._lib.ptr(tA), spn._lib.ptr(tB), spn._lib.ptr(D), N_, K,
                                                       spn._lib.stream()), "selftest")
    torch.cuda.synchronize()
    ref = bf16r(A).astype(np.float64) @ bf16r(B).astype(np.float64).T
    np.save("gpurun_out/selftest_D_%d_%d.npy" % (N_, K), N(D)) if __import__("os").path.isdir("gpurun_out") else None
    close(D, ref, rtol=1e-4, atol=1e-4)


def test_mlp_bf16_forward_matches_emulation_and_oracle():
    g, x90, x6 = _mlp_case()
    net, p = make_net(11, spn.PREC_BF16)
    with torch.no_grad():
        raw = N(net(T(x6)))
    emu, _ = mlp_forward_bf16(p, x6)
    scale = np.abs(g["raw"]).max()
    if __import__("os").path.isdir("gpurun_out"):
        np.savez("gpurun_out/mlp_bf16_dbg.npz", raw=raw, emu=emu, ref=g["raw"])
    assert np.abs(raw - emu).max() <= 2e-3 * scale, np.abs(raw - emu).max() / scale      # same arithmetic, other sum order
    assert np.abs(raw - g["raw"]).max() <= 2e-2 * scale                                  # vs the fp32 reference


@pytest.mark.parametrize("m", [1, 127, 128, 129, 257, 1000, 40000])
def test_mlp_bf16_ragged_and_multi_tile(m):
    net, p = make_net(11, spn.PREC_BF16)
    rng = np.random.default_rng(m)
    x6 = np.concatenate([rng.standard_normal((m, 3)) * 2, rng.standard_normal((m, 3))], -1).astype(np.float32)
    x6[:, 3:] /= np.linalg.norm(x6[:, 3:], axis=1, keepdims=True)
    with torch.no_grad():
        raw = N(net(T(x6)))
    sel = rng.choice(m, min(m, 512), replace=False)
    emu, _ = mlp_forward_bf16(p, x6[sel])
    assert np.abs(raw[sel] - emu).max() <= 3e-3 * max(1.0, np.abs(emu).max())


def mlp_backward_bf16(p, acts, masks_from, draw):
    """The backward kernels' arithmetic on the CPU: bf16 operands (transposed weights, d(pre-activations),
    stashed activations), fp32 accumulation; heads use fp32 d_raw and fp32 head weights."""
    W = lambda k: bf16r(p[k])
    d_rgb, d_alpha = draw[:, :3], draw[:, 3:4]
    g, dpre = {}, {}
    hv, h7 = acts["hv"], acts["h7"]
    g["rgb_linear.weight"] = d_rgb.T @ hv; g["rgb_linear.bias"] = d_rgb.sum(0)
    g["alpha_linear.weight"] = d_alpha.T @ h7; g["alpha_linear.bias"] = d_alpha.sum(0)
    d_hv = bf16r((d_rgb @ p["rgb_linear.weight"]) * (masks_from["hv"] > 0)); dpre["hv"] = d_hv
    wv = W("views_linears.0.weight")
    g["views_linears.0.weight"] = np.concatenate([d_hv.T @ acts["feat"], d_hv.T @ acts["xd"]], 1)
    g["views_linears.0.bias"] = d_hv.sum(0)
    d_feat = bf16r(d_hv @ wv[:, :256]); dpre["feat"] = d_feat
    g["feature_linear.weight"] = d_feat.T @ h7; g["feature_linear.bias"] = d_feat.sum(0)
    d = bf16r((d_feat @ W("feature_linear.weight") + d_alpha * p["alpha_linear.weight"]) * (masks_from["h7"] > 0))
    for i in range(7, -1, -1):
        dpre[f"h{i}"] = d
        inp = acts["xp"] if i == 0 else (np.concatenate([acts["xp"], acts["h4"]], 1) if i == 5 else acts[f"h{i-1}"])
        g[f"pts_linears.{i}.weight"] = d.T @ inp
        g[f"pts_linears.{i}.bias"] = d.sum(0)
        if i > 0:
            w = W(f"pts_linears.{i}.weight")
            w = w[:, 63:] if i == 5 else w
            d = bf16r((d @ w) * (masks_from[f"h{i-1}"] > 0))
    return {k: v.astype(np.float32) for k, v in g.items()}, dpre


def decode_tiles(buf, tile_bytes, atom0, natoms, ntiles, rows):
    """bf16 SWIZZLE_128B tile images -> float32 [rows, 64*natoms] (spin-nerf_b200/csrc/mlp_tc.cuh layouts)."""
    raw = N(buf[:ntiles * tile_bytes]).reshape(ntiles, tile_bytes)
    out = np.zeros((ntiles * 128, 64 * natoms), np.float32)
    r = np.arange(128)
    for t in range(ntiles):
        for a in range(natoms):
            atom = raw[t, (atom0 + a) * 16384:(atom0 + a + 1) * 16384].reshape(128, 8, 16)
            for j in range(8):
                chunk = atom[r, j ^ (r & 7)]                                # [128,16] bytes
                vals = chunk.copy().view(np.uint16).astype(np.uint32) << 16
                out[t * 128:(t + 1) * 128, a * 64 + j * 8:a * 64 + j * 8 + 8] = vals.view(np.float32)
    return out[:rows]


def _e4m3_table():
    """all 256 E4M3 (fp8, 4 exponent bits bias 7, 3 mantissa bits, no infinities) values"""
    t = np.zeros(256, np.float32)
    for b in range(256):
        sgn = -1.0 if b & 0x80 else 1.0
        e, m = (b >> 3) & 15, b & 7
        if e == 15 and m == 7:
            v = np.nan
        elif e == 0:
            v = m / 8.0 * 2.0 ** -6
        else:
            v = (1.0 + m / 8.0) * 2.0 ** (e - 7)
        t[b] = sgn * v
    return t


E4M3 = _e4m3_table()
STASH_TILE_BYTES, DSTASH_TILE_BYTES, STASH_ATOMS = 397312, 622592, 22          # mlp_tc.cuh
SA_ENC, SA_X0, SA_HV, SA_DENC = 0, 1, 19, 21


def decode_x8(buf, layer, ntiles, rows):
    """E4M3 stash image of a 256-wide layer output (layer 0..7 = h0..h7, 8 = feature) -> float32 [rows, 256]:
    atom SA_X0 + 2*layer + half, row r = 128 bytes, chunk j' at position j' ^ (r & 7) = features 64*half + 16*(j'&3) + 128*(j'>>2) ..+15"""
    raw = N(buf[:ntiles * STASH_TILE_BYTES]).reshape(ntiles, STASH_TILE_BYTES)
    out = np.zeros((ntiles * 128, 256), np.float32)
    r = np.arange(128)
    for t in range(ntiles):
        for half in range(2):
            a = SA_X0 + 2 * layer + half
            atom = raw[t, a * 16384:(a + 1) * 16384].reshape(128, 8, 16)
            for j in range(8):
                f0 = 64 * half + 16 * (j & 3) + 128 * (j >> 2)
                out[t * 128:(t + 1) * 128, f0:f0 + 16] = E4M3[atom[r, j ^ (r & 7)]]
    return out[:rows]


def decode_masks(buf, slot, ntiles, rows, width=256):
    """ReLU mask stash (after the atoms): slot x 128 rows x 8 words, column 32w + c at bit ((c & 1) << 4) | (c >> 1) of word w"""
    raw = N(buf[:ntiles * STASH_TILE_BYTES]).reshape(ntiles, STASH_TILE_BYTES)
    out = np.zeros((ntiles * 128, width), bool)
    c = np.arange(32)
    bit = ((c & 1) << 4) | (c >> 1)
    for t in range(ntiles):
        words = raw[t, STASH_ATOMS * 16384:].copy().view(np.uint32).reshape(9, 128, 8)[slot]
        for w in range(width // 32):
            out[t * 128:(t + 1) * 128, 32 * w:32 * w + 32] = (words[:, w:w + 1] >> bit[None, :]) & 1
    return out[:rows]


def test_e4m3_decode_all_codes():
    """the weight-gradient kernel's five-instruction E4M3 -> bf16 widening is exact for every finite code, subnormals included"""
    codes = torch.arange(256, dtype=torch.uint8, device=DEV)
    out = torch.zeros(256, dtype=torch.int16, device=DEV)
    spn._lib.check(spn._lib.lib().spn_tc_e4m3_decode(spn._lib.ptr(codes), spn._lib.ptr(out), 256, spn._lib.stream()), "e4m3 decode")
    torch.cuda.synchronize()
    got = (N(out).view(np.uint16).astype(np.uint32) << 16).view(np.float32)
    finite = ~np.isnan(E4M3)
    assert np.array_equal(got[finite], E4M3[finite]), np.nonzero(got[finite] != E4M3[finite])


@pytest.mark.parametrize("m", [192, 1000])
def test_mlp_bf16_backward(m):
    net, p = make_net(11, spn.PREC_BF16)
    rng = np.random.default_rng(m)
    if m == 192:
        g, x90, x6 = _mlp_case(); draw = g["draw"]
    else:
        x6 = np.concatenate([rng.standard_normal((m, 3)) * 2, rng.standard_normal((m, 3))], -1).astype(np.float32)
        x6[:, 3:] /= np.linalg.norm(x6[:, 3:], axis=1, keepdims=True)
        draw = rng.standard_normal((m, 4)).astype(np.float32)
    flat, packed = net._sync()
    stash = ops.mlp_stash(m, spn.PREC_BF16, DEV); stash.zero_()
    raw, _ = ops.mlp_forward_points(flat, packed, T(x6), spn.PREC_BF16, stash)
    ws = ops.mlp_bwd_workspace(m, spn.PREC_BF16, DEV); ws.zero_()
    grads = torch.zeros(spn.MLP_NPARAMS, device=DEV)
    ops.mlp_backward(flat, packed, stash, T(draw), grads, spn.PREC_BF16, workspace=ws)
    torch.cuda.synchronize()
    emu_raw, acts = mlp_forward_bf16(p, x6)
    ntiles = (m + 127) // 128
    SB, DB = STASH_TILE_BYTES, DSTASH_TILE_BYTES
    # forward stash == emulated activations: E4M3 carries 3 mantissa bits (half an ulp = 2^-4 relative)
    h3 = decode_x8(stash, 3, ntiles, m)
    assert np.abs(h3 - acts["h3"]).max() <= 7e-2 * max(1.0, np.abs(acts["h3"]).max())
    assert np.mean(np.abs(h3 - acts["h3"])) <= 3e-2 * np.mean(np.abs(acts["h3"])) + 1e-6     # E4M3: ~2.2 % mean relative rounding error
    gpu_acts = {"xp": decode_tiles(stash, SB, SA_ENC, 1, ntiles, m)[:, :63], "xd": decode_tiles(stash, SB, SA_DENC, 1, ntiles, m)[:, :27],
                "feat": decode_x8(stash, 8, ntiles, m), "hv": decode_tiles(stash, SB, SA_HV, 2, ntiles, m)}
    for i in range(8):
        gpu_acts[f"h{i}"] = decode_x8(stash, i, ntiles, m)
    # the ReLU masks the dgrad chain applies are the stashed bit masks (a small activation may round to zero in E4M3)
    gpu_masks = {f"h{i}": decode_masks(stash, i, ntiles, m) for i in range(8)}
    gpu_masks["hv"] = decode_masks(stash, 8, ntiles, m, 128)
    for i in (0, 7):     # mask bits == sign of the emulated activations, up to last-bit flips around zero
        assert np.mean(gpu_masks[f"h{i}"] != (acts[f"h{i}"] > 0)) < 2e-3
    # emulate the backward on the GPU's own stashed activations (so only the backward is under test)
    emu_g, dpre = mlp_backward_bf16(p, gpu_acts, gpu_masks, draw)
    d_h7 = decode_tiles(ws, DB, 6, 4, ntiles, m); d_h0 = decode_tiles(ws, DB, 34, 4, ntiles, m)
    d_hv = decode_tiles(ws, DB, 0, 2, ntiles, m)
    off = spn._lib.param_offsets()
    names = [k for k, _ in O.PARAM_SHAPES]
    G = {k: N(grads[off[i]:off[i + 1]]).reshape(p[k].shape) for i, k in enumerate(names)}
    if __import__("os").path.isdir("gpurun_out"):
        np.savez(f"gpurun_out/mlp_bwd_dbg_{m}.npz", d_h7=d_h7, d_h0=d_h0, d_hv=d_hv, e_h7=dpre["h7"], e_h0=dpre["h0"],
                 e_hv=dpre["hv"], **{"G_" + k: v for k, v in G.items()}, **{"E_" + k: v for k, v in emu_g.items()})
    for got, want in ((d_hv, dpre["hv"]), (d_h7, dpre["h7"]), (d_h0, dpre["h0"])):
        assert np.abs(got - want).max() <= 2e-2 * max(1e-6, np.abs(want).max()), np.abs(got - want).max()
    for k in names:
        scale = max(1e-6, np.abs(emu_g[k]).max())
        assert np.abs(G[k] - emu_g[k]).max() <= 2e-2 * scale, (k, np.abs(G[k] - emu_g[k]).max() / scale)
    if m == 192:        # and against the fp32 reference autograd (golden)
        g = load_golden("mlp")
        for k in names:
            ref_abs = float(g["g_abs__" + k])
            assert abs(np.abs(G[k]).sum(dtype=np.float64) - ref_abs) <= 6e-2 * ref_abs + 1e-4, k
            # gradients are sums of +-terms over samples; each bf16-rounded term carries ~4e-3 relative noise, so
            # entries are compared at 10% of the tensor's scale (the 2e-2 check vs the bf16 emulation above is the
            # strict one)
            close_mostly(G[k].reshape(-1)[::97], g["g_sub__" + k], rtol=5e-2, atol=1e-1 * np.abs(G[k]).max() + 1e-6,
                         max_frac=0.05, hard=1.0)


def test_train_step_bf16_tracks_reference():
    """Two Adam steps in bf16 mode: loss within 1%% of the reference's fp32 autograd run."""
    g = load_golden("render"); tr = load_golden("train_step")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    netc, _ = make_net(11, spn.PREC_BF16, 1.0)
    netf, _ = make_net(12, spn.PREC_BF16, 1.0)
    opt = torch.optim.Adam(list(netc.parameters()) + list(netf.parameters()), lr=5e-4, betas=(0.9, 0.999))
    target, tdisp = T(tr["target"]), T(tr["tdisp"])
    mse = lambda a, b: torch.mean((a - b) ** 2)
    for it in range(2):
        rgb, disp, acc, depth, ex = spn.render(H, W, f, chunk=32768, rays=T(g["rays"]), retraw=True, use_viewdirs=True,
                                               network_query_fn=None, network_fn=netc, network_fine=netf, N_samples=64,
                                               N_importance=64, ndc=False, lindisp=True, white_bkgd=True, perturb=0.,
                                               raw_noise_std=0., near=1.2, far=8.0)
        opt.zero_grad()
        loss = mse(rgb, target) + mse(ex["rgb0"], target) + mse(disp, tdisp) + mse(ex["disp0"], tdisp)
        loss.backward()
        assert abs(loss.item() - float(tr[f"loss{it}"])) <= 1e-2 * abs(float(tr[f"loss{it}"])), (it, loss.item())
        opt.step()


def test_mlp_bf16_full_size_is_chunk_invariant():
    """BASELINE size (3 x 1024 rays x 128 fine samples = 393 216 evaluations, many rounds of every CTA pair): a sample's
    output must not depend on which tile / CTA / round processed it, and the weight gradient of the whole batch must be
    the sum of the gradients of its parts (size-independent properties; the oracle is too slow at this size)."""
    net, p = make_net(11, spn.PREC_BF16)
    M = 3 * 1024 * 128
    gen = torch.Generator(device=DEV); gen.manual_seed(7)
    x6 = torch.randn(M, 6, device=DEV, generator=gen)
    x6[:, 3:] = x6[:, 3:] / x6[:, 3:].norm(dim=1, keepdim=True)
    draw = torch.randn(M, 4, device=DEV, generator=gen) / M
    flat, packed = net._sync()
    cut = 1000 * 128 + 77                                    # ragged split: tiles of the parts do not line up with the whole
    outs, grads = [], []
    for lo, hi in ((0, M), (0, cut), (cut, M)):
        n = hi - lo
        stash = ops.mlp_stash(n, spn.PREC_BF16, DEV)
        raw, _ = ops.mlp_forward_points(flat, packed, x6[lo:hi].contiguous(), spn.PREC_BF16, stash)
        g = torch.zeros(spn.MLP_NPARAMS, device=DEV)
        ops.mlp_backward(flat, packed, stash, draw[lo:hi].contiguous(), g, spn.PREC_BF16,
                         workspace=ops.mlp_bwd_workspace(n, spn.PREC_BF16, DEV))
        outs.append(raw); grads.append(g)
        del stash
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], torch.cat([outs[1], outs[2]], 0))          # bit-exact, row by row
    gsum = grads[1] + grads[2]
    scale = grads[0].abs().max()
    assert scale > 0 and (grads[0] - gsum).abs().max() <= 2e-3 * scale, float((grads[0] - gsum).abs().max() / scale)


@pytest.mark.parametrize("prec_name", ["fp32", "bf16"])
def test_trainer_batched_step_equals_three_render_calls(prec_name):
    """Trainer.step renders the step's three ray batches as ONE chunk (per-ray-range losses and detach_weights);
    it must produce the gradients, loss and parameters of the three separate render calls of run_nerf.py:1455-1521."""
    trainer_mod = __import__("importlib").import_module("spin-nerf_b200.trainer")
    prec = spn.PREC_FP32 if prec_name == "fp32" else spn.PREC_BF16
    g = load_golden("render")
    rng = np.random.default_rng(5)
    rays = g["rays"]                                           # [2, n, 3]
    n = rays.shape[1]
    sel = [rng.permutation(n)[:m] for m in (96, 64, 80)]       # three batches of different sizes
    batches = []
    for i, ix in enumerate(sel):
        batches.append(T(np.ascontiguousarray(rays[:, ix])))
        batches.append(T(rng.uniform(0, 1, (len(ix), 3) if i < 2 else (len(ix),)).astype(np.float32)))
    out = []
    for mode in ("batched", "three"):
        netc, _ = make_net(11, prec, 1.0); netf, _ = make_net(12, prec, 1.0)
        tr = trainer_mod.Trainer(netc, netf, lr=5e-4, N_samples=64, N_importance=64, lindisp=True, white_bkgd=True,
                                 perturb=0.0, raw_noise_std=0.0, near=1.2, far=8.0)
        fn = tr.step if mode == "batched" else tr.step_three_calls
        loss, psnr = fn(*batches)
        torch.cuda.synchronize()
        out.append((float(loss), float(psnr), N(tr.grads[0]), N(tr.grads[1]), N(netc.flat_params()), N(netf.flat_params())))
    a, b = out
    assert abs(a[0] - b[0]) <= 1e-5 * abs(b[0]) and abs(a[1] - b[1]) <= 1e-4 * abs(b[1]), (a[:2], b[:2])
    tol = 1e-5 if prec_name == "fp32" else 2e-3              # bf16: tile boundaries move, tensor-core sums reorder
    for ga, gb in ((a[2], b[2]), (a[3], b[3])):
        assert np.abs(gb).max() > 0
        assert np.abs(ga - gb).max() <= tol * np.abs(gb).max(), np.abs(ga - gb).max() / np.abs(gb).max()
    for pa, pb in ((a[4], b[4]), (a[5], b[5])):
        assert np.abs(pa - pb).max() <= 2e-3 * 5e-4 + (0 if prec_name == "fp32" else 1e-3)   # Adam steps are +-lr sized


def test_trainer_graphed_steps_equal_eager_steps():
    """Trainer.step_graphed (eager pass, capture, replays with the Adam schedule on the device) walks the same parameter
    trajectory as Trainer.step with the host-side schedule."""
    trainer_mod = __import__("importlib").import_module("spin-nerf_b200.trainer")
    g = load_golden("render")
    rng = np.random.default_rng(3)
    rays = g["rays"]
    n = rays.shape[1]
    steps = []
    for _ in range(5):
        b = []
        for i in range(3):
            ix = rng.permutation(n)[:40]
            b.append(torch.from_numpy(np.ascontiguousarray(rays[:, ix])).pin_memory())
            b.append(torch.from_numpy(rng.uniform(0, 1, (len(ix), 3) if i < 2 else (len(ix),)).astype(np.float32)).pin_memory())
        steps.append(b)
    res = []
    for mode in ("graph", "eager"):
        netc, _ = make_net(11, spn.PREC_BF16, 1.0); netf, _ = make_net(12, spn.PREC_BF16, 1.0)
        tr = trainer_mod.Trainer(netc, netf, lr=5e-4, lrate_decay=1, N_samples=64, N_importance=64, lindisp=True,
                                 white_bkgd=True, perturb=0.0, raw_noise_std=0.0, near=1.2, far=8.0)
        losses = []
        for b in steps:
            if mode == "graph":
                loss, _ = tr.step_graphed(*b)
            else:
                loss, _ = tr.step(*[t.to(DEV) for t in b])
            losses.append(float(loss))
        torch.cuda.synchronize()
        res.append((losses, N(netc.flat_params()), N(netf.flat_params()), tr.global_step))
    assert res[0][3] == res[1][3] == 5
    assert np.allclose(res[0][0], res[1][0], rtol=2e-3), (res[0][0], res[1][0])
    for a, b in ((res[0][1], res[1][1]), (res[0][2], res[1][2])):
        assert np.abs(a - b).max() <= 1.5e-3          # 5 Adam steps of <= 5e-4 each; bf16 tensor-core sums may reorder


def test_trainer_step_from_pool_equals_explicit_batches():
    """The device-resident sampler path (one gather kernel) feeds the step the same rays / targets as explicit batches."""
    trainer_mod = __import__("importlib").import_module("spin-nerf_b200.trainer")
    g = load_golden("render")
    rng = np.random.default_rng(9)
    pool = T(np.ascontiguousarray(g["rays"]))                     # [2, M, 3]
    M = pool.shape[1]
    rgb_pool = T(rng.uniform(0, 1, (M, 3)).astype(np.float32)); disp_pool = T(rng.uniform(0, 1, M).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, M, (3, 72))).to(DEV)
    res = []
    for mode in ("pool", "explicit"):
        netc, _ = make_net(11, spn.PREC_FP32, 1.0); netf, _ = make_net(12, spn.PREC_FP32, 1.0)
        tr = trainer_mod.Trainer(netc, netf, N_samples=64, N_importance=64, lindisp=True, white_bkgd=True, perturb=0.0,
                                 raw_noise_std=0.0, near=1.2, far=8.0)
        if mode == "pool":
            loss, _ = tr.step_from_pool(pool, rgb_pool, disp_pool, idx)
        else:
            loss, _ = tr.step(pool[:, idx[0]], rgb_pool[idx[0]], pool[:, idx[1]], rgb_pool[idx[1]], pool[:, idx[2]], disp_pool[idx[2]])
        torch.cuda.synchronize()
        res.append((float(loss), N(tr.grads[0]), N(tr.grads[1])))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0])
    for a, b in ((res[0][1], res[1][1]), (res[0][2], res[1][2])):     # same rays, same kernels: only atomic-add order differs
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()


def test_render_bf16_psnr_vs_reference():
    g = load_golden("render")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    netc, _ = make_net(11, spn.PREC_BF16, 1.0)
    netf, _ = make_net(12, spn.PREC_BF16, 1.0)
    with torch.no_grad():
        rgb, disp, acc, depth, ex = spn.render(H, W, f, chunk=32768, rays=T(g["rays"]), retraw=True, use_viewdirs=True,
                                               ndc=False, near=1.2, far=8.0, network_query_fn=None, network_fn=netc,
                                               network_fine=netf, N_samples=64, N_importance=64, lindisp=True,
                                               white_bkgd=True, perturb=0., raw_noise_std=0.)
    ref = g["det_lindisp_white__rgb"]
    mse = float(((N(rgb) - ref) ** 2).mean())
    assert -10 * np.log10(mse) > 35.0, -10 * np.log10(mse)          # bf16 render vs the fp32 reference render
    assert np.abs(N(ex["rgb0"]) - g["det_lindisp_white__rgb0"]).max() < 3e-2


def test_adam_flat_matches_torch():
    rng = np.random.default_rng(4)
    p = rng.standard_normal(10007).astype(np.float32); gr = rng.standard_normal(10007).astype(np.float32)
    tp = torch.nn.Parameter(T(p.copy())); opt = torch.optim.Adam([tp], lr=3e-3)
    mp, m, v = T(p.copy()), torch.zeros(10007, device=DEV), torch.zeros(10007, device=DEV)
    for step in range(1, 4):
        tp.grad = T(gr * step); opt.step()
        ops.adam_step(mp, T(gr * step), m, v, step, 3e-3)
    close(mp, N(tp), rtol=1e-5, atol=1e-6)
    po, m0, v0 = O.adam_step(p, gr, np.zeros_like(p), np.zeros_like(p), 1, 3e-3)
    mp2, m2, v2 = T(p.copy()), torch.zeros(10007, device=DEV), torch.zeros(10007, device=DEV)
    ops.adam_step(mp2, T(gr), m2, v2, 1, 3e-3)
    close(mp2, po, rtol=1e-5, atol=1e-6)


def test_errors_are_loud():
    with pytest.raises(RuntimeError):
        ops.embed(torch.zeros(4, 3), 10)                      # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        ops.sample_pdf(torch.zeros(2, 300, device=DEV), torch.zeros(2, 299, device=DEV), 8, det=True)   # nb > 256
    with pytest.raises(NotImplementedError):
        spn.NeRF(D=4, W=128, input_ch=63, input_ch_views=27, use_viewdirs=True)
