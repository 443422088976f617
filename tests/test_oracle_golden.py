"""Pins oracle/nerf_oracle.py against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import nerf_oracle as O
from conftest import load_golden


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def close_mostly(a, b, rtol, atol, max_frac=0.01, hard=5e-2):
    """sample_pdf amplifies 1-ulp differences where a cdf step is ~1e-5 (see
    test_sample_pdf_indices_bit_exact); everything downstream of the fine z_vals is therefore
    compared as: all but a small fraction within tolerance, and nothing wildly off."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    err = np.abs(a - b); tol = atol + rtol * np.abs(b)
    assert np.mean(err > tol) <= max_frac, (np.mean(err > tol), err.max())
    assert err.max() <= hard * max(1.0, np.abs(b).max()), err.max()


def test_embed_matches_reference():
    g = load_golden("embed")
    # fp32 sin/cos of arguments up to 512*x: libm vs torch vectorised sin differ in the last ulps
    close(O.embed(g["x"], 10), g["e10"], rtol=0, atol=2e-6)
    close(O.embed(g["x"], 4), g["e4"], rtol=0, atol=2e-6)
    assert O.embed(g["x"], 10).shape[-1] == 63 and O.embed(g["x"], 4).shape[-1] == 27


def test_mlp_forward_and_param_grads():
    g = load_golden("mlp")
    p = O.init_params(11)
    raw, acts = O.mlp_forward(p, g["x90"], keep=True)
    close(raw, g["raw"], rtol=1e-4, atol=2e-5)
    grads = O.mlp_backward(p, acts, g["draw"])
    assert sum(v.size for v in p.values()) == 595844
    for k, gv in grads.items():
        assert gv.shape == p[k].shape
        close(gv.reshape(-1)[::97], g["g_sub__" + k], rtol=2e-3, atol=2e-4)
        assert abs(np.abs(gv).sum(dtype=np.float64) - g["g_abs__" + k]) <= 1e-3 * g["g_abs__" + k] + 1e-4


@pytest.mark.parametrize("tag,white,detach", [("plain", False, False), ("white", True, False),
                                              ("noise_detach", True, True)])
def test_raw2outputs_forward_backward(tag, white, detach):
    g = {k.split("__", 1)[1]: v for k, v in load_golden("raw2outputs").items() if k.startswith(tag + "__")}
    rgb, disp, acc, w, depth, alpha = O.raw2outputs(g["raw"], g["z"], g["rd"], g["noise"], white, True)
    close(w, g["w"], atol=1e-6); close(alpha, g["alpha"], atol=1e-6)
    close(rgb, g["rgb"], atol=2e-6); close(acc, g["acc"], atol=2e-6)
    close(depth, g["depth"], rtol=1e-5); close(disp, g["disp"], rtol=1e-4)
    d_raw = O.raw2outputs_backward(g["raw"], g["z"], g["rd"], g["g_rgb"], g["g_disp"], g["g_acc"], g["g_w"],
                                   g["g_depth"], g["noise"], white, detach)
    close(d_raw, g["d_raw"], rtol=2e-3, atol=2e-4 * np.abs(g["d_raw"]).max())


def test_raw2outputs_hand_case():
    g = load_golden("raw2outputs")
    raw = np.zeros((1, 4, 4), np.float32); raw[..., 3] = 1
    out = O.raw2outputs(raw, np.array([[1, 2, 3, 4]], np.float32), np.array([[2, 0, 0]], np.float32))
    close(out[3], g["hand__w"]); close(out[2], g["hand__acc"]); close(out[4], g["hand__depth"])
    close(out[3][0], [0.86466, 0.11702, 0.015837, 0.0024788], rtol=1e-4)


def test_sample_pdf_indices_bit_exact():
    g = load_golden("sample_pdf")
    cdf = O.pdf_to_cdf(g["weights"])
    # torch's vectorised fp32 `sum` rounds differently from numpy's (1 ulp on the row total), so the
    # oracle's own cdf is within 2 ulp of the reference's; the prefix-sum emulation itself is exact:
    close(cdf, g["cdf"], rtol=0, atol=4e-7)
    w = g["weights"] + np.float32(1e-5)
    import torch
    tot = torch.sum(torch.from_numpy(w), -1, keepdim=True).numpy()
    assert np.array_equal(np.cumsum((w / tot).astype(np.float64), -1).astype(np.float32), g["cdf"][:, 1:])
    # bit-exact indices GIVEN identical cdf and u
    s_det, i_det = O.sample_pdf_from_cdf(g["bins"], g["cdf"], g["u_det"])
    s_sto, i_sto = O.sample_pdf_from_cdf(g["bins"], g["cdf"], g["u_sto"])
    assert np.array_equal(i_det, g["inds_det"]) and np.array_equal(i_sto, g["inds_sto"])
    close(s_det, g["det"], rtol=1e-6, atol=1e-6); close(s_sto, g["sto"], rtol=1e-6, atol=1e-6)
    # and through the oracle's own cdf the samples still agree to fp32 tolerance
    # ... except where denom = cdf[i]-cdf[i-1] is ~1e-5: there t = (u-cdf)/denom amplifies a 1-ulp
    # change of u or cdf by up to 6e-8/1e-5; those outliers are rare and bounded by the bin width
    mine = O.sample_pdf(g["bins"], g["weights"], 64, det=False, u=g["u_det"])[0]
    err = np.abs(mine - g["det"])
    assert np.mean(err > 1e-5) < 0.01 and err.max() < 2e-2
    assert np.abs(O.linspace01(64) - g["u_det"][0]).max() <= 6e-8


def test_searchsorted_kat():
    g = load_golden("searchsorted_kat")
    _, inds = O.sample_pdf_from_cdf(np.zeros_like(g["cdf"]), g["cdf"], g["u"])
    assert inds.tolist() == g["inds"].tolist() == [[1, 3, 4, 4, 5, 5]]


def test_rays():
    g = load_golden("rays")
    ro, rd = O.get_rays(12, 16, 14.4, g["c2w"])
    close(ro, g["ro"]); close(rd, g["rd"]); close(ro, g["ro_np"]); close(rd, g["rd_np"])
    no, nd = O.ndc_rays(12, 16, 14.4, 1.0, ro, rd)
    close(no, g["ndc_o"], rtol=1e-5, atol=1e-6); close(nd, g["ndc_d"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["det_lindisp_white", "det_ndc", "sto_lindisp_white", "coarse_only"])
def test_render_end_to_end(tag):
    g = load_golden("render")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    pc = O.init_params(11); pc["alpha_linear.bias"] = pc["alpha_linear.bias"] + 1.0
    pf = O.init_params(12); pf["alpha_linear.bias"] = pf["alpha_linear.bias"] + 1.0
    ndc = tag == "det_ndc"
    near, far = (0.0, 1.0) if ndc else (1.2, 8.0)
    rb = O.make_ray_batch(g["rays"][0], g["rays"][1], near, far, ndc=ndc, H=H, W=W, focal=f)
    kw = dict(lindisp="lindisp" in tag, white_bkgd="white" in tag)
    n_imp = 0 if tag == "coarse_only" else 64
    if tag.startswith("sto"):
        # pytest=True re-seeds numpy before every draw (run_nerf.py:662-666, helpers:320-327,376-380)
        def draw(*shape):
            np.random.seed(0)
            return np.random.rand(*shape).astype(np.float32)
        kw.update(t_rand=draw(48, 64), u=draw(48, 64), noise0=draw(48, 64) * 1.0, noise1=draw(48, 128) * 1.0)
    else:
        kw.update(u=np.broadcast_to(g["lin64"], (48, 64)))
    out = O.render_rays(rb, pc, pf if n_imp else None, 64, n_imp, retraw=True, need_alpha=bool(n_imp),
                        t_vals=g["lin64"], **kw)
    G = lambda k: g[f"{tag}__{k}"]
    cm = close if n_imp == 0 else close_mostly
    cm(out["z_vals"], G("z_vals"), rtol=2e-5, atol=1e-5)
    cm(out["raw"], G("raw"), rtol=1e-3, atol=1e-3)
    cm(out["weights"], G("weights"), rtol=0, atol=2e-4)
    cm(out["rgb_map"], G("rgb"), rtol=0, atol=2e-4); cm(out["acc_map"], G("acc"), rtol=0, atol=2e-4)
    cm(out["depth_map"], G("depth"), rtol=2e-4, atol=2e-4); cm(out["disp_map"], G("disp"), rtol=5e-4, atol=0)
    if n_imp:
        close(out["rgb0"], G("rgb0"), atol=2e-4)
        cm(out["z_std"], G("z_std"), rtol=1e-3, atol=1e-4)
        cm(out["alpha"], G("alpha"), rtol=0, atol=2e-4); close(out["alpha0"], G("alpha0"), atol=2e-4)


@pytest.mark.parametrize("tag", ["depths", "c2w_patch", "staticcam", "rgb_net", "no_coarse"])
def test_render_call_variants(tag):
    """render()'s other call forms (tests/golden/make_variants_golden.py ran the unmodified reference): a depth column,
    rays from c2w with a patch window, c2w_staticcam, a NeRF_RGB fine network over a frozen density provider, --no_coarse."""
    g = load_golden("render_variants")
    H, W, f = 12, 16, 14.4
    pc = O.init_params(11); pc["alpha_linear.bias"] = pc["alpha_linear.bias"] + 1.0
    pf = O.init_params(12); pf["alpha_linear.bias"] = pf["alpha_linear.bias"] + 1.0
    pr = O.init_params(13)
    ro, rd = g["rays"][0], g["rays"][1]
    kw = {}
    if tag == "depths":
        rb = O.make_ray_batch(ro, rd, 1.2, 8.0, depths=g["depths"])
        assert rb.shape[1] == 12
    elif tag == "c2w_patch":
        fo, fd = O.get_rays(H, W, f, g["pose_a"])
        rb = O.make_ray_batch(fo[3:9, 5:13].reshape(-1, 3), fd[3:9, 5:13].reshape(-1, 3), 1.2, 8.0)
    elif tag == "staticcam":
        _, view_d = O.get_rays(H, W, f, g["pose_a"])                 # view directions from c2w ...
        so, sd = O.get_rays(H, W, f, g["pose_b"])                    # ... rays from the static camera (run_nerf.py:128-133)
        rb = O.make_ray_batch(so.reshape(-1, 3), sd.reshape(-1, 3), 1.2, 8.0)
        vd = view_d.reshape(-1, 3)
        rb[:, -3:] = vd / np.linalg.norm(vd, axis=-1, keepdims=True)
    else:
        rb = O.make_ray_batch(ro, rd, 1.2, 8.0)
    n = rb.shape[0]
    kw = dict(lindisp=True, white_bkgd=True, u=np.broadcast_to(O.linspace01(64), (n, 64)), retraw=True)
    if tag == "rgb_net":
        out = O.render_rays(rb, pc, pr, 64, 64, need_alpha=True, p_alpha=pf, **kw)
    elif tag == "no_coarse":
        out = O.render_rays(rb, None, pr, 64, 64, p_alpha=pf, **kw)
    else:
        out = O.render_rays(rb, pc, pf, 64, 64, **kw)
    G = lambda k: g[f"{tag}__{k}"].reshape(n, *g[f"{tag}__{k}"].shape[(2 if tag in ("c2w_patch", "staticcam") else 1):])
    close_mostly(out["rgb_map"], G("rgb"), rtol=0, atol=2e-4); close_mostly(out["acc_map"], G("acc"), rtol=0, atol=2e-4)
    close_mostly(out["depth_map"], G("depth"), rtol=2e-4, atol=2e-4); close_mostly(out["disp_map"], G("disp"), rtol=5e-4, atol=0)
    close(out["rgb0"], G("rgb0"), atol=2e-4); close(out["disp0"], G("disp0"), rtol=5e-4)
    close_mostly(out["z_std"], G("z_std"), rtol=1e-3, atol=1e-4)
    if tag in ("depths", "rgb_net", "no_coarse"):
        close_mostly(out["z_vals"], G("z_vals"), rtol=2e-5, atol=1e-5)
        close_mostly(out["raw"], G("raw"), rtol=1e-3, atol=1e-3)
        close_mostly(out["weights"], G("weights"), rtol=0, atol=2e-4)
    if tag == "rgb_net":
        close_mostly(out["alpha"], G("alpha"), rtol=0, atol=2e-4); close(out["alpha0"], G("alpha0"), atol=2e-4)
    if tag == "depths":      # the depth column is carried, not used (sigma_loss aside): same render as without it
        ref = O.render_rays(O.make_ray_batch(ro, rd, 1.2, 8.0), pc, pf, 64, 64, **kw)
        np.testing.assert_array_equal(out["rgb_map"], ref["rgb_map"])


def test_train_step_oracle_matches_reference_autograd():
    """oracle/train_oracle.py (loss, both nets' gradients, two Adam steps) vs autograd + torch.optim.Adam run
    through the reference's render() (golden train_step)."""
    from oracle import train_oracle as TO
    g = load_golden("render"); tr = load_golden("train_step")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    pc = O.init_params(11); pc["alpha_linear.bias"] = pc["alpha_linear.bias"] + 1.0
    pf = O.init_params(12); pf["alpha_linear.bias"] = pf["alpha_linear.bias"] + 1.0
    rb = O.make_ray_batch(g["rays"][0], g["rays"][1], 1.2, 8.0)
    sc, sf = TO.AdamState(pc), TO.AdamState(pf)
    u = np.broadcast_to(g["lin64"], (48, 64))
    for it in range(2):
        loss, gc, gf = TO.train_step(rb, tr["target"], tr["tdisp"], pc, pf, sc, sf, 5e-4, t_vals=g["lin64"], u=u)
        assert abs(loss - float(tr[f"loss{it}"])) <= 1e-4 * abs(float(tr[f"loss{it}"]))
        if it == 0:
            for nm, gr in (("c", gc), ("f", gf)):
                for k, gv in gr.items():
                    ref_abs = float(tr[f"g_abs__{nm}__{k}"])
                    assert abs(np.abs(gv).sum(dtype=np.float64) - ref_abs) <= 5e-3 * ref_abs + 1e-7, (nm, k)
                    close_mostly(gv.reshape(-1)[::997], tr[f"g_sub__{nm}__{k}"], rtol=1e-2,
                                 atol=1e-3 * np.abs(gv).max() + 1e-9, max_frac=0.03, hard=1.0)
    for nm, p in (("c", pc), ("f", pf)):
        for k, v in p.items():
            close_mostly(v.reshape(-1)[::997], tr[f"p_sub__{nm}__{k}"], rtol=0, atol=2e-4, max_frac=0.03, hard=1.0)


def test_one_chunk_step_equals_three_render_calls():
    """The reformulation the GPU trainer relies on (rays are independent; losses and detach_weights per ray range): in the
    numpy oracle the gradients of one concatenated render equal the summed gradients of the reference's three calls."""
    from oracle import train_oracle as TO
    g = load_golden("render")
    rng = np.random.default_rng(1)
    rays = g["rays"]
    n = rays.shape[1]
    batches = []
    for i, m in enumerate((10, 7, 9)):
        ix = rng.permutation(n)[:m]
        rb = O.make_ray_batch(rays[0][ix], rays[1][ix], 1.2, 8.0)
        batches.append((rb, rng.uniform(0, 1, (len(ix), 3) if i < 2 else (len(ix),)).astype(np.float32)))
    pc, pf = O.init_params(11), O.init_params(12)
    for p in (pc, pf):
        p["alpha_linear.bias"] = p["alpha_linear.bias"] + np.float32(1.0)
    l3, gc3, gf3 = TO.spin_step_grads(batches, pc, pf, one_chunk=False)
    l1, gc1, gf1 = TO.spin_step_grads(batches, pc, pf, one_chunk=True)
    assert abs(l1 - l3) <= 1e-6 * abs(l3)
    for a, b in ((gc1, gc3), (gf1, gf3)):
        for k in a:
            scale = np.abs(b[k]).max()
            assert np.abs(a[k] - b[k]).max() <= 2e-5 * max(scale, 1e-12), k


def test_spin_step_trajectory_starts_like_the_references():
    """tests/golden/convergence.npz: the unmodified reference trained for 300 steps of the SPIn-NeRF step (three render
    calls, six MSE terms, Adam, lr decay) on a small analytic scene.  The oracle's one-chunk formulation of the step + Adam
    walks the same trajectory (first steps here; over all 300 steps the oracle's final loss was within 0.05 % when the
    fixture was made).  The GPU trainer is held to the same trajectory in tests/test_zz_gpu_next_rows.py."""
    import importlib.util
    import os
    from conftest import GOLDEN
    from oracle import train_oracle as TO
    spec = importlib.util.spec_from_file_location("make_convergence_golden", os.path.join(GOLDEN, "make_convergence_golden.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    gold = load_golden("convergence")
    ro, rd, rgb_t, disp_t, idx = gen.problem()
    assert int(idx.sum()) == int(gold["idx_checksum"][0])          # the numpy stream reproduced the fixture's batches
    pc, pf = gen.params()
    sc, sf = TO.AdamState(pc), TO.AdamState(pf)
    for it in range(3):
        b = [(O.make_ray_batch(ro[idx[it, k]], rd[idx[it, k]], gen.NEAR, gen.FAR), (rgb_t if k < 2 else disp_t)[idx[it, k]])
             for k in range(3)]
        loss, gc, gf = TO.spin_step_grads(b, pc, pf, one_chunk=True)
        assert abs(loss - float(gold["loss"][it])) <= 1e-4 * float(gold["loss"][it]), (it, loss, float(gold["loss"][it]))
        lr = gen.LR * 0.1 ** (max(it - 1, 0) / (gen.DECAY * 1000))     # step k = it+1 runs at lr0 * 0.1^((k-2)/decay_steps)
        for p, g, st in ((pc, gc, sc), (pf, gf, sf)):
            st.step += 1
            for k in p:
                p[k], st.m[k], st.v[k] = O.adam_step(p[k], g[k], st.m[k], st.v[k], st.step, lr)


@pytest.mark.parametrize("one_chunk", [False, True])
def test_spin_step_with_sparse_depth_rays_matches_reference(one_chunk):
    """The shipped config's step (colmap_depth + depth_loss, DS_NeRF/configs/config.txt): four reference render() calls and
    loss += 0.1 * img2mse(depth_col, target_depth) (tests/golden/train_step_depth.npz) against the oracle, as four calls
    and as the one-chunk formulation Trainer.step launches."""
    import importlib.util
    import os
    from conftest import GOLDEN
    from oracle import train_oracle as TO
    spec = importlib.util.spec_from_file_location("make_depth_step_golden", os.path.join(GOLDEN, "make_depth_step_golden.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    gold = load_golden("train_step_depth")
    pc, pf = gen.params()
    batches = [(O.make_ray_batch(r[0], r[1], gen.NEAR, gen.FAR), t) for r, t in gen.problem()]
    loss, gc, gf = TO.spin_step_grads(batches, pc, pf, one_chunk=one_chunk, depth_lambda=gen.DEPTH_LAMBDA)
    assert abs(loss - float(gold["loss"])) <= 1e-4 * float(gold["loss"]), (loss, float(gold["loss"]))
    for nm, grads in (("c", gc), ("f", gf)):
        for k, gv in grads.items():
            ref_abs = float(gold[f"g_abs__{nm}__{k}"])
            assert abs(np.abs(gv).sum(dtype=np.float64) - ref_abs) <= 5e-3 * ref_abs + 1e-7, (nm, k)
            close_mostly(gv.reshape(-1)[::997], gold[f"g_sub__{nm}__{k}"], rtol=1e-2, atol=1e-3 * np.abs(gv).max() + 1e-9,
                         max_frac=0.03, hard=1.0)


def test_torch_port_render_matches_reference():
    """oracle/torch_port.py (the PyTorch restatement bench.py times on the GPU as `gpu_reference_port`) against the
    unmodified reference's render golden."""
    import torch
    from oracle import torch_port as TP
    g = load_golden("render")
    pc = O.init_params(11); pc["alpha_linear.bias"] = pc["alpha_linear.bias"] + 1.0
    pf = O.init_params(12); pf["alpha_linear.bias"] = pf["alpha_linear.bias"] + 1.0
    with torch.no_grad():
        out = TP.render_rays(torch.from_numpy(g["rays"][0]), torch.from_numpy(g["rays"][1]), 1.2, 8.0,
                             TP.make_params(pc, "cpu"), TP.make_params(pf, "cpu"))
    G = lambda k: g[f"det_lindisp_white__{k}"]
    N = lambda t: t.numpy()
    close_mostly(N(out["z_vals"]), G("z_vals"), rtol=2e-5, atol=1e-5)
    close_mostly(N(out["weights"]), G("weights"), rtol=0, atol=2e-4)
    close_mostly(N(out["rgb_map"]), G("rgb"), rtol=0, atol=2e-4); close_mostly(N(out["disp_map"]), G("disp"), rtol=5e-4, atol=0)
    close(N(out["rgb0"]), G("rgb0"), atol=2e-4); close(N(out["disp0"]), G("disp0"), rtol=5e-4)


def test_torch_port_walks_the_references_training_trajectory():
    """... and its train step (autograd + torch.optim.Adam + the reference's schedule) against the first steps of the
    reference's 300-step run (tests/golden/convergence.npz)."""
    import importlib.util
    import os
    import torch
    from conftest import GOLDEN
    from oracle import torch_port as TP
    spec = importlib.util.spec_from_file_location("make_convergence_golden", os.path.join(GOLDEN, "make_convergence_golden.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    gold = load_golden("convergence")
    ro, rd, rgb_t, disp_t, idx = gen.problem()
    pc, pf = (TP.make_params(p, "cpu") for p in gen.params())
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=gen.LR, betas=(0.9, 0.999))
    T = torch.from_numpy
    for it in range(4):
        b = [(T(np.stack([ro[idx[it, k]], rd[idx[it, k]]], 0)), T((rgb_t if k < 2 else disp_t)[idx[it, k]])) for k in range(3)]
        opt.zero_grad()
        loss, psnr = TP.spin_step_loss(b, pc, pf, gen.NEAR, gen.FAR)
        loss.backward()
        opt.step()
        for group in opt.param_groups:
            group["lr"] = gen.LR * (0.1 ** (it / (gen.DECAY * 1000)))
        assert abs(float(loss) - float(gold["loss"][it])) <= 1e-4 * float(gold["loss"][it]), (it, float(loss))
        assert abs(float(psnr) - float(gold["psnr"][it])) <= 1e-3


def test_torch_port_full_size_step_matches_reference():
    """BASELINE configs[1] size (3 x 1024 rays, coarse + fine): the PyTorch restatement that bench.py times as the CPU arm
    reproduces the unmodified reference's maps and step loss (tests/golden/make_fullsize_golden.py)."""
    import torch
    from conftest import load_golden
    from oracle import torch_port as TP
    g = load_golden("fullsize_step")
    H, W, f, near, far, n_rand, seed_c, seed_f = [float(x) for x in g["cfg"]]
    ps = []
    for seed in (int(seed_c), int(seed_f)):
        p = O.init_params(seed)
        p["alpha_linear.bias"] = p["alpha_linear.bias"] + np.float32(1.0)
        ps.append(TP.make_params(p, "cpu"))
    torch.set_num_threads(8)
    rays = torch.from_numpy(g["rays"])
    with torch.no_grad():
        out = TP.render_rays(rays[0, 0], rays[0, 1], near, far, ps[0], ps[1], lindisp=True, white_bkgd=True)
    for k, rk in (("rgb_map", "rgb"), ("rgb0", "rgb0"), ("acc_map", "acc")):
        assert np.abs(out[k].numpy() - g["clf__" + rk]).max() <= 2e-4, k
    batches = [(rays[0], torch.from_numpy(g["target_clf"])), (rays[1], torch.from_numpy(g["target_s"])),
               (rays[2], torch.from_numpy(g["depth_inp"]))]
    with torch.no_grad():
        loss, _ = TP.spin_step_loss(batches, ps[0], ps[1], near, far, perturb=False, raw_noise_std=0.0)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * float(g["loss"])
