"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference/DS_NeRF) on CPU.  Run in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, the fixtures do.  MLP weights are NOT stored:
they come from oracle.nerf_oracle.init_params(seed) (numpy PCG64, stable across machines)
and are loaded into the reference nn.Module with load_state_dict.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O      # noqa: E402
from oracle import ref_loader            # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
T = torch.from_numpy


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in kw.items()})
    print("wrote", name, len(kw), "arrays")


def main():
    torch.set_num_threads(8)
    H_, R_ = ref_loader.load()
    rng = np.random.default_rng(1234)

    # ---- a5 embed --------------------------------------------------------------------
    x = (rng.standard_normal((64, 3)) * 3).astype(np.float32)
    e10, d10 = H_.get_embedder(10, 0)
    e4, d4 = H_.get_embedder(4, 0)
    assert d10 == 63 and d4 == 27
    save("embed", x=x, e10=e10(T(x)).numpy(), e4=e4(T(x)).numpy())

    # ---- a6 MLP forward + autograd parameter grads -------------------------------------
    pc = O.init_params(11, scale=1.0)
    net, = ref_loader.reference_nets(pc, None)[:1]
    x90 = np.concatenate([O.embed((rng.standard_normal((192, 3)) * 2).astype(np.float32), 10),
                          O.embed(_unit(rng.standard_normal((192, 3))), 4)], -1)
    raw = net(T(x90))
    draw = rng.standard_normal(raw.shape).astype(np.float32)
    raw.backward(T(draw))
    gsum = {}
    for k, v in net.named_parameters():
        g = v.grad.numpy()
        gsum["g_sum__" + k] = g.sum(dtype=np.float64)
        gsum["g_abs__" + k] = np.abs(g).sum(dtype=np.float64)
        gsum["g_sub__" + k] = g.reshape(-1)[::97].copy()
    save("mlp", x90=x90, raw=raw.detach().numpy(), draw=draw, **gsum)

    # ---- a7 raw2outputs forward + backward ---------------------------------------------
    cases = {}
    n, s = 24, 64
    for tag, (noise_std, white, detach) in {"plain": (0.0, False, False), "white": (0.0, True, False),
                                            "noise_detach": (1.0, True, True)}.items():
        rawc = (rng.standard_normal((n, s, 4)) * 1.5).astype(np.float32)
        rawc[..., 3] += 0.5
        z = np.sort(rng.uniform(1.2, 7.9, (n, s)).astype(np.float32), -1)
        rd = (rng.standard_normal((n, 3))).astype(np.float32)
        rt = T(rawc).requires_grad_(True)
        out = H_.raw2outputs(rt, T(z), T(rd), raw_noise_std=noise_std, white_bkgd=white,
                             pytest=True, need_alpha=True, detach_weights=detach)
        rgb, disp, acc, w, depth, alpha = out
        ups = [rng.standard_normal(tuple(o.shape)).astype(np.float32) for o in (rgb, disp, acc, w, depth)]
        loss = sum((o * T(u)).sum() for o, u in zip((rgb, disp, acc, w, depth), ups))
        loss.backward()
        np.random.seed(0)
        noise = (np.random.rand(n, s) * noise_std).astype(np.float32) if noise_std > 0 else np.zeros((n, s), np.float32)
        for k, v in dict(raw=rawc, z=z, rd=rd, noise=noise, rgb=rgb, disp=disp, acc=acc, w=w, depth=depth,
                         alpha=alpha, g_rgb=ups[0], g_disp=ups[1], g_acc=ups[2], g_w=ups[3], g_depth=ups[4],
                         d_raw=rt.grad).items():
            cases[f"{tag}__{k}"] = v.detach().numpy() if torch.is_tensor(v) else v
    # the hand case of SURVEY.md 8 a7
    rawh = np.zeros((1, 4, 4), np.float32); rawh[..., 3] = 1
    zh = np.array([[1, 2, 3, 4]], np.float32); rdh = np.array([[2, 0, 0]], np.float32)
    oh = H_.raw2outputs(T(rawh), T(zh), T(rdh))
    cases.update(hand__w=oh[3].numpy(), hand__acc=oh[2].numpy(), hand__depth=oh[4].numpy(), hand__disp=oh[1].numpy())
    save("raw2outputs", **cases)

    # ---- a8 sample_pdf -----------------------------------------------------------------
    n = 96
    bins = np.sort(rng.uniform(1.0, 8.0, (n, 63)).astype(np.float32), -1)
    wts = (rng.uniform(0, 1, (n, 62)) ** 4).astype(np.float32)
    wts[:8] = 0.0                                  # degenerate rows: uniform pdf
    wts[8:16, 5:] = 0.0                            # mass concentrated in few bins
    det = H_.sample_pdf(T(bins), T(wts), 64, det=True)
    sto = H_.sample_pdf(T(bins), T(wts), 64, det=False, pytest=True)
    np.random.seed(0)
    u_sto = np.random.rand(n, 64).astype(np.float32)
    # the reference's own cdf / inds (recomputed with the same torch ops, helpers:306-331)
    wt = T(wts) + 1e-5
    cdf = torch.cumsum(wt / torch.sum(wt, -1, keepdim=True), -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u_det = torch.linspace(0., 1., steps=64).expand(n, 64).contiguous()
    inds_det = torch.searchsorted(cdf, u_det, right=True)
    inds_sto = torch.searchsorted(cdf, T(u_sto).contiguous(), right=True)
    save("sample_pdf", bins=bins, weights=wts, cdf=cdf.numpy(), det=det.numpy(), sto=sto.numpy(),
         u_sto=u_sto, u_det=u_det.numpy(), inds_det=inds_det.numpy(), inds_sto=inds_sto.numpy())
    # searchsorted right=True KAT of SURVEY.md 8 a8
    kc = torch.tensor([[0, .2, .2, .7, 1.]]); ku = torch.tensor([[0, .2, .7, .99, 1, 1.1]])
    save("searchsorted_kat", cdf=kc.numpy(), u=ku.numpy(), inds=torch.searchsorted(kc, ku, right=True).numpy())

    # ---- a9 rays -----------------------------------------------------------------------
    c2w = _pose(rng)
    ro, rdd = H_.get_rays(12, 16, 14.4, T(c2w))
    ro_np, rd_np = H_.get_rays_np(12, 16, 14.4, c2w)
    no, nd = H_.ndc_rays(12, 16, 14.4, 1., ro, rdd)
    save("rays", c2w=c2w, ro=ro.numpy(), rd=rdd.numpy(), ro_np=ro_np, rd_np=rd_np, ndc_o=no.numpy(), ndc_d=nd.numpy())

    # ---- a1-a3 render() end to end through the reference -------------------------------
    pf = O.init_params(12, scale=1.0)
    # bias sigma up so transmittance really decays
    pc2 = dict(pc); pc2["alpha_linear.bias"] = pc["alpha_linear.bias"] + 1.0
    pf["alpha_linear.bias"] = pf["alpha_linear.bias"] + 1.0
    netc, netf = ref_loader.reference_nets(pc2, pf)
    e10, _ = H_.get_embedder(10, 0); e4, _ = H_.get_embedder(4, 0)
    qfn = lambda inputs, viewdirs, fn: R_.run_network(inputs, viewdirs, fn, embed_fn=e10, embeddirs_fn=e4, netchunk=65536)
    Hh, Ww, f = 378, 504, 0.9 * 504
    ro, rdd = H_.get_rays_np(Hh, Ww, f, _pose(rng))
    sel = rng.choice(Hh * Ww, 48, replace=False)
    rays = np.stack([ro.reshape(-1, 3)[sel], rdd.reshape(-1, 3)[sel]], 0).astype(np.float32)
    rcases = {"rays": rays, "H": Hh, "W": Ww, "focal": f, "lin64": torch.linspace(0., 1., steps=64).numpy()}
    for tag, kw in {
        "det_lindisp_white": dict(ndc=False, lindisp=True, white_bkgd=True, perturb=0., raw_noise_std=0., near=1.2, far=8.0),
        "det_ndc": dict(ndc=True, white_bkgd=False, perturb=0., raw_noise_std=0., near=0., far=1.),
        "sto_lindisp_white": dict(ndc=False, lindisp=True, white_bkgd=True, perturb=1., raw_noise_std=1., near=1.2, far=8.0, pytest=True),
        "coarse_only": dict(ndc=False, lindisp=False, white_bkgd=False, perturb=0., raw_noise_std=0., near=1.2, far=8.0, N_importance=0),
    }.items():
        kw = dict(kw)
        n_imp = kw.pop("N_importance", 64)
        with torch.no_grad():
            rgb, disp, acc, depth, ex = R_.render(Hh, Ww, f, chunk=32768, rays=T(rays), retraw=True, use_viewdirs=True,
                                                  network_query_fn=qfn, network_fn=netc, network_fine=netf if n_imp else None,
                                                  N_samples=64, N_importance=n_imp, need_alpha=bool(n_imp), **kw)
        rcases.update({f"{tag}__rgb": rgb, f"{tag}__disp": disp, f"{tag}__acc": acc, f"{tag}__depth": depth})
        for k, v in ex.items():
            rcases[f"{tag}__{k}"] = v
    rcases = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in rcases.items()}
    save("render", **rcases)

    # ---- train-step golden: loss + grads through render() + 2 Adam steps -----------------
    netc, netf = ref_loader.reference_nets(pc2, pf)
    params = list(netc.parameters()) + list(netf.parameters())
    opt = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.999))
    target = rng.uniform(0, 1, (48, 3)).astype(np.float32)
    tdisp = rng.uniform(0, 1, (48,)).astype(np.float32)
    tr = {"target": target, "tdisp": tdisp}
    for it in range(2):
        rgb, disp, acc, depth, ex = R_.render(Hh, Ww, f, chunk=32768, rays=T(rays), retraw=True, use_viewdirs=True,
                                              network_query_fn=qfn, network_fn=netc, network_fine=netf, N_samples=64,
                                              N_importance=64, ndc=False, lindisp=True, white_bkgd=True, perturb=0.,
                                              raw_noise_std=0., near=1.2, far=8.0)
        opt.zero_grad()
        loss = H_.img2mse(rgb, T(target)) + H_.img2mse(ex["rgb0"], T(target)) \
            + torch.nn.MSELoss()(disp, T(tdisp)) + torch.nn.MSELoss()(ex["disp0"], T(tdisp))
        loss.backward()
        tr[f"loss{it}"] = loss.item()
        if it == 0:
            for nm, net in (("c", netc), ("f", netf)):
                for k, v in net.named_parameters():
                    g = v.grad.numpy()
                    tr[f"g_sum__{nm}__{k}"] = g.sum(dtype=np.float64)
                    tr[f"g_abs__{nm}__{k}"] = np.abs(g).sum(dtype=np.float64)
                    tr[f"g_sub__{nm}__{k}"] = g.reshape(-1)[::997].copy()
        opt.step()
    for nm, net in (("c", netc), ("f", netf)):
        for k, v in net.named_parameters():
            tr[f"p_sub__{nm}__{k}"] = v.detach().numpy().reshape(-1)[::997].copy()
    save("train_step", **tr)


def _unit(v):
    v = np.asarray(v, np.float32)
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)


def _pose(rng):
    z = _unit(np.array([0.1, -0.05, 1.0]) + rng.standard_normal(3) * 0.05)
    up = np.array([0, 1, 0], np.float32)
    x = _unit(np.cross(up, z)); y = np.cross(z, x)
    pos = np.array([rng.uniform(-.5, .5), rng.uniform(-.5, .5), 0.0], np.float32)
    return np.stack([x, y, z, pos], 1).astype(np.float32)


if __name__ == "__main__":
    main()
