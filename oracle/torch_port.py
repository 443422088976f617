"""PyTorch restatement of the reference's train step on the render hot path — TEST INFRASTRUCTURE AND BASELINE ONLY
(see oracle/nerf_oracle.py header: nothing under spin-nerf_b200/ imports this).

Why it exists next to the numpy oracle: the reference is Python + PyTorch and cannot travel to the GPU box
(/root/reference is absent there), so the like-for-like baseline SURVEY.md section 8d asks for — the reference's own
formulation (eager PyTorch ops, fp32 GEMMs with TF32 off, autograd, torch.optim.Adam) timed on the same B200 —
needs a restatement that can.  This module is that: the ops of DS_NeRF/run_nerf.py:593-737 and
DS_NeRF/run_nerf_helpers.py:22-127, 304-401 written against torch tensors on any device, pinned on CPU against the same
golden fixtures as the numpy oracle (tests/test_oracle_golden.py::test_torch_port_*).  bench.py times it as
`gpu_reference_port`; it is never part of the product path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def embed(x, n_freqs):
    """helpers:22-52: [x, sin(2^k x), cos(2^k x)] for k = 0..L-1 (3-vector blocks)."""
    out = [x]
    for k in range(n_freqs):
        out += [torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))]
    return torch.cat(out, -1)


def mlp(p, x):
    """NeRF.forward, D=8 W=256 skips=[4] use_viewdirs (helpers:104-127).  p: dict name -> tensor (state_dict names)."""
    pts, views = x[..., :63], x[..., 63:]
    h = pts
    for i in range(8):
        h = F.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i == 4:
            h = torch.cat([pts, h], -1)
    alpha = F.linear(h, p["alpha_linear.weight"], p["alpha_linear.bias"])
    feature = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])
    h = F.relu(F.linear(torch.cat([feature, views], -1), p["views_linears.0.weight"], p["views_linears.0.bias"]))
    rgb = F.linear(h, p["rgb_linear.weight"], p["rgb_linear.bias"])
    return torch.cat([rgb, alpha], -1)


def run_network(p, pts, viewdirs, netchunk=65536):
    """run_nerf.py:56-71 (embedding materialised, evaluated in slices of netchunk rows like the reference)."""
    flat = pts.reshape(-1, 3)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    x = torch.cat([embed(flat, 10), embed(dirs, 4)], -1)
    out = torch.cat([mlp(p, x[i:i + netchunk]) for i in range(0, x.shape[0], netchunk)], 0)
    return out.reshape(*pts.shape[:-1], 4)


def raw2outputs(raw, z_vals, rays_d, noise=None, white_bkgd=False, detach_weights=False):
    """helpers:350-401."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1) * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-F.relu(sigma) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    w_rgb = weights.detach() if detach_weights else weights
    rgb_map = torch.sum(w_rgb[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    disp_map = 1. / torch.clamp(depth_map / acc_map, min=1e-10)        # NaN (acc == 0) propagates like torch.max (helpers:391)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


def sample_pdf(bins, weights, n_samples, u=None):
    """helpers:304-347; u=None is the deterministic linspace (perturb == 0)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0., 1., steps=n_samples, device=bins.device).expand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp(min=0)                       # bin edges on either side of u ...
    hi = hi.clamp(max=cdf.shape[-1] - 1)             # ... clamped to the table (helpers:332-333)
    c_lo, c_hi = cdf.gather(-1, lo), cdf.gather(-1, hi)
    b_lo, b_hi = bins.gather(-1, lo), bins.gather(-1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)       # flat stretches of the cdf (helpers:342-343)
    return b_lo + (u - c_lo) / span * (b_hi - b_lo)


def render_rays(rays_o, rays_d, near, far, pc, pf, n_samples=64, n_importance=64, lindisp=True, white_bkgd=True,
                perturb=False, raw_noise_std=0., detach_weights=False):
    """run_nerf.py:593-737 for [n,3] origins / directions (viewdirs = normalised directions, run_nerf.py:128-134)."""
    n, dev = rays_o.shape[0], rays_o.device
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    t_vals = torch.linspace(0., 1., steps=n_samples, device=dev)
    near_, far_ = near * torch.ones(n, 1, device=dev), far * torch.ones(n, 1, device=dev)
    z = 1. / (1. / near_ * (1. - t_vals) + 1. / far_ * t_vals) if lindisp else near_ * (1. - t_vals) + far_ * t_vals
    if perturb:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper, lower = torch.cat([mids, z[..., -1:]], -1), torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * torch.rand(z.shape, device=dev)
    noise = (lambda s: torch.randn(n, s, device=dev) * raw_noise_std) if raw_noise_std > 0. else (lambda s: None)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    rgb0, disp0, acc0, w0, _ = raw2outputs(run_network(pc, pts, viewdirs), z, rays_d, noise(n_samples), white_bkgd, detach_weights)
    z_mid = .5 * (z[..., 1:] + z[..., :-1])
    u = torch.rand(n, n_importance, device=dev) if perturb else None
    z_samples = sample_pdf(z_mid, w0[..., 1:-1], n_importance, u).detach()
    z, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    rgb, disp, acc, w, depth = raw2outputs(run_network(pf, pts, viewdirs), z, rays_d, noise(n_samples + n_importance), white_bkgd,
                                           detach_weights)
    return dict(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, weights=w, z_vals=z, rgb0=rgb0, disp0=disp0, acc0=acc0)


def spin_step_loss(batches, pc, pf, near, far, **kw):
    """The step's three render calls and six MSE terms (run_nerf.py:1455-1521, default flags).
    batches = [(rays [2,n,3], rgb target), (rays, rgb target), (rays, disparity target)].  Returns (loss, psnr)."""
    mse = lambda a, b: torch.mean((a - b) ** 2)
    (r1, t1), (r2, t2), (r3, t3) = batches
    o1 = render_rays(r1[0], r1[1], near, far, pc, pf, **kw)
    o2 = render_rays(r2[0], r2[1], near, far, pc, pf, detach_weights=True, **kw)
    o3 = render_rays(r3[0], r3[1], near, far, pc, pf, **kw)
    img = mse(o1["rgb_map"], t1)
    psnr = -10. * torch.log10(img)
    loss = img + mse(o2["rgb_map"], t2) + mse(o2["rgb0"], t2) + mse(o1["rgb0"], t1)
    inp = mse(o3["disp_map"], t3) + mse(o3["disp0"], t3)
    return loss + torch.where(torch.isnan(inp), torch.zeros_like(inp), inp), psnr


def make_params(np_params, device):
    return {k: torch.tensor(v, device=device, requires_grad=True) for k, v in np_params.items()}
