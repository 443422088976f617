"""Torch-facing operators of the B200 hot path.  Each wraps one C-ABI call (include/spinnerf_b200.h);
tensors cross the boundary as device pointers only.  Names/semantics mirror the reference's
DS_NeRF/run_nerf_helpers.py (file:line in each docstring)."""
from __future__ import annotations

import torch

from . import _lib as L
from ._lib import check, f32, lib, ptr, stream


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


# ---------------------------------------------------------------------------------------------
# rays (helpers:249-300)
# ---------------------------------------------------------------------------------------------
def get_rays(H, W, focal, c2w, patch=None):
    """get_rays (helpers:249-260); `patch=(i, j, len1, len2)` slices like render() (run_nerf.py:120-123)."""
    c2w = f32(torch.as_tensor(c2w)[:3, :4])
    H, W = int(H), int(W)
    i0, j0, h, w = (0, 0, H, W) if patch is None else map(int, patch)
    h = max(0, min(h, H - i0)); w = max(0, min(w, W - j0))     # python slicing clamps
    ro = _empty((h, w, 3), c2w); rd = _empty((h, w, 3), c2w)
    if h * w:
        check(lib().spn_get_rays(ptr(c2w), H, W, float(focal), i0, j0, h, w, ptr(ro), ptr(rd), stream()), "spn_get_rays")
    return ro, rd


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """ndc_rays (helpers:283-300)."""
    sh = rays_o.shape
    o = f32(rays_o).reshape(-1, 3); d = f32(rays_d).reshape(-1, 3)
    oo = torch.empty_like(o); od = torch.empty_like(d)
    check(lib().spn_ndc_rays(o.shape[0], int(H), int(W), float(focal), float(near), ptr(o), ptr(d), ptr(oo), ptr(od),
                             stream()), "spn_ndc_rays")
    return oo.reshape(sh), od.reshape(sh)


def build_ray_batch(rays_o, rays_d, near, far, ndc=False, H=0, W=0, focal=1.0):
    """render()'s ray matrix [n,11] = [o, d, near, far, viewdir] (run_nerf.py:126-153)."""
    o = f32(rays_o).reshape(-1, 3); d = f32(rays_d).reshape(-1, 3)
    rays = _empty((o.shape[0], 11), o)
    check(lib().spn_build_ray_batch(o.shape[0], ptr(o), ptr(d), float(near), float(far), int(bool(ndc)), int(H), int(W),
                                    float(focal), ptr(rays), stream()), "spn_build_ray_batch")
    return rays


def embed(x, n_freqs):
    """Positional encoding gamma(x) (helpers:22-70): [..., 3] -> [..., 3 + 6*n_freqs]."""
    sh = x.shape
    xf = f32(x).reshape(-1, 3)
    out = _empty((xf.shape[0], 3 + 6 * n_freqs), xf)
    check(lib().spn_embed(ptr(xf), xf.shape[0], int(n_freqs), ptr(out), stream()), "spn_embed")
    return out.reshape(*sh[:-1], 3 + 6 * n_freqs)


def sample_z(rays, n_samples, lindisp=False, t_rand=None):
    """Stratified depths along each ray (run_nerf.py:646-668). rays [n, >=8]; t_rand [n,S] or None."""
    rays = f32(rays)
    z = _empty((rays.shape[0], n_samples), rays)
    tr = None if t_rand is None else f32(t_rand)
    check(lib().spn_sample_z(ptr(rays), rays.shape[0], rays.shape[1], int(n_samples), int(bool(lindisp)), ptr(tr),
                             ptr(z), stream()), "spn_sample_z")
    return z


# ---------------------------------------------------------------------------------------------
# raw2outputs (helpers:350-401)
# ---------------------------------------------------------------------------------------------
class _Raw2Outputs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise, white_bkgd, need_alpha, detach_weights):
        raw = f32(raw); z = f32(z_vals); rd = f32(rays_d)
        n, s = z.shape
        rgb = _empty((n, 3), raw); disp = _empty((n,), raw); acc = _empty((n,), raw)
        w = _empty((n, s), raw); depth = _empty((n,), raw)
        alpha = _empty((n, s), raw) if need_alpha else None
        nz = None if noise is None else f32(noise)
        check(lib().spn_raw2outputs_fwd(ptr(raw), ptr(z), ptr(rd), 3, ptr(nz), n, s, int(white_bkgd), ptr(rgb),
                                        ptr(disp), ptr(acc), ptr(w), ptr(depth), ptr(alpha), stream()),
              "spn_raw2outputs_fwd")
        ctx.save_for_backward(raw, z, rd, nz)
        ctx.cfg = (int(white_bkgd), int(detach_weights))
        outs = (rgb, disp, acc, w, depth) + ((alpha,) if need_alpha else ())
        if need_alpha:
            ctx.mark_non_differentiable(alpha)
        return outs

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_w, g_depth, *_):
        raw, z, rd, nz = ctx.saved_tensors
        n, s = z.shape
        white, detach = ctx.cfg
        gs = [None if g is None else f32(g) for g in (g_rgb, g_disp, g_acc, g_w, g_depth)]
        d_raw = torch.empty_like(raw)
        check(lib().spn_raw2outputs_bwd(ptr(raw), ptr(z), ptr(rd), 3, ptr(nz), n, s, white, detach, ptr(gs[0]),
                                        ptr(gs[1]), ptr(gs[2]), ptr(gs[3]), ptr(gs[4]), ptr(d_raw), stream()),
              "spn_raw2outputs_bwd")
        return d_raw, None, None, None, None, None, None


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, need_alpha=False,
                detach_weights=False, noise=None):
    """Same signature/returns as the reference (helpers:350-401).  `noise` (unscaled N(0,1) [n,S]) may be
    injected for deterministic tests; pytest=True reproduces the reference's numpy stream (helpers:376-380)."""
    nz = None
    if raw_noise_std > 0.:
        if pytest:
            import numpy as np
            np.random.seed(0)
            nz = torch.as_tensor(np.random.rand(*list(raw[..., 3].shape)), dtype=torch.float32, device=raw.device)
        elif noise is not None:
            nz = noise
        else:
            nz = torch.randn(raw[..., 3].shape, device=raw.device)
        nz = nz * raw_noise_std
    outs = _Raw2Outputs.apply(raw, z_vals, rays_d, nz, bool(white_bkgd), bool(need_alpha), bool(detach_weights))
    rgb, disp, acc, w, depth = outs[:5]
    return rgb, disp, acc, w, depth, (outs[5] if need_alpha else None)


# ---------------------------------------------------------------------------------------------
# sample_pdf (helpers:304-347), sort(cat) (run_nerf.py:702)
# ---------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, det=False, pytest=False, u=None, return_inds=False, return_cdf=False):
    """Inverse-CDF sampling; no gradient (the caller detaches, run_nerf.py:700)."""
    bins = f32(bins.detach()); weights = f32(weights.detach())
    n, nb = bins.shape
    if pytest:      # helpers:318-327
        import numpy as np
        np.random.seed(0)
        if det:
            u = torch.as_tensor(np.broadcast_to(np.linspace(0., 1., N_samples), (n, N_samples)).copy(),
                                dtype=torch.float32, device=bins.device)
        else:
            u = torch.as_tensor(np.random.rand(n, N_samples), dtype=torch.float32, device=bins.device)
    elif u is None and not det:
        u = torch.rand((n, N_samples), device=bins.device)
    uu = None if u is None else f32(u)
    samples = _empty((n, N_samples), bins)
    inds = _empty((n, N_samples), bins, torch.int64) if return_inds else None
    cdf = _empty((n, nb), bins) if return_cdf else None
    check(lib().spn_sample_pdf_cdf(ptr(bins), ptr(weights), ptr(uu), n, nb, int(N_samples), ptr(samples), ptr(inds),
                                   ptr(cdf), stream()), "spn_sample_pdf")
    if return_inds or return_cdf:
        return samples, inds, cdf
    return samples


def searchsorted(a, v, out=None, side='left'):
    """torchsearchsorted.searchsorted (DS_NeRF/torchsearchsorted/src/torchsearchsorted/searchsorted.py:20-53): row-wise
    np.searchsorted of v [Bv,V] in the sorted rows of a [Ba,A] (Ba == Bv or one of them 1) -> int64 [max(Ba,Bv), V]."""
    assert len(a.shape) == 2, "input `a` must be 2-D."
    assert len(v.shape) == 2, "input `v` must be 2-D."
    assert a.shape[0] == v.shape[0] or a.shape[0] == 1 or v.shape[0] == 1, \
        "`a` and `v` must have the same number of rows or one of them must have only one"
    assert a.device == v.device, '`a` and `v` must be on the same device'
    shape = (max(a.shape[0], v.shape[0]), v.shape[1])
    if out is not None:
        assert out.device == a.device and out.dtype == torch.long and tuple(out.shape) == shape, \
            "`out` must be a torch.long tensor of the result shape on the device of `a`"
    else:
        out = torch.empty(shape, device=v.device, dtype=torch.long)
    if a.dtype != torch.float32 or v.dtype != torch.float32:
        raise NotImplementedError("spinnerf_b200.searchsorted is built for float32 (the dtype of the render path); a silent "
                                  "cast could reorder ties")
    af, vf = a.contiguous(), v.contiguous()
    check(lib().spn_searchsorted(ptr(af), ptr(vf), ptr(out), a.shape[0], v.shape[0], a.shape[1], v.shape[1],
                                 1 if side == 'left' else 0, stream()), "spn_searchsorted")
    return out


def merge_sorted(a, b):
    """sort(cat([a, b], -1)) values (run_nerf.py:702)."""
    a = f32(a); b = f32(b)
    out = _empty((a.shape[0], a.shape[1] + b.shape[1]), a)
    check(lib().spn_merge_sorted(ptr(a), ptr(b), a.shape[0], a.shape[1], b.shape[1], ptr(out), stream()),
          "spn_merge_sorted")
    return out


def resample(z_vals, weights, n_importance, u=None, want_samples=False, want_inds=False):
    """run_nerf.py:696-702,726 fused: returns (z_merged [n,S+n_imp], z_std [n], z_samples|None, inds|None)."""
    z = f32(z_vals.detach()); w = f32(weights.detach())
    n, S = z.shape
    uu = None if u is None else f32(u)
    out = _empty((n, S + n_importance), z); std = _empty((n,), z)
    zs = _empty((n, n_importance), z) if want_samples else None
    inds = _empty((n, n_importance), z, torch.int64) if want_inds else None
    check(lib().spn_resample(ptr(z), ptr(w), ptr(uu), n, S, int(n_importance), ptr(out), ptr(zs), ptr(std), ptr(inds),
                             stream()), "spn_resample")
    return out, std, zs, inds


# ---------------------------------------------------------------------------------------------
# flat Adam (run_nerf.py:433-434,1611-1622)
# ---------------------------------------------------------------------------------------------
def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    n = param.numel()
    check(lib().spn_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), n, float(lr), float(betas[0]),
                              float(betas[1]), float(eps), int(step), float(grad_scale), stream()), "spn_adam_step")


# ---------------------------------------------------------------------------------------------
# MLP (helpers:74-127) fused with encoding / run_network (run_nerf.py:56-71)
# ---------------------------------------------------------------------------------------------
def mlp_packed_bytes():
    return int(lib().spn_mlp_packed_bytes())


def mlp_pack(flat_params, out=None):
    flat = f32(flat_params)
    if out is None:
        out = torch.empty(mlp_packed_bytes(), dtype=torch.uint8, device=flat.device)
    check(lib().spn_mlp_pack_weights(ptr(flat), ptr(out), stream()), "spn_mlp_pack_weights")
    return out


def mlp_stash(m, precision, device):
    return torch.empty(int(lib().spn_mlp_stash_bytes(int(m), int(precision))), dtype=torch.uint8, device=device)


def mlp_bwd_workspace(m, precision, device):
    return torch.empty(int(lib().spn_mlp_bwd_workspace_bytes(int(m), int(precision))), dtype=torch.uint8, device=device)


def mlp_forward_points(flat, packed, x6, precision, stash=None):
    x6 = f32(x6)
    raw = _empty((x6.shape[0], 4), x6)
    if precision == L.PREC_FP32 and stash is None:
        stash = mlp_stash(x6.shape[0], precision, x6.device)       # fp32 mode materialises activations
    check(lib().spn_mlp_fwd_points(ptr(flat), ptr(packed), ptr(x6), x6.shape[0], ptr(raw), ptr(stash), int(precision),
                                   stream()), "spn_mlp_fwd_points")
    return raw, stash


def mlp_forward_rays(flat, packed, rays, z, precision, stash=None):
    rays = f32(rays); z = f32(z)
    n, S = z.shape
    raw = _empty((n, S, 4), z)
    if precision == L.PREC_FP32 and stash is None:
        stash = mlp_stash(n * S, precision, z.device)
    check(lib().spn_mlp_fwd_rays(ptr(flat), ptr(packed), ptr(rays), rays.shape[1], ptr(z), n, S, ptr(raw), ptr(stash),
                                 int(precision), stream()), "spn_mlp_fwd_rays")
    return raw, stash


def mlp_backward(flat, packed, stash, d_raw, grads_flat, precision, workspace=None):
    d_raw = f32(d_raw).reshape(-1, 4)
    m = d_raw.shape[0]
    if workspace is None:
        workspace = mlp_bwd_workspace(m, precision, d_raw.device)
    check(lib().spn_mlp_bwd(ptr(flat), ptr(packed), ptr(stash), ptr(d_raw), m, ptr(grads_flat), ptr(workspace),
                            int(precision), stream()), "spn_mlp_bwd")
    return grads_flat
