"""render / batchify_rays / render_rays / render_path with the reference's signatures and return
structure (DS_NeRF/run_nerf.py:74-307, 593-737), executed by the fused CUDA chunk pipeline
(spn_render_rays_fwd / spn_render_rays_bwd): sample -> encode -> MLP -> composite -> resample ->
encode -> MLP -> composite in one C call per ray chunk, with its own autograd node.
"""
from __future__ import annotations

import ctypes as C
import os
import random

import numpy as np
import torch

from . import _lib as L
from . import ops
from ._lib import check, f32, lib, ptr, stream
from .frame_io import FrameWriter, dump_frame, to8b  # noqa: F401  (to8b re-exported: run_nerf.py uses it next to render_path)
from .embed import get_embedder
from .nerf import NeRF, NeRF_RGB

_DIFF = ("rgb_map", "disp_map", "acc_map", "depth_map", "weights", "rgb0", "disp0", "acc0")


def _flags(lindisp, white_bkgd, detach_weights, perturb, need_alpha):
    return ((L.F_LINDISP if lindisp else 0) | (L.F_WHITE_BKGD if white_bkgd else 0) |
            (L.F_DETACH_WEIGHTS if detach_weights else 0) | (L.F_PERTURB if perturb else 0) |
            (L.F_NEED_ALPHA if need_alpha else 0))


_GRAD_NAMES = (("g_rgb", "rgb_map"), ("g_disp", "disp_map"), ("g_acc", "acc_map"), ("g_depth", "depth_map"),
               ("g_weights", "weights"), ("g_rgb0", "rgb0"), ("g_disp0", "disp0"), ("g_acc0", "acc0"))


def _bind_io(keep, net_c, net_f):
    flat_c, packed_c = net_c._sync()
    flat_f, packed_f = (net_f._sync() if net_f is not None else (None, None))
    io = L.RenderIO()
    for k, v in keep.items():
        setattr(io, k, ptr(v))
    io.params_coarse, io.packed_coarse = ptr(flat_c), ptr(packed_c)
    io.params_fine, io.packed_fine = ptr(flat_f), ptr(packed_f)
    return io, (flat_c, packed_c, flat_f, packed_f)


def chunk_forward(opts, rays, net_c, net_f, t_rand=None, u=None, noise0=None, noise1=None, train=False, pool=None):
    """spn_render_rays_fwd on one ray chunk.  Returns (cfg, keep): `keep` holds every output / saved buffer
    by its spn_render_io field name.  `pool(name, shape, dtype)` may supply reusable buffers."""
    rays = f32(rays)
    n, ncols = rays.shape
    S, NI = opts["N_samples"], opts["N_importance"]
    S2 = S + NI
    prec = net_c.precision
    dev = rays.device
    if pool is None:
        E = lambda name, *sh: torch.empty(sh, device=dev, dtype=torch.float32)
        stash = lambda name, m: ops.mlp_stash(m, prec, dev)
    else:
        E = lambda name, *sh: pool(name, sh, torch.float32)
        stash = lambda name, m: pool(name, (int(lib().spn_mlp_stash_bytes(int(m), int(prec))),), torch.uint8)
    cfg = L.RenderCfg(n, ncols, S, NI, _flags(opts["lindisp"], opts["white_bkgd"], opts["detach_weights"],
                                             opts["perturb"], opts["need_alpha"]), prec, float(opts["raw_noise_std"]))
    keep = dict(rays=rays, t_rand=t_rand, u=u, noise0=noise0, noise1=noise1,
                rgb_map=E("rgb_map", n, 3), disp_map=E("disp_map", n), acc_map=E("acc_map", n),
                depth_map=E("depth_map", n), weights=E("weights", n, S2), z_vals=E("z_vals", n, S2),
                raw=E("raw", n, S2, 4))
    if NI > 0:
        keep.update(rgb0=E("rgb0", n, 3), disp0=E("disp0", n), acc0=E("acc0", n), z_std=E("z_std", n),
                    z_coarse=E("z_coarse", n, S), raw_coarse=E("raw_coarse", n, S, 4))
    if opts["need_alpha"]:
        keep.update(alpha=E("alpha", n, S2), alpha0=E("alpha0", n, S))
    if train or prec == L.PREC_FP32:          # fp32 mode materialises activations even for inference
        keep["stash_coarse"] = stash("stash_coarse", n * S)
        if NI > 0:
            keep["stash_fine"] = stash("stash_fine", n * S2)
    io, hold = _bind_io(keep, net_c, net_f)
    check(lib().spn_render_rays_fwd(C.byref(cfg), C.byref(io), stream()), "spn_render_rays_fwd")
    return cfg, keep


def chunk_backward(cfg, keep, net_c, net_f, g, gc, gf, scratch=None, ws=None, detach_range=(0, 0)):
    """spn_render_rays_bwd: accumulates d(loss)/d(params) into the flat buffers gc / gf.
    g: dict output-name -> upstream gradient tensor (missing = zero).  detach_range: rays [a, b) of the chunk that the
    caller rendered with detach_weights=True (run_nerf_helpers.py:385-388) when several render calls share the chunk."""
    dev = keep["rays"].device
    n, S2 = cfg.n_rays, cfg.n_samples + cfg.n_importance
    io, hold = _bind_io(keep, net_c, net_f)
    gr = L.RenderGrads()
    live = []
    for cname, k in _GRAD_NAMES:
        t = g.get(k)
        if t is not None:
            t = f32(t); live.append(t)
            setattr(gr, cname, ptr(t))
    if scratch is None:
        scratch = torch.empty(n * S2 * 4, device=dev)
    if ws is None:
        ws = ops.mlp_bwd_workspace(n * S2, cfg.precision, dev)
    gr.grads_coarse, gr.grads_fine = ptr(gc), ptr(gf)
    gr.d_raw_scratch, gr.workspace = ptr(scratch), ptr(ws)
    gr.detach_begin, gr.detach_end = int(detach_range[0]), int(detach_range[1])
    check(lib().spn_render_rays_bwd(C.byref(cfg), C.byref(io), C.byref(gr), stream()), "spn_render_rays_bwd")


class RenderChunk(torch.autograd.Function):
    """One ray chunk through spn_render_rays_fwd; backward = spn_render_rays_bwd into flat grads."""

    @staticmethod
    def forward(ctx, opts, rays, net_c, net_f, t_rand, u, noise0, noise1, *params):
        train = opts["train"]       # decided by the caller: grad mode is off inside Function.forward
        cfg, keep = chunk_forward(opts, rays, net_c, net_f, t_rand, u, noise0, noise1, train)
        names = ["rgb_map", "disp_map", "acc_map", "depth_map", "weights", "z_vals", "raw"]
        if opts["N_importance"] > 0:
            names += ["rgb0", "disp0", "acc0", "z_std"]
        if opts["need_alpha"]:
            names += ["alpha", "alpha0"]
        ctx.names = names
        outs = tuple(keep[k] for k in names)
        # The Function's own outputs go through save_for_backward: kept as plain attributes they would close the cycle
        # output -> grad_fn -> ctx -> output, which Python's collector cannot see through the C++ node — a forward whose backward
        # never runs would then leak the whole chunk's activation stash.  Only non-output state stays on ctx.
        ctx.save_for_backward(*outs)
        ctx.state = (cfg, {k: v for k, v in keep.items() if k not in names}, net_c, net_f, train)
        ctx.mark_non_differentiable(*[keep[k] for k in names if k not in _DIFF])
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        cfg, rest, net_c, net_f, train = ctx.state
        if not train:
            raise RuntimeError("RenderChunk.backward without a training forward")
        keep = dict(rest)
        keep.update(zip(ctx.names, ctx.saved_tensors))
        dev = keep["rays"].device
        gc = torch.zeros(L.MLP_NPARAMS, device=dev)
        gf = torch.zeros(L.MLP_NPARAMS, device=dev) if net_f is not None else None
        chunk_backward(cfg, keep, net_c, net_f, dict(zip(ctx.names, gouts)), gc, gf)
        ctx.state = None
        grads = [gc[o:o + p.numel()].view(p.shape) for o, p in zip(net_c._offsets, net_c._flat_params())]
        if net_f is not None:
            grads += [gf[o:o + p.numel()].view(p.shape) for o, p in zip(net_f._offsets, net_f._flat_params())]
        return (None,) * 8 + tuple(grads)


def batchify(fn, chunk):
    """run_nerf.py:44-53: `fn` applied to slices of `chunk` rows."""
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_nerf.py:56-71 with the reference's arguments: flatten the points, embed them and the per-ray view directions
    (broadcast over the samples), concatenate, evaluate `fn` in slices of `netchunk`, restore the leading shape.  With the
    lazy embedders of embed.get_embedder the concatenation is the raw [pt, viewdir] 6-vector and NeRF.forward encodes and
    evaluates it in one kernel; materialised embeddings (90 columns) are accepted as well."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


def default_query_fn(netchunk=1024 * 64):
    """The `network_query_fn` create_nerf builds (run_nerf.py:427-430) for the reference's default embedders."""
    embed_fn, _ = get_embedder(10, 0)
    embeddirs_fn, _ = get_embedder(4, 0)
    return lambda inputs, viewdirs, network_fn: run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn,
                                                            embeddirs_fn=embeddirs_fn, netchunk=netchunk)


def create_nerf(args, device=None):
    """run_nerf.py:380-496 for the `--no_tcnn` model: embedders, coarse / fine networks (NeRF, or NeRF_RGB over a frozen
    density provider loaded from --alpha_model_path, optionally --no_coarse), the query function, torch.optim.Adam over
    the trainable parameters, checkpoint reload from --ft_path or the newest *.tar of basedir/expname, and the train /
    test render kwargs.  Returns (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer) like the reference;
    checkpoints interchange with it (same state_dict keys).  `--sigma_loss` is outside this path and raises."""
    if device is None:
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    shape = dict(input_ch=input_ch, output_ch=output_ch, skips=[4], input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs)
    alpha_model = None
    if getattr(args, "alpha_model_path", None) is None:
        model = NeRF(D=args.netdepth, W=args.netwidth, **shape).to(device)
        grad_vars = list(model.parameters())
    else:
        alpha_model = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, **shape).to(device)
        alpha_model.load_state_dict(torch.load(args.alpha_model_path, map_location=device)['network_fine_state_dict'])
        if not args.no_coarse:
            model = NeRF_RGB(D=args.netdepth, W=args.netwidth, alpha_model=alpha_model, **shape).to(device)
            grad_vars = list(model.parameters())
        else:
            model, grad_vars = None, []
    model_fine = None
    if args.N_importance > 0:
        if alpha_model is None:
            model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, **shape).to(device)
        else:
            model_fine = NeRF_RGB(D=args.netdepth_fine, W=args.netwidth_fine, alpha_model=alpha_model, **shape).to(device)
        grad_vars += list(model_fine.parameters())

    def network_query_fn(inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=args.netchunk)
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    logdir = os.path.join(args.basedir, args.expname)
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        ckpts = [os.path.join(logdir, f) for f in sorted(os.listdir(logdir)) if 'tar' in f] if os.path.isdir(logdir) else []
    if len(ckpts) > 0 and not args.no_reload:
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        if model is not None:
            model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

    render_kwargs_train = {'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
                           'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model,
                           'use_viewdirs': args.use_viewdirs, 'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std}
    if args.dataset_type != 'llff' or args.no_ndc:      # NDC only suits forward-facing LLFF captures (run_nerf.py:478-483)
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    else:
        render_kwargs_train['ndc'] = True
    render_kwargs_test = dict(render_kwargs_train, perturb=False, raw_noise_std=0.)
    if getattr(args, "sigma_loss", False):
        raise NotImplementedError("sigma_loss (DS_NeRF/loss.py) is outside the B200 hot path")
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


def _np_rand(*shape, device):
    """The reference's pytest=True stream: re-seeded before every draw (run_nerf.py:662-666)."""
    np.random.seed(0)
    return torch.as_tensor(np.random.rand(*shape), dtype=torch.float32, device=device)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., pytest=False, sigma_loss=None,
                verbose=False, need_alpha=False, detach_weights=False):
    """Volumetric rendering of one chunk — same arguments and returned dict as run_nerf.py:593-737.
    `network_query_fn` is not called: encoding + MLP run inside the fused kernels."""
    if sigma_loss is not None:
        raise NotImplementedError("sigma_loss (DS_NeRF/loss.py) is outside the B200 hot path")
    if need_alpha and N_importance <= 0:
        # the reference builds ret['alpha0'] from a name that only exists after a fine pass (run_nerf.py:696, 719-721)
        raise NameError("name 'alpha0' is not defined (need_alpha=True requires N_importance > 0, as in the reference)")
    fused = type(network_fn) is NeRF and (network_fine is None or type(network_fine) is NeRF)
    if not fused:
        return render_rays_composed(ray_batch, network_fn, network_query_fn, N_samples, retraw, lindisp, perturb,
                                    N_importance, network_fine, white_bkgd, raw_noise_std, pytest, need_alpha,
                                    detach_weights)
    if ray_batch.shape[-1] <= 9:
        raise NotImplementedError("spinnerf_b200 renders with use_viewdirs=True ray batches (11|12 columns)")
    n = ray_batch.shape[0]
    dev = ray_batch.device
    S2 = N_samples + N_importance
    t_rand = u = noise0 = noise1 = None
    if perturb > 0.:
        t_rand = _np_rand(n, N_samples, device=dev) if pytest else torch.rand(n, N_samples, device=dev)
        if N_importance > 0:
            u = _np_rand(n, N_importance, device=dev) if pytest else torch.rand(n, N_importance, device=dev)
    if raw_noise_std > 0.:
        noise0 = _np_rand(n, N_samples, device=dev) if pytest else torch.randn(n, N_samples, device=dev)
        if N_importance > 0:
            noise1 = _np_rand(n, S2, device=dev) if pytest else torch.randn(n, S2, device=dev)
    opts = dict(N_samples=int(N_samples), N_importance=int(N_importance), lindisp=bool(lindisp),
                white_bkgd=bool(white_bkgd), detach_weights=bool(detach_weights), perturb=perturb > 0.,
                need_alpha=bool(need_alpha), raw_noise_std=float(raw_noise_std))
    params = network_fn._flat_params() + (network_fine._flat_params() if network_fine is not None else [])
    opts["train"] = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    outs = RenderChunk.apply(opts, ray_batch, network_fn, network_fine, t_rand, u, noise0, noise1, *params)
    names = ["rgb_map", "disp_map", "acc_map", "depth_map", "weights", "z_vals", "raw"]
    if N_importance > 0:
        names += ["rgb0", "disp0", "acc0", "z_std"]
    if need_alpha:
        names += ["alpha", "alpha0"]
    ret = dict(zip(names, outs))
    if not retraw:
        ret.pop("raw")
    return ret


def render_rays_composed(ray_batch, network_fn, network_query_fn, N_samples, retraw, lindisp, perturb, N_importance,
                         network_fine, white_bkgd, raw_noise_std, pytest, need_alpha, detach_weights):
    """Operator-level composition for the variants the fused chunk does not cover (NeRF_RGB /
    --no_coarse / alpha_model, run_nerf.py:680-692): same CUDA ops, Python control flow."""
    n = ray_batch.shape[0]
    dev = ray_batch.device
    if network_query_fn is None:
        network_query_fn = default_query_fn()
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 9 else None
    t_rand = None
    if perturb > 0.:
        t_rand = _np_rand(n, N_samples, device=dev) if pytest else torch.rand(n, N_samples, device=dev)
    z_vals = ops.sample_z(ray_batch, N_samples, lindisp, t_rand)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    if network_fn is not None:
        coarse = network_fn
    else:
        coarse = network_fine.alpha_model if getattr(network_fine, "alpha_model", None) is not None else network_fine
    raw = network_query_fn(pts, viewdirs, coarse)
    rgb_map, disp_map, acc_map, weights, depth_map, alpha = ops.raw2outputs(
        raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest=pytest, need_alpha=need_alpha, detach_weights=detach_weights)
    ret = {}
    if N_importance > 0:
        rgb0, disp0, acc0, alpha0 = rgb_map, disp_map, acc_map, alpha
        u = None
        if perturb > 0.:
            u = _np_rand(n, N_importance, device=dev) if pytest else torch.rand(n, N_importance, device=dev)
        z_vals, z_std, _, _ = ops.resample(z_vals, weights, N_importance, u)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        run_fn = network_fn if network_fine is None else network_fine
        raw = network_query_fn(pts, viewdirs, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map, alpha = ops.raw2outputs(
            raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest=pytest, need_alpha=need_alpha,
            detach_weights=detach_weights)
        ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, z_std=z_std)
        if need_alpha:
            ret["alpha0"] = alpha0
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map, depth_map=depth_map, weights=weights, z_vals=z_vals)
    if retraw:
        ret["raw"] = raw
    if need_alpha:
        ret["alpha"] = alpha
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, need_alpha=False, detach_weights=False, **kwargs):
    """run_nerf.py:74-87."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, depths=None, need_alpha=False, detach_weights=False, patch=None, **kwargs):
    """run_nerf.py:90-165 — same arguments, same [rgb, disp, acc, depth, extras] return."""
    if not use_viewdirs:
        raise NotImplementedError("spinnerf_b200 implements the use_viewdirs=True path")
    if c2w is not None:
        c2w = torch.as_tensor(c2w)
        if not c2w.is_cuda:     # the reference keeps its poses on the default (CUDA) device; a host pose is 48 bytes to move
            net = kwargs.get("network_fn") or kwargs.get("network_fine")
            c2w = c2w.to(next(net.parameters()).device)
        rays_o, rays_d = ops.get_rays(H, W, focal, c2w, patch)
    else:
        rays_o, rays_d = rays
    view_src = rays_d
    if c2w_staticcam is not None:
        rays_o, rays_d = ops.get_rays(H, W, focal, c2w_staticcam)
    sh = rays_d.shape
    if c2w_staticcam is None and depths is None and not torch.is_tensor(near) and not torch.is_tensor(far):
        ray_mat = ops.build_ray_batch(rays_o, rays_d, near, far, ndc, H, W, focal)
    else:   # general form (run_nerf.py:128-153)
        viewdirs = f32(view_src).reshape(-1, 3)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        if ndc:
            rays_o, rays_d = ops.ndc_rays(H, W, focal, 1., rays_o, rays_d)
        rays_o = f32(rays_o).reshape(-1, 3); rays_d = f32(rays_d).reshape(-1, 3)
        cols = [rays_o, rays_d, near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])]
        if depths is not None:
            cols.append(depths.reshape(-1, 1).float())
        ray_mat = torch.cat(cols + [viewdirs], -1)
    all_ret = batchify_rays(ray_mat, chunk, need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map', 'depth_map']
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


def render_path(render_poses, hwf, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                disp_require_grad=False, need_alpha=False, rgb_require_grad=False, detach_weights=False,
                patch_len=None, masks=None):
    """run_nerf.py:168-307: per-pose render loop (rays generated on the device from c2w), optional
    LPIPS patch sampling inside the mask bbox (:197-211), optional per-frame dumps (:231-295).

    Without gradients (video / test renders) frames leave through frame_io.FrameWriter: device->host copies on a side
    stream into double-buffered pinned memory and file writes on a worker thread, overlapping the next frame's kernels
     — the reference blocks on `.cpu()` and on the files after every frame (:221-295)."""
    H, W, focal = hwf
    if render_factor != 0:
        H = H // render_factor; W = W // render_factor; focal = focal / render_factor
    H, W = int(H), int(W)
    if savedir is not None:
        np.savetxt(os.path.join(savedir, 'intrinsics.txt'), np.array([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1]]))
    with_grad = disp_require_grad or rgb_require_grad
    writer = None if with_grad else FrameWriter(savedir, gt_imgs, need_alpha)
    rgbs, disps, Xs, Ys = [], [], [], []
    for i, c2w in enumerate(render_poses):
        c2w = torch.as_tensor(c2w)
        if with_grad:
            patch = None
            if patch_len is not None:
                masked = np.where(masks[i] != 0)
                masked = (masked[0] // render_factor, masked[1] // render_factor)
                Xs.append(random.randint(masked[0].min(), max(masked[0].max() - patch_len[0], masked[0].min())))
                Ys.append(random.randint(masked[1].min(), max(masked[1].max() - patch_len[1], masked[1].min())))
                patch = (Xs[-1], Ys[-1], patch_len[0], patch_len[1])
            rgb, disp, acc, depth, extras = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], retraw=True,
                                                   need_alpha=need_alpha, detach_weights=detach_weights, patch=patch,
                                                   **render_kwargs)
            disps.append(disp if disp_require_grad else disp.detach().cpu().numpy())
            rgbs.append(rgb if rgb_require_grad else rgb.detach().cpu().numpy())
            if savedir is not None:
                N = lambda t: t.detach().cpu().numpy()
                dump_frame(savedir, i, N(rgb), None if gt_imgs is None else gt_imgs[i], N(depth), N(disp),
                           N(extras['weights']), N(extras['z_vals']), N(extras['alpha']) if need_alpha else None, N(c2w))
        else:
            with torch.no_grad():
                rgb, disp, acc, depth, extras = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], retraw=True,
                                                       need_alpha=need_alpha, **render_kwargs)
            writer.submit(i, c2w, dict(rgb=rgb, disp=disp, depth=depth, weights=extras['weights'], z_vals=extras['z_vals'],
                                       alpha=extras.get('alpha')))
    if writer is not None:
        rgbs, disps = writer.close()
        return rgbs, disps, (Xs, Ys)
    disps = torch.stack(disps, 0) if disp_require_grad else np.stack(disps, 0)
    rgbs = torch.stack(rgbs, 0) if rgb_require_grad else np.stack(rgbs, 0)
    return rgbs, disps, (Xs, Ys)


def render_path_sharded(render_poses, hwf, chunk, render_kwargs, render_factor=0, group=None, dst=None):
    """Config 5 (BASELINE.json): novel-view video over several GPUs.  Frames are independent, so rank r renders frames
    r, r+W, ... (dist.frames_for_rank) and keeps them ON THE DEVICE as [rgb | disp] rows; one collective
    (dist.gather_frames: all-gather over NVLink, 12.2 MB per 1008x756 frame) assembles the video, then ONE device->host
    copy hands it over — the reference's per-frame `.cpu()` (run_nerf.py:221-229) happens once per video.
    dst=None: every rank returns the whole video; dst=r: only rank r does (the others return None, None).
    Single process: the same code without the collective."""
    import torch.distributed as dist
    from .dist import frames_for_rank, gather_frames
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    H, W, focal = hwf
    if render_factor != 0:
        H = H // render_factor; W = W // render_factor; focal = focal / render_factor
    H, W = int(H), int(W)
    n = len(render_poses)
    mine = frames_for_rank(n, rank, world)
    net = render_kwargs.get("network_fn") or render_kwargs.get("network_fine")
    dev = next(net.parameters()).device
    local = torch.zeros((len(frames_for_rank(n, 0, world)), H, W, 4), device=dev)      # rank 0 has the most frames
    with torch.no_grad():
        for j, i in enumerate(mine):
            c2w = torch.as_tensor(render_poses[i])
            rgb, disp, _, _, _ = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)
            local[j, ..., :3] = rgb
            local[j, ..., 3] = disp
    video = gather_frames(local, n, rank, world, group, dst)
    if video is None:
        return None, None
    video = video.cpu().numpy()
    return np.ascontiguousarray(video[..., :3]), np.ascontiguousarray(video[..., 3])
