"""CPU oracle of one SPIn-NeRF train step on the render hot path (numpy; test infrastructure only, see
oracle/nerf_oracle.py header).  Restates what autograd + torch.optim.Adam do in DS_NeRF/run_nerf.py:1455-1622
for the terms that involve the hot path: MSE on rgb / rgb0 (img2mse, helpers:15) and MSE on disp / disp0
(nn.MSELoss, run_nerf.py:1517-1519).  Pinned by tests/golden/train_step.npz.
"""
from __future__ import annotations

import numpy as np

from . import nerf_oracle as O

F32 = np.float32


def render_with_grads(rb, pc, pf, n_samples, n_importance, lindisp, white_bkgd, g_out, detach_weights=False,
                      t_vals=None, u=None, t_rand=None, noise0=None, noise1=None):
    """Forward through render_rays (run_nerf.py:593-737) and backward to both networks' parameters.

    g_out(outputs) -> dict of upstream gradients for any of rgb_map, disp_map, acc_map, depth_map, rgb0,
    disp0, acc0.  Returns (outputs, grads_coarse, grads_fine)."""
    rb = np.asarray(rb, F32)
    o, d, vd = rb[:, 0:3], rb[:, 3:6], rb[:, -3:]
    z0 = O.sample_z(rb[:, 6], rb[:, 7], n_samples, lindisp, t_rand, t_vals)

    def net(p, z):
        n, s = z.shape
        pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).astype(F32)
        x = np.concatenate([O.embed(pts.reshape(-1, 3), 10),
                            O.embed(np.broadcast_to(vd[:, None, :], pts.shape).reshape(-1, 3), 4)], -1)
        raw, acts = O.mlp_forward(p, x, keep=True)
        return raw.reshape(n, s, 4), acts

    raw0, acts0 = net(pc, z0)
    rgb0, disp0, acc0, w0, depth0, _ = O.raw2outputs(raw0, z0, d, noise0, white_bkgd)
    mid = F32(0.5) * (z0[:, 1:] + z0[:, :-1])
    zs, _ = O.sample_pdf(mid, w0[:, 1:-1], n_importance, det=(u is None), u=u)
    z1 = O.merge_sorted(z0, zs)
    raw1, acts1 = net(pf, z1)
    rgb, disp, acc, w, depth, _ = O.raw2outputs(raw1, z1, d, noise1, white_bkgd)
    outs = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, weights=w, z_vals=z1, rgb0=rgb0, disp0=disp0,
                acc0=acc0)
    g = g_out(outs)
    d_raw1 = O.raw2outputs_backward(raw1, z1, d, g.get("rgb_map"), g.get("disp_map"), g.get("acc_map"), None,
                                    g.get("depth_map"), noise1, white_bkgd, detach_weights)
    gf = O.mlp_backward(pf, acts1, d_raw1.reshape(-1, 4))
    # z_samples are detached (run_nerf.py:700): the coarse net only sees rgb0 / disp0 / acc0
    d_raw0 = O.raw2outputs_backward(raw0, z0, d, g.get("rgb0"), g.get("disp0"), g.get("acc0"), None, None, noise0,
                                    white_bkgd, detach_weights)
    gc = O.mlp_backward(pc, acts0, d_raw0.reshape(-1, 4))
    return outs, gc, gf


def mse_grad(x, y):
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    return float(np.mean((x - y) ** 2)), (2.0 * (x - y) / x.size).astype(F32)


class AdamState:
    def __init__(self, params):
        self.m = {k: np.zeros_like(v) for k, v in params.items()}
        self.v = {k: np.zeros_like(v) for k, v in params.items()}
        self.step = 0


def train_step(rb, target, tdisp, pc, pf, st_c, st_f, lr, n_samples=64, n_importance=64, lindisp=True,
               white_bkgd=True, t_vals=None, u=None):
    """loss = mse(rgb,t) + mse(rgb0,t) + mse(disp,td) + mse(disp0,td); Adam(lr, betas=(0.9,0.999)) on both nets
    (the golden's train_step case).  Updates pc / pf / states in place, returns (loss, grads_c, grads_f)."""
    box = {}

    def g_out(o):
        l1, g1 = mse_grad(o["rgb_map"], target); l2, g2 = mse_grad(o["rgb0"], target)
        l3, g3 = mse_grad(o["disp_map"], tdisp); l4, g4 = mse_grad(o["disp0"], tdisp)
        box["loss"] = l1 + l2 + l3 + l4
        return dict(rgb_map=g1, rgb0=g2, disp_map=g3, disp0=g4)

    _, gc, gf = render_with_grads(rb, pc, pf, n_samples, n_importance, lindisp, white_bkgd, g_out, t_vals=t_vals, u=u)
    for p, g, st in ((pc, gc, st_c), (pf, gf, st_f)):
        st.step += 1
        for k in p:
            p[k], st.m[k], st.v[k] = O.adam_step(p[k], g[k], st.m[k], st.v[k], st.step, lr)
    return box["loss"], gc, gf


def spin_step_grads(batches, pc, pf, one_chunk, n_samples=64, n_importance=64, lindisp=True, white_bkgd=True,
                    depth_lambda=0.1):
    """Loss and parameter gradients of the SPIn-NeRF step (run_nerf.py:1455-1521, default flags) for
    batches = [(rb_unmasked, rgb_target), (rb_masked, rgb_target), (rb_inpainted, disp_target)] and, optionally, a fourth
    (rb_sparse_depth, depth_target) for `--colmap_depth --depth_loss` (run_nerf.py:1475-1477, 1491-1506: loss +=
    depth_lambda * img2mse(depth_map, target_depth), fine depth only):
      one_chunk=False: three render calls as in the reference — MSE on rgb/rgb0 for the first two (the second with
                       detach_weights=True), MSE on disp/disp0 for the third — gradients summed;
      one_chunk=True:  the three ray batches concatenated into one render call, every loss term and detach_weights
                       applied to its ray range (what spin-nerf_b200/trainer.py:Trainer.step launches on the GPU).
    Deterministic sampling (perturb=0, no noise).  Returns (loss, grads_coarse, grads_fine)."""
    (rb1, t1), (rb2, t2), (rb3, t3) = batches[:3]
    n1, n2, n3 = len(rb1), len(rb2), len(rb3)
    rb4, t4 = batches[3] if len(batches) > 3 else (np.zeros((0, rb1.shape[1]), F32), np.zeros((0,), F32))
    n4 = len(rb4)
    if not one_chunk:
        total, gc_sum, gf_sum = 0.0, None, None
        for i, (rb, tgt) in enumerate(batches):
            box = {}

            def g_out(o, i=i, tgt=tgt, box=box):
                if i == 3:
                    l, g = mse_grad(o["depth_map"], tgt)
                    box["loss"] = depth_lambda * l
                    return {"depth_map": (F32(depth_lambda) * g).astype(F32)}
                a, b = ("rgb_map", "rgb0") if i < 2 else ("disp_map", "disp0")
                la, ga = mse_grad(o[a], tgt); lb, gb = mse_grad(o[b], tgt)
                box["loss"] = la + lb
                return {a: ga, b: gb}
            _, gc, gf = render_with_grads(rb, pc, pf, n_samples, n_importance, lindisp, white_bkgd, g_out,
                                          detach_weights=(i == 1))
            total += box["loss"]
            gc_sum = gc if gc_sum is None else {k: gc_sum[k] + gc[k] for k in gc}
            gf_sum = gf if gf_sum is None else {k: gf_sum[k] + gf[k] for k in gf}
        return total, gc_sum, gf_sum
    rb = np.concatenate([rb1, rb2, rb3, rb4], 0)
    box = {}

    def g_out(o):
        g = {k: np.zeros_like(o[k]) for k in ("rgb_map", "rgb0", "disp_map", "disp0", "depth_map")}
        loss = 0.0
        for lo, hi, tgt, keys in ((0, n1, t1, ("rgb_map", "rgb0")), (n1, n1 + n2, t2, ("rgb_map", "rgb0")),
                                  (n1 + n2, n1 + n2 + n3, t3, ("disp_map", "disp0"))):
            for k in keys:
                l, gk = mse_grad(o[k][lo:hi], tgt)
                loss += l
                g[k][lo:hi] = gk
        if n4:
            m = n1 + n2 + n3
            l, gk = mse_grad(o["depth_map"][m:], t4)
            loss += depth_lambda * l
            g["depth_map"][m:] = F32(depth_lambda) * gk
        box["loss"] = loss
        return g
    detach = np.zeros(n1 + n2 + n3 + n4, bool)
    detach[n1:n1 + n2] = True
    _, gc, gf = render_with_grads(rb, pc, pf, n_samples, n_importance, lindisp, white_bkgd, g_out, detach_weights=detach)
    return box["loss"], gc, gf
