"""Fused train step of the SPIn-NeRF hot path (DS_NeRF/run_nerf.py:1455-1521, 1611-1622) without autograd:
three render calls (unmasked rays -> rgb loss; masked rays of the kept view -> rgb loss with detached
weights; inpainted-disparity rays -> disparity loss), analytic loss gradients, backward through the fused
CUDA chunk pipeline into ONE flat fp32 gradient vector per network, optional NCCL all-reduce of that
vector (rays are sharded across ranks), one flat Adam launch per network.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from . import peer
from .dist import RaySharder, allreduce_sum_, allreduce_sum_async
from .render import chunk_backward, chunk_forward


class BufferPool:
    """Reusable device buffers keyed by name (stash / outputs of the step's render calls): the train loop
    allocates once and launches kernels only."""

    def __init__(self, device):
        self.device, self.bufs = device, {}
        self.frozen = False      # set once a CUDA graph has baked the buffers' addresses in: growing one would free memory the graph uses

    def __call__(self, name, shape, dtype):
        need = 1
        for s in shape:
            need *= int(s)
        b = self.bufs.get(name)
        if b is None or b.numel() < need or b.dtype != dtype:
            if self.frozen:
                raise RuntimeError(f"BufferPool: '{name}' would have to grow to {need} elements, but a captured CUDA graph replays on "
                                   "the current buffers; use a second Trainer (or eager steps only) for larger / other work")
            b = torch.empty(max(need, 1), device=self.device, dtype=dtype)
            self.bufs[name] = b
        return b[:need].view(*shape) if need else b[:0].view(*shape)


class Trainer:
    def __init__(self, net_c, net_f, lr=5e-4, lrate_decay=250, N_samples=64, N_importance=64, lindisp=True,
                 white_bkgd=True, perturb=1.0, raw_noise_std=1.0, near=1.2, far=8.0, ndc=False, hwf=None,
                 process_group=None, sharder=None, betas=(0.9, 0.999), eps=1e-8, seed=None):
        self.net_c, self.net_f = net_c, net_f
        self.lr0, self.lrate_decay = float(lr), lrate_decay
        self.betas, self.eps = betas, eps
        self.cfg = dict(N_samples=int(N_samples), N_importance=int(N_importance), lindisp=bool(lindisp),
                        white_bkgd=bool(white_bkgd), perturb=perturb > 0, need_alpha=False,
                        raw_noise_std=float(raw_noise_std))
        self.near, self.far, self.ndc, self.hwf = float(near), float(far), bool(ndc), hwf
        self.pg = process_group
        self.sharder = sharder or RaySharder()
        dev = net_c.flat_params().device
        self.device = dev
        z = lambda: torch.zeros(L.MLP_NPARAMS, device=dev)
        # both networks' flat gradients live in ONE buffer (the fine network's vector starts at the next 512-byte boundary,
        # i.e. with the alignment of a fresh allocation): one memset and one all-reduce per step
        stride = (L.MLP_NPARAMS + 127) // 128 * 128
        self.grad_all = torch.zeros(2 * stride, device=dev)
        self.grads = [self.grad_all[:L.MLP_NPARAMS], self.grad_all[stride:stride + L.MLP_NPARAMS]]
        self.peer = None
        if self.sharder.world > 1 and peer.enabled():
            # opt-in (SPN_P2P_ALLREDUCE=1): the gradient vector lives in an IPC-exported region and the step ends with the
            # fused reduce-scatter / all-gather / Adam over NVLink peer memory (csrc/peer_reduce.cu) instead of NCCL + Adam
            self.peer = peer.PeerGradExchange(L.MLP_NPARAMS, dev, process_group)
            self.grad_all, self.grads = self.peer.grad_all, self.peer.grads
        self.m = [z(), z()]
        self.v = [z(), z()]
        self.global_step = 0
        self.pools = [BufferPool(dev)] * 3      # each render is consumed (backward) before the next starts
        self.shared = BufferPool(dev)
        self.adam_state = None                  # device-resident {step, lr/bc1, sqrt(bc2), lr}: CUDA-graph mode
        # seed: a private random stream for the jitter / resampling / density noise.  With it, a W-rank run consumes the
        # stream of the single-process run: every rank draws the step's GLOBAL random tensors and keeps the rows of its rays
        # (SURVEY.md section 8e "the union of shards equals the single-GPU stream"), so the two runs differ only by the
        # summation order of the gradient all-reduce.  None: torch's default generator, one local draw per rank.
        self.gen = None
        if seed is not None:
            self.gen = torch.Generator(device=dev)
            self.gen.manual_seed(int(seed))
        self._graph = None
        self._segments = None                   # multi-GPU CUDA-graph mode: [graph to the fine backward, coarse backward, Adam]
        self._pending = []

    # ---------------------------------------------------------------------------------------
    def _global_rows(self, sizes_global):
        """Row ids, in the single-process chunk [group 1 | group 2 | ...] of the GLOBAL batch, of this rank's rays."""
        rows, off = [], 0
        for n in sizes_global:
            lo, hi = self.sharder.bounds(int(n))
            rows.append(torch.arange(off + lo, off + hi, device=self.device))
            off += int(n)
        return torch.cat(rows), off

    def _forward(self, idx, rays_od, detach_weights, rb=None, test_kwargs=False, sizes_global=None):
        """rays_od [2,n,3] (origin, direction) or a ready ray matrix rb [n,11] -> chunk state; random draws made on
        the device.  test_kwargs: render like render_kwargs_test (perturb=False, raw_noise_std=0, run_nerf.py:485-488)."""
        if rb is None:
            o, d = rays_od[0], rays_od[1]
            H, W, f = self.hwf if self.hwf is not None else (0, 0, 1.0)
            rb = ops.build_ray_batch(o, d, self.near, self.far, self.ndc, H, W, f)
        n = rb.shape[0]
        S, NI = self.cfg["N_samples"], self.cfg["N_importance"]
        opts = dict(self.cfg, detach_weights=detach_weights)
        if test_kwargs:
            opts.update(perturb=False, raw_noise_std=0.0)
        t_rand = u = n0 = n1 = None
        if self.gen is None:
            draw = lambda fn, cols: fn(n, cols, device=self.device)
        elif self.sharder.world == 1 or sizes_global is None:
            draw = lambda fn, cols: fn(n, cols, device=self.device, generator=self.gen)
        else:
            rows, n_global = self._global_rows(sizes_global)
            draw = lambda fn, cols: fn(n_global, cols, device=self.device, generator=self.gen)[rows]
        if opts["perturb"]:
            t_rand = draw(torch.rand, S)
            u = draw(torch.rand, NI) if NI else None
        if opts["raw_noise_std"] > 0:
            n0 = draw(torch.randn, S)
            n1 = draw(torch.randn, S + NI) if NI else None
        return chunk_forward(opts, rb, self.net_c, self.net_f, t_rand, u, n0, n1, train=True, pool=self.pools[idx])

    def step(self, rays_clf, target_clf, rays_s, target_s, rays_inp, depth_inp, rays_depth=None, target_depth=None,
             depth_lambda=0.1, _apply=True):
        """One optimisation step on this rank's shard of the three ray batches.  Returns the (local) loss and
        psnr like run_nerf.py:1481-1521 defines them for the default flags.

        The reference renders the three batches with three render() calls (run_nerf.py:1455-1470); rays are
        independent, so here they are concatenated into ONE chunk: one sampling / MLP / compositing pass forward and
        one backward over 3x the rays (fewer, fuller kernel launches).  What differs between the calls — which loss
        a ray feeds and detach_weights=True for the masked rays — is applied per ray range (spn_train_losses,
        spn_render_grads.detach_begin/end).

        rays_depth [2,n4,3] / target_depth [n4]: the optional fourth render call of `--colmap_depth --depth_loss`
        (run_nerf.py:1400-1413, 1475-1477, 1491-1506: rays through COLMAP's sparse points, loss += depth_lambda *
        img2mse(depth_map, target_depth)); they ride in the same chunk as a fourth ray range."""
        sh = self.sharder
        sizes_global = [rays_clf.shape[1], rays_s.shape[1], rays_inp.shape[1]] + ([rays_depth.shape[1]] if rays_depth is not None else [])
        rays_clf, target_clf = sh.shard(rays_clf, 1), sh.shard(target_clf)
        rays_s, target_s = sh.shard(rays_s, 1), sh.shard(target_s)
        rays_inp, depth_inp = sh.shard(rays_inp, 1), sh.shard(depth_inp)
        n1, n2, n3 = rays_clf.shape[1], rays_s.shape[1], rays_inp.shape[1]
        if target_clf.shape[0] != n1 or target_s.shape[0] != n2 or depth_inp.shape[0] != n3:
            raise RuntimeError("Trainer.step: every ray batch needs one target per ray")
        groups, n4 = [rays_clf, rays_s, rays_inp], 0
        if rays_depth is not None:
            rays_depth, target_depth = sh.shard(rays_depth, 1), L.f32(sh.shard(target_depth))
            n4 = rays_depth.shape[1]
            if target_depth.shape[0] != n4:
                raise RuntimeError("Trainer.step: every ray batch needs one target per ray")
            groups.append(rays_depth)
        n = n1 + n2 + n3 + n4
        P = self.shared
        rays = P("rays_cat", (2, n, 3), torch.float32)
        torch.cat(groups, 1, out=rays)
        tgt_rgb = P("tgt_rgb", (n1 + n2, 3), torch.float32)
        torch.cat([target_clf, target_s], 0, out=tgt_rgb)
        cfg, k = self._forward(0, rays, False, sizes_global=sizes_global)
        return self._finish_step(cfg, k, tgt_rgb, L.f32(depth_inp), n1, n2, n3, _apply, n4, target_depth, depth_lambda)

    def step_from_pool(self, pool_od, rgb_pool, disp_pool, idx, _apply=True):
        """The same step fed by a device-resident ray pool: pool_od [2,M,3] all rays of the
        scene, rgb_pool [M,3] / disp_pool [M] their colour / inpainted-disparity targets, idx [3,N] int64 the rays
        sampled for the three groups (unmasked, masked, inpainted) of this step.  One gather kernel assembles this
        rank's ray matrix and targets (instead of the reference's per-ray Dataset/DataLoader path)."""
        lo, hi = self.sharder.bounds(idx.shape[1])
        m = hi - lo
        ids = idx if (lo == 0 and hi == idx.shape[1]) else idx[:, lo:hi].contiguous()
        ids = ids.reshape(-1)
        n = 3 * m
        P = self.shared
        rb = P("rays_gathered", (n, 11), torch.float32)
        tgt_rgb, tgt_disp = P("tgt_rgb", (2 * m, 3), torch.float32), P("tgt_disp", (m,), torch.float32)
        H, W, f = self.hwf if self.hwf is not None else (0, 0, 1.0)
        L.check(L.lib().spn_gather_ray_batch(n, L.ptr(pool_od[0]), L.ptr(pool_od[1]), L.ptr(ids), self.near, self.far,
                                             int(self.ndc), int(H), int(W), float(f), L.ptr(rb), L.ptr(rgb_pool),
                                             L.ptr(tgt_rgb), 2 * m, L.ptr(disp_pool), L.ptr(tgt_disp), L.stream()),
                "spn_gather_ray_batch")
        cfg, k = self._forward(0, None, False, rb=rb, sizes_global=[idx.shape[1]] * 3)
        return self._finish_step(cfg, k, tgt_rgb, tgt_disp, m, m, m, _apply)

    def _finish_step(self, cfg, k, tgt_rgb, depth_inp, n1, n2, n3, apply=True, n4=0, target_depth=None, depth_lambda=0.1):
        n = n1 + n2 + n3 + n4
        P = self.shared
        g_rgb, g_rgb0 = P("g_rgb", (n, 3), torch.float32), P("g_rgb0", (n, 3), torch.float32)
        g_disp, g_disp0 = P("g_disp", (n,), torch.float32), P("g_disp0", (n,), torch.float32)
        sums, out = P("loss_sums", (8,), torch.float32), P("loss_out", (8,), torch.float32)
        sums.zero_()
        L.check(L.lib().spn_train_losses(L.ptr(k["rgb_map"]), L.ptr(k["rgb0"]), L.ptr(k["disp_map"]), L.ptr(k["disp0"]),
                                         L.ptr(tgt_rgb), L.ptr(depth_inp), n1, n2, n3, L.ptr(sums), L.ptr(g_rgb),
                                         L.ptr(g_rgb0), L.ptr(g_disp), L.ptr(g_disp0), L.ptr(out), L.stream()),
                "spn_train_losses")
        g = {"rgb_map": g_rgb, "rgb0": g_rgb0, "disp_map": g_disp, "disp0": g_disp0}
        depth_term = None
        if n4:      # sparse-depth rays: only the fine depth_map is supervised (run_nerf.py:1506); the loss kernel filled rows < m
            m = n1 + n2 + n3
            for t in (g_rgb, g_rgb0, g_disp, g_disp0):
                t[m:].zero_()
            diff = k["depth_map"][m:n] - target_depth
            depth_term = float(depth_lambda) * torch.mean(diff * diff)
            g_depth = P("g_depth", (n,), torch.float32)
            g_depth[:m].zero_()
            torch.mul(diff, 2.0 * float(depth_lambda) / n4, out=g_depth[m:])
            g["depth_map"] = g_depth
        self.grad_all.zero_()
        gc, gf = self.grads
        self._pending = []
        split = self.sharder.world > 1 and apply and self.peer is None and cfg.n_importance > 0
        if split:
            # The fine network's backward runs first and is independent of the coarse one (the resampled depths are detached,
            # run_nerf.py:700): its gradient all-reduce is issued as soon as it is complete and travels over NVLink (NCCL's
            # own stream) while the coarse backward computes.
            fine = {kk: v for kk, v in g.items() if kk in ("rgb_map", "disp_map", "depth_map", "acc_map", "weights")}
            coarse = {kk: v for kk, v in g.items() if kk in ("rgb0", "disp0", "acc0")}
            chunk_backward(cfg, k, self.net_c, self.net_f, fine, gc, gf, self._scratch(cfg), self._ws(cfg),
                           detach_range=(n1, n1 + n2))
            self._coarse_args = (cfg, k, coarse, (n1, n1 + n2))
            if self._segments is None:                # eager; in segmented-graph mode the caller runs the rest (step_graphed)
                self._after_fine_backward()
        else:
            chunk_backward(cfg, k, self.net_c, self.net_f, g, gc, gf,
                           self._scratch(cfg), self._ws(cfg), detach_range=(n1, n1 + n2))
            if apply:
                self.apply_gradients()
        res = out[:2].clone()          # `out` is a pooled buffer: hand back copies of (loss, psnr)
        return (res[0] if depth_term is None else res[0] + depth_term), res[1]

    def _after_fine_backward(self):
        """multi-GPU tail of the step: all-reduce(fine) || coarse backward -> all-reduce(coarse) -> Adam"""
        gc, gf = self.grads
        self._pending.append(allreduce_sum_async(gf, self.pg))
        self._coarse_backward()
        self._pending.append(allreduce_sum_async(gc, self.pg))
        self.apply_gradients()

    def _coarse_backward(self):
        cfg, k, coarse, dr = self._coarse_args
        gc, gf = self.grads
        chunk_backward(cfg, k, self.net_c, self.net_f, coarse, gc, gf, self._scratch(cfg), self._ws(cfg), detach_range=dr)

    # ---------------------------------------------------------------------------------------
    # perceptual-loss branch (run_nerf.py:1523-1561, `--lpips`)
    def lpips_patch_backward(self, poses, patches, targets, lpips_fn, hwf_s, batch_size=None, weight=0.01):
        """Renders this rank's share of the step's LPIPS patches and accumulates d(lpips term)/d(params) into the flat
        gradient buffers; returns this rank's part of the term  sum_i LPIPS(pred_i, target_i).mean() / batch_size / 100
        (run_nerf.py:1551-1559) as a device scalar.  Call between `step(..., _apply=False)` and `apply_gradients()`
        (what `step_with_lpips` does).

        poses [B,3,4|5] camera-to-world; patches: B windows (X, Y, len0, len1) on the down-scaled grid hwf_s = (H_s, W_s,
        focal_s) (lpips_patch.patch_geometry / draw_origins); targets: B tensors [1,3,h,w] in [-1,1]
        (lpips_patch.PatchSampler.target_patches); lpips_fn(pred, target) any differentiable torch callable.

        Like the reference the patches are rendered with the TEST kwargs and detach_weights=True (:1541-1549), but the
        rays of each window are generated on the device from c2w (spn_get_rays' patch window) instead of slicing a
        full-frame ray grid (run_nerf.py:119-123), all B patches share ONE ray chunk, and the chunk runs through the
        fused forward / backward without autograd: torch autograd only sees LPIPS itself (d term / d rgb).
        Multi-GPU: whole patches are dealt round-robin to the ranks (LPIPS needs a complete patch on one device); each
        rank's term is scaled by the world size so that the gradient mean over ranks is the global term's gradient."""
        B = len(patches)
        batch_size = B if batch_size is None else int(batch_size)
        Hs, Ws, fs = hwf_s
        mine = list(range(self.sharder.rank, B, self.sharder.world))
        if not mine:
            return torch.zeros((), device=self.device)
        ros, rds, shapes = [], [], []
        for i in mine:
            ro, rd = ops.get_rays(Hs, Ws, fs, torch.as_tensor(poses[i], device=self.device)[:3, :4], patch=patches[i])
            ros.append(ro.reshape(-1, 3)); rds.append(rd.reshape(-1, 3)); shapes.append(tuple(ro.shape[:2]))
        rb = ops.build_ray_batch(torch.cat(ros), torch.cat(rds), self.near, self.far, self.ndc, Hs, Ws, fs)
        cfg, k = self._forward(0, None, True, rb=rb, test_kwargs=True)
        with torch.enable_grad():
            rgb = k["rgb_map"].detach().clone().requires_grad_(True)    # [n,3]: the only tensor torch autograd sees
            term, off = 0.0, 0
            for i, (h, w) in zip(mine, shapes):
                pred = ((rgb[off:off + h * w].view(h, w, 3) - 0.5) * 2).permute(2, 0, 1)[None, ...]      # :1552
                term = term + lpips_fn(pred, targets[i].to(self.device)).mean()
                off += h * w
            term = term * (float(weight) / batch_size)
            g_rgb, = torch.autograd.grad(term * float(self.sharder.world), rgb)
        gc, gf = self.grads
        chunk_backward(cfg, k, self.net_c, self.net_f, {"rgb_map": g_rgb.contiguous()}, gc, gf,
                       self._scratch(cfg), self._ws(cfg))
        return term.detach()

    def step_with_lpips(self, batch, poses, patches, targets, lpips_fn, hwf_s, batch_size=None, from_pool=False):
        """The `--lpips` train step (iterations > 300, run_nerf.py:1523): `batch` = the six tensors of `step` (or the
        four arguments of `step_from_pool` with from_pool=True), then the patch branch, then ONE optimiser step."""
        loss, psnr = (self.step_from_pool if from_pool else self.step)(*batch, _apply=False)
        lp = self.lpips_patch_backward(poses, patches, targets, lpips_fn, hwf_s, batch_size)
        self.apply_gradients()
        return loss + lp, psnr

    def step_three_calls(self, rays_clf, target_clf, rays_s, target_s, rays_inp, depth_inp):
        """The same step as three separate render calls, exactly in the reference's order (kept for parity tests)."""
        sh = self.sharder
        rays_clf, target_clf = sh.shard(rays_clf, 1), sh.shard(target_clf)
        rays_s, target_s = sh.shard(rays_s, 1), sh.shard(target_s)
        rays_inp, depth_inp = sh.shard(rays_inp, 1), sh.shard(depth_inp)
        self.grad_all.zero_()
        gc, gf = self.grads
        mse_g = lambda x, t: (2.0 / x.numel()) * (x - t)
        losses = []
        # 1) unmasked rays: rgb + rgb0 vs target (run_nerf.py:1455, 1481, 1511-1513)
        cfg, k = self._forward(0, rays_clf, False)
        losses += [torch.mean((k["rgb_map"] - target_clf) ** 2), torch.mean((k["rgb0"] - target_clf) ** 2)]
        chunk_backward(cfg, k, self.net_c, self.net_f, {"rgb_map": mse_g(k["rgb_map"], target_clf),
                                                        "rgb0": mse_g(k["rgb0"], target_clf)}, gc, gf,
                       self._scratch(cfg), self._ws(cfg))
        # 2) masked rays of the kept view, weights detached (run_nerf.py:1465, 1484-1489)
        cfg, k = self._forward(1, rays_s, True)
        losses += [torch.mean((k["rgb_map"] - target_s) ** 2), torch.mean((k["rgb0"] - target_s) ** 2)]
        chunk_backward(cfg, k, self.net_c, self.net_f, {"rgb_map": mse_g(k["rgb_map"], target_s),
                                                        "rgb0": mse_g(k["rgb0"], target_s)}, gc, gf,
                       self._scratch(cfg), self._ws(cfg))
        # 3) inpainted-disparity rays (run_nerf.py:1470, 1516-1521)
        cfg, k = self._forward(2, rays_inp, False)
        l_d, l_d0 = torch.mean((k["disp_map"] - depth_inp) ** 2), torch.mean((k["disp0"] - depth_inp) ** 2)
        # `if not inp_loss.isnan(): loss += inp_loss` (run_nerf.py:1520) without a host sync: a NaN disparity loss
        # (a ray with zero accumulated opacity) contributes neither loss nor gradient
        bad = torch.isnan(l_d + l_d0)
        zero = torch.zeros((), device=self.device)
        losses += [torch.where(bad, zero, l_d), torch.where(bad, zero, l_d0)]
        g_d = torch.where(bad, torch.zeros_like(depth_inp), mse_g(k["disp_map"], depth_inp))
        g_d0 = torch.where(bad, torch.zeros_like(depth_inp), mse_g(k["disp0"], depth_inp))
        chunk_backward(cfg, k, self.net_c, self.net_f, {"disp_map": g_d, "disp0": g_d0}, gc, gf,
                       self._scratch(cfg), self._ws(cfg))
        self.apply_gradients()
        loss = sum(losses)
        return loss, -10.0 * torch.log10(losses[0])

    def _scratch(self, cfg):
        return self.shared("d_raw", (cfg.n_rays * (cfg.n_samples + cfg.n_importance) * 4,), torch.float32)

    def _ws(self, cfg):
        nbytes = int(L.lib().spn_mlp_bwd_workspace_bytes(cfg.n_rays * (cfg.n_samples + cfg.n_importance), cfg.precision))
        return self.shared("bwd_ws", (nbytes,), torch.uint8)

    # ---------------------------------------------------------------------------------------
    def apply_gradients(self):
        """NCCL all-reduce (mean over ranks) of the two flat gradient vectors, then one Adam launch per network;
        learning-rate schedule of run_nerf.py:1616-1622.  With `self.adam_state` set (CUDA-graph mode) the step counter
        and schedule live on the device (spn_adam_tick / spn_adam_step_dev) so the launches are replayable."""
        if self.peer is not None and self.adam_state is None:
            self.global_step += 1
            lr = self.lr0 * (0.1 ** (max(self.global_step - 2, 0) / (self.lrate_decay * 1000)))
            self.peer.allreduce_adam(self.global_step, self.net_c, self.net_f, self.m, self.v, lr, self.betas, self.eps,
                                     self.global_step)
            self.net_c.mark_params_changed(); self.net_f.mark_params_changed()
            return
        pending = getattr(self, "_pending", None)
        if pending:                                   # the two per-network all-reduces were issued behind their backward passes
            for w in pending:
                if w is not None:
                    w.wait()                          # stream-side wait: the current stream continues after the collective
            self._pending = []
            scale = 1.0 / self.sharder.world
        else:
            scale = allreduce_sum_([self.grad_all], self.pg) if self.sharder.world > 1 else 1.0
        self.global_step += 1
        if self.adam_state is not None:
            L.check(L.lib().spn_adam_tick(L.ptr(self.adam_state), self.lr0, 0.1, float(self.lrate_decay * 1000),
                                          self.betas[0], self.betas[1], L.stream()), "spn_adam_tick")
        # the reference sets the rate AFTER optimizer.step() from its 0-based global_step (run_nerf.py:1611-1622, 1703): step k
        # (1-based) runs at lr0 * 0.1^((k-2)/decay_steps), steps 1 and 2 both at lr0
        lr = self.lr0 * (0.1 ** (max(self.global_step - 2, 0) / (self.lrate_decay * 1000)))
        for net, g, m, v in zip((self.net_c, self.net_f), self.grads, self.m, self.v):
            if net is None:
                continue
            if self.adam_state is not None:
                L.check(L.lib().spn_adam_step_dev(L.ptr(net.flat_params()), L.ptr(g), L.ptr(m), L.ptr(v), g.numel(),
                                                  L.ptr(self.adam_state), self.betas[0], self.betas[1], self.eps,
                                                  float(scale), L.stream()), "spn_adam_step_dev")
            else:
                ops.adam_step(net.flat_params(), g, m, v, self.global_step, lr, self.betas, self.eps, grad_scale=scale)
            net.mark_params_changed()

    # ---------------------------------------------------------------------------------------
    # checkpoints in the reference's layout (run_nerf.py:443-461, 1626-1636)
    def checkpoint(self):
        """{'global_step', 'network_fn_state_dict', 'network_fine_state_dict', 'optimizer_state_dict'} exactly as the
        reference trainer saves it: the optimizer entry is a torch.optim.Adam state_dict over the 48 parameter tensors
        (coarse then fine, registration order) cut out of the flat moment vectors, so either side can resume the other's run."""
        lr = self.lr0 * (0.1 ** (max(self.global_step - 1, 0) / (self.lrate_decay * 1000)))     # the rate the NEXT step runs at
        state, i = {}, 0
        for net, m, v in zip((self.net_c, self.net_f), self.m, self.v):
            net.flat_params()
            for o, p in zip(net._offsets, net._flat_params()):
                state[i] = {"step": torch.tensor(float(self.global_step)),
                            "exp_avg": m[o:o + p.numel()].view(p.shape).clone(),
                            "exp_avg_sq": v[o:o + p.numel()].view(p.shape).clone()}
                i += 1
        group = {"lr": lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(i))}
        # the reference saves its 0-based loop counter, i.e. (optimizer steps done - 1), next to the optimizer state (:1629, 1703)
        return {"global_step": max(self.global_step - 1, 0), "network_fn_state_dict": self.net_c.state_dict(),
                "network_fine_state_dict": self.net_f.state_dict(), "optimizer_state_dict": {"state": state, "param_groups": [group]}}

    def load_checkpoint(self, ckpt):
        """Resume from a checkpoint written by `checkpoint()` or by the reference trainer (create_nerf's reload,
        run_nerf.py:452-461): weights, Adam moments, step counter."""
        self.net_c.load_state_dict(ckpt["network_fn_state_dict"])
        self.net_f.load_state_dict(ckpt["network_fine_state_dict"])
        self.net_c.mark_params_changed(); self.net_f.mark_params_changed()
        st = ckpt["optimizer_state_dict"]["state"]
        # optimizer steps done: Adam's own counter where there is one (the reference's `global_step` lags it by one, :1703)
        self.global_step = int(st[0]["step"]) if 0 in st else int(ckpt["global_step"])
        i = 0
        for net, m, v in zip((self.net_c, self.net_f), self.m, self.v):
            net.flat_params()
            for o, p in zip(net._offsets, net._flat_params()):
                if i in st:        # a fresh optimizer has no state yet
                    m[o:o + p.numel()].copy_(st[i]["exp_avg"].reshape(-1))
                    v[o:o + p.numel()].copy_(st[i]["exp_avg_sq"].reshape(-1))
                else:
                    m[o:o + p.numel()].zero_(); v[o:o + p.numel()].zero_()
                i += 1
        if self.adam_state is not None:
            self.adam_state[0] = float(self.global_step)
        return self

    # ---------------------------------------------------------------------------------------
    def step_graphed(self, rays_clf, target_clf, rays_s, target_s, rays_inp, depth_inp):
        """`step` replayed as CUDA graphs (the whole step is ~45 small and 6 large launches; replaying it removes the launch
        latency a caller that reads the loss back every step would otherwise expose).  Inputs may live on the host (pinned)
        or the device: they are copied into static device buffers, then the graph is replayed.  The first call runs
        eagerly (allocates every pooled buffer), the second captures, later ones replay; shapes must not change.  Returns
        (loss, psnr) as views of static buffers (valid until the next call).

        One GPU: ONE graph.  Several GPUs: THREE graphs with the two gradient all-reduces between them — [forward, losses,
        fine backward] -> all-reduce(fine) on NCCL's stream || [coarse backward] -> all-reduce(coarse) -> [Adam] — because a
        CUDA-graph capture that contains the NCCL collectives hung on this stack (8 GPUs in round 1, 2 GPUs in round 2,
        TORCH_NCCL_ASYNC_ERROR_HANDLING=0 notwithstanding): the collectives stay eager, everything else is replayed."""
        ins = (rays_clf, target_clf, rays_s, target_s, rays_inp, depth_inp)
        multi = self.sharder.world > 1
        if multi and self.peer is not None:
            raise NotImplementedError("Trainer.step_graphed: the peer-memory gradient exchange is not graph-captured; "
                                      "use Trainer.step / step_from_pool")
        if self._graph is None:
            self._static_in = [torch.empty(t.shape, dtype=torch.float32, device=self.device) for t in ins]
            if self.adam_state is None:
                self.adam_state = torch.zeros(4, device=self.device)
                self.adam_state[0] = float(self.global_step)
            self._graph = "warm"
        for s, t in zip(self._static_in, ins):
            s.copy_(t, non_blocking=True)
        if self._graph == "warm":                      # eager pass: sizes every pooled buffer, sets kernel attributes
            out = self.step(*self._static_in)
            self._graph = "capture"
            return out
        if self._graph == "capture":
            torch.cuda.synchronize()
            n0 = L.lib().spn_launch_count(0)
            if self.gen is not None and multi:
                raise NotImplementedError("Trainer.step_graphed: Trainer(seed=...) with the multi-GPU segmented graphs is not supported")
            if not multi:
                graph = torch.cuda.CUDAGraph()
                if self.gen is not None:               # the private random stream must advance with every replay
                    graph.register_generator_state(self.gen)
                with torch.cuda.graph(graph):
                    self._static_out = self.step(*self._static_in)
                self._graph = graph
            else:
                g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                self._segments = "capturing"
                with torch.cuda.graph(g1):
                    self._static_out = self.step(*self._static_in)     # stops behind the fine backward (self._segments is set)
                pool = g1.pool()
                with torch.cuda.graph(g2, pool=pool):
                    self._coarse_backward()
                self._pending = []
                with torch.cuda.graph(g3, pool=pool):
                    self._graph_adam(1.0 / self.sharder.world)
                self._segments = (g1, g2, g3)
                self._graph = "segments"
            self.graph_launches = int(L.lib().spn_launch_count(0) - n0)   # our kernels per replay
            for pool in set(self.pools) | {self.shared}:
                pool.frozen = True
            self.global_step -= 1                      # capture only recorded the step, it did not run
        if multi:
            g1, g2, g3 = self._segments
            gc, gf = self.grads
            g1.replay()
            w1 = allreduce_sum_async(gf, self.pg)
            g2.replay()
            w2 = allreduce_sum_async(gc, self.pg)
            for w in (w1, w2):
                if w is not None:
                    w.wait()
            g3.replay()
        else:
            self._graph.replay()
        self.global_step += 1
        return self._static_out

    def _graph_adam(self, scale):
        """the optimiser tail with a fixed gradient scale (segmented-graph mode: the all-reduces ran outside the graphs)"""
        self.global_step += 1
        L.check(L.lib().spn_adam_tick(L.ptr(self.adam_state), self.lr0, 0.1, float(self.lrate_decay * 1000),
                                      self.betas[0], self.betas[1], L.stream()), "spn_adam_tick")
        for net, g, m, v in zip((self.net_c, self.net_f), self.grads, self.m, self.v):
            if net is None:
                continue
            L.check(L.lib().spn_adam_step_dev(L.ptr(net.flat_params()), L.ptr(g), L.ptr(m), L.ptr(v), g.numel(),
                                              L.ptr(self.adam_state), self.betas[0], self.betas[1], self.eps,
                                              float(scale), L.stream()), "spn_adam_step_dev")
            net.mark_params_changed()
