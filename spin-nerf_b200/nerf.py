"""NeRF / NeRF_RGB modules with the reference's constructor, parameter names and state_dict layout
(DS_NeRF/run_nerf_helpers.py:74-245), evaluated by the fused CUDA MLP (csrc/mlp_tc.cu, mlp_fp32.cu).

All 24 parameter tensors are views into ONE flat fp32 buffer in nn.Module registration order
(spn_mlp_param_offsets), so the kernels, the fused Adam and the NCCL gradient all-reduce see a
single 595 844-float vector while `state_dict()` / `parameters()` look exactly like the reference's.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L
from . import ops

_DEFAULT_PRECISION = {"bf16": L.PREC_BF16, "fp32": L.PREC_FP32}[os.environ.get("SPN_PRECISION", "bf16").lower()]


def default_precision():
    return _DEFAULT_PRECISION


def set_default_precision(p):
    global _DEFAULT_PRECISION
    _DEFAULT_PRECISION = {"bf16": L.PREC_BF16, "fp32": L.PREC_FP32}.get(p, p)


class _MLPFunction(torch.autograd.Function):
    """raw = NeRF(x6); gradient w.r.t. the 24 parameter tensors only (inputs carry none:
    z_samples are detached, run_nerf.py:700)."""

    @staticmethod
    def forward(ctx, net, need_grad, x6, *params):
        # (grad mode is always off inside Function.forward, so the caller decides whether to stash)
        flat, packed = net._sync()
        stash = ops.mlp_stash(x6.shape[0], net.precision, x6.device) if need_grad else None
        raw, stash = ops.mlp_forward_points(flat, packed, x6, net.precision, stash)
        ctx.net, ctx.stash, ctx.m = net, stash if need_grad else None, x6.shape[0]
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        net = ctx.net
        flat, packed = net._sync()
        g = torch.zeros(L.MLP_NPARAMS, device=d_raw.device, dtype=torch.float32)
        ops.mlp_backward(flat, packed, ctx.stash, d_raw.contiguous(), g, net.precision)
        ctx.stash = None
        grads = [g[o:o + p.numel()].view(p.shape) for o, p in zip(net._offsets, net._flat_params())]
        return (None, None, None) + tuple(grads)


class NeRF(nn.Module):
    """Same constructor as the reference (helpers:75-102).  The CUDA path implements the configuration
    BASELINE.json names: D=8, W=256, skips=[4], use_viewdirs=True, 63/27-dim encodings."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        if not (D == 8 and W == 256 and list(skips) == [4] and use_viewdirs and input_ch == 63 and input_ch_views == 27):
            raise NotImplementedError(
                "spinnerf_b200 implements the D=8, W=256, skips=[4], use_viewdirs=True, multires=10/4 network "
                f"(got D={D} W={W} skips={skips} use_viewdirs={use_viewdirs} input_ch={input_ch}/{input_ch_views}); "
                "there is no generic PyTorch fallback.")
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W)
                                        for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self.precision = default_precision()
        self._flat = None
        self._packed = None
        self._packed_version = None
        self._offsets = None

    # ---- flat storage --------------------------------------------------------------------
    def _flat_params(self):
        # registration order == spn_mlp_param_offsets order (sub-modules such as alpha_model excluded)
        mods = list(self.pts_linears) + [self.views_linears[0], self.feature_linear, self._alpha_head(), self.rgb_linear]
        return [p for m in mods for p in (m.weight, m.bias)]

    def _alpha_head(self):
        return self.alpha_linear

    def _flatten(self):
        ps = self._flat_params()
        total = sum(p.numel() for p in ps)
        assert total == L.MLP_NPARAMS, total
        dev = ps[0].device
        flat = torch.empty(total, device=dev, dtype=torch.float32)
        offs, o = [], 0
        for p in ps:
            flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[o:o + p.numel()].view(p.shape)
            offs.append(o); o += p.numel()
        self._flat, self._offsets = flat, offs
        self._packed_version = None

    def _aliased(self):
        if self._flat is None:
            return False
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * o and p.device == self._flat.device
                   for o, p in zip(self._offsets, self._flat_params()))

    def flat_params(self):
        """The flat fp32 parameter vector (re-established if .to()/load_state_dict broke the aliasing)."""
        if not self._aliased():
            self._flatten()
        return self._flat

    def _sync(self):
        flat = self.flat_params()
        if not flat.is_cuda:
            raise RuntimeError("spinnerf_b200.NeRF runs on CUDA tensors only (no CPU implementation)")
        packed = None
        if self.precision == L.PREC_BF16:
            ver = sum(p._version for p in self._flat_params()) + flat._version
            if self._packed is None or self._packed_version != ver or self._packed.device != flat.device:
                self._packed = ops.mlp_pack(flat, self._packed if self._packed is not None and self._packed.device == flat.device else None)
                self._packed_version = ver
            packed = self._packed
        return flat, packed

    def mark_params_changed(self):
        """Call after writing the flat vector directly (fused Adam) so the bf16 image is re-packed."""
        self._packed_version = None

    # ---- forward (helpers:104-127) -------------------------------------------------------
    def _as_points(self, x):
        c = x.shape[-1]
        if c == 6:                      # lazy embedder: [pt, viewdir]
            return x
        if c == self.input_ch + self.input_ch_views:
            # gamma() keeps its input as the first 3 columns (include_input=True, helpers:30-32), so the raw
            # point / direction are recoverable and the kernel re-encodes them on the fly
            return torch.cat([x[..., 0:3], x[..., self.input_ch:self.input_ch + 3]], -1)
        raise RuntimeError(f"NeRF.forward: expected 6 or {self.input_ch + self.input_ch_views} input columns, got {c}")

    def forward(self, x):
        sh = x.shape
        x6 = self._as_points(x).reshape(-1, 6).float().contiguous()
        params = self._flat_params()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        raw = _MLPFunction.apply(self, need_grad, x6, *params)
        return raw.reshape(*sh[:-1], 4)

    def seeded_init_(self, seed, scale=1.0):
        """nn.Linear-style uniform init U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (what helpers:86-102 get from torch's default
        initialiser) drawn from a numpy PCG64 stream in registration order, so that every rank, every machine and the
        test-side checker derive the same synthetic weights from a seed without shipping them.  Returns self."""
        import numpy as np
        rng = np.random.default_rng(seed)
        with torch.no_grad():
            for p in self._flat_params():          # weight then bias of each Linear, both bounded by the layer's fan-in
                if p.dim() == 2:
                    bound = scale / np.sqrt(p.shape[1])
                p.copy_(torch.from_numpy(rng.uniform(-bound, bound, size=tuple(p.shape)).astype(np.float32)))
        self.mark_params_changed()
        return self

    def load_weights_from_keras(self, weights):
        """helpers:129-156."""
        import numpy as np
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        def put(lin, i):
            lin.weight.data.copy_(torch.from_numpy(np.transpose(weights[i])))
            lin.bias.data.copy_(torch.from_numpy(np.transpose(weights[i + 1])))
        for i in range(self.D):
            put(self.pts_linears[i], 2 * i)
        put(self.feature_linear, 2 * self.D)
        put(self.views_linears[0], 2 * self.D + 2)
        put(self.rgb_linear, 2 * self.D + 4)
        put(self._alpha_head(), 2 * self.D + 6)
        self.mark_params_changed()


class NeRF_RGB(NeRF):
    """helpers:159-245: colour head trained on top of a frozen density provider `alpha_model` (its sigma replaces this
    network's, under no_grad, helpers:202-203).  Like the reference module it owns no alpha_linear: parameters(),
    state_dict() and the optimizer see the reference's 22 tensors (+ the registered alpha_model's); the kernels' flat
    parameter layout keeps a zero density head that is not a registered parameter."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 alpha_model=None):
        super().__init__(D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs)
        head = self.alpha_linear
        del self.alpha_linear
        head.weight.requires_grad_(False); head.bias.requires_grad_(False)
        with torch.no_grad():
            head.weight.zero_(); head.bias.zero_()
        object.__setattr__(self, "_zero_alpha_head", head)      # plain attribute: not a sub-module, not in state_dict
        self.alpha_model = alpha_model

    def _alpha_head(self):
        return self._zero_alpha_head

    def forward(self, x):
        raw = super().forward(x)
        with torch.no_grad():
            alpha = self.alpha_model(x)[..., 3][..., None]
        return torch.cat([raw[..., :3], alpha], -1)
