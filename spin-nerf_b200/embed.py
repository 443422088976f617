"""Positional-encoding front end with the reference's interface (DS_NeRF/run_nerf_helpers.py:22-70).

`get_embedder(multires, i)` returns `(callable, out_dim)` like the reference.  By default the callable is LAZY: it hands
the raw 3-vector through and reports the encoded width (63 / 27), so that create_nerf builds reference-shaped layers while
NeRF.forward receives [pts, viewdir] and encodes inside the fused MLP kernel (the encoding never exists in HBM).
SPN_LAZY_EMBED=0 materialises gamma(x) with spn_embed instead (NeRF.forward accepts both)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops

LAZY_EMBED = os.environ.get("SPN_LAZY_EMBED", "1") != "0"


class Embedder:
    """helpers:22-52 (include_input=True, log-sampled frequencies 2^0 .. 2^(L-1), sin / cos)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if not (kwargs['include_input'] and kwargs['input_dims'] == 3 and kwargs['log_sampling']):
            raise NotImplementedError("spinnerf_b200 implements the reference's default embedder configuration")
        self.n_freqs = kwargs['num_freqs']
        self.out_dim = 3 + 6 * self.n_freqs

    def embed(self, inputs):
        if LAZY_EMBED:
            return inputs
        return ops.embed(inputs, self.n_freqs)


def get_embedder(multires, i=0):
    """helpers:55-70."""
    if i == -1:
        return nn.Identity(), 3
    embedder_obj = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                            log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    embed = lambda x, eo=embedder_obj: eo.embed(x)
    return embed, embedder_obj.out_dim
