"""Drop-in replacement for DS_NeRF/run_nerf_helpers.py — the reference's only plugin seam
(`from run_nerf_helpers import *`, DS_NeRF/run_nerf.py:23).  Every public name of the reference
module is exported with the same signature; the work is done by spinnerf_b200's CUDA library.

Use (python >= 3.11):
    PYTHONPATH=<repo>/spin-nerf_b200/dropin:<repo>/spin-nerf_b200/compat:/path/to/SPIn-NeRF/DS_NeRF \
        python -P /path/to/SPIn-NeRF/DS_NeRF/run_nerf.py --config ... --no_tcnn
"""
import importlib
import os
import sys

import cv2                                   # noqa: F401  (re-exported like the reference, helpers:1-12)
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F             # noqa: F401
import torchvision                           # noqa: F401  run_nerf.py:1536 gets it through the star import
from torch import searchsorted               # noqa: F401

try:                                         # helpers:12 — pyplot is only used by visualize_sigma
    from matplotlib import pyplot as plt
except Exception:                            # pragma: no cover
    plt = None

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_spn = importlib.import_module("spin-nerf_b200")
_ops = _spn.ops

# NOTE: unlike the reference (helpers:5) autograd anomaly mode is NOT switched on at import.

# Misc (helpers:15-18)
img2mse = lambda x, y: torch.mean((x - y) ** 2)
img2l1 = lambda x, y: torch.mean(torch.abs(x - y))
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device if torch.is_tensor(x) else None))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)

_embed = importlib.import_module("spin-nerf_b200.embed")
LAZY_EMBED = _embed.LAZY_EMBED
Embedder = _embed.Embedder             # helpers:22-52 (lazy by default: NeRF.forward encodes inside the fused MLP kernel)
get_embedder = _embed.get_embedder     # helpers:55-70


NeRF = _spn.NeRF
NeRF_RGB = _spn.NeRF_RGB


# Ray helpers (helpers:249-300)
def get_rays(H, W, focal, c2w):
    return _ops.get_rays(H, W, focal, c2w)


def get_rays_np(H, W, focal, c2w):
    """helpers:263-272 — host-side (numpy in, numpy out), used to pre-compute training rays."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def get_rays_by_coord_np(H, W, focal, c2w, coords):
    """helpers:275-280."""
    i, j = (coords[:, 0] - W * 0.5) / focal, -(coords[:, 1] - H * 0.5) / focal
    dirs = np.stack([i, j, -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    return _ops.ndc_rays(H, W, focal, near, rays_o, rays_d)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """helpers:304-347."""
    return _ops.sample_pdf(bins, weights, N_samples, det=det, pytest=pytest)


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, need_alpha=False,
                detach_weights=False):
    """helpers:350-401."""
    return _ops.raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest, need_alpha, detach_weights)


def sample_sigma(rays_o, rays_d, viewdirs, network, z_vals, network_query):
    """helpers:404-418 (the reference unpacks 5 of raw2outputs' 6 returns and raises; fixed here)."""
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    raw = network_query(pts, viewdirs, network)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = F.relu(raw[..., 3])
    depth_map = raw2outputs(raw, z_vals, rays_d)[4]
    return rgb, sigma, depth_map


def visualize_sigma(sigma, z_vals, filename):
    """helpers:421-425."""
    plt.plot(z_vals, sigma); plt.xlabel('z_vals'); plt.ylabel('sigma'); plt.savefig(filename)


device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
