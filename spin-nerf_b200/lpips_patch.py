"""Host side of the perceptual-loss branch of the SPIn-NeRF train step (DS_NeRF/run_nerf.py:1523-1561, the
`--lpips` flag): which patches are rendered and what they are compared with.

Every step after iteration 300 the reference picks `lpips_batch_size` training views, draws one patch origin per view
inside the bounding box of that view's mask (render_path, run_nerf.py:197-211), renders the 47x63 patch at
`lpips_render_factor` through render() with the TEST kwargs (perturb=0, raw_noise_std=0) and detach_weights=True, and
adds  sum_i LPIPS(2*(rgb_i-0.5), resized_target_i[patch]).mean() / batch_size / 100  to the loss.  The LPIPS network
itself is a fixed differentiable input (out of scope); this module reproduces the sampling and the targets:

* `patch_geometry`   the integer arithmetic of run_nerf.py:1534-1535 and render_path's render_factor scaling (:171-176)
* `MaskBoxes`        the per-view mask bounding boxes on the down-scaled grid, computed ONCE (the reference runs
                     np.where over the full-resolution mask every step and view, :198-200)
* `draw_origins`     the two random.randint draws per view in the reference's order (:201-208)
* `resize_targets`   torchvision.transforms.Resize of the [-1,1] target frames (:1537-1539, 1556-1557), done ONCE for all
                     views instead of once per step and view
* `PatchSampler`     the three together with the view shuffle of :1529-1532

Rendering the patches and pushing the LPIPS gradient through the fused chunk pipeline is Trainer.lpips_patch_backward.
"""
from __future__ import annotations

import copy
import random as _random

import numpy as np
import torch


def patch_geometry(hwf, lpips_render_factor, patch_len_factor):
    """-> (H_s, W_s, focal_s, (len0, len1)): render_path's down-scaled camera (run_nerf.py:171-176) and the patch size
    of run_nerf.py:1534-1535 (1008x756, factor 2, patch_len_factor 8 -> 378x504 and 47x63 = 2961 rays)."""
    H, W, focal = hwf
    H, W = int(H), int(W)
    rf = int(lpips_render_factor)
    Hs, Ws, fs = H, W, float(focal)
    if rf != 0:
        Hs, Ws, fs = H // rf, W // rf, float(focal) / rf
    plen = (H // rf // int(patch_len_factor), W // rf // int(patch_len_factor))
    return Hs, Ws, fs, plen


class MaskBoxes:
    """Bounding box of every view's mask on the render_factor grid: rows [x0, x1], columns [y0, y1] (inclusive), i.e.
    min / max of `np.where(mask != 0)[k] // render_factor` (run_nerf.py:198-200; floor division is monotone, so the
    extremes of the divided indices are the divided extremes)."""

    def __init__(self, masks, render_factor):
        masks = np.asarray(masks)
        rf = int(render_factor)
        boxes = np.zeros((masks.shape[0], 4), np.int64)
        for v in range(masks.shape[0]):
            nz = masks[v] != 0
            rows, cols = np.flatnonzero(nz.any(1)), np.flatnonzero(nz.any(0))
            if rows.size == 0:
                boxes[v] = -1          # e.g. a held-out ground-truth view: only an error if a patch is ever drawn from it
                continue
            boxes[v] = (rows[0] // rf, rows[-1] // rf, cols[0] // rf, cols[-1] // rf)
        self.boxes = boxes

    def __getitem__(self, v):
        if self.boxes[v][0] < 0:
            raise ValueError(f"view {v} has an empty mask: the reference's np.where(...).min() raises here too")
        return tuple(int(b) for b in self.boxes[v])


def draw_origins(boxes, views, patch_len, rand=_random):
    """One (X, Y) patch origin per view, drawn exactly like render_path (run_nerf.py:201-208): X first, then Y, with
    `rand.randint(lo, max(hi - len, lo))` (both ends inclusive).  `rand` is the `random` module or a random.Random."""
    Xs, Ys = [], []
    for v in views:
        x0, x1, y0, y1 = boxes[v]
        Xs.append(rand.randint(x0, max(x1 - patch_len[0], x0)))
        Ys.append(rand.randint(y0, max(y1 - patch_len[1], y0)))
    return Xs, Ys


def resize_targets(images, Hs, Ws):
    """[V,H,W,3] frames in [0,1] -> [V,3,Hs,Ws] in [-1,1], resized like torchvision.transforms.Resize((Hs, Ws)) applied
    to the (x-0.5)*2 NCHW tensor (run_nerf.py:1537-1539, 1553-1557): bilinear, antialiased, half-pixel centres — what Resize
    does to a float tensor in torchvision >= 0.17 (antialias on by default; older releases did not antialias tensors, so the
    reference's own targets depend on the installed torchvision; the fixture pins this image's 0.26)."""
    x = torch.as_tensor(images, dtype=torch.float32)
    x = ((x - 0.5) * 2).permute(0, 3, 1, 2)
    if tuple(x.shape[-2:]) == (int(Hs), int(Ws)):
        return x.contiguous()
    return torch.nn.functional.interpolate(x, size=(int(Hs), int(Ws)), mode="bilinear", align_corners=False,
                                           antialias=True)


def crop(frames, view, X, Y, patch_len):
    """target[:, :, X:X+len0, Y:Y+len1] of one resized frame (run_nerf.py:1556-1557) -> [1,3,<=len0,<=len1]."""
    return frames[view:view + 1, :, X:X + patch_len[0], Y:Y + patch_len[1]]


class PatchSampler:
    """The per-step choices of the LPIPS branch.  `sample()` follows the reference's order of random draws: a numpy
    shuffle of the training-view list (run_nerf.py:1529-1532), then X and Y per chosen view from `random`."""

    def __init__(self, hwf, masks, images, i_train, lpips_render_factor=2, patch_len_factor=8, lpips_batch_size=4,
                 device=None):
        self.Hs, self.Ws, self.focal_s, self.patch_len = patch_geometry(hwf, lpips_render_factor, patch_len_factor)
        self.render_factor = int(lpips_render_factor)
        self.i_train = np.asarray(i_train).copy()
        self.batch_size = int(lpips_batch_size)
        self.boxes = MaskBoxes(masks, self.render_factor)
        self.targets = resize_targets(images, self.Hs, self.Ws)
        if device is not None:
            self.targets = self.targets.to(device)

    def sample(self, np_random=np.random, py_random=_random):
        idx = copy.deepcopy(self.i_train)
        np_random.shuffle(idx)
        idx = [int(v) for v in idx[:self.batch_size]]
        Xs, Ys = draw_origins(self.boxes, idx, self.patch_len, py_random)
        return idx, Xs, Ys

    def target_patches(self, idx, Xs, Ys):
        return [crop(self.targets, v, x, y, self.patch_len) for v, x, y in zip(idx, Xs, Ys)]
