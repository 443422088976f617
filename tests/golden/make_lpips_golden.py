"""Generate tests/golden/lpips_patches.npz from the UNMODIFIED reference (build container only):

    python tests/golden/make_lpips_golden.py

What is recorded is the host side of the `--lpips` branch (DS_NeRF/run_nerf.py:1523-1561): the reference's own
`render_path(..., rgb_require_grad=True, patch_len=..., masks=...)` is run with `run_nerf.render` replaced by a recorder
(so the patch windows it asks for are captured, not the pixels), after the same view shuffle the train loop does, and the
targets come from torchvision.transforms.Resize exactly as in the train loop.  Images and masks are NOT stored: the test
regenerates them with `scene(seed, ...)` below (numpy PCG64, stable across machines).
"""
import copy
import os
import random
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader            # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def scene(seed, V, H, W):
    rng = np.random.default_rng(seed)
    images = rng.uniform(0, 1, (V, H, W, 3)).astype(np.float32)
    masks = np.zeros((V, H, W), np.float32)
    for v in range(V):        # one blob per view; view 1's is smaller than a patch and sits in the corner, view 2's touches the border
        h0, w0 = int(rng.integers(0, H // 2)), int(rng.integers(0, W // 2))
        hh, ww = int(rng.integers(H // 4, H // 2)), int(rng.integers(W // 4, W // 2))
        if v == 1:
            hh, ww = 5, 7
            h0, w0 = H - hh, W - ww       # in the corner: its patch sticks out of the frame and is clamped
        if v == 2:
            h0, w0 = H - hh, W - ww
        masks[v, h0:h0 + hh, w0:w0 + ww] = 1.0
        masks[v, (h0 + hh // 2) % H, (w0 + 3) % W] = 0.5     # any non-zero value counts (masks[i] != 0)
    poses = rng.standard_normal((V, 3, 5)).astype(np.float32)
    return images, masks, poses


def main():
    _, R = ref_loader.load()
    out = {}
    for case, (H, W, rf, plf, bs, steps) in enumerate([(96, 128, 2, 8, 4, 6), (75, 101, 1, 4, 3, 5), (60, 80, 3, 2, 2, 4)]):
        V = 6
        images, masks, poses = scene(100 + case, V, H, W)
        hwf = [H, W, 0.9 * W]
        i_train = np.array([0, 1, 2, 4, 5])
        calls = []

        def fake_render(Hs, Ws, focal, chunk=0, c2w=None, patch=None, **kw):
            calls.append((Hs, Ws, focal, patch, kw.get("detach_weights"), kw.get("retraw")))
            n0 = len(range(Hs)[patch[0]:patch[0] + patch[2]]); n1 = len(range(Ws)[patch[1]:patch[1] + patch[3]])
            z = torch.zeros(n0, n1, 3)
            return z, z[..., 0], z[..., 0], z[..., 0], {}
        saved = R.render
        R.render = fake_render
        np.random.seed(7 + case); random.seed(11 + case)
        rec_idx, rec_X, rec_Y, rec_shape, rec_tgt = [], [], [], [], []
        try:
            for _ in range(steps):
                # run_nerf.py:1529-1539
                idx = copy.deepcopy(i_train); np.random.shuffle(idx); idx = idx[:bs]
                patch_len = (hwf[0] // rf // plf, hwf[1] // rf // plf)
                transform = torchvision.transforms.Resize((hwf[0] // rf, hwf[1] // rf))
                calls.clear()
                try:
                    rgbs, disps, (Xs, Ys) = R.render_path(torch.from_numpy(poses[idx]), hwf, 1024, {}, render_factor=rf,
                                                          rgb_require_grad=True, need_alpha=False, detach_weights=True,
                                                          patch_len=patch_len, masks=masks[idx])
                except (RuntimeError, ValueError):   # np/torch.stack of clamped (unequal) patches raises in the reference:
                    # the draws still happened; ours renders the clamped patch instead of failing
                    Xs, Ys = [c[3][0] for c in calls], [c[3][1] for c in calls]
                assert all(c[3][2:] == patch_len and c[4] is True for c in calls) and len(calls) == len(idx)
                rec_idx.append(idx); rec_X.append(Xs); rec_Y.append(Ys)
                rec_shape.append([calls[0][0], calls[0][1]])
                for j in range(len(idx)):     # run_nerf.py:1553-1557
                    target = ((torch.from_numpy(images[idx[j]]) - 0.5) * 2).permute(2, 0, 1)[None, ...]
                    target = transform(target)[:, :, Xs[j]:Xs[j] + patch_len[0], Ys[j]:Ys[j] + patch_len[1]]
                    full = np.zeros((3, patch_len[0], patch_len[1]), np.float32)
                    t = target[0].numpy()
                    full[:, :t.shape[1], :t.shape[2]] = t
                    rec_tgt.append(full)
                    rec_shape.append([t.shape[1], t.shape[2]])
        finally:
            R.render = saved
        out.update({f"c{case}_cfg": np.array([H, W, rf, plf, bs, steps, 7 + case, 11 + case]), f"c{case}_i_train": i_train, f"c{case}_idx": np.array(rec_idx),
                    f"c{case}_X": np.array(rec_X), f"c{case}_Y": np.array(rec_Y), f"c{case}_shapes": np.array(rec_shape),
                    f"c{case}_targets": np.stack(rec_tgt), f"c{case}_focal_s": np.array([calls[0][2]])})
    np.savez_compressed(os.path.join(OUT, "lpips_patches.npz"), **out)
    print("wrote lpips_patches.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
