"""Ray pools of a loaded scene: what `train()` prepares between loading the data and the first
optimisation step (DS_NeRF/run_nerf.py:1225-1325) and what RayDataset / DataLoader then sample from (data.py:4-15,
run_nerf.py:1336-1348, 1367-1413), as ONE resident pool plus index sets instead of three materialised [M,3,4] arrays:

    every pixel of every training view -> origin, direction (get_rays_np, run_nerf_helpers.py:263-272), colour of the
    (inpainted) frame, mask label, inpainted disparity;
    idx_clf   pixels whose label is 0        (`rays_rgb_clf`: unmasked rays, rgb loss)              run_nerf.py:1307-1310
    idx_inp   pixels whose label is not 0    (`rays_inp`: rays supervised by the inpainted disparity)    :1311
    idx_rgb   pixels whose label is 1        (`rays_rgb`: masked rays of the view that keeps its mask)   :1317-1318

Pixel order is (view in i_train order, row, column), i.e. the order of the reference's arrays, so `pool[idx_*]` reproduces
them row for row.  Host numpy; `RayPools.to(device)` uploads once — `Trainer.step_from_pool` then samples on the device.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def camera_rays(H, W, focal, c2w):
    """[H,W,3] origins and directions of one view: integer pixel centres, dirs ((i-W/2)/f, -(j-H/2)/f, -1) rotated by
    c2w[:3,:3], origin c2w[:3,3] (run_nerf_helpers.py:263-272).  float32 like the reference (np.arange(dtype=float32))."""
    c2w = np.asarray(c2w)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -np.ones_like(i)], -1)
    d = np.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    return np.broadcast_to(c2w[:3, -1], d.shape), d


@dataclass
class RayPools:
    o: np.ndarray          # [M,3] float32
    d: np.ndarray          # [M,3] float32
    rgb: np.ndarray        # [M,3] float32   colour target of every pixel
    label: np.ndarray      # [M]   float32   mask label (0 background, 1 object of the kept view, -1 object of LPIPS views / no label)
    disp: np.ndarray       # [M]   float32   inpainted disparity target
    idx_clf: np.ndarray    # int64 index sets into the pool (see the module docstring)
    idx_inp: np.ndarray
    idx_rgb: np.ndarray

    def to(self, device):
        """Torch tensors on `device`: pool_od [2,M,3], rgb [M,3], disp [M] and the three int64 index sets."""
        import torch
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        return dict(pool_od=torch.stack([t(self.o), t(self.d)], 0), rgb=t(self.rgb), disp=t(self.disp),
                    idx_clf=t(self.idx_clf), idx_inp=t(self.idx_inp), idx_rgb=t(self.idx_rgb))

    def sample_indices(self, n_rand, rng):
        """[3, n_rand] pool indices of one step: one draw without replacement from each group (a DataLoader batch of a
        shuffled RayDataset, run_nerf.py:1336-1348), in the order (unmasked, masked, inpainted) Trainer.step_from_pool takes."""
        def pick(idx):
            if len(idx) == 0:
                raise ValueError("RayPools.sample_indices: a ray group is empty (no view with that mask label among i_train)")
            return idx[rng.choice(len(idx), n_rand, replace=len(idx) < n_rand)]
        return np.stack([pick(self.idx_clf), pick(self.idx_rgb), pick(self.idx_inp)], 0)


def build_ray_pools(images, poses, hwf, masks, inpainted_depths, i_train, prepare=False, train_gt=False):
    """images [N,H,W,3], poses [N,3,>=4], masks / inpainted_depths [N,H,W] as returned by scene_io.load_scene; i_train the
    training view indices.  `prepare` / `train_gt` widen the groups like run_nerf.py:1307-1318 (all pixels feed the rgb loss)."""
    H, W, focal = int(hwf[0]), int(hwf[1]), hwf[2]
    o, d, rgb, lab, dsp = [], [], [], [], []
    for i in i_train:
        ro, rd = camera_rays(H, W, focal, poses[i, :3, :4])
        o.append(ro.reshape(-1, 3)); d.append(rd.reshape(-1, 3))
        rgb.append(np.asarray(images[i]).reshape(-1, 3))
        lab.append(np.asarray(masks[i]).reshape(-1)); dsp.append(np.asarray(inpainted_depths[i]).reshape(-1))
    f32 = lambda parts: np.concatenate(parts, 0).astype(np.float32)
    o, d, rgb, lab, dsp = f32(o), f32(d), f32(rgb), f32(lab), f32(dsp)
    everything = np.arange(lab.shape[0], dtype=np.int64)
    idx_clf = everything if (train_gt or prepare) else np.flatnonzero(lab == 0)
    idx_inp = np.flatnonzero(lab != 0)
    idx_rgb = everything if prepare else np.flatnonzero(lab == 1)
    return RayPools(o, d, rgb, lab, dsp, idx_clf, idx_inp, idx_rgb)


def sparse_depth_rays(depth_gts, poses, hwf, masks, i_train, prepare=False):
    """The fourth pool of `--colmap_depth` (run_nerf.py:1266-1300): one ray through every COLMAP point a training view
    observes OUTSIDE its mask (all points with `prepare`), with that point's depth and confidence weight.
    depth_gts: scene_io.colmap_depth_rays(...) (one {"depth", "coord", "weight"} per view); masks [N,H,W].
    Returns (rays [2,M,3] float32, depth [M] float32, weight [M] float32) in the reference's order (views in i_train order, points
    in file order); `np.concatenate([rays.transpose(1,0,2), repeat(depth), repeat(weight)], 1)` is its `rays_depth` [M,4,3]."""
    H, W, focal = int(hwf[0]), int(hwf[1]), hwf[2]
    o, d, dep, wgt = [], [], [], []
    for i in i_train:
        coord, depth, weight = (np.asarray(depth_gts[i][k]) for k in ("coord", "depth", "weight"))
        if not prepare:      # keep points whose (clamped, truncated) pixel carries mask label 0 (:1272-1283)
            m = np.asarray(masks[i])
            rows = np.minimum(coord[:, 1].astype(np.int64), m.shape[0] - 1)
            cols = np.minimum(coord[:, 0].astype(np.int64), m.shape[1] - 1)
            keep = m[rows, cols] == 0
            coord, depth, weight = coord[keep], depth[keep], weight[keep]
        c2w = np.asarray(poses[i])[:3, :4]
        # get_rays_by_coord_np (run_nerf_helpers.py:275-280): no rounding of the sub-pixel COLMAP coordinates
        px, py = (coord[:, 0] - W * 0.5) / focal, -(coord[:, 1] - H * 0.5) / focal
        dirs = np.stack([px, py, -np.ones_like(px)], -1)
        rd = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
        o.append(np.broadcast_to(c2w[:3, -1], rd.shape)); d.append(rd); dep.append(depth); wgt.append(weight)
    cat = lambda parts: np.concatenate(parts, 0).astype(np.float32)
    return np.stack([cat(o), cat(d)], 0), cat(dep), cat(wgt)


def draw_step_indices(dev_pools, n_rand, generator=None):
    """[3, n_rand] int64 pool indices for one step, drawn ON THE DEVICE from the three index sets of RayPools.to(device)
    (with replacement — a uniform draw per step instead of the DataLoader's epoch-wise shuffle), ready for
    Trainer.step_from_pool(dev_pools["pool_od"], dev_pools["rgb"], dev_pools["disp"], idx)."""
    import torch
    rows = []
    for key in ("idx_clf", "idx_rgb", "idx_inp"):
        ids = dev_pools[key]
        if ids.numel() == 0:
            raise ValueError(f"draw_step_indices: {key} is empty")
        rows.append(ids[torch.randint(0, ids.numel(), (n_rand,), device=ids.device, generator=generator)])
    return torch.stack(rows, 0)
