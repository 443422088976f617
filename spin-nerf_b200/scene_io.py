"""On-disk scene contract of the SPIn-NeRF trainer: reader + synthetic-scene writer.

What `DS_NeRF/run_nerf.py:982` consumes comes from `DS_NeRF/load_llff.py` (`_load_data` :68-190, `load_llff_data` :315-433):

    <scene>/poses_bounds.npy                N x 17 float64: a 3x5 matrix [-up, right, back | position | (H, W, focal)] row-major, then (near, far)
    <scene>/images/*.png|jpg                full-resolution frames (only the first one is opened, for its shape)
    <scene>/images_<f>/lama_images/*.png    inpainted frames at 1/f resolution (what training uses; `prepare` reads images_<f>/*)
    <scene>/images_<f>/label/<stem>.png     object masks (any positive value = masked), dilated 5 x (5x5) on load
    <scene>/images_<f>/depth/<stem>.png     inpainted disparities, 8 bit, read as value / 255

`load_scene` returns the same 8-tuple as the reference loader (images, poses [N,3,5], bds, render_poses [120,3,5], i_test, masks,
inpainted_depths, mask_indices); `write_scene` produces a directory both loaders accept, so synthetic "statue-shaped" data can
stand in for the Google-Drive dataset.  Pure numpy / OpenCV host code, written from the contract above — no CUDA on this path.
tests/test_scene_io.py pins it against the unmodified reference loader (imported from /root/reference where that exists) and
against a committed digest of its outputs.
"""
from __future__ import annotations

import os

import cv2
import numpy as np

_EXT = (".JPG", ".jpg", ".jpeg", ".png")


# ---------------------------------------------------------------------------------------------------------------------
# small pose algebra (float32 in, float32 out — the reference keeps poses in float32 after load_llff.py:335)
# ---------------------------------------------------------------------------------------------------------------------
def _unit(v):
    return v / np.linalg.norm(v)


def look_at(z, up, pos):
    """3x4 camera-to-world from a viewing axis, an approximate up vector and a position (load_llff.py:197-203)."""
    z = _unit(z)
    x = _unit(np.cross(up, z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, pos], 1)


def average_pose(poses):
    """3x5 'mean camera' of poses [N,3,5] (load_llff.py:211-219): mean position, summed axes, hwf of the first pose."""
    z = _unit(poses[:, :3, 2].sum(0))         # normalised here AND in look_at, like the reference (one float32 ulp matters)
    up = poses[:, :3, 1].sum(0)
    return np.concatenate([look_at(z, up, poses[:, :3, 3].mean(0)), poses[0, :3, -1:]], 1)


def _to_h(p34):
    """[..., 3, 4] -> [..., 4, 4] in float64 (the reference's homogeneous row is a float64 array, so its pose algebra runs
    in double and is rounded to float32 once, on assignment)"""
    row = np.broadcast_to(np.array([0, 0, 0, 1.]), p34.shape[:-2] + (1, 4))
    return np.concatenate([p34, row], -2)


def recenter(poses):
    """Express every pose in the frame of the average pose (load_llff.py:235-247)."""
    out = poses.copy()
    ref = _to_h(average_pose(poses)[:3, :4])
    out[:, :3, :4] = (np.linalg.inv(ref) @ _to_h(poses[:, :3, :4]))[:, :3, :4]
    return out


def spiral_path(c2w, up, rads, focal, zrate, rots, n):
    """Spiral of n poses around c2w (load_llff.py:222-232), all angles at once."""
    hwf = c2w[:, 4:5]
    rads = np.append(np.asarray(rads), 1.)
    theta = np.linspace(0., 2. * np.pi * rots, n + 1)[:-1]
    offs = np.stack([np.cos(theta), -np.sin(theta), -np.sin(theta * zrate), np.ones_like(theta)], -1) * rads   # [n,4]
    centres = offs @ c2w[:3, :4].T                                                                             # [n,3]
    target = c2w[:3, :4] @ np.array([0, 0, -focal, 1.])
    return [np.concatenate([look_at(c - target, up, c), hwf], 1) for c in centres]


def spherify(poses, bds):
    """load_llff.py:252-312 restated: rigid + scale normalisation of a capture that orbits a point.  Returns
    (poses_reset [N,3,5], circle_poses [120,3,5], bds * sc, sc, world->sphere 4x4); `bds` is NOT modified in place (the
    reference does, see load_scene)."""
    d = poses[:, :3, 2:3]
    o = poses[:, :3, 3:4]
    proj = np.eye(3) - d * np.transpose(d, [0, 2, 1])                      # projector orthogonal to each optical axis
    rhs = -proj @ o
    centre = np.squeeze(-np.linalg.inv((np.transpose(proj, [0, 2, 1]) @ proj).mean(0)) @ rhs.mean(0))
    up = _unit((poses[:, :3, 3] - centre).mean(0))
    a = _unit(np.cross([.1, .2, .3], up))
    b = _unit(np.cross(up, a))
    to_sphere = np.linalg.inv(_to_h(np.stack([a, b, up, centre], 1)[None]))
    reset = to_sphere @ _to_h(poses[:, :3, :4])
    rad = np.sqrt(np.mean(np.sum(np.square(reset[:, :3, 3]), -1)))
    sc = 1. / rad
    reset[:, :3, 3] *= sc
    rad = rad * sc
    zh = np.mean(reset[:, :3, 3], 0)[2]
    rc = np.sqrt(rad ** 2 - zh ** 2)
    ring = []
    for th in np.linspace(0., 2. * np.pi, 120):
        cam = np.array([rc * np.cos(th), rc * np.sin(th), zh])
        z = _unit(cam)
        x = _unit(np.cross(z, np.array([0, 0, -1.])))
        ring.append(np.stack([x, _unit(np.cross(z, x)), z, cam], 1))
    ring = np.stack(ring, 0)
    hwf = poses[0, :3, -1:]
    ring = np.concatenate([ring, np.broadcast_to(hwf, ring[:, :3, -1:].shape)], -1)
    reset = np.concatenate([reset[:, :3, :4], np.broadcast_to(hwf, reset[:, :3, -1:].shape)], -1)
    return reset, ring, (bds * sc).astype(bds.dtype), sc, to_sphere


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------
def _read(path):
    """Image file -> array in RGB(A) channel order, raw integer values (what imageio.imread returns for 8/16-bit files)."""
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    return img


def _frames(d):
    return [f for f in sorted(os.listdir(d)) if f.endswith(_EXT)]


def _plane(img, hw, scale):
    """First channel of a mask / depth image, divided by `scale`, nearest-resized to the frame size if it differs."""
    m = img / scale
    if m.ndim > 2:
        m = m[:, :, 0]
    if m.shape != hw:
        m = cv2.resize(m, (hw[1], hw[0]), interpolation=cv2.INTER_NEAREST)
    return m


def load_scene(basedir, factor=8, recenter_poses=True, bd_factor=.75, spherify_poses=False, path_zflat=False,
               spherify_hack=True, prepare=False, lpips=False):
    """Same return value as `load_llff_data(basedir, factor, recenter, bd_factor, spherify, path_zflat, spherify_hack, prepare,
    args=<lpips flag>)` (load_llff.py:315-433) for a scene whose images_<factor> directory exists (no ImageMagick minify)."""
    arr = np.load(os.path.join(basedir, "poses_bounds.npy"))
    poses = arr[:, :15].reshape(-1, 3, 5)                                  # [N,3,5] float64
    bds = arr[:, 15:17]
    sfx = "" if factor is None else f"_{factor}"
    root = os.path.join(basedir, "images" + sfx)
    imgdir = root if prepare else os.path.join(root, "lama_images")
    if not os.path.isdir(imgdir):
        raise FileNotFoundError(f"{imgdir} does not exist (write_scene creates it; the reference would shell out to mogrify)")
    names = _frames(imgdir)
    if poses.shape[0] > len(names):
        poses = poses[:len(names)]
    if poses.shape[0] != len(names):
        raise ValueError(f"{len(names)} images but {poses.shape[0]} poses in {basedir}")
    imgs = np.stack([_read(os.path.join(imgdir, f))[..., :3] / 255. for f in names], 0)        # [N,H,W,3] float64
    hw = imgs.shape[1:3]
    poses = poses.copy()
    poses[:, 0, 4], poses[:, 1, 4] = hw[0], hw[1]                                             # load_llff.py:131-132
    poses[:, 2, 4] = poses[:, 2, 4] * 1. / (1 if factor is None else factor)

    # masks: label/<stem>.png for every frame that is not a 'cutout' / 'pseudo' view (load_llff.py:110-165)
    mask_names = [f for f in names if "cutout" not in f and "pseudo" not in f]
    masks, mask_indices = [], []
    for i, f in enumerate(mask_names):
        p = os.path.join(root, "label", f.split(".")[0] + ".png")
        if os.path.isfile(p):
            raw = _read(p)
            m = cv2.dilate(_plane(raw, hw, raw.max()), np.ones((5, 5), np.uint8), iterations=5)
            if lpips and not prepare and i != len(mask_names) - 5:          # sign marks the views rendered for LPIPS (:161-162)
                m = -m
            masks.append(m); mask_indices.append(i)
        else:
            masks.append(-np.ones(hw))
    masks = np.stack(masks, 0)
    masks = masks / np.max(masks)
    depthdir = os.path.join(root, "depth")
    depth_paths = ([os.path.join(depthdir, f.split(".")[0] + ".png") for f in _frames(depthdir)] if os.path.isdir(depthdir)
                   else [os.path.join(root, "label", f.split(".")[0] + ".png") for f in mask_names])
    depths = []
    for p in depth_paths:
        depths.append(_plane(_read(p), hw, 255.) if os.path.isfile(p) else -np.ones(hw))
    depths = np.stack(depths, 0)

    # [-up, right, back] -> [right, up, back]; float32 from here on (load_llff.py:329-340)
    poses = np.concatenate([poses[:, :, 1:2], -poses[:, :, 0:1], poses[:, :, 2:]], 2).astype(np.float32)
    images = imgs.astype(np.float32)
    masks = np.squeeze(masks).astype(np.float32)
    depths = np.squeeze(depths).astype(np.float32)
    bds = bds.astype(np.float32)
    sc = 1. if bd_factor is None else 1. / (bds.min() * bd_factor)
    poses[:, :3, 3] *= sc
    bds = bds * sc
    if recenter_poses:
        poses = recenter(poses)
    if spherify_poses:
        poses, _, bds, _, _ = spherify(poses, bds)
    elif spherify_hack:
        # the reference scales bds in place inside spherify_poses and divides the result by the same factor again
        # (load_llff.py:356-359): numerically a no-op up to two float32 roundings, which are reproduced here
        _, _, bds_s, s2, _ = spherify(poses, bds)
        bds = bds_s / s2
    # spiral render path — computed for every branch (load_llff.py:377-410)
    c2w = average_pose(poses)
    up = _unit(poses[:, :3, 1].sum(0))
    close, inf = bds.min() * .9, bds.max() * 5.
    focal = 1. / ((1. - .75) / close + .75 / inf)
    rads = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
    n_views, rots = 120, 2
    if path_zflat:
        c2w[:3, 3] = c2w[:3, 3] + (-close * .1) * c2w[:3, 2]
        rads[2] = 0.
        rots, n_views = 1, 60
    render_poses = np.array(spiral_path(c2w, up, rads, focal, zrate=.5, rots=rots, n=n_views)).astype(np.float32)
    c2w = average_pose(poses)
    i_test = int(np.argmin(np.sum(np.square(c2w[:3, 3] - poses[:, :3, 3]), -1)))               # hold-out view (:420-422)
    poses = poses.astype(np.float32)
    if masks.shape[-1] == 3 and masks.ndim == 4:
        masks = masks[..., 0].squeeze()
    if depths.shape[-1] == 3 and depths.ndim == 4:
        depths = depths[..., 0].squeeze()
    return images, poses, bds, render_poses, i_test, masks, depths, mask_indices


# ---------------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------------
def poses_bounds(c2w, hwf, bounds):
    """[N,3,4] camera-to-world in the trainer's [right, up, back | position] convention + (H, W, focal) + [N,2] depth bounds ->
    the N x 17 array of poses_bounds.npy (LLFF stores [-up, right, back])."""
    c2w = np.asarray(c2w, np.float64)
    n = c2w.shape[0]
    m = np.zeros((n, 3, 5))
    m[:, :, 0], m[:, :, 1], m[:, :, 2], m[:, :, 3] = -c2w[:, :, 1], c2w[:, :, 0], c2w[:, :, 2], c2w[:, :, 3]
    m[:, :, 4] = np.asarray(hwf, np.float64)
    return np.concatenate([m.reshape(n, 15), np.asarray(bounds, np.float64).reshape(n, 2)], 1)


def _u8(x):
    x = np.asarray(x)
    return x if x.dtype == np.uint8 else (255 * np.clip(x, 0, 1) + .5).astype(np.uint8)


def _save(path, img):
    img = _u8(img)
    cv2.imwrite(path, img[..., ::-1] if img.ndim == 3 else img)


def write_scene(basedir, images, c2w, focal, bounds, factor=2, masks=None, inpainted=None, depths=None, stem="IMG_{:04d}"):
    """Write a scene directory both loaders accept.  images [N,H,W,3] at the training resolution (float in [0,1] or uint8);
    full-resolution frames are written `factor` x larger (nearest up-sampling: only their shape is ever read), focal is the
    full-resolution focal length.  masks [N,H,W] (non-zero = object, or None for frames without a label file: pass a list
    with None entries), inpainted [N,H,W,3] (default: images), depths [N,H,W] in [0,1] (default: none -> the loaders fall
    back to the label files)."""
    images = np.asarray(images)
    n, h, w = images.shape[:3]
    os.makedirs(os.path.join(basedir, "images"), exist_ok=True)
    sub = os.path.join(basedir, f"images_{factor}")
    for d in ("", "lama_images", "label") + (("depth",) if depths is not None else ()):
        os.makedirs(os.path.join(sub, d), exist_ok=True)
    np.save(os.path.join(basedir, "poses_bounds.npy"), poses_bounds(c2w, (h * factor, w * factor, focal), bounds))
    for i in range(n):
        name = stem.format(i) + ".png"
        frame = _u8(images[i])
        _save(os.path.join(basedir, "images", name), np.repeat(np.repeat(frame, factor, 0), factor, 1))
        _save(os.path.join(sub, name), frame)
        _save(os.path.join(sub, "lama_images", name), frame if inpainted is None else inpainted[i])
        if masks is not None and masks[i] is not None:
            _save(os.path.join(sub, "label", name), (np.asarray(masks[i]) != 0).astype(np.uint8) * 255)
        if depths is not None:
            _save(os.path.join(sub, "depth", name), depths[i])
    return basedir


def synthetic_scene(basedir, n_views=8, hw=(48, 64), factor=2, seed=0, n_unlabelled=1):
    """A small forward-facing scene with the statue dataset's structure: an arc of cameras looking at the origin region,
    random textures, one blob mask per labelled view, smooth inpainted disparities.  Returns what was written."""
    rng = np.random.default_rng(seed)
    h, w = hw
    focal = 0.9 * w * factor
    c2w = []
    for i in range(n_views):
        pos = np.array([rng.uniform(-.5, .5), rng.uniform(-.5, .5), rng.uniform(-.05, .05)])
        z = _unit(pos - np.array([rng.normal(0, .05), rng.normal(0, .05), -4.]))           # cameras look down -z
        c2w.append(look_at(z, np.array([0, 1., 0]), pos))
    c2w = np.stack(c2w, 0)
    bounds = np.stack([rng.uniform(1.2, 1.5, n_views), rng.uniform(7., 9., n_views)], 1)
    images = rng.uniform(0, 1, (n_views, h, w, 3))
    inpainted = np.clip(images + rng.normal(0, .05, images.shape), 0, 1)
    yy, xx = np.mgrid[0:h, 0:w]
    masks = []
    for i in range(n_views):
        if i >= n_views - n_unlabelled:
            masks.append(None)
            continue
        cy, cx, r = rng.uniform(.3, .7) * h, rng.uniform(.3, .7) * w, rng.uniform(.08, .16) * min(h, w)
        masks.append(((yy - cy) ** 2 + (xx - cx) ** 2 < r * r).astype(np.uint8))
    depths = np.clip(.5 + .3 * np.sin(xx / w * 3 + np.arange(n_views)[:, None, None]) * np.cos(yy / h * 2), 0, 1)
    write_scene(basedir, images, c2w, focal, bounds, factor=factor, masks=masks, inpainted=inpainted, depths=depths)
    return dict(c2w=c2w, focal=focal, bounds=bounds, images=images, inpainted=inpainted, masks=masks, depths=depths)


# ---------------------------------------------------------------------------------------------------------------------
# COLMAP sparse model -> depth-supervision rays (load_llff.py:436-501; binary layouts: colmapUtils/read_write_model.py
# :225-257 images.bin, :336-363 points3D.bin)
# ---------------------------------------------------------------------------------------------------------------------
_IMG_HEAD = np.dtype([("id", "<i4"), ("q", "<f8", 4), ("t", "<f8", 3), ("cam", "<i4")])          # 64 bytes
_OBS = np.dtype([("x", "<f8"), ("y", "<f8"), ("p3d", "<i8")])                                  # 24 bytes
_PT_HEAD = np.dtype([("id", "<u8"), ("xyz", "<f8", 3), ("rgb", "u1", 3), ("err", "<f8")])      # 43 bytes (packed)


def read_colmap_images(path):
    """images.bin -> list (file order) of dicts: id, qvec [4] (w, x, y, z), tvec [3], camera_id, name, xys [P,2], point3D_ids [P]."""
    buf = open(path, "rb").read()
    n = int(np.frombuffer(buf, "<u8", 1, 0)[0])
    off, out = 8, []
    for _ in range(n):
        h = np.frombuffer(buf, _IMG_HEAD, 1, off)[0]
        off += _IMG_HEAD.itemsize
        end = buf.index(b"\x00", off)
        name = buf[off:end].decode("utf-8")
        off = end + 1
        npts = int(np.frombuffer(buf, "<u8", 1, off)[0])
        off += 8
        obs = np.frombuffer(buf, _OBS, npts, off)
        off += npts * _OBS.itemsize
        out.append(dict(id=int(h["id"]), qvec=np.array(h["q"]), tvec=np.array(h["t"]), camera_id=int(h["cam"]), name=name,
                        xys=np.stack([obs["x"], obs["y"]], 1) if npts else np.zeros((0, 2)), point3D_ids=obs["p3d"].copy()))
    return out


def read_colmap_points(path):
    """points3D.bin -> dict of arrays in file order: ids [M] uint64, xyz [M,3], rgb [M,3] uint8, error [M] (tracks are skipped)."""
    buf = open(path, "rb").read()
    n = int(np.frombuffer(buf, "<u8", 1, 0)[0])
    ids, xyz, rgb, err = np.zeros(n, np.uint64), np.zeros((n, 3)), np.zeros((n, 3), np.uint8), np.zeros(n)
    off = 8
    for i in range(n):
        h = np.frombuffer(buf, _PT_HEAD, 1, off)[0]
        ids[i], xyz[i], rgb[i], err[i] = h["id"], h["xyz"], h["rgb"], h["err"]
        off += _PT_HEAD.itemsize
        off += 8 + 8 * int(np.frombuffer(buf, "<u8", 1, off)[0])
    return dict(ids=ids, xyz=xyz, rgb=rgb, error=err)


def quat_to_rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def rot_to_quat(R):
    """Unit quaternion (w, x, y, z), w >= 0, of a rotation matrix: Shepperd's method (divide by the largest component)."""
    R = np.asarray(R, np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    cand = np.array([tr, R[0, 0], R[1, 1], R[2, 2]])
    i = int(np.argmax(cand))
    if i == 0:
        w = np.sqrt(1. + tr) / 2
        q = np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    else:
        j, k = i % 3, (i + 1) % 3            # i - 1 is the dominant axis; j, k the other two in cyclic order
        a = i - 1
        v = np.sqrt(1. + R[a, a] - R[j, j] - R[k, k]) / 2
        q = np.zeros(4)
        q[1 + a] = v
        q[0] = (R[k, j] - R[j, k]) / (4 * v)
        q[1 + j] = (R[j, a] + R[a, j]) / (4 * v)
        q[1 + k] = (R[k, a] + R[a, k]) / (4 * v)
    q = q / np.linalg.norm(q)
    return -q if q[0] < 0 else q


def colmap_depth_rays(basedir, factor=8, bd_factor=.75):
    """Sparse depth supervision from <basedir>/sparse/0/{images,points3D}.bin (load_llff.py:448-501): for every registered image
    the depth along its optical axis of each triangulated point it observes (rescaled like the poses), the pixel it was seen at
    (at 1/factor resolution) and a confidence weight 2 exp(-(reprojection error / mean error)^2); points outside the image's
    [near, far] bounds are dropped, images without any are skipped.  Returns a list of {"depth", "coord", "weight"} dicts."""
    images = read_colmap_images(os.path.join(basedir, "sparse", "0", "images.bin"))
    pts = read_colmap_points(os.path.join(basedir, "sparse", "0", "points3D.bin"))
    row_of = {int(i): k for k, i in enumerate(pts["ids"])}
    err_mean = np.mean(pts["error"])
    bds = np.load(os.path.join(basedir, "poses_bounds.npy"))[:, 15:17].astype(np.float32)
    sc = np.float32(1.) if bd_factor is None else 1. / (bds.min() * bd_factor)
    by_id = {im["id"]: im for im in images}
    out = []
    for k in range(len(images)):
        im = by_id[k + 1]                                   # the reference indexes poses / bounds by COLMAP image id - 1
        w2c = np.eye(4)
        w2c[:3, :3], w2c[:3, 3] = quat_to_rot(images[k]["qvec"]), images[k]["tvec"]     # poses are stacked in FILE order (:436-445)
        c2w = np.linalg.inv(w2c)
        seen = im["point3D_ids"] != -1
        rows = np.array([row_of[int(i)] for i in im["point3D_ids"][seen]], np.int64)
        if rows.size == 0:
            continue
        depth = ((pts["xyz"][rows] - c2w[:3, 3]) @ c2w[:3, 2]) * sc
        keep = ~((depth < bds[k, 0] * sc) | (depth > bds[k, 1] * sc))
        if not keep.any():
            continue
        out.append({"depth": depth[keep], "coord": im["xys"][seen][keep] / factor,
                    "weight": 2 * np.exp(-(pts["error"][rows][keep] / err_mean) ** 2)})
    return out


def write_colmap_model(basedir, c2w, focal_full, hw_full, points, errors, rng, p_miss=0.2, p_unmatched=0.1):
    """Write sparse/0/images.bin + points3D.bin for cameras c2w [N,3,4] in the trainer's [right, up, back | position]
    convention: every point is projected into every camera (pinhole, principal point at the image centre), a fraction
    p_miss of the observations is dropped and a fraction p_unmatched of extra keypoints without a 3D point (id -1) added."""
    d = os.path.join(basedir, "sparse", "0")
    os.makedirs(d, exist_ok=True)
    n, m = len(c2w), len(points)
    tracks = [[] for _ in range(m)]
    with open(os.path.join(d, "images.bin"), "wb") as f:
        f.write(np.uint64(n).tobytes())
        for k in range(n):
            R = np.asarray(c2w[k])[:, :3] * np.array([1., -1., -1.])       # COLMAP cameras look down +z with y pointing down
            w2c_R = R.T
            t = -w2c_R @ np.asarray(c2w[k])[:, 3]
            cam = (points - np.asarray(c2w[k])[:, 3]) @ R                    # points in the COLMAP camera frame
            xy = cam[:, :2] / cam[:, 2:3] * focal_full + np.array([hw_full[1] / 2, hw_full[0] / 2])
            obs = [(xy[j, 0], xy[j, 1], j + 1) for j in range(m) if rng.uniform() > p_miss]
            obs += [(rng.uniform(0, hw_full[1]), rng.uniform(0, hw_full[0]), -1) for _ in range(int(p_unmatched * m))]
            order = rng.permutation(len(obs))
            obs = [obs[i] for i in order]
            head = np.zeros(1, _IMG_HEAD)
            head["id"], head["q"], head["t"], head["cam"] = k + 1, rot_to_quat(w2c_R), t, 1
            f.write(head.tobytes()); f.write(f"IMG_{k:04d}.png".encode() + b"\x00"); f.write(np.uint64(len(obs)).tobytes())
            rec = np.zeros(len(obs), _OBS)
            for i, (x, y, pid) in enumerate(obs):
                rec[i] = (x, y, pid)
                if pid > 0:
                    tracks[pid - 1].append((k + 1, i))
            f.write(rec.tobytes())
    with open(os.path.join(d, "points3D.bin"), "wb") as f:
        f.write(np.uint64(m).tobytes())
        for j in range(m):
            head = np.zeros(1, _PT_HEAD)
            head["id"], head["xyz"], head["rgb"], head["err"] = j + 1, points[j], rng.integers(0, 256, 3), errors[j]
            f.write(head.tobytes()); f.write(np.uint64(len(tracks[j])).tobytes())
            f.write(np.asarray(tracks[j], "<i4").reshape(-1, 2).tobytes())
