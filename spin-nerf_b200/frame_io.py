"""Frame sink of render_path (DS_NeRF/run_nerf.py:221-295).

The reference ends every frame with blocking `.cpu().numpy()` calls on the default stream (12 MB for a 1008x756 RGB+disp
frame, 0.8 GB more when the per-sample `weights` / `z_vals` dumps are on) and writes the PNG / NPY files before the next
frame's first kernel is launched.  Here the render thread only enqueues: a frame's tensors are copied device->host on a
side stream into one of `depth` pinned staging sets (double buffering), a worker thread waits for that copy's event, files
the frame (same directory layout, file names and array contents as the reference) and hands the staging set back, so
frame i's transfer and file writes overlap frame i+1's kernels.

CPU tensors / numpy arrays are accepted too (no staging): the file formats are testable without a GPU.
"""
from __future__ import annotations

import os
import queue
import threading

import numpy as np
import torch


def to8b(x):
    """run_nerf_helpers.py:18."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def dump_frame(savedir, i, rgb, gt_img, depth, disp, weights, z_vals, alpha, c2w):
    """One frame's files exactly as run_nerf.py:231-295 lays them out: rgb/%06d.png (8-bit), images/%06d.png (ground
    truth, if given), depth|disp|weight|z|alpha/%06d.npy (float32 arrays as rendered), pose/%06d.txt (4x4 c2w).
    All arguments are host arrays.  PNGs are written with OpenCV (imageio is not in the image): same pixels."""
    import cv2
    sub = lambda d: os.path.join(savedir, d)
    for d in ['rgb', 'depth', 'images', 'weight', 'z', 'pose', 'disp'] + (['alpha'] if alpha is not None else []):
        os.makedirs(sub(d), exist_ok=True)
    name = '{:06d}'.format(i)
    rgb8 = to8b(np.nan_to_num(rgb))          # the reference's `rgb8[np.isnan(rgb8)] = 0` (a no-op on uint8) made effective
    cv2.imwrite(os.path.join(sub('rgb'), name + '.png'), np.ascontiguousarray(rgb8[..., ::-1]))
    if gt_img is not None:
        cv2.imwrite(os.path.join(sub('images'), name + '.png'), np.ascontiguousarray(to8b(_np(gt_img))[..., ::-1]))
    np.save(os.path.join(sub('depth'), name + '.npy'), depth)
    np.save(os.path.join(sub('disp'), name + '.npy'), disp)
    np.save(os.path.join(sub('weight'), name + '.npy'), weights)
    np.save(os.path.join(sub('z'), name + '.npy'), z_vals)
    if alpha is not None:
        np.save(os.path.join(sub('alpha'), name + '.npy'), alpha)
    pose = np.concatenate([np.asarray(c2w, dtype=np.float64)[:3, :4], np.array([[0, 0, 0, 1]])], axis=0)
    np.savetxt(os.path.join(sub('pose'), name + '.txt'), pose)


def write_video(path, frames, fps=30):
    """imageio.mimwrite(path, to8b(frames), fps=30, quality=8) of run_nerf.py:1211-1217, 1661-1671 with OpenCV's mp4
    writer.  frames: [P,H,W,3] RGB or [P,H,W] in [0,1] (or already uint8)."""
    import cv2
    frames = np.asarray(frames)
    if frames.dtype != np.uint8:
        frames = to8b(np.nan_to_num(frames))
    if frames.ndim == 3:
        frames = np.repeat(frames[..., None], 3, -1)
    h, w = frames.shape[1:3]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'mp4v'), float(fps), (w, h))
    if not vw.isOpened():
        raise RuntimeError(f"write_video: OpenCV cannot open an mp4 writer for {path}")
    for f in frames:
        vw.write(np.ascontiguousarray(f[..., ::-1]))
    vw.release()


class FrameWriter:
    """submit(i, c2w, frame) per rendered frame, then close() -> (rgbs [P,H,W,3], disps [P,H,W]) float32 numpy stacks in
    frame order, like render_path returns them.  `frame` maps names to tensors: rgb, disp (always), depth, weights,
    z_vals, alpha (needed only when savedir is set).  Frames may be submitted in any order and must all have the same
    shapes."""

    KEYS = ("rgb", "disp", "depth", "weights", "z_vals", "alpha")

    def __init__(self, savedir=None, gt_imgs=None, need_alpha=False, depth=2):
        self.savedir, self.gt_imgs, self.need_alpha = savedir, gt_imgs, bool(need_alpha)
        self.depth = max(1, int(depth))
        self.free = queue.Queue()
        for s in range(self.depth):
            self.free.put({})                       # staging sets (name -> pinned host tensor), filled lazily
        self.jobs = queue.Queue()
        self.results = {}
        self.error = None
        self.copy_stream = None
        self.worker = threading.Thread(target=self._run, name="spn-frame-writer", daemon=True)
        self.worker.start()

    def _wanted(self, frame):
        keys = ["rgb", "disp"]
        if self.savedir is not None:
            keys += ["depth", "weights", "z_vals"] + (["alpha"] if self.need_alpha else [])
        missing = [k for k in keys if frame.get(k) is None]
        if missing:
            raise KeyError(f"FrameWriter.submit: frame lacks {missing}")
        return keys

    def submit(self, i, c2w, frame):
        if self.error is not None:
            raise RuntimeError("frame writer failed") from self.error
        keys = self._wanted(frame)
        c2w = _np(c2w).astype(np.float64)
        on_gpu = torch.is_tensor(frame["rgb"]) and frame["rgb"].is_cuda
        if not on_gpu:
            self.jobs.put((int(i), c2w, {k: _np(frame[k]) for k in keys}, None, None, None))
            return
        dev = frame["rgb"].device
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=dev)
        staging = self.free.get()                   # blocks while all `depth` staging sets are in flight
        src = {k: frame[k].detach() for k in keys}  # referenced by the job until its copy has completed
        self.copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.copy_stream):
            for k, t in src.items():
                buf = staging.get(k)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = staging[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                buf.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.jobs.put((int(i), c2w, None, staging, done, src))

    def _run(self):
        while True:
            job = self.jobs.get()
            if job is None:
                return
            i, c2w, host, staging, done, src = job
            try:
                if staging is not None:
                    torch.cuda.set_device(next(iter(src.values())).device)
                    done.synchronize()
                    keys = list(src.keys())
                    host = {k: staging[k].numpy() for k in keys}
                # the returned stacks own their memory (the staging set is reused two frames later)
                self.results[i] = (np.array(host["rgb"], dtype=np.float32, copy=True),
                                   np.array(host["disp"], dtype=np.float32, copy=True))
                if self.savedir is not None:
                    gt = None if self.gt_imgs is None else self.gt_imgs[i]
                    dump_frame(self.savedir, i, host["rgb"], gt, host["depth"], host["disp"], host["weights"], host["z_vals"],
                               host.get("alpha") if self.need_alpha else None, c2w)
            except BaseException as e:              # surfaced by the next submit() / close()
                self.error = e
            finally:
                if staging is not None:
                    self.free.put(staging)

    def close(self):
        self.jobs.put(None)
        self.worker.join()
        if self.error is not None:
            raise RuntimeError("frame writer failed") from self.error
        order = sorted(self.results)
        if not order:
            return np.zeros((0,), np.float32), np.zeros((0,), np.float32)
        return np.stack([self.results[i][0] for i in order], 0), np.stack([self.results[i][1] for i in order], 0)
