"""cv2-backed stand-in for the few imageio calls DS_NeRF uses (imread / imwrite / mimwrite)."""
import cv2
import numpy as np


def imread(path, *a, **k):
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    return img


def imwrite(path, img, *a, **k):
    img = np.asarray(img)
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    cv2.imwrite(path, img)


def mimwrite(path, frames, fps=30, quality=8, **k):
    frames = [np.asarray(f) for f in frames]
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    for f in frames:
        if f.ndim == 2:
            f = np.repeat(f[..., None], 3, -1)
        vw.write(f[..., ::-1].astype(np.uint8))
    vw.release()
