"""Minimal imageio stand-in (absent from this image) for the reference driver and loader: imread / imwrite through
cv2, mimwrite records the stack as .npy next to the requested path."""
import cv2
import numpy as np


def imread(path, **_kw):
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    return img[..., ::-1].copy() if img.ndim == 3 else img


def imwrite(path, img, **_kw):
    img = np.asarray(img)
    cv2.imwrite(str(path), img[..., ::-1] if img.ndim == 3 else img)


def mimwrite(path, frames, **_kw):
    np.save(str(path) + ".frames.npy", np.asarray(frames))


v2 = None
