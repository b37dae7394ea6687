"""Synthetic grayscale frames for the BASELINE.json configs (SURVEY.md Appendix B recipe)."""
from __future__ import annotations

import numpy as np


def synth(W: int, H: int, seed: int) -> np.ndarray:
    """Gaussian blobs of random size / contrast on a grey background plus low-pass
    noise; float32 in [0, 255].  Deterministic for a given (W, H, seed)."""
    import cv2

    r = np.random.default_rng(seed)
    img = np.full((H, W), 128.0, np.float32)
    n = round(W * H / 400)
    cx = r.uniform(0, W, n)
    cy = r.uniform(0, H, n)
    sg = np.exp(r.uniform(np.log(1.2), np.log(12.0), n))
    amp = r.uniform(8, 64, n) * r.choice([-1.0, 1.0], n)
    for k in range(n):
        R = int(np.ceil(3 * sg[k]))
        x0 = int(np.floor(cx[k]))
        y0 = int(np.floor(cy[k]))
        xa, xb = max(0, x0 - R), min(W, x0 + R + 1)
        ya, yb = max(0, y0 - R), min(H, y0 + R + 1)
        if xa >= xb or ya >= yb:
            continue
        xs = np.arange(xa, xb, dtype=np.float32) - np.float32(cx[k])
        ys = np.arange(ya, yb, dtype=np.float32) - np.float32(cy[k])
        img[ya:yb, xa:xb] += np.float32(amp[k]) * np.exp(
            -(ys[:, None] ** 2 + xs[None, :] ** 2) / np.float32(2 * sg[k] * sg[k])).astype(np.float32)
    noise = r.random((H, W), dtype=np.float32) - np.float32(0.5)
    noise = cv2.GaussianBlur(noise, (0, 0), 1.5)
    noise *= np.float32(2.0) / max(1e-9, float(np.abs(noise).max()))
    return np.clip(img + noise, 0, 255).astype(np.float32)
