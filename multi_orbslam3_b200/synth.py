"""Seeded synthetic grayscale frames.  numpy only; shared by tests and bench.

S-rects: mid-gray canvas, 400 filled rectangles of random size 6-90 px and random intensity,
plus N(0,4) noise.  Frame t+1 of a stream is frame t's canvas translated by (3,2) px with fresh noise.
S-noise: uniform noise box-smoothed and min-max normalised (corner-dense worst case).
"""
import numpy as np


def _canvas(rng, w, h, n_rect=400, margin=64):
    W, H = w + 2 * margin, h + 2 * margin
    img = np.full((H, W), 128, np.float32)
    for _ in range(n_rect):
        rw, rh = int(rng.integers(6, 91)), int(rng.integers(6, 91))
        x0, y0 = int(rng.integers(0, W - 6)), int(rng.integers(0, H - 6))
        img[y0:y0 + rh, x0:x0 + rw] = float(rng.integers(0, 256))
    return img


def rects_stream(w, h, n_frames, seed=0, shift=(3, 2), noise_sigma=4.0, n_rect=400):
    """n_frames frames [n,h,w] u8: one canvas panned by `shift` px per frame, fresh noise per frame."""
    rng = np.random.default_rng(seed)
    margin = 64 + max(abs(shift[0]), abs(shift[1])) * n_frames
    canvas = _canvas(rng, w, h, n_rect, margin)
    out = np.empty((n_frames, h, w), np.uint8)
    for t in range(n_frames):
        x0, y0 = margin + shift[0] * t, margin + shift[1] * t
        f = canvas[y0:y0 + h, x0:x0 + w] + rng.normal(0.0, noise_sigma, (h, w)).astype(np.float32)
        out[t] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return out


def rects_frame(w, h, seed=0, **kw):
    return rects_stream(w, h, 1, seed, **kw)[0]


def noise_frame(w, h, seed=0):
    """Corner-dense worst case: uniform noise, 5x5 box smoothing, min-max normalised."""
    rng = np.random.default_rng(seed)
    a = rng.random((h + 4, w + 4)).astype(np.float32)
    c = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), 0), 1)
    s = c[5:, 5:] - c[:-5, 5:] - c[5:, :-5] + c[:-5, :-5]
    s = (s - s.min()) / (s.max() - s.min())
    return np.clip(np.rint(s * 255), 0, 255).astype(np.uint8)


def stereo_pair(w, h, seed=0, disparity=12, n_rect=400):
    """Left/right views of one canvas: the right view is the left shifted by `disparity` px."""
    rng = np.random.default_rng(seed)
    margin = 64 + disparity
    canvas = _canvas(rng, w, h, n_rect, margin)
    out = []
    for dx in (0, disparity):
        f = canvas[margin:margin + h, margin + dx:margin + dx + w] + rng.normal(0.0, 4.0, (h, w)).astype(np.float32)
        out.append(np.clip(np.rint(f), 0, 255).astype(np.uint8))
    return out[0], out[1]


def random_descriptors(n, seed=0, dup_frac=0.0):
    """n x 32 random descriptor bytes; dup_frac of rows are near-duplicates (<= 8 flipped bits) of earlier rows."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ndup = int(n * dup_frac)
    if ndup and n > 1:
        dst = rng.choice(np.arange(1, n), ndup, replace=False)
        for j in dst:
            src = int(rng.integers(0, j))
            row = d[src].copy()
            for _ in range(int(rng.integers(0, 9))):
                b = int(rng.integers(0, 256))
                row[b >> 3] ^= np.uint8(1 << (b & 7))
            d[j] = row
    return d
