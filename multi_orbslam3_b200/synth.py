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


def random_vocabulary(k=10, L=3, seed=0, irregular=False, shuffle=False, stop_frac=0.02):
    """Synthetic DBoW2 ORB vocabulary as a node table in text-file order (ORBvoc.txt itself is not redistributable here):
    returns (parent int32[n], is_leaf uint8[n], desc uint8[n,32], weight float64[n]); node 0 is the root, parent[i] < i.
    A child's descriptor is its parent's with ~24 random bits flipped; leaf weights look like idf values, a few are 0
    (stopped words).  irregular: early leaves and nodes with fewer than k children; shuffle: siblings not contiguous."""
    rng = np.random.default_rng(seed)
    parent = [np.array([-1], np.int32)]; desc = [rng.integers(0, 256, (1, 32), dtype=np.uint8)]
    leaf = [np.array([0], np.uint8)]
    prev_ids = np.array([0], np.int64); prev_desc = desc[0]; n = 1
    for level in range(1, L + 1):
        cnt = np.full(len(prev_ids), k, np.int64)
        if irregular and level > 1:
            cnt = rng.integers(2, k + 1, len(prev_ids))
            cnt[rng.random(len(prev_ids)) < 0.1] = 0                      # early leaf
        par = np.repeat(prev_ids, cnt)
        m = len(par)
        if m == 0:
            break
        flips = np.zeros((m, 256), np.uint8)
        cols = rng.integers(0, 256, (m, 24))
        np.put_along_axis(flips, cols, 1, axis=1)
        d = np.repeat(prev_desc, cnt, axis=0) ^ np.packbits(flips, axis=1, bitorder="little")
        ids = n + np.arange(m)
        parent.append(par.astype(np.int32)); desc.append(d); leaf.append(np.zeros(m, np.uint8))
        if irregular and level > 1:
            early = prev_ids[cnt == 0]
            flat = np.concatenate(leaf); flat[early] = 1
            leaf = [flat]
        prev_ids, prev_desc, n = ids, d, n + m
    parent = np.concatenate(parent); desc = np.concatenate(desc); leaf = np.concatenate(leaf)
    has_child = np.zeros(n, bool); has_child[parent[1:]] = True
    leaf = (~has_child).astype(np.uint8); leaf[0] = 0
    weight = np.where(leaf == 1, rng.uniform(0.5, 9.0, n), 0.0)
    stopped = (leaf == 1) & (rng.random(n) < stop_frac)
    weight[stopped] = 0.0
    if shuffle:
        # random order that keeps every parent before its children
        pri = np.zeros(n); step = rng.uniform(0.01, 1.0, n)
        for i in range(1, n):
            pri[i] = pri[parent[i]] + step[i]
        order = np.argsort(pri, kind="stable")
        new_id = np.empty(n, np.int64); new_id[order] = np.arange(n)
        parent = np.where(parent[order] >= 0, new_id[np.maximum(parent[order], 0)], -1).astype(np.int32)
        desc, leaf, weight = desc[order], leaf[order], weight[order]
    return parent, leaf, np.ascontiguousarray(desc), weight.astype(np.float64)


def vocabulary_queries(vocab, n, seed=0):
    """n descriptors: noisy copies of random leaf descriptors (most) and uniformly random ones (some)."""
    parent, leaf, desc, _ = vocab
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(leaf)[0]
    q = desc[rng.choice(leaves, n)].copy()
    flips = np.zeros((n, 256), np.uint8)
    np.put_along_axis(flips, rng.integers(0, 256, (n, 12)), 1, axis=1)
    q ^= np.packbits(flips, axis=1, bitorder="little")
    rnd = rng.random(n) < 0.1
    q[rnd] = rng.integers(0, 256, (int(rnd.sum()), 32), dtype=np.uint8)
    return np.ascontiguousarray(q)
