"""Deterministic synthetic 4:2:0 video: the reference ships no sequences
(cfg/per-sequence/*.cfg point at absent files), so every test, bench and encode in this
repo uses this generator.

Frame n = 8x8-block random luma texture blurred with a 5x5 box, translated (+3,-2) px per
frame, a 48x48 inverted box moving (+5,+2) px per frame, per-frame Gaussian noise sigma=2,
clipped to the bit depth.  Cb = Y(2x2 subsample)/4 + 96 (scaled to bit depth), Cr = mid-grey.
"""
import numpy as np


def _texture(rng, w, h):
    bw, bh = (w + 7) // 8 + 16, (h + 7) // 8 + 16
    blocks = rng.integers(0, 256, size=(bh, bw)).astype(np.float64)
    tex = np.kron(blocks, np.ones((8, 8)))
    # 5x5 box blur (separable, edge-replicated)
    k = 5
    pad = np.pad(tex, k // 2, mode="edge")
    cs = np.cumsum(np.pad(pad, ((1, 0), (0, 0))), axis=0)
    v = (cs[k:, :] - cs[:-k, :]) / k
    cs = np.cumsum(np.pad(v, ((0, 0), (1, 0))), axis=1)
    return (cs[:, k:] - cs[:, :-k]) / k


def luma_frames(width, height, n_frames, bit_depth=8, seed=1234):
    """-> uint16 array [n_frames, height, width]"""
    rng = np.random.default_rng(seed)
    tex = _texture(rng, width, height)
    th, tw = tex.shape
    out = np.empty((n_frames, height, width), dtype=np.uint16)
    maxv = (1 << bit_depth) - 1
    scale = 1 << (bit_depth - 8)
    for n in range(n_frames):
        ox, oy = 64 + 3 * n, 64 - 2 * n
        ys = (np.arange(height) + oy) % th
        xs = (np.arange(width) + ox) % tw
        f = tex[np.ix_(ys, xs)].copy()
        bx, by = (20 + 5 * n) % max(1, width - 48), (16 + 2 * n) % max(1, height - 48)
        f[by:by + 48, bx:bx + 48] = 255.0 - f[by:by + 48, bx:bx + 48]
        f = f + rng.normal(0.0, 2.0, size=f.shape)
        out[n] = np.clip(np.rint(f * scale), 0, maxv).astype(np.uint16)
    return out


def yuv420_frames(width, height, n_frames, bit_depth=8, seed=1234):
    """-> (Y, Cb, Cr) uint16 arrays"""
    y = luma_frames(width, height, n_frames, bit_depth, seed)
    scale = 1 << (bit_depth - 8)
    cb = (y[:, ::2, ::2] // 4 + 96 * scale).astype(np.uint16)
    cr = np.full_like(cb, 128 * scale)
    return y, cb, cr


def write_yuv(path, width, height, n_frames, bit_depth=8, seed=1234):
    """planar 4:2:0 file: 1 byte/sample at 8 bit, 2 bytes little-endian above"""
    y, cb, cr = yuv420_frames(width, height, n_frames, bit_depth, seed)
    dt = np.uint8 if bit_depth == 8 else np.dtype("<u2")
    with open(path, "wb") as f:
        for n in range(n_frames):
            for p in (y[n], cb[n], cr[n]):
                f.write(np.ascontiguousarray(p.astype(dt)).tobytes())
    return path
