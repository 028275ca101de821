"""f3, second half: the deblocking filter's edge filtering on the device, hmgpu_deblock, against the golden
pictures of the instrumented reference decoder (tests/golden/deblock_golden.npz) and against the oracle on random inputs."""
import os

import numpy as np
import pytest

import hmgpu
from oracle import binding as B

pytestmark = pytest.mark.gpu


def _golden():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deblock_golden.npz"))
    for i in range(int(z["n_pictures"][0])):
        yield i, {k[len("p%d_" % i):]: z[k] for k in z.files if k.startswith("p%d_" % i)}


def test_deblock_matches_reference_decoder_pictures():
    n = 0
    for i, p in _golden():
        w, h, bdl, bdc, beta, tc, cbo, cro, poc = [int(v) for v in p["params"]]
        with hmgpu.Context(w, h, bdl, 1) as ctx:
            y, cb, cr = ctx.deblock(p["pre_y"], p["pre_cb"], p["pre_cr"], p["bs_ver"], p["bs_hor"], p["qp"], p["nofilter"], beta, tc, cbo, cro)
        assert np.array_equal(y, p["post_y"]), (i, poc, np.argwhere(y != p["post_y"])[:4])
        assert np.array_equal(cb, p["post_cb"]) and np.array_equal(cr, p["post_cr"]), (i, poc)
        n += 1
    assert n >= 9


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_deblock_random_maps_match_oracle(bit_depth):
    """random pictures with random boundary strengths, QPs (0..51), no-filter flags and offsets: every branch of the filters (strong,
    weak with and without the second sample, thresholds at their edges), picture sizes that are not multiples of 8"""
    O = B.oracle()
    rng = np.random.default_rng(300 + bit_depth)
    mx = (1 << bit_depth) - 1
    for it in range(6):
        w, h = [(416, 240), (200, 120), (136, 72), (64, 64), (1920, 1080), (36, 20)][it]
        base = rng.integers(0, mx + 1, ((h + 15) // 16, (w + 15) // 16)).astype(np.int64)
        y = np.kron(base, np.ones((16, 16), np.int64))[:h, :w]
        y = np.clip(y // 4 + mx // 3 + rng.integers(-3, 4, (h, w)) * (1 + it % 3), 0, mx).astype(np.int16)     # blocky and smooth: the filters switch on
        if it == 3:
            y = rng.integers(0, mx + 1, (h, w)).astype(np.int16)
        cb = np.clip(y[::2, ::2] // 2 + rng.integers(-2, 3, (h // 2, w // 2)), 0, mx).astype(np.int16)
        cr = rng.integers(0, mx + 1, (h // 2, w // 2)).astype(np.int16)
        uw, uh = (w + 3) // 4, (h + 3) // 4
        bs_ver = rng.integers(0, 3, (uh, uw)).astype(np.uint8)
        bs_hor = rng.integers(0, 3, (uh, uw)).astype(np.uint8)
        bs_ver[:, 0] = 0; bs_hor[0, :] = 0                     # no edge at the picture boundary
        qp = rng.integers(0, 52, (uh, uw)).astype(np.int8) if it % 2 else np.full((uh, uw), 30 + it, np.int8)
        nf = (rng.integers(0, 12, (uh, uw)) == 0).astype(np.uint8)
        beta, tc, cbo, cro = int(rng.integers(-6, 7)), int(rng.integers(-6, 7)), int(rng.integers(-12, 13)), int(rng.integers(-12, 13))
        ey, ecb, ecr = y.copy(), cb.copy(), cr.copy()
        O.hmo_deblock_picture(ey.ctypes.data, ecb.ctypes.data, ecr.ctypes.data, w, h, bit_depth, bit_depth, bs_ver.ctypes.data, bs_hor.ctypes.data,
                              qp.ctypes.data, nf.ctypes.data, beta, tc, cbo, cro)
        with hmgpu.Context(max(w, 64), max(h, 64), bit_depth, 1) as ctx:
            gy, gcb, gcr = ctx.deblock(y, cb, cr, bs_ver, bs_hor, qp, nf, beta, tc, cbo, cro)
            if it == 0:
                with pytest.raises(hmgpu.HmGpuError):
                    ctx.deblock(y, cb, cr, bs_ver + 3, bs_hor, qp, nf)
        assert it == 3 or (ey != y).any()             # (pure noise, case 3: every edge fails the activity test, nothing is filtered)
        assert np.array_equal(gy, ey), (it, np.argwhere(gy != ey)[:4])
        assert np.array_equal(gcb, ecb) and np.array_equal(gcr, ecr), it
