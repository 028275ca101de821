"""f3, first half: SAO statistics of a picture component on the device, hmgpu_sao_stats, against the oracle
(which tests/test_oracle_vs_ref.py pins to the reference's TEncSampleAdaptiveOffset::getBlkStats)."""
import numpy as np
import pytest

import hmgpu
from oracle import binding as B

pytestmark = pytest.mark.gpu


def _oracle_picture(rec, org, ctu_w, ctu_h, skip_r, skip_b, bit_depth, ctu_flags=None):
    O = B.oracle()
    h, w = rec.shape
    cx, cy = (w + ctu_w - 1) // ctu_w, (h + ctu_h - 1) // ctu_h
    # a frame around the picture: the reference never reads it where a neighbour is unavailable, the oracle neither
    pr = np.pad(rec, 2, mode="constant", constant_values=-777).astype(np.int16)
    po = np.pad(org, 2, mode="constant").astype(np.int16)
    st = w + 4
    out = np.zeros((cx * cy, 5, 2, 32), np.int64)
    for c in range(cx * cy):
        x0, y0 = (c % cx) * ctu_w, (c // cx) * ctu_h
        bw, bh = min(ctu_w, w - x0), min(ctu_h, h - y0)
        left, above = x0 > 0, y0 > 0
        al = left and above
        if ctu_flags is not None:
            f = int(ctu_flags[c]); left, above, al = bool(f & 1), bool(f & 4), bool(f & 16)
        right, below = x0 + ctu_w < w, y0 + ctu_h < h
        ar = y0 > 0 and right
        flags = int(left) | int(right) << 1 | int(above) << 2 | int(below) << 3 | int(al) << 4 | int(ar) << 5
        d, n = np.zeros((5, 32), np.int64), np.zeros((5, 32), np.int64)
        off = (y0 + 2) * st + x0 + 2
        O.hmo_sao_blk_stats(B.ptr(pr, off), st, B.ptr(po, off), st, bw, bh, flags, skip_r, skip_b, bit_depth, d, n)
        out[c, :, 0], out[c, :, 1] = d, n
    return out


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("geom", [(416, 240, 64), (208, 120, 32), (200, 136, 64)])
def test_sao_stats_match_oracle(bit_depth, geom):
    w, h, ctu = geom
    rng = np.random.default_rng(w + bit_depth)
    mx = (1 << bit_depth) - 1
    rec = rng.integers(0, mx + 1, (h, w)).astype(np.int16)
    rec[: h // 2] = (rec[: h // 2] // 32 * 32 + rng.integers(0, 3, (h // 2, w))).astype(np.int16)    # flat areas: ties in the edge classes
    org = np.clip(rec + rng.integers(-8, 9, rec.shape), 0, mx).astype(np.int16)
    cases = [(np.array([5, 5, 5, 5, 5], np.int32), np.array([4, 4, 4, 4, 4], np.int32), None),
             (np.array([3, 3, 3, 3, 3], np.int32), np.array([2, 2, 2, 2, 2], np.int32), None),
             (np.array([5, 5, 5, 5, 4], np.int32), np.array([3, 3, 3, 3, 4], np.int32), "random")]
    with hmgpu.Context(w, h, bit_depth, 1) as ctx:
        for skip_r, skip_b, fl in cases:
            n_ctus = ((w + ctu - 1) // ctu) * ((h + ctu - 1) // ctu)
            flags = None
            if fl is not None:      # slice / tile boundaries: left, above, above-left switched off at random (never switched on at the picture edge)
                cx = (w + ctu - 1) // ctu
                flags = np.zeros(n_ctus, np.uint8)
                for c in range(n_ctus):
                    left, above = c % cx > 0, c // cx > 0
                    f = (1 if left and rng.integers(0, 4) else 0) | (4 if above and rng.integers(0, 4) else 0)
                    f |= 16 if (left and above and rng.integers(0, 4)) else 0
                    flags[c] = f
            got = ctx.sao_stats(rec, org, ctu, ctu, skip_r, skip_b, flags)
            exp = _oracle_picture(rec, org, ctu, ctu, skip_r, skip_b, bit_depth, flags)
            bad = np.argwhere(got != exp)
            assert bad.size == 0, "ctu %d type %d %s class %d: %d vs %d" % (bad[0][0], bad[0][1], ["diff", "count"][bad[0][2]], bad[0][3],
                                                                            got[tuple(bad[0])], exp[tuple(bad[0])])
            assert int(got[:, 4, 1].sum()) > 0 and int(got[:, :4, 1].sum()) > 0
        with pytest.raises(hmgpu.HmGpuError):
            ctx.sao_stats(rec, org, 4, 64, cases[0][0], cases[0][1])


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("geom", [(416, 240, 64), (208, 120, 32)])
def test_sao_apply_matches_oracle(bit_depth, geom):
    """SAO applied to a component: every CTU its own type (or off) and offsets, picture-edge CTUs, partial CTUs, random slice / tile
    boundary flags; against the oracle block by block."""
    O = B.oracle()
    w, h, ctu = geom
    rng = np.random.default_rng(7 * w + bit_depth)
    mx = (1 << bit_depth) - 1
    rec = rng.integers(0, mx + 1, (h, w)).astype(np.int16)
    rec[:, : w // 2] = (rec[:, : w // 2] // 32 * 32 + rng.integers(0, 3, (h, w // 2))).astype(np.int16)
    rec[: h // 4] = rng.choice(np.array([0, 1, mx - 1, mx], np.int16), (h // 4, w))                 # clipping at both ends
    cx, cy = (w + ctu - 1) // ctu, (h + ctu - 1) // ctu
    n = cx * cy
    types = rng.integers(-1, 5, n).astype(np.int8)
    offsets = np.zeros((n, 32), np.int32)
    for c in range(n):
        if 0 <= types[c] < 4:
            offsets[c, :5] = [rng.integers(0, 8), rng.integers(0, 8), 0, -rng.integers(0, 8), -rng.integers(0, 8)]
        elif types[c] == 4:
            b0 = int(rng.integers(0, 29)); offsets[c, b0:b0 + 4] = rng.integers(-7, 8, 4)
    for use_flags in (False, True):
        flags = None
        if use_flags:
            flags = np.zeros(n, np.uint8)
            for c in range(n):
                l, r, a, b = c % cx > 0, c % cx < cx - 1, c // cx > 0, c // cx < cy - 1
                geo = [l, r, a, b, l and a, r and a, l and b, r and b]
                flags[c] = sum((1 << k) for k in range(8) if geo[k] and rng.integers(0, 4))
        with hmgpu.Context(w, h, bit_depth, 1) as ctx:
            got = ctx.sao_apply(rec, ctu, ctu, types, offsets, flags)
        exp = rec.copy()
        pr = np.pad(rec, 2, mode="constant", constant_values=-999).astype(np.int16)
        pe = np.pad(exp, 2, mode="constant").astype(np.int16)
        st = w + 4
        for c in range(n):
            if types[c] < 0:
                continue
            x0, y0 = (c % cx) * ctu, (c // cx) * ctu
            bw, bh = min(ctu, w - x0), min(ctu, h - y0)
            if flags is None:
                l, r, a, b = x0 > 0, x0 + ctu < w, y0 > 0, y0 + ctu < h
                f = int(l) | int(r) << 1 | int(a) << 2 | int(b) << 3 | int(l and a) << 4 | int(r and a) << 5 | int(l and b) << 6 | int(r and b) << 7
            else:
                f = int(flags[c])
            off = (y0 + 2) * st + x0 + 2
            O.hmo_sao_offset_block(int(types[c]), offsets[c], B.ptr(pr, off), st, B.ptr(pe, off), st, bw, bh, f, bit_depth)
        assert np.array_equal(got, pe[2:-2, 2:-2]), np.argwhere(got != pe[2:-2, 2:-2])[:5]
