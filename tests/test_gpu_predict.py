"""GPU parity: PU motion compensation (luma + 4:2:0 chroma, uni / bi) and the prediction-error costs of
merge estimation / AMVP against the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

import hmgpu
from oracle import binding as B
from util import M

pytestmark = pytest.mark.gpu

W, H = 416, 240
CM = 40  # chroma padding


def _pad(plane, m):
    h, w = plane.shape
    out = np.zeros((h + 2 * m, w + 2 * m), np.int16)
    B.oracle().hmo_extend_border(np.ascontiguousarray(plane, np.int16), w, h, m, out)
    return out


def _pictures(bit_depth, n, seed):
    rng = np.random.default_rng(seed)
    mx = 1 << bit_depth
    return [(rng.integers(0, mx, (H, W)).astype(np.int16), rng.integers(0, mx, (H // 2, W // 2)).astype(np.int16),
             rng.integers(0, mx, (H // 2, W // 2)).astype(np.int16)) for _ in range(n)]


def _jobs(rng, n, n_refs, bi_every=3):
    shapes = [(8, 8), (16, 16), (64, 64), (4, 8), (8, 4), (12, 16), (16, 12), (32, 24), (24, 32), (64, 32), (16, 4), (4, 16), (32, 8)]
    jobs = np.zeros(n, hmgpu.PRED_JOB)
    off = 0
    for i in range(n):
        w, h = shapes[i % len(shapes)]
        x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
        y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
        j = jobs[i]
        j["pu_x"], j["pu_y"], j["pu_w"], j["pu_h"] = x, y, w, h
        lists = [0, 1] if i % bi_every == 0 else [int(rng.integers(0, 2))]
        j["ref_slot"] = (-1, -1)
        for l in lists:
            j["ref_slot"][l] = int(rng.integers(0, n_refs))
            # clipMv-like range: the block may hang up to ~70 samples outside the picture
            j["mv_x"][l] = int(rng.integers(max(-300, (-70 - x) * 4), min(300, (W + 6 - x - w) * 4)))
            j["mv_y"][l] = int(rng.integers(max(-300, (-70 - y) * 4), min(300, (H + 6 - y - h) * 4)))
            if i % 7 == 0:
                j["mv_x"][l] &= ~3          # integer / half phases exercise the filterCopy branches
            if i % 5 == 0:
                j["mv_y"][l] &= ~7
        j["dst_offset"] = off
        off += w * h * 3 // 2
    return jobs, off


def _oracle_component(O, pads, j, comp, bit_depth):
    w, h = int(j["pu_w"]), int(j["pu_h"])
    cw, ch = (w, h) if comp == 0 else (w // 2, h // 2)
    used = [l for l in range(2) if j["ref_slot"][l] >= 0]
    bi = len(used) == 2
    preds = []
    for l in used:
        pad = pads[int(j["ref_slot"][l])][comp]
        m = M if comp == 0 else CM
        pw = pad.shape[1]
        x, y = (int(j["pu_x"]), int(j["pu_y"])) if comp == 0 else (int(j["pu_x"]) // 2, int(j["pu_y"]) // 2)
        d = np.zeros((ch, cw), np.int16)
        O.hmo_pred_inter_blk(int(comp != 0), B.ptr(pad, (y + m) * pw + x + m), pw, int(j["mv_x"][l]), int(j["mv_y"][l]), cw, ch,
                             int(bi), bit_depth, B.ptr(d), cw)
        preds.append(d)
    if not bi:
        return preds[0]
    out = np.zeros((ch, cw), np.int16)
    O.hmo_add_avg(B.ptr(preds[0]), cw, B.ptr(preds[1]), cw, cw, ch, bit_depth, B.ptr(out), cw)
    return out


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_predict_matches_oracle(bit_depth):
    O = B.oracle()
    rng = np.random.default_rng(23)
    pics = _pictures(bit_depth, 2, 31)
    pads = [(_pad(y, M), _pad(cb, CM), _pad(cr, CM)) for (y, cb, cr) in pics]
    jobs, n_dst = _jobs(rng, 160, 2)
    with hmgpu.Context(W, H, bit_depth, 2) as ctx:
        for s, (y, cb, cr) in enumerate(pics):
            ctx.ref_upload(s, y, cb, cr)
        got = ctx.predict(jobs, n_dst, with_chroma=True)
        got_y = ctx.predict(jobs, n_dst, with_chroma=False)
    for i, j in enumerate(jobs):
        w, h = int(j["pu_w"]), int(j["pu_h"])
        off = int(j["dst_offset"])
        sizes = [w * h, w * h // 4, w * h // 4]
        for comp in range(3):
            exp = _oracle_component(O, pads, j, comp, bit_depth).ravel()
            blk = got[off:off + sizes[comp]]
            assert np.array_equal(blk, exp), (i, comp, j)
            if comp == 0:
                assert np.array_equal(got_y[off:off + sizes[0]], exp), (i, "luma only")
            off += sizes[comp]


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_pred_error_matches_oracle(bit_depth):
    """merge-estimation cost (xGetInterPredictionError: MC + HADS) and the AMVP template distortion (MC + SAD)"""
    O = B.oracle()
    rng = np.random.default_rng(29)
    pics = _pictures(bit_depth, 2, 37)
    org = _pictures(bit_depth, 1, 41)[0][0]
    pads = [(_pad(y, M), None, None) for (y, cb, cr) in pics]
    jobs, _ = _jobs(rng, 200, 2, bi_every=4)
    with hmgpu.Context(W, H, bit_depth, 2) as ctx:
        for s, (y, cb, cr) in enumerate(pics):
            ctx.ref_upload(s, y)
        ctx.org_upload(org)
        got_sad = ctx.pred_error(jobs, hmgpu.DF_SAD)
        got_had = ctx.pred_error(jobs, hmgpu.DF_HADS)
        with pytest.raises(hmgpu.HmGpuError, match="without chroma"):
            ctx.predict(jobs[:1], 96 * 64, with_chroma=True)
    for i, j in enumerate(jobs):
        w, h = int(j["pu_w"]), int(j["pu_h"])
        pred = np.ascontiguousarray(_oracle_component(O, pads, j, 0, bit_depth))
        blk = np.ascontiguousarray(org[int(j["pu_y"]):int(j["pu_y"]) + h, int(j["pu_x"]):int(j["pu_x"]) + w])
        assert int(got_sad[i]) == O.hmo_sad(B.ptr(blk), w, B.ptr(pred), w, w, h, 0, bit_depth, 0), (i, j)
        assert int(got_had[i]) == O.hmo_hads(B.ptr(blk), w, B.ptr(pred), w, w, h, bit_depth), (i, j)


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_merge_skip_dist_matches_oracle(bit_depth):
    """f2: every merge candidate of a CU -- motion compensation of Y, Cb, Cr (uni / bi) and the SSE of the skip
    reconstruction per component (xCheckRDCostMerge2Nx2N + the bSkipRes branch of encodeResAndCalcRdInterCU) -- in one call."""
    O = B.oracle()
    rng = np.random.default_rng(53)
    pics = _pictures(bit_depth, 2, 59)
    src = _pictures(bit_depth, 1, 61)[0]
    pads = [(_pad(y, M), _pad(cb, CM), _pad(cr, CM)) for (y, cb, cr) in pics]
    # CUs of 8..64 samples, five merge candidates each (same CU, different motion)
    cus = [(int(rng.integers(0, (W - s) // 8 + 1)) * 8, int(rng.integers(0, (H - s) // 8 + 1)) * 8, s) for s in (8, 16, 32, 64, 16, 8, 32) for _ in range(3)]
    jobs = np.zeros(len(cus) * 5, hmgpu.PRED_JOB)
    org_off = np.zeros(len(jobs), np.uint32)
    org_blocks, off_o, off_p = [], 0, 0
    for c, (x, y, s) in enumerate(cus):
        blk = [src[0][y:y + s, x:x + s], src[1][y // 2:(y + s) // 2, x // 2:(x + s) // 2], src[2][y // 2:(y + s) // 2, x // 2:(x + s) // 2]]
        org_blocks += [np.ascontiguousarray(b).ravel() for b in blk]
        for k in range(5):
            j = jobs[c * 5 + k]
            j["pu_x"], j["pu_y"], j["pu_w"], j["pu_h"] = x, y, s, s
            j["ref_slot"] = (-1, -1)
            for l in ([0, 1] if k % 2 == 0 else [k % 2 - 1 + 1 - (k // 3)]):
                l = int(l) & 1
                j["ref_slot"][l] = int(rng.integers(0, 2))
                j["mv_x"][l] = int(rng.integers(max(-200, (-70 - x) * 4), min(200, (W + 6 - x - s) * 4)))
                j["mv_y"][l] = int(rng.integers(max(-200, (-70 - y) * 4), min(200, (H + 6 - y - s) * 4)))
            j["dst_offset"] = off_p
            org_off[c * 5 + k] = off_o
            off_p += s * s * 3 // 2
        off_o += s * s * 3 // 2
    org_blocks = np.concatenate(org_blocks).astype(np.int16)
    with hmgpu.Context(W, H, bit_depth, 2) as ctx:
        for s_, (y_, cb, cr) in enumerate(pics):
            ctx.ref_upload(s_, y_, cb, cr)
        pred, sse = ctx.merge_skip_dist(jobs, org_off, org_blocks, off_p)
        with pytest.raises(hmgpu.HmGpuError):
            ctx.merge_skip_dist(jobs[:1], np.array([org_blocks.size], np.uint32), org_blocks, off_p)
    for i, j in enumerate(jobs):
        s = int(j["pu_w"])
        po, oo = int(j["dst_offset"]), int(org_off[i])
        for comp in range(3):
            cs = s if comp == 0 else s // 2
            exp = np.ascontiguousarray(_oracle_component(O, pads, j, comp, bit_depth))
            assert np.array_equal(pred[po:po + cs * cs], exp.ravel()), (i, comp)
            o = np.ascontiguousarray(org_blocks[oo:oo + cs * cs])
            assert int(sse[i, comp]) == O.hmo_sse(B.ptr(o), cs, B.ptr(exp), cs, cs, cs, bit_depth), (i, comp, j)
            po += cs * cs; oo += cs * cs
