"""Pin the oracle against the real reference (oracle/_ref/libhmref.so, compiled from the
unmodified HM sources) on fresh random inputs.  Skipped where the reference build is absent."""
import numpy as np
import pytest

import hmgpu
import synth
from oracle import binding as B
from util import padded_ref

pytestmark = pytest.mark.skipif(not B.have_ref(), reason="oracle/_ref/libhmref.so not built (needs /root/reference)")


@pytest.mark.parametrize("bd", [8, 10])
def test_distortions_random(bd):
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(1)
    mx = (1 << bd) - 1
    for it in range(120):
        w = int(rng.choice([4, 8, 12, 16, 24, 32, 48, 64]))
        h = int(rng.choice([4, 8, 12, 16, 24, 32, 48, 64]))
        o = (rng.integers(-mx, 2 * mx + 1, (64, 80)) if it % 3 == 0 else rng.integers(0, mx + 1, (64, 80))).astype(np.int16)
        c = rng.integers(0, mx + 1, (64, 96)).astype(np.int16)
        po, pc = B.ptr(o), B.ptr(c)
        for ss in (0, 1):
            assert R.ref_sad_me(po, 80, pc, 96, w, h, ss, bd) == O.hmo_sad(po, 80, pc, 96, w, h, ss, bd, 0)
        assert R.ref_dist_subpel(po, 80, pc, 96, w, h, 1, bd) == O.hmo_hads(po, 80, pc, 96, w, h, bd)
        assert R.ref_calc_had(po, 80, pc, 96, w, h, bd) == O.hmo_hads(po, 80, pc, 96, w, h, bd)
        assert R.ref_dist_subpel(po, 80, pc, 96, w, h, 0, bd) == O.hmo_sad(po, 80, pc, 96, w, h, 0, bd, 0)
        assert R.ref_sse(pc, 96, po, 80, w, h, bd) == O.hmo_sse(po, 80, pc, 96, w, h, bd)


def test_border_and_rdcost():
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (24, 40)).astype(np.int16)
    d1 = np.zeros((24 + 160, 40 + 160), np.int16)
    d2 = np.zeros_like(d1)
    assert R.ref_extend_border(src, 40, 24, d1, d1.size) == 80
    O.hmo_extend_border(src, 40, 24, 80, d2)
    assert np.array_equal(d1, d2)
    for lam in rng.uniform(0.1, 5000, 100):
        b, d = int(rng.integers(0, 60)), int(rng.integers(0, 100000))
        assert R.ref_calc_rd_cost_sad(float(lam), b, d) == O.hmo_calc_rd_cost_sad(float(lam), b, d)


@pytest.mark.parametrize("bd", [8, 10])
def test_motion_compensation(bd):
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(3)
    mx = (1 << bd) - 1
    for it in range(150):
        chroma = it % 2
        w = int(rng.choice([2, 4, 8, 16, 32])) if chroma else int(rng.choice([4, 8, 12, 16, 64]))
        h = int(rng.choice([2, 4, 8, 16])) if chroma else int(rng.choice([4, 8, 16, 64]))
        plane = rng.integers(0, mx + 1, (120, 140)).astype(np.int16)
        mvx, mvy = [int(v) for v in rng.integers(-60, 60, 2)]
        for bi in (0, 1):
            d1 = np.zeros((h, w), np.int16)
            d2 = np.zeros_like(d1)
            R.ref_pred_inter_blk(chroma, B.ptr(plane, 30 * 140 + 30), 140, mvx, mvy, w, h, bi, bd, B.ptr(d1), w)
            O.hmo_pred_inter_blk(chroma, B.ptr(plane, 30 * 140 + 30), 140, mvx, mvy, w, h, bi, bd, B.ptr(d2), w)
            assert np.array_equal(d1, d2)


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("noise", [False, True])
def test_search_batches(bd, noise):
    """hmo_me_batch vs ref_me_batch over HM-shaped job lists (TZ and full search + fractional)"""
    import worklist
    w, h = 416, 240
    if noise:
        fr = np.random.default_rng(9).integers(0, 1 << bd, (4, h, w)).astype(np.int16)
    else:
        fr = synth.luma_frames(w, h, 4, bd).astype(np.int16)
    pads = [padded_ref(fr[k]) for k in range(3)]
    tz = worklist.frame_jobs(w, h, n_refs=3, seed=5, ref_dist=[3, 2, 1])
    tz = tz[:: max(1, len(tz) // 1500)]
    fs = worklist.frame_jobs(w, h, n_refs=3, seed=6, full_search=True, search_range=12, ref_dist=[3, 2, 1])
    fs = fs[:: max(1, len(fs) // 150)]
    for jobs in (tz, fs):
        a, _ = B.me_batch(B.oracle().hmo_me_batch, jobs, pads, fr[3], bd)
        b, _ = B.me_batch(B.ref().ref_me_batch, jobs, pads, fr[3], bd)
        a = a.view(hmgpu.ME_RESULT).reshape(-1)
        b = b.view(hmgpu.ME_RESULT).reshape(-1)
        for f in hmgpu.ME_RESULT.names:
            if f != "n_cand":
                assert np.array_equal(a[f], b[f]), f


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("noise", [False, True])
def test_selective_search(bd, noise):
    """xTZSearchSelective (FastSearch=2): oracle vs the compiled reference, job by job"""
    import ctypes as C
    import worklist
    from util import M
    O, R = B.oracle(), B.ref()
    w, h = 416, 240
    rng = np.random.default_rng(31)
    if noise:
        fr = rng.integers(0, 1 << bd, (2, h, w)).astype(np.int16)
    else:
        fr = synth.luma_frames(w, h, 2, bd).astype(np.int16)
    pad = padded_ref(fr[0])
    pw = pad.shape[1]
    jobs = worklist.frame_jobs(w, h, n_refs=1, seed=8, search_range=int(16 if noise else 64))
    jobs = jobs[:: max(1, len(jobs) // (40 if noise else 160))]
    for k, j in enumerate(jobs):
        pu_w, pu_h, x0, y0 = int(j["pu_w"]), int(j["pu_h"]), int(j["pu_x"]), int(j["pu_y"])
        blk = np.ascontiguousarray(fr[1][y0:y0 + pu_h, x0:x0 + pu_w])
        cu_x = -(int(j["clip_hmin"]) // 4) - 71
        cu_y = -(int(j["clip_vmin"]) // 4) - 71
        sel = rng.integers(-60, 61, 6).astype(np.int32)
        if k % 3 == 0:
            sel[:2] = (int(j["pred_x"]) + 8, int(j["pred_y"]) - 4)
        has2n = int(bool(int(j["flags"]) & hmgpu.F_HAS_2NX2N))
        s = B.SearchT()
        s.org = B.ptr(blk); s.org_stride = pu_w; s.w = pu_w; s.h = pu_h
        s.ref = B.ptr(pad, (y0 + M) * pw + x0 + M); s.ref_stride = pw
        s.l, s.t, s.r, s.b = int(j["win_l"]), int(j["win_t"]), int(j["win_r"]), int(j["win_b"])
        s.ui_cost = int(j["ui_cost"]); s.pred_x = int(j["pred_x"]); s.pred_y = int(j["pred_y"])
        s.bit_depth = bd; s.cu_x = cu_x; s.cu_y = cu_y; s.pic_w = w; s.pic_h = h
        s.search_range = int(j["search_range"])
        s.start_x, s.start_y = int(j["start_x"]), int(j["start_y"])
        s.has_2nx2n = has2n; s.i2n_x = int(j["i2n_x"]); s.i2n_y = int(j["i2n_y"])
        for i in range(3):
            s.sel_pred[i][0] = int(sel[2 * i]); s.sel_pred[i][1] = int(sel[2 * i + 1])
        O.hmo_tz_selective(C.byref(s))
        mv = np.array([s.start_x, s.start_y], np.int32)
        sad = np.zeros(1, np.uint32)
        R.ref_tz_selective(B.ptr(blk), pu_w, pu_w, pu_h, B.ptr(pad, (y0 + M) * pw + x0 + M), pw, s.l, s.t, s.r, s.b,
                           s.ui_cost, s.pred_x, s.pred_y, bd, w, h, cu_x, cu_y, s.search_range, has2n, s.i2n_x, s.i2n_y,
                           sel, mv, sad)
        assert (s.mv_x, s.mv_y, s.sad) == (int(mv[0]), int(mv[1]), int(sad[0])), (k, j)


def _intra_line(rng, n, bd, kind):
    """4n+1 reference samples, bottom-left -> top-left -> above-right"""
    mx = (1 << bd) - 1
    if kind == 0:
        return rng.integers(0, mx + 1, 4 * n + 1).astype(np.int16)
    if kind == 1:   # smooth ramp with noise: the interpolating modes see small differences
        return np.clip(np.linspace(mx // 4, 3 * mx // 4, 4 * n + 1) + rng.integers(-3, 4, 4 * n + 1), 0, mx).astype(np.int16)
    return rng.choice(np.array([0, mx], np.int16), 4 * n + 1)    # extremes: the edge filters clip


@pytest.mark.parametrize("bd", [8, 10])
def test_intra_prediction_all_modes(bd):
    """f4: hmo_intra_pred against the reference's xPredIntraPlanar / xPredIntraAng / xDCPredFiltering for every mode, size,
    availability and edge-filter combination; the filtered / unfiltered rule against filteringIntraReferenceSamples."""
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(40 + bd)
    for n in (4, 8, 16, 32, 64):
        for mode in range(35):
            for ns in (0, 1):
                assert O.hmo_intra_use_filtered(mode, n, ns) == R.ref_intra_use_filtered(mode, n, ns), (mode, n, ns)
        for kind in range(3):
            line = _intra_line(rng, n, bd, kind)
            for mode in range(35):
                for above, left, edge in ((1, 1, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (0, 0, 0)):
                    a = np.zeros((n, n), np.int16)
                    b = np.full((n, n), -1, np.int16)
                    R.ref_intra_pred(bd, B.ptr(line), n, mode, above, left, edge, B.ptr(a))
                    O.hmo_intra_pred(B.ptr(line), n, mode, bd, above, left, edge, B.ptr(b))
                    assert np.array_equal(a, b), (n, mode, above, left, edge, kind)


@pytest.mark.parametrize("bd", [8, 10])
def test_sao_block_statistics(bd):
    """f3 (first half): hmo_sao_blk_stats against the reference's TEncSampleAdaptiveOffset::getBlkStats for every combination
    of neighbour availability, the skip-line sets HM uses, whole and partial CTUs, luma and chroma block sizes."""
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(90 + bd)
    mx = (1 << bd) - 1
    skips = [([5, 5, 5, 5, 5], [4, 4, 4, 4, 4]), ([3, 3, 3, 3, 3], [2, 2, 2, 2, 2]), ([5, 5, 5, 5, 4], [3, 3, 3, 3, 4]), ([0, 0, 0, 0, 0], [0, 0, 0, 0, 0])]
    for it in range(160):
        w, h = [(64, 64), (32, 32), (64, 56), (32, 28), (16, 16), (48, 64), (8, 8)][it % 7]
        pic = rng.integers(0, mx + 1, (h + 8, w + 8)).astype(np.int16)
        if it % 3 == 0:     # smooth content: many flat / monotone neighbourhoods (edge classes 0 ties)
            pic = (pic // 64 * 64 + rng.integers(0, 3, pic.shape)).astype(np.int16)
        org = np.clip(pic + rng.integers(-6, 7, pic.shape), 0, mx).astype(np.int16)
        flags = int(rng.integers(0, 64)) if it >= 8 else [0, 63, 1, 2, 4, 8, 16, 32][it]
        sr, sb = skips[it % 4]
        sr, sb = np.array(sr, np.int32), np.array(sb, np.int32)
        st = w + 8
        d0, c0 = np.zeros((5, 32), np.int64), np.zeros((5, 32), np.int64)
        d1, c1 = np.zeros((5, 32), np.int64), np.zeros((5, 32), np.int64)
        R.ref_sao_blk_stats(B.ptr(pic, 4 * st + 4), st, B.ptr(org, 4 * st + 4), st, w, h, flags, sr, sb, bd, d0, c0)
        O.hmo_sao_blk_stats(B.ptr(pic, 4 * st + 4), st, B.ptr(org, 4 * st + 4), st, w, h, flags, sr, sb, bd, d1, c1)
        assert np.array_equal(c0, c1), (it, w, h, flags, np.argwhere(c0 != c1)[:4])
        assert np.array_equal(d0, d1), (it, w, h, flags)


@pytest.mark.parametrize("bd", [8, 10])
def test_sao_offset_block(bd):
    """f3: hmo_sao_offset_block against the reference's TComSampleAdaptiveOffset::offsetBlock, all five types, every availability
    combination (first / last line rules of the diagonal types), offsets that clip at both ends of the sample range."""
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(120 + bd)
    mx = (1 << bd) - 1
    for it in range(300):
        w, h = [(64, 64), (32, 32), (64, 56), (32, 28), (16, 16), (8, 8)][it % 6]
        pic = rng.integers(0, mx + 1, (h + 8, w + 8)).astype(np.int16)
        if it % 3 == 0:
            pic = (pic // 64 * 64 + rng.integers(0, 3, pic.shape)).astype(np.int16)
        if it % 5 == 0:
            pic = rng.choice(np.array([0, 1, mx - 1, mx], np.int16), pic.shape)
        typ = it % 5
        off = np.zeros(32, np.int32)
        if typ < 4:
            off[:5] = [int(rng.integers(0, 8)), int(rng.integers(0, 8)), 0, -int(rng.integers(0, 8)), -int(rng.integers(0, 8))]
        else:
            b0 = int(rng.integers(0, 29))
            off[b0:b0 + 4] = rng.integers(-7, 8, 4)
        flags = int(rng.integers(0, 256)) if it >= 10 else [0, 255, 1, 2, 4, 8, 16, 32, 64, 128][it]
        st = w + 8
        a = np.full_like(pic, -5)
        b = np.full_like(pic, -5)
        R.ref_sao_offset_block(typ, off, B.ptr(pic, 4 * st + 4), st, B.ptr(a, 4 * st + 4), st, w, h, flags, bd)
        O.hmo_sao_offset_block(typ, off, B.ptr(pic, 4 * st + 4), st, B.ptr(b, 4 * st + 4), st, w, h, flags, bd)
        assert np.array_equal(a, b), (it, typ, w, h, flags, np.argwhere(a != b)[:4])


def test_deblock_against_instrumented_decoder_live():
    """f3, second half: a fresh encode (content and QP not among the golden pictures) decoded by the instrumented reference decoder
    (oracle/_ref/TAppDecoderDbk, oracle/Makefile target `dbk`); hmo_deblock_picture reproduces every picture it deblocks."""
    import os
    import subprocess
    import tempfile
    import dbkdump
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    enc = os.path.join(root, "oracle", "_ref", "TAppEncoderRef")
    dec = os.path.join(root, "oracle", "_ref", "TAppDecoderDbk")
    cfg = os.path.join(root, "oracle", "_ref", "cfg", "encoder_lowdelay_main.cfg")
    if not (os.path.exists(enc) and os.path.exists(dec) and os.path.exists(cfg)):
        pytest.skip("instrumented decoder not built (make -C oracle dbk, needs /root/reference)")
    O = B.oracle()
    with tempfile.TemporaryDirectory(prefix="hmdbk_") as tmp:
        yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), 208, 120, 4, 8, seed=4242)
        bits, dump = os.path.join(tmp, "s.bin"), os.path.join(tmp, "s.dump")
        subprocess.run([enc, "-c", cfg, "-i", yuv, "-wdt", "208", "-hgt", "120", "-fr", "30", "-f", "4", "-q", "29", "-b", bits], check=True, capture_output=True)
        subprocess.run([dec, "-b", bits], check=True, capture_output=True, env=dict(os.environ, HM_DBK_DUMP=dump))
        pics = dbkdump.read(dump)
    assert len(pics) == 4
    for p in pics:
        y, cb, cr = [a.copy() for a in p["pre"]]
        O.hmo_deblock_picture(y.ctypes.data, cb.ctypes.data, cr.ctypes.data, p["w"], p["h"], p["bd_luma"], p["bd_chroma"], p["bs_ver"].ctypes.data,
                              p["bs_hor"].ctypes.data, p["qp"].ctypes.data, p["nofilter"].ctypes.data, p["beta_offset_div2"], p["tc_offset_div2"],
                              p["cb_qp_offset"], p["cr_qp_offset"])
        assert np.array_equal(y, p["post"][0]) and np.array_equal(cb, p["post"][1]) and np.array_equal(cr, p["post"][2]), p["poc"]


@pytest.mark.parametrize("bd", [8, 10])
def test_inverse_transform(bd):
    """hmo_inv_transform against the reference's xITrMxN (partialButterflyInverse4/8/16/32, fastInverseDst): coefficients as the forward
    transform + quantiser leave them, and extreme ones that hit both clips"""
    O, R = B.oracle(), B.ref()
    rng = np.random.default_rng(500 + bd)
    for n in (4, 8, 16, 32):
        for it in range(40):
            if it % 3 == 0:
                c = rng.integers(-32768, 32768, (n, n)).astype(np.int32)
            elif it % 3 == 1:
                c = np.zeros((n, n), np.int32); c[: max(1, n // 4), : max(1, n // 4)] = rng.integers(-2000, 2000, (max(1, n // 4), max(1, n // 4)))
            else:
                c = rng.choice(np.array([-32768, 32767, 0], np.int32), (n, n))
            for dst in ((0, 1) if n == 4 else (0,)):
                a = np.zeros((n, n), np.int32); b = np.zeros((n, n), np.int32)
                R.ref_inv_transform(bd, c.copy(), a, n, dst)
                O.hmo_inv_transform(bd, c, b, n, dst)
                assert np.array_equal(a, b), (n, it, dst)


def test_rdoq_against_instrumented_encoder_live():
    """f1: a fresh encode (content, size and QP not among the golden calls) by the instrumented reference encoder
    (oracle/_ref/TAppEncoderRdoq, oracle/Makefile target `rdoq`); hmo_rdoq reproduces every sampled call of xRateDistOptQuant and
    hmo_dequant every sampled call of xDeQuant."""
    import os
    import subprocess
    import tempfile
    import rdoqdump
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    enc = os.path.join(root, "oracle", "_ref", "TAppEncoderRdoq")
    cfg = os.path.join(root, "oracle", "_ref", "cfg", "encoder_lowdelay_main.cfg")
    if not (os.path.exists(enc) and os.path.exists(cfg)):
        pytest.skip("instrumented encoder not built (make -C oracle rdoq, needs /root/reference)")
    with tempfile.TemporaryDirectory(prefix="hmrdoq_") as tmp:
        yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), 208, 120, 3, 8, seed=991)
        dump = os.path.join(tmp, "s.dump")
        deq_dump = os.path.join(tmp, "d.dump")
        subprocess.run([enc, "-c", cfg, "-i", yuv, "-wdt", "208", "-hgt", "120", "-fr", "30", "-f", "3", "-q", "26", "-b", os.path.join(tmp, "s.bin"),
                        "-o", os.path.join(tmp, "rec.yuv")],
                       check=True, capture_output=True, env=dict(os.environ, HM_RDOQ_DUMP=dump, HM_RDOQ_EVERY="11", HM_DEQ_DUMP=deq_dump, HM_DEQ_EVERY="7"))
        calls = rdoqdump.read(dump)
        deq_calls = rdoqdump.read_dequant(deq_dump)
    assert len(calls) > 5000
    coded = 0
    for i, c in enumerate(calls):
        assert rdoqdump.supported(c)
        tu, bits = rdoqdump.to_tu_and_bits(c, B.RDOQ_TU, B.RDOQ_BITS)
        level, abs_sum = B.rdoq(tu, bits, c["coef"])
        assert abs_sum == c["abs_sum"] and np.array_equal(level, c["level"]), (i, {k: c[k] for k in rdoqdump.HDR})
        coded += int(abs_sum > 0)
    assert coded > 1000
    # the same encode's calls of xDeQuant (hook hm_deq_after) against hmo_dequant
    assert len(deq_calls) > 1000
    for i, c in enumerate(deq_calls):
        assert rdoqdump.dequant_supported(c)
        assert np.array_equal(B.dequant(c["level"], c["log2"], c["per"], c["rem"], c["bit_depth"]), c["coef"]), (i, {k: c[k] for k in rdoqdump.DEQ_HDR})
