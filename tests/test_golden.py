"""The CPU oracle (oracle/hm_oracle.c) against the committed golden vectors, which are outputs
of the unmodified reference (tests/golden/make_golden.py).  No GPU, no /root/reference."""
import os

import numpy as np
import pytest

import hmgpu
import synth
from oracle import binding as B
from util import padded_ref

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hm162_golden.npz"))


@pytest.mark.parametrize("bd", [8, 10])
def test_distortion_family(bd):
    O = B.oracle()
    org, cur = G["dist_org_%d" % bd], G["dist_cur_%d" % bd]
    for (w, h, oo, co, kind, ss, exp) in G["dist_rows_%d" % bd]:
        w, h, oo, co, ss = int(w), int(h), int(oo), int(co), int(ss)
        po, pc = B.ptr(org, oo), B.ptr(cur, co)
        if kind == 0:
            got = O.hmo_sad(po, 80, pc, 96, w, h, ss, bd, 0)
        elif kind == 1:
            got = O.hmo_hads(po, 80, pc, 96, w, h, bd)
        elif kind == 2:
            got = O.hmo_sse(po, 80, pc, 96, w, h, bd)
        else:  # generic setDistParam: widths 12/24/48 fall to xGetSAD, which ignores iSubShift
            got = O.hmo_sad(po, 80, pc, 96, w, h, ss, bd, 1 if w in (12, 24, 48) else 0)
        assert got == exp, (w, h, kind, ss)


def test_mv_cost_and_lambda():
    O = B.oracle()
    for (uc, px, py, sc, x, y, bits, cost) in G["cost_rows"]:
        a = [int(v) for v in (px, py, sc, x, y)]
        assert O.hmo_mv_bits(*a) == bits
        assert O.hmo_mv_cost(int(uc), *a) == cost
        assert hmgpu.lib().hmgpu_mv_bits(*a) == bits          # host-side helpers of the product
        assert hmgpu.lib().hmgpu_mv_cost(int(uc), *a) == cost
    for lam, ui in zip(G["cost_lambda"], G["cost_lambda_ui"]):
        assert O.hmo_lambda_to_cost(float(lam)) == ui


def test_clip_and_search_range():
    O = B.oracle()
    for r in G["clip_rows"]:
        pw, ph, cx, cy, px, py, sr = [int(v) for v in r[:7]]
        out = np.zeros(4, np.int32)
        O.hmo_set_search_range(pw, ph, cx, cy, px, py, sr, out)
        assert out.tolist() == r[7:11].tolist()
        mv = np.array([px, py], np.int32)
        O.hmo_clip_mv(pw, ph, cx, cy, mv)
        assert mv.tolist() == r[11:13].tolist()
        bd = hmgpu.clip_bounds(pw, ph, cx, cy)                # product host helpers
        assert hmgpu.search_range(bd, px, py, sr).tolist() == r[7:11].tolist()


@pytest.mark.parametrize("bd", [8, 10])
def test_interpolation_filters(bd):
    O = B.oracle()
    src, mid = G["filt_src_%d" % bd], G["filt_mid_%d" % bd]
    for chroma in (0, 1):
        for frac in range(8 if chroma else 4):
            for last in (0, 1):
                d = np.zeros((17, 33), np.int16)
                O.hmo_filter_hor(chroma, B.ptr(src, 8 * 64 + 8), 64, B.ptr(d), 33, 33, 17, frac, last, bd)
                assert np.array_equal(d, G["filt_%d_h_%d_%d_%d" % (bd, chroma, frac, last)])
                for first in (0, 1):
                    s = src if first else mid
                    d = np.zeros((17, 33), np.int16)
                    O.hmo_filter_ver(chroma, B.ptr(s, 8 * 64 + 8), 64, B.ptr(d), 33, 33, 17, frac, first, last, bd)
                    assert np.array_equal(d, G["filt_%d_v_%d_%d_%d_%d" % (bd, chroma, frac, first, last)])


@pytest.mark.parametrize("bd", [8, 10])
def test_forward_transform(bd):
    O = B.oracle()
    for n in (4, 8, 16, 32):
        blk = G["tr_%d_blk_%d" % (bd, n)]
        for dst in ((0, 1) if n == 4 else (0,)):
            exp = G["tr_%d_coef_%d_%d" % (bd, n, dst)]
            for t in range(len(blk)):
                c = np.zeros((n, n), np.int32)
                O.hmo_fwd_transform(bd, np.ascontiguousarray(blk[t]), c, n, n, dst)
                assert np.array_equal(c, exp[t]), (n, dst, t)


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("mode", ["tz", "fs"])
def test_searches(bd, mode):
    """xTZSearch / xPatternSearch + xPatternSearchFracDIF answers of the reference"""
    jobs = np.ascontiguousarray(G["search_%s_jobs_%d" % (mode, bd)]).view(hmgpu.ME_JOB).reshape(-1)
    exp = np.ascontiguousarray(G["search_%s_res_%d" % (mode, bd)]).view(hmgpu.ME_RESULT).reshape(-1)
    fr = synth.luma_frames(416, 240, 4, bd).astype(np.int16)
    pads = [padded_ref(fr[k]) for k in range(3)]
    got, _ = B.me_batch(B.oracle().hmo_me_batch, jobs, pads, fr[3], bd)
    got = got.view(hmgpu.ME_RESULT).reshape(-1)
    for f in hmgpu.ME_RESULT.names:
        if f == "n_cand":      # the reference does not count candidates
            continue
        assert np.array_equal(got[f], exp[f]), f
    assert (got["n_cand"] >= 18).all()


@pytest.mark.parametrize("bd", [8, 10])
def test_selective_search(bd):
    """xTZSearchSelective (FastSearch=2) answers of the reference: integer MV and SAD"""
    from util import oracle_me
    jobs = np.ascontiguousarray(G["sel_jobs_%d" % bd]).view(hmgpu.ME_JOB).reshape(-1).copy()
    jobs["flags"] &= np.uint8(~hmgpu.F_FRAC & 0xff)          # the fixture holds the integer search only
    side, rows = G["sel_side_%d" % bd], G["sel_rows_%d" % bd]
    fr = synth.luma_frames(416, 240, 4, bd).astype(np.int16)
    pads = [padded_ref(fr[k]) for k in range(3)]
    got = oracle_me(jobs, pads, fr[3], bd, side)
    assert np.array_equal(got["int_x"], rows[:, 0]) and np.array_equal(got["int_y"], rows[:, 1])
    assert np.array_equal(got["int_sad"].astype(np.int64), rows[:, 2])


def _deblock_pictures():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deblock_golden.npz"))
    for i in range(int(z["n_pictures"][0])):
        yield i, {k[len("p%d_" % i):]: z[k] for k in z.files if k.startswith("p%d_" % i)}


def test_deblock_oracle_matches_reference_decoder_pictures():
    """f3, second half: hmo_deblock_picture on the pictures the instrumented reference decoder dumped (before / after
    TComLoopFilter::loopFilterPic with the boundary strengths, QPs and no-filter flags it used; tests/golden/make_deblock_golden.py):
    intra, P and B pictures, 8 and 10 bit, deblocking and chroma QP offsets, picture sizes that are not multiples of the CTU."""
    O = B.oracle()
    n = 0
    for i, p in _deblock_pictures():
        w, h, bdl, bdc, beta, tc, cbo, cro, poc = [int(v) for v in p["params"]]
        y, cb, cr = p["pre_y"].copy(), p["pre_cb"].copy(), p["pre_cr"].copy()
        O.hmo_deblock_picture(y.ctypes.data, cb.ctypes.data, cr.ctypes.data, w, h, bdl, bdc,
                              np.ascontiguousarray(p["bs_ver"]).ctypes.data, np.ascontiguousarray(p["bs_hor"]).ctypes.data,
                              np.ascontiguousarray(p["qp"]).ctypes.data, np.ascontiguousarray(p["nofilter"]).ctypes.data, beta, tc, cbo, cro)
        assert np.array_equal(y, p["post_y"]) and np.array_equal(cb, p["post_cb"]) and np.array_equal(cr, p["post_cr"]), (i, poc)
        assert (p["pre_y"] != p["post_y"]).any(), "picture %d: the reference filtered nothing" % i
        n += 1
    assert n >= 9


def rdoq_golden_calls():
    """tests/golden/rdoq_golden.npz -> the dumped calls of the reference's xRateDistOptQuant (tests/golden/make_rdoq_golden.py)"""
    import rdoqdump
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rdoq_golden.npz"))
    calls = []
    for i in range(len(z["abs_sum"])):
        c = {k: int(v) for k, v in zip(rdoqdump.HDR, z["hdr"][i])}
        c["err_scale"], c["lambda"] = float(z["scale_lambda"][i, 0]), float(z["scale_lambda"][i, 1])
        c["bits"] = z["bits"][z["bits_index"][i]]
        a, b = int(z["offset"][i]), int(z["offset"][i + 1])
        c["coef"], c["level"], c["abs_sum"] = z["coef"][a:b], z["level"][a:b], int(z["abs_sum"][i])
        calls.append(c)
    return calls


def test_rdoq_oracle_matches_reference_encoder_calls():
    """f1: hmo_rdoq on the calls of TComTrQuant::xRateDistOptQuant the instrumented reference encoder dumped: 4x4 .. 32x32 TUs, luma
    and chroma, the three scans, inter (root cbf) and intra TUs, transform skip, sign-bit hiding on and off, 8 and 10 bit, P / B /
    intra slices -- levels and uiAbsSum identical for every call (the decisions compare sums of doubles: this also pins the order
    of the floating-point operations)."""
    import rdoqdump
    calls = rdoq_golden_calls()
    assert len(calls) >= 1000
    seen = set()
    hidden = 0
    for i, c in enumerate(calls):
        assert rdoqdump.supported(c)
        tu, bits = rdoqdump.to_tu_and_bits(c, B.RDOQ_TU, B.RDOQ_BITS)
        level, abs_sum = B.rdoq(tu, bits, c["coef"])
        assert abs_sum == c["abs_sum"] and np.array_equal(level, c["level"]), (i, {k: c[k] for k in rdoqdump.HDR})
        seen.add((c["log2"], c["channel"], c["scan"], c["sign_hide"], c["bit_depth"], c["root_cbf"], c["tskip"]))
        hidden += int(c["sign_hide"] and np.abs(c["level"]).sum() != c["abs_sum"])
    assert {s[0] for s in seen} == {2, 3, 4, 5} and {s[1] for s in seen} == {0, 1} and {s[2] for s in seen} == {0, 1, 2}
    assert {s[3] for s in seen} == {0, 1} and {s[4] for s in seen} == {8, 10} and {s[5] for s in seen} == {0, 1} and {s[6] for s in seen} == {0, 1}
    assert hidden > 20          # calls in which the sign-hiding pass changed a level


def test_scan_orders():
    """hmo_scan_order: every scan is a permutation, groups of 16 stay inside one 4x4 block, and the three 4x4 orders are the
    textbook ones (TComRom.cpp:53-137)"""
    O = B.oracle()
    for log2 in (2, 3, 4, 5):
        n = 1 << log2
        for st in (0, 1, 2):
            scan = np.zeros(n * n, np.uint16)
            cg = np.zeros(max(1, n * n // 16), np.uint16)
            O.hmo_scan_order(log2, st, scan.ctypes.data, cg.ctypes.data)
            assert sorted(scan.tolist()) == list(range(n * n)) and sorted(cg.tolist()) == list(range(n * n // 16))
            for g in range(n * n // 16):
                blk = scan[16 * g:16 * g + 16]
                assert len({(int(p) // n // 4, int(p) % n // 4) for p in blk}) == 1
                assert (int(blk[0]) // n // 4) * (n // 4) + int(blk[0]) % n // 4 == int(cg[g])
    scan = np.zeros(16, np.uint16)
    cg = np.zeros(1, np.uint16)
    O.hmo_scan_order(2, 0, scan.ctypes.data, cg.ctypes.data)
    assert scan.tolist() == [0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15]
    O.hmo_scan_order(2, 1, scan.ctypes.data, cg.ctypes.data)
    assert scan.tolist() == list(range(16))
    O.hmo_scan_order(2, 2, scan.ctypes.data, cg.ctypes.data)
    assert scan.tolist() == [0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15]


def dequant_golden_calls():
    """tests/golden/dequant_golden.npz -> the dumped calls of the reference's xDeQuant (tests/golden/make_dequant_golden.py)"""
    import rdoqdump
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dequant_golden.npz"))
    calls = []
    for i in range(len(z["hdr"])):
        c = {k: int(v) for k, v in zip(rdoqdump.DEQ_HDR, z["hdr"][i])}
        a, b = int(z["offset"][i]), int(z["offset"][i + 1])
        c["level"], c["coef"] = z["level"][a:b], z["coef"][a:b]
        calls.append(c)
    return calls


def test_dequant_oracle_matches_reference_encoder_calls():
    """hmo_dequant against what the instrumented reference encoder's xDeQuant produced: every TU size, 8 and 10 bit, QPs on both
    sides of the point where the right shift turns into a left shift"""
    from oracle import binding as B
    calls = dequant_golden_calls()
    shifts = set()
    for i, c in enumerate(calls):
        got = B.dequant(c["level"], c["log2"], c["per"], c["rem"], c["bit_depth"])
        assert np.array_equal(got, c["coef"]), (i, {k: c[k] for k in ("w", "per", "rem", "bit_depth")})
        shifts.add(6 - (15 - c["bit_depth"] - c["log2"] + c["per"]))
    assert len(calls) >= 300 and min(shifts) < 0 < max(shifts)
