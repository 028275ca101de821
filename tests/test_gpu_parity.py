"""GPU parity tests: libhmgpu (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bit-exact: all of this is integer arithmetic."""
import numpy as np
import pytest

import hmgpu
import synth
import worklist
from oracle import binding as B
from util import M, assert_results_equal, oracle_me, padded_ref

pytestmark = pytest.mark.gpu

W, H = 416, 240


def _frames(bit_depth, n=5, noise=False, seed=1234):
    if noise:
        return np.random.default_rng(seed).integers(0, 1 << bit_depth, (n, H, W)).astype(np.int16)
    return synth.luma_frames(W, H, n, bit_depth, seed).astype(np.int16)


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_phase_planes_match_oracle(bit_depth):
    fr = _frames(bit_depth, 1, noise=True)
    pad = padded_ref(fr[0])
    exp = np.zeros((16,) + pad.shape, np.int16)
    B.oracle().hmo_phase_planes(pad, pad.shape[1], pad.shape[0], bit_depth, exp)
    with hmgpu.Context(W, H, bit_depth, 2) as ctx:
        ctx.ref_upload(1, fr[0])
        for fy in range(4):
            for fx in range(4):
                got = ctx.ref_plane(1, fx, fy)
                assert np.array_equal(got, exp[fy * 4 + fx]), (fx, fy)


def _random_jobs(rng, n, bit_depth, mode, n_refs, org_blocks=None):
    """jobs over all PU shapes; mode in {'tz','fs','frac'}"""
    shapes = [(64, 64), (32, 32), (16, 16), (8, 8), (32, 64), (64, 32), (16, 32), (32, 16), (8, 16), (16, 8),
              (8, 4), (4, 8), (12, 16), (16, 12), (24, 32), (32, 24), (4, 16), (16, 4), (32, 8), (8, 32),
              (64, 16), (16, 64), (64, 48), (48, 64)]
    jobs = np.zeros(n, hmgpu.ME_JOB)
    blocks = []
    off = 0
    for i in range(n):
        w, h = shapes[i % len(shapes)]
        cus = 8 if max(w, h) <= 8 else 16 if max(w, h) <= 16 else 32 if max(w, h) <= 32 else 64
        cx = int(rng.integers(0, W // cus)) * cus
        cy = int(rng.integers(0, H // cus)) * cus
        px = cx + int(rng.integers(0, (cus - w) // 4 + 1)) * 4
        py = cy + int(rng.integers(0, (cus - h) // 4 + 1)) * 4
        j = jobs[i]
        j["pu_x"], j["pu_y"], j["pu_w"], j["pu_h"] = px, py, w, h
        j["ref_slot"] = int(rng.integers(0, n_refs))
        k = i % 11
        if k == 0:
            pred = (int(rng.integers(-3000, 3000)), int(rng.integers(-2000, 2000)))
        elif k < 4:
            pred = (int(rng.integers(-400, 400)), int(rng.integers(-400, 400)))
        else:
            pred = (int(rng.integers(-40, 40)), int(rng.integers(-40, 40)))
        j["pred_x"], j["pred_y"] = pred
        bd = hmgpu.clip_bounds(W, H, cx, cy)
        j["clip_hmin"], j["clip_hmax"], j["clip_vmin"], j["clip_vmax"] = bd
        sr = 64 if mode != "fs" else int(rng.choice([4, 8, 24, 64]))
        j["search_range"] = sr
        j["win_l"], j["win_t"], j["win_r"], j["win_b"] = hmgpu.search_range(bd, pred[0], pred[1], sr)
        j["ui_cost"] = worklist.lambda_to_cost(float(rng.uniform(4, 200)))
        fl = hmgpu.F_FRAC
        if i % 5:
            fl |= hmgpu.F_FEN
        if i % 6:
            fl |= hmgpu.F_HADME
        if i % 13 == 0:
            fl |= hmgpu.F_LOSSLESS
        if mode == "frac":
            ltrb = [int(v) for v in (j["win_l"], j["win_t"], j["win_r"], j["win_b"])]
            j["start_x"] = int(rng.integers(ltrb[0], ltrb[2] + 1))
            j["start_y"] = int(rng.integers(ltrb[1], ltrb[3] + 1))
        else:
            fl |= hmgpu.F_INTEGER
            j["start_x"], j["start_y"] = pred
            if mode == "fs":
                fl |= hmgpu.F_FULL
            elif i % 2:
                fl |= hmgpu.F_HAS_2NX2N
                j["i2n_x"], j["i2n_y"] = int(rng.integers(-30, 30)), int(rng.integers(-30, 30))
        if org_blocks is not None and i % 3 == 0:
            fl |= hmgpu.F_ORG_BLOCK
            j["org_offset"] = off
            blocks.append((i, w, h))
            off += w * h
        j["flags"] = fl
    return jobs, blocks, off


def _run(bit_depth, mode, n, noise=False, with_blocks=False, seed=5, chunk=None):
    rng = np.random.default_rng(seed)
    fr = _frames(bit_depth, 4, noise)
    org = fr[3]
    jobs, blocks, n_elems = _random_jobs(rng, n, bit_depth, mode, 3, org_blocks=True if with_blocks else None)
    org_blocks = None
    if with_blocks:
        # bi-pred style key pattern 2*org - pred: leaves the pixel range (TComYuv.cpp:415)
        org_blocks = np.zeros(max(1, n_elems), np.int16)
        mx = (1 << bit_depth) - 1
        for (i, w, h) in blocks:
            j = jobs[i]
            o = org[int(j["pu_y"]):int(j["pu_y"]) + h, int(j["pu_x"]):int(j["pu_x"]) + w].astype(np.int32)
            p = rng.integers(0, mx + 1, (h, w))
            org_blocks[int(j["org_offset"]):int(j["org_offset"]) + w * h] = (2 * o - p).astype(np.int16).ravel()
    pads = [padded_ref(fr[k]) for k in range(3)]
    exp = oracle_me(jobs, pads, org, bit_depth, org_blocks)
    with hmgpu.Context(W, H, bit_depth, 3) as ctx:
        for k in range(3):
            ctx.ref_upload(k, fr[k])
        ctx.org_upload(org)
        if chunk is None:
            got = ctx.me_search(jobs, org_blocks)
        else:
            # small calls take the fused low-latency kernel (me_single.cu): 1..chunk jobs per call
            parts, i, k = [], 0, 1
            while i < len(jobs):
                sub_jobs = jobs[i:i + k].copy()
                sub_blocks = None
                if org_blocks is not None:
                    # re-base the key-pattern blocks of this call
                    bl, off = [], 0
                    for j in sub_jobs:
                        if j["flags"] & hmgpu.F_ORG_BLOCK:
                            sz = int(j["pu_w"]) * int(j["pu_h"])
                            bl.append(org_blocks[int(j["org_offset"]):int(j["org_offset"]) + sz])
                            j["org_offset"] = off
                            off += sz
                    sub_blocks = np.concatenate(bl) if bl else None
                parts.append(ctx.me_search(sub_jobs, sub_blocks))
                i += k
                k = k % chunk + 1
            got = np.concatenate(parts)
    assert_results_equal(got, exp, jobs)
    return got


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("noise", [False, True])
def test_tz_search_and_frac(bit_depth, noise):
    _run(bit_depth, "tz", 600, noise=noise)


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_full_search_and_frac(bit_depth):
    _run(bit_depth, "fs", 96)


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_frac_only(bit_depth):
    _run(bit_depth, "frac", 400)


def test_full_search_window_by_tma():
    """xPatternSearch with the window staged by TMA (cp.async.bulk.tensor, me_full_impl.cuh) and by per-thread loads: same
    bytes, equal to the oracle -- windows clipped at every picture border, every PU shape, search ranges 4 .. 64."""
    rng = np.random.default_rng(31)
    fr = _frames(8, 3, noise=True)
    jobs, _, _ = _random_jobs(rng, 240, 8, "fs", 2)
    pads = [padded_ref(fr[k]) for k in range(2)]
    exp = oracle_me(jobs, pads, fr[2], 8)
    with hmgpu.Context(W, H, 8, 2) as ctx:
        for k in range(2):
            ctx.ref_upload(k, fr[k])
        ctx.org_upload(fr[2])
        got = ctx.me_search(jobs)
        assert_results_equal(got, exp, jobs)
        ctx.set_option("fs_tma", 0)
        assert ctx.me_search(jobs).tobytes() == got.tobytes()


@pytest.mark.parametrize("noise", [False, True])
def test_fractional_kernels_agree(noise):
    """The packed fractional kernels (me_frac2.cu) against the oracle and against the generic kernel (me_frac.cu): all PU shapes
    incl. 64x64, SAD and SATD, blocks hanging over every picture border, fractional-only jobs next to full searches."""
    rng = np.random.default_rng(123)
    fr = _frames(8, 4, noise)
    org = fr[3]
    jobs, _, _ = _random_jobs(rng, 3000, 8, "frac", 3)
    tz, _, _ = _random_jobs(rng, 600, 8, "tz", 3)
    jobs = np.concatenate([jobs, tz])
    # integer positions at the clip bounds: the candidate blocks reach the outermost samples the padding holds
    for i in range(0, 3000, 9):
        j = jobs[i]
        j["start_x"] = int([j["win_l"], j["win_r"]][(i // 9) & 1])
        j["start_y"] = int([j["win_t"], j["win_b"]][(i // 18) & 1])
    pads = [padded_ref(fr[k]) for k in range(3)]
    exp = oracle_me(jobs, pads, org, 8)
    with hmgpu.Context(W, H, 8, 3) as ctx:
        for k in range(3):
            ctx.ref_upload(k, fr[k])
        ctx.org_upload(org)
        got = ctx.me_search(jobs)
        assert_results_equal(got, exp, jobs)
        ctx.set_option("frac_overlap", 0)
        assert ctx.me_search(jobs).tobytes() == got.tobytes()
        ctx.set_option("frac_v1", 1)
        assert ctx.me_search(jobs).tobytes() == got.tobytes()
        # the CTU-group kernel with TMA-staged windows (me_fracw.cu; by default only for batches of >= 4096 jobs): random vectors
        # put most of these jobs on its slow path (footprint outside the group's window), the rest on the fast path
        ctx.set_option("frac_v1", 0)
        ctx.set_option("frac_win_min", 1)
        assert ctx.me_search(jobs).tobytes() == got.tobytes()


@pytest.mark.parametrize("spread", [0, 3, 40])
def test_fractional_window_kernel(spread):
    """me_fracw.cu on the jobs of whole pictures (every PU of every CTU x 3 references, fractional-only at integer vectors spread
    around a global motion): windows staged by TMA, groups larger than one pass (duplicated jobs), picture-border CTUs, SAD-mode
    and lossless jobs mixed in; against the oracle and against the per-tile kernels (me_frac2.cu)."""
    rng = np.random.default_rng(5 + spread)
    fr = _frames(8, 4, True)
    org = fr[3]
    jobs = worklist.frame_jobs(W, H, n_refs=3, seed=3 + spread)
    jobs["flags"] &= ~np.uint8(hmgpu.F_INTEGER)
    n = len(jobs)
    mvx = -3 * (jobs["ref_slot"].astype(np.int64) + 1) + rng.integers(-spread, spread + 1, n)
    mvy = 2 * (jobs["ref_slot"].astype(np.int64) + 1) + rng.integers(-spread, spread + 1, n)
    jobs["start_x"] = np.clip(mvx, jobs["win_l"], jobs["win_r"])
    jobs["start_y"] = np.clip(mvy, jobs["win_t"], jobs["win_b"])
    # SAD instead of SATD, lossless: the slow path of the kernel
    jobs["flags"][::17] &= ~np.uint8(hmgpu.F_HADME)
    jobs["flags"][5::29] |= np.uint8(hmgpu.F_LOSSLESS)
    # one CTU with more jobs than a pass holds
    first_ctu = jobs[(jobs["pu_x"] < 64) & (jobs["pu_y"] < 64) & (jobs["ref_slot"] == 0)]
    jobs = np.concatenate([jobs, first_ctu, first_ctu, first_ctu])
    assert len(jobs) >= 4096
    with hmgpu.Context(W, H, 8, 3) as ctx:
        for k in range(3):
            ctx.ref_upload(k, fr[k])
        ctx.org_upload(org)
        got = ctx.me_search(jobs)
        ctx.set_option("frac_win", 0)
        ref = ctx.me_search(jobs)
    bad = np.nonzero(got.view(np.uint8).reshape(len(jobs), -1) != ref.view(np.uint8).reshape(len(jobs), -1))[0]
    assert bad.size == 0, "first differing job %d: %s vs %s (%s)" % (bad[0], got[bad[0]], ref[bad[0]], jobs[bad[0]])
    sample = rng.choice(len(jobs), 1500, replace=False)
    pads = [padded_ref(fr[k]) for k in range(3)]
    assert_results_equal(got[sample], oracle_me(jobs[sample], pads, org, 8), jobs[sample])


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("mode", ["tz", "fs", "frac"])
def test_bipred_key_pattern_blocks(bit_depth, mode):
    _run(bit_depth, mode, 120 if mode != "fs" else 48, with_blocks=True)


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("mode", ["tz", "fs", "frac"])
@pytest.mark.parametrize("with_blocks", [False, True])
def test_low_latency_path(bit_depth, mode, with_blocks):
    _run(bit_depth, mode, 96 if mode != "fs" else 40, with_blocks=with_blocks, chunk=5, seed=21)


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_dist_batch_matches_oracle(bit_depth):
    rng = np.random.default_rng(11)
    O = B.oracle()
    mx = (1 << bit_depth) - 1
    org = rng.integers(-mx, 2 * mx + 1, 64 * 80).astype(np.int16)
    cur = rng.integers(0, mx + 1, 64 * 96).astype(np.int16)
    items, exp = [], []
    sizes = [4, 8, 12, 16, 24, 32, 48, 64]
    for w in sizes:
        for h in sizes:
            for func in (hmgpu.DF_SAD, hmgpu.DF_SAD_GENERIC, hmgpu.DF_HADS, hmgpu.DF_SSE):
                for ss in ((0, 1, 2) if func in (hmgpu.DF_SAD, hmgpu.DF_SAD_GENERIC) else (0,)):
                    oo, co = int(rng.integers(0, 16)), int(rng.integers(0, 32))
                    items.append((oo, co, 80, 96, w, h, func, ss))
                    po, pc = B.ptr(org, oo), B.ptr(cur, co)
                    if func == hmgpu.DF_SAD:
                        exp.append(O.hmo_sad(po, 80, pc, 96, w, h, ss, bit_depth, 0))
                    elif func == hmgpu.DF_SAD_GENERIC:
                        exp.append(O.hmo_sad(po, 80, pc, 96, w, h, ss, bit_depth, 1))
                    elif func == hmgpu.DF_HADS:
                        exp.append(O.hmo_hads(po, 80, pc, 96, w, h, bit_depth))
                    else:
                        exp.append(O.hmo_sse(po, 80, pc, 96, w, h, bit_depth))
    # 2x2-tiled SATD (chroma-sized blocks) and ragged sizes
    for (w, h) in [(2, 2), (6, 2), (2, 6), (6, 6), (10, 14)]:
        items.append((0, 0, 80, 96, w, h, hmgpu.DF_HADS, 0))
        exp.append(O.hmo_hads(B.ptr(org), 80, B.ptr(cur), 96, w, h, bit_depth))
    for (w, h) in [(1, 1), (3, 5), (7, 64), (63, 1)]:
        items.append((5, 9, 80, 96, w, h, hmgpu.DF_SAD_GENERIC, 0))
        exp.append(O.hmo_sad(B.ptr(org, 5), 80, B.ptr(cur, 9), 96, w, h, 0, bit_depth, 1))
        items.append((5, 9, 80, 96, w, h, hmgpu.DF_SSE, 0))
        exp.append(O.hmo_sse(B.ptr(org, 5), 80, B.ptr(cur, 9), 96, w, h, bit_depth))
    it = np.array(items, dtype=hmgpu.DIST_ITEM)
    with hmgpu.Context(W, H, bit_depth, 1) as ctx:
        got = ctx.dist_batch(org, cur, it)
        assert len(ctx.dist_batch(org, cur, it[:0])) == 0      # empty batch
    assert np.array_equal(got, np.array(exp, np.uint32))


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_fwd_transform_and_quant(bit_depth):
    rng = np.random.default_rng(13)
    O = B.oracle()
    with hmgpu.Context(W, H, bit_depth, 1) as ctx:
        for n in (4, 8, 16, 32):
            for use_dst in ((False, True) if n == 4 else (False,)):
                resi = rng.integers(-(1 << bit_depth) + 1, 1 << bit_depth, (37, n, n)).astype(np.int16)
                resi[0] = (1 << bit_depth) - 1      # extremes
                resi[1] = -(1 << bit_depth) + 1
                got = ctx.fwd_transform(resi, n, use_dst)
                exp = np.zeros_like(got)
                for t in range(len(resi)):
                    O.hmo_fwd_transform(bit_depth, np.ascontiguousarray(resi[t], np.int32), exp[t], n, n, int(use_dst))
                assert np.array_equal(got, exp), (n, use_dst)
                for (per, rem, intra) in [(5, 2, 0), (4, 3, 1), (6, 0, 0)]:
                    lv, du, sm = ctx.quant(got, n, per, rem, intra)
                    log2n = n.bit_length() - 1
                    tshift = 15 - bit_depth - log2n
                    for t in range(len(resi)):
                        el = np.zeros(n * n, np.int32)
                        ed = np.zeros(n * n, np.int32)
                        es = O.hmo_quant(np.ascontiguousarray(got[t].ravel()), n * n, per, rem, tshift, intra, el, ed)
                        assert np.array_equal(lv[t].ravel(), el) and np.array_equal(du[t].ravel(), ed) and sm[t] == es


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_inverse_transform(bit_depth):
    """hmgpu_inv_transform (xITrMxN) against the oracle: quantised-looking coefficients, full-range noise, the clip extremes; and the
    round trip forward -> inverse of a residual stays within the transform's rounding"""
    rng = np.random.default_rng(19)
    O = B.oracle()
    with hmgpu.Context(W, H, bit_depth, 1) as ctx:
        for n in (4, 8, 16, 32):
            for use_dst in ((False, True) if n == 4 else (False,)):
                c = rng.integers(-32768, 32768, (41, n, n)).astype(np.int32)
                c[0] = 32767; c[1] = -32768
                c[2:12] = 0
                c[2:12, : max(1, n // 4), : max(1, n // 4)] = rng.integers(-3000, 3000, (10, max(1, n // 4), max(1, n // 4)))
                got = ctx.inv_transform(c, n, use_dst)
                exp = np.zeros(c.shape, np.int32)
                for t in range(len(c)):
                    O.hmo_inv_transform(bit_depth, np.ascontiguousarray(c[t]), exp[t], n, int(use_dst))
                assert np.array_equal(got.astype(np.int32), exp), (n, use_dst)
                resi = rng.integers(-(1 << bit_depth) + 1, 1 << bit_depth, (8, n, n)).astype(np.int16)
                back = ctx.inv_transform(ctx.fwd_transform(resi, n, use_dst), n, use_dst)
                assert np.abs(back.astype(np.int32) - resi).max() <= 2 << (bit_depth - 6), (n, use_dst)    # the integer transform pair is not exactly orthogonal


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_mc_luma_matches_oracle(bit_depth):
    rng = np.random.default_rng(17)
    O = B.oracle()
    fr = _frames(bit_depth, 2, noise=True)
    pad = padded_ref(fr[0])
    pw = pad.shape[1]
    jobs = np.zeros(200, hmgpu.MC_JOB)
    off = 0
    exp = []
    for i in range(len(jobs)):
        w, h = [(8, 8), (16, 16), (64, 64), (4, 8), (12, 16), (32, 24)][i % 6]
        x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
        y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
        mvx = int(rng.integers(max(-280, (-70 - x) * 4), min(280, (W + 6 - x - w) * 4)))
        mvy = int(rng.integers(max(-280, (-70 - y) * 4), min(280, (H + 6 - y - h) * 4)))
        jobs[i] = (x, y, w, h, 0, 0, mvx, mvy, off)
        e = np.zeros((h, w), np.int16)
        O.hmo_pred_inter_blk(0, B.ptr(pad, (y + M) * pw + x + M), pw, mvx, mvy, w, h, 0, bit_depth, B.ptr(e), w)
        exp.append(e.ravel())
        off += w * h
    with hmgpu.Context(W, H, bit_depth, 1) as ctx:
        ctx.ref_upload(0, fr[0])
        got = ctx.mc_luma(jobs, off)
    assert np.array_equal(got, np.concatenate(exp))


def test_errors_are_reported_not_thrown():
    with hmgpu.Context(W, H, 8, 2) as ctx:
        jobs = worklist.frame_jobs(W, H, n_refs=1)[:4]
        with pytest.raises(hmgpu.HmGpuError, match="not uploaded"):
            ctx.me_search(jobs)
        ctx.ref_upload(0, np.zeros((H, W), np.int16))
        ctx.org_upload(np.zeros((H, W), np.int16))
        bad = jobs.copy()
        bad["pu_w"][0] = 5
        with pytest.raises(hmgpu.HmGpuError, match="unsupported"):
            ctx.me_search(bad)
        # shapes with more than 64 SATD tiles (not HEVC PU shapes) would overflow the (job, tile) packing of the
        # fractional stage: refused on the small-batch path and by the device-side scan of the pipelined path alike
        for w_, h_ in ((36, 32), (60, 64), (64, 60), (44, 28)):
            bad = jobs.copy()
            bad["pu_x"][1], bad["pu_y"][1], bad["pu_w"][1], bad["pu_h"][1] = 0, 0, w_, h_
            with pytest.raises(hmgpu.HmGpuError, match="job 1: PU %dx%d unsupported" % (w_, h_)):
                ctx.me_search(bad)
        big = np.tile(jobs, 20000)
        big["pu_x"][70001], big["pu_y"][70001], big["pu_w"][70001], big["pu_h"][70001] = 0, 0, 36, 32
        with pytest.raises(hmgpu.HmGpuError, match="job 70001: PU size unsupported"):
            ctx.me_search(big)
        # a negative element count of the key-pattern array is an argument error, not a huge memcpy
        blk = np.zeros(64, np.int16)
        rc = ctx.L.hmgpu_me_search(ctx.h, jobs.ctypes.data, len(jobs), blk.ctypes.data, -5, np.zeros(len(jobs), hmgpu.ME_RESULT).ctypes.data)
        assert rc == -1 and b"negative" in ctx.L.hmgpu_last_error(ctx.h)
        assert len(ctx.me_search(jobs[:0])) == 0
    with pytest.raises(hmgpu.HmGpuError):
        hmgpu.Context(W + 1, H, 8, 2)
    with pytest.raises(hmgpu.HmGpuError, match="int16"):
        hmgpu.Context(8188, 64, 8, 1)      # (w + 7) * 4 would not fit the int16 quarter-pel clip bounds


def test_full_size_1080p_properties():
    """BASELINE cfg-2 size: size-independent properties instead of a (slow) full oracle pass."""
    w, h = 1920, 1080
    fr = synth.luma_frames(w, h, 2, 8).astype(np.int16)
    jobs = worklist.frame_jobs(w, h, n_refs=1)
    with hmgpu.Context(w, h, 8, 1) as ctx:
        ctx.ref_upload(0, fr[0])
        # 1) a picture searched against itself finds the zero vector with zero SAD everywhere
        ctx.org_upload(fr[0])
        z = jobs.copy()
        z["pred_x"] = z["pred_y"] = z["start_x"] = z["start_y"] = 0
        bd = worklist.clip_bounds_np(w, h, -(z["clip_hmin"].astype(np.int32) // 4) - 71, -(z["clip_vmin"].astype(np.int32) // 4) - 71)
        z["win_l"], z["win_t"], z["win_r"], z["win_b"] = worklist.search_range_np(bd, 0, 0, 64)
        r = ctx.me_search(z)
        assert (r["int_x"] == 0).all() and (r["int_y"] == 0).all() and (r["int_sad"] == 0).all()
        assert (r["half_x"] == 0).all() and (r["qter_y"] == 0).all()
        # 2) real motion: spot-check a seeded sample of the full job list against the oracle
        ctx.org_upload(fr[1])
        r = ctx.me_search(jobs)
        idx = np.random.default_rng(3).choice(len(jobs), 300, replace=False)
        exp = oracle_me(jobs[idx], [padded_ref(fr[0])], fr[1], 8)
        assert_results_equal(r[idx], exp, jobs[idx])
        # 3) idempotence: the same batch again gives the same bytes
        r2 = ctx.me_search(jobs)
        assert r.tobytes() == r2.tobytes()
        # 4) the chunked two-lane pipeline (large batches) and the single-shot path give the same bytes
        ctx.set_option("pipeline", 0)
        r3 = ctx.me_search(jobs)
        ctx.set_option("pipeline", 1)
        assert r.tobytes() == r3.tobytes()


@pytest.mark.parametrize("pinned", [False, True])
def test_pipelined_batch_with_key_blocks(pinned):
    """>= 65536 jobs take the two-lane pipeline; bi-pred key-pattern blocks travel once for all chunks.
    Checked against the oracle on the distinct jobs the batch is tiled from."""
    rng = np.random.default_rng(5)
    fr = _frames(8, 4)
    n_base = 96
    base, blocks, n_elems = _random_jobs(rng, n_base, 8, "tz", 2, org_blocks=True)
    org_blocks = np.zeros(max(1, n_elems), np.int16)
    for (i, w, h) in blocks:
        j = base[i]
        o = fr[2][int(j["pu_y"]):int(j["pu_y"]) + h, int(j["pu_x"]):int(j["pu_x"]) + w].astype(np.int32)
        org_blocks[int(j["org_offset"]):int(j["org_offset"]) + w * h] = (2 * o - rng.integers(0, 256, (h, w))).astype(np.int16).ravel()
    exp = oracle_me(base, [padded_ref(fr[0]), padded_ref(fr[1])], fr[2], 8, org_blocks)
    reps = 70000 // n_base + 1
    jobs = np.tile(base, reps)
    with hmgpu.Context(W, H, 8, 2) as ctx:
        ctx.ref_upload(0, fr[0]); ctx.ref_upload(1, fr[1]); ctx.org_upload(fr[2])
        if pinned:
            hj = ctx.host_array(jobs.shape, hmgpu.ME_JOB); hj[...] = jobs
            hr = ctx.host_array(jobs.shape, hmgpu.ME_RESULT)
            got = ctx.me_search(hj, org_blocks, out=hr).copy()   # the pinned buffer dies with the context
        else:
            got = ctx.me_search(jobs, org_blocks)
    assert_results_equal(got, np.tile(exp, reps), jobs)


def test_pipelined_batch_reports_bad_job():
    """large batches are range-checked on the device (job_scan_kernel): same error contract as the host check"""
    fr = _frames(8, 2)
    jobs = worklist.frame_jobs(W, H, n_refs=1)
    jobs = np.tile(jobs, 70000 // len(jobs) + 1)
    bad = jobs.copy()
    bad[66001]["pu_w"] = 7
    worse = jobs.copy()
    worse[40000]["ref_slot"] = 1          # slot never uploaded
    with hmgpu.Context(W, H, 8, 2) as ctx:
        ctx.ref_upload(0, fr[0]); ctx.org_upload(fr[1])
        with pytest.raises(hmgpu.HmGpuError, match="job 66001"):
            ctx.me_search(bad)
        with pytest.raises(hmgpu.HmGpuError, match="job 40000"):
            ctx.me_search(worse)
        r = ctx.me_search(jobs)            # the context stays usable
        assert int(r["n_cand"].min()) >= 18


def test_full_size_4k_main10_properties():
    """BASELINE cfg-5 size (3840x2160, 10-bit, SearchRange 128): size-independent properties + a sampled oracle check."""
    w, h, bd = 3840, 2160, 10
    fr = synth.luma_frames(w, h, 2, bd).astype(np.int16)
    pus = worklist.pu_list(w, h)
    pus = pus[::7]                                   # every 7th PU of the quadtree: ~170 k jobs, all shapes
    jobs = worklist.frame_jobs(w, h, n_refs=1, search_range=128, pus=pus)
    with hmgpu.Context(w, h, bd, 1) as ctx:
        ctx.ref_upload(0, fr[0])
        ctx.org_upload(fr[0])
        z = jobs.copy()
        z["pred_x"] = z["pred_y"] = z["start_x"] = z["start_y"] = 0
        bdn = worklist.clip_bounds_np(w, h, -(z["clip_hmin"].astype(np.int32) // 4) - 71, -(z["clip_vmin"].astype(np.int32) // 4) - 71)
        z["win_l"], z["win_t"], z["win_r"], z["win_b"] = worklist.search_range_np(bdn, 0, 0, 128)
        r = ctx.me_search(z)
        assert (r["int_x"] == 0).all() and (r["int_y"] == 0).all() and (r["int_sad"] == 0).all()
        assert (r["half_x"] == 0).all() and (r["qter_y"] == 0).all()
        ctx.org_upload(fr[1])
        r = ctx.me_search(jobs)
        idx = np.random.default_rng(4).choice(len(jobs), 200, replace=False)
        exp = oracle_me(jobs[idx], [padded_ref(fr[0])], fr[1], bd)
        assert_results_equal(r[idx], exp, jobs[idx])
        assert r.tobytes() == ctx.me_search(jobs).tobytes()


def test_mailbox_server_restarts_and_uploads():
    """small calls go through the resident mailbox server: it must survive its own idle exit (a pause longer than the idle
    limit between calls), picture uploads in between (the host stops it, the planes change) and bursts of calls"""
    import time
    rng = np.random.default_rng(77)
    fr = _frames(8, 4)
    jobs, _, _ = _random_jobs(rng, 64, 8, "tz", 2)
    pads = [padded_ref(fr[k]) for k in range(3)]
    with hmgpu.Context(W, H, 8, 2) as ctx:
        ctx.ref_upload(0, fr[0]); ctx.ref_upload(1, fr[1]); ctx.org_upload(fr[3])
        exp = oracle_me(jobs, [pads[0], pads[1]], fr[3], 8)
        for i in range(0, 32, 4):                       # burst
            assert_results_equal(ctx.me_search(jobs[i:i + 4]), exp[i:i + 4], jobs[i:i + 4])
        time.sleep(0.02)                                # the server leaves (idle), the next call starts a new generation
        assert_results_equal(ctx.me_search(jobs[32:35]), exp[32:35], jobs[32:35])
        ctx.ref_upload(1, fr[2])                        # new picture in slot 1: the server is stopped and restarted
        exp2 = oracle_me(jobs, [pads[0], pads[2]], fr[3], 8)
        for i in range(32, 64, 8):
            assert_results_equal(ctx.me_search(jobs[i:i + 8]), exp2[i:i + 8], jobs[i:i + 8])
            time.sleep(0.001)
        ctx.org_upload(fr[2])
        exp3 = oracle_me(jobs[:16], [pads[0], pads[2]], fr[2], 8)
        assert_results_equal(ctx.me_search(jobs[:16]), exp3, jobs[:16])


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("path", ["batch", "small_calls"])
def test_selective_search(bit_depth, path):
    """xTZSearchSelective (FastSearch = 2): job.kind = KIND_SELECTIVE, the three spatial MV predictors in the
    side array; batch kernels and the low-latency path against the oracle (itself pinned to the reference)"""
    rng = np.random.default_rng(61)
    noise = bit_depth == 10
    fr = _frames(bit_depth, 4, noise)
    n = 72 if path == "batch" else 30
    jobs, _, _ = _random_jobs(rng, n, bit_depth, "tz", 2)
    side = np.zeros(6 * n, np.int16)
    for i in range(n):
        jobs[i]["kind"] = hmgpu.KIND_SELECTIVE
        jobs[i]["org_offset"] = 6 * i
        jobs[i]["search_range"] = int(rng.choice([8, 16, 64]))
        bd = [int(jobs[i][k]) for k in ("clip_hmin", "clip_hmax", "clip_vmin", "clip_vmax")]
        jobs[i]["win_l"], jobs[i]["win_t"], jobs[i]["win_r"], jobs[i]["win_b"] = hmgpu.search_range(
            bd, int(jobs[i]["pred_x"]), int(jobs[i]["pred_y"]), int(jobs[i]["search_range"]))
        side[6 * i:6 * i + 6] = rng.integers(-80, 81, 6)
        if i % 4 == 0:
            side[6 * i:6 * i + 2] = (int(jobs[i]["pred_x"]) + 12, int(jobs[i]["pred_y"]) - 8)
    pads = [padded_ref(fr[k]) for k in range(2)]
    exp = oracle_me(jobs, pads, fr[3], bit_depth, side)
    with hmgpu.Context(W, H, bit_depth, 2) as ctx:
        ctx.ref_upload(0, fr[0]); ctx.ref_upload(1, fr[1]); ctx.org_upload(fr[3])
        if path == "batch":
            got = ctx.me_search(jobs, side)
        else:
            parts = []
            for i in range(0, n, 3):
                sub = jobs[i:i + 3].copy()
                sub_side = np.concatenate([side[int(o):int(o) + 6] for o in sub["org_offset"]])
                sub["org_offset"] = np.arange(len(sub)) * 6
                parts.append(ctx.me_search(sub, sub_side))
            got = np.concatenate(parts)
        with pytest.raises(hmgpu.HmGpuError, match="predictors"):
            ctx.me_search(jobs[:2], side[:6])
    assert_results_equal(got, exp, jobs)


@pytest.mark.parametrize("noise", [False, True])
@pytest.mark.parametrize("p2", ["0", "1"])
def test_tz_thread_per_job_kernels(noise, p2):
    """The one-thread-per-job TZ kernels (me_tz_thread.cu: per-shape launches, per-job windows in shared memory, hand-over of
    refinement / raster / far-start jobs to the warp-per-job kernel or, with HMGPU_TZ_P2=1, to their second pass) against the
    oracle, forced onto a batch far below the size from which the library picks them.  The job mix stresses what decides their
    control flow: 2Nx2N integer MVs next to and far from the predictor, predictors at the picture edge (windows touching the
    padded border are handed over), small search ranges (rings cut short), FEN on and off (without it the tall shapes go to
    the warp-per-job kernel), every PU shape."""
    rng = np.random.default_rng(77)
    fr = _frames(8, 4, noise)
    org = fr[3]
    jobs, _, _ = _random_jobs(rng, 3000, 8, "tz", 3)
    for i in range(len(jobs)):
        j = jobs[i]
        if i % 3 == 0:      # 2Nx2N integer MV within a few samples of the predictor (the common case in the encoder)
            j["flags"] |= hmgpu.F_HAS_2NX2N
            j["i2n_x"] = (int(j["pred_x"]) >> 2) + int(rng.integers(-3, 4))
            j["i2n_y"] = (int(j["pred_y"]) >> 2) + int(rng.integers(-3, 4))
        if i % 7 == 0:      # adaptive search range smaller than the far rings
            sr = int(rng.choice([1, 2, 4, 8, 16]))
            bd = [int(j[k]) for k in ("clip_hmin", "clip_hmax", "clip_vmin", "clip_vmax")]
            j["search_range"] = sr
            j["win_l"], j["win_t"], j["win_r"], j["win_b"] = hmgpu.search_range(bd, int(j["pred_x"]), int(j["pred_y"]), sr)
    pads = [padded_ref(fr[k]) for k in range(3)]
    exp = oracle_me(jobs, pads, org, 8)
    with hmgpu.Context(W, H, 8, 3) as ctx:
        for k in range(3):
            ctx.ref_upload(k, fr[k])
        ctx.org_upload(org)
        ctx.set_option("tz_thread_min", 1)
        ctx.set_option("tz_p2", int(p2))
        got = ctx.me_search(jobs)
        assert_results_equal(got, exp, jobs)
        # the warp-per-job mapping gives the same bytes
        ctx.set_option("tz_thread", 0)
        got0 = ctx.me_search(jobs)
        assert got.tobytes() == got0.tobytes()
        # ... and so does the warp-per-job kernel without its merged passes (hand-over jobs resume in it either way)
        ctx.set_option("tz_merge", 0)
        assert ctx.me_search(jobs).tobytes() == got.tobytes()
        ctx.set_option("tz_thread", 1)
        assert ctx.me_search(jobs).tobytes() == got.tobytes()
