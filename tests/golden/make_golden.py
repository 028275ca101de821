"""Generate the golden vectors of tests/golden/*.npz from the UNMODIFIED reference
(oracle/_ref/libhmref.so, compiled from /root/reference by oracle/Makefile).

The reference ships no known-answer tests of its own, so these fixtures are
outputs of the reference itself, run in the build container; they travel with the repo so the
oracle can be pinned anywhere.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hmgpu  # noqa: E402
import synth  # noqa: E402
from oracle import binding as B  # noqa: E402
from util import M, padded_ref  # noqa: E402

SHAPES = [(64, 64), (32, 32), (16, 16), (8, 8), (32, 64), (64, 32), (16, 32), (32, 16), (8, 16), (16, 8),
          (8, 4), (4, 8), (12, 16), (16, 12), (24, 32), (32, 24), (4, 16), (16, 4), (32, 8), (8, 32),
          (64, 16), (16, 64), (64, 48), (48, 64)]


def golden_dist(rng, bit_depth):
    R = B.ref()
    mx = (1 << bit_depth) - 1
    org = rng.integers(-mx, 2 * mx + 1, 64 * 80).astype(np.int16)
    cur = rng.integers(0, mx + 1, 64 * 96).astype(np.int16)
    rows = []
    for w in (4, 8, 12, 16, 24, 32, 48, 64):
        for h in (4, 8, 12, 16, 24, 32, 48, 64):
            oo, co = int(rng.integers(0, 16)), int(rng.integers(0, 32))
            po, pc = B.ptr(org, oo), B.ptr(cur, co)
            for ss in (0, 1):
                rows.append((w, h, oo, co, 0, ss, R.ref_sad_me(po, 80, pc, 96, w, h, ss, bit_depth)))
            rows.append((w, h, oo, co, 1, 0, R.ref_dist_subpel(po, 80, pc, 96, w, h, 1, bit_depth)))
            rows.append((w, h, oo, co, 2, 0, R.ref_sse(pc, 96, po, 80, w, h, bit_depth)))
            rows.append((w, h, oo, co, 3, 1, R.ref_dist_generic(po, 80, pc, 96, w, h, 0, bit_depth, 1)))
    return org, cur, np.array(rows, np.int64)


def golden_cost(rng):
    R = B.ref()
    rows = []
    for i in range(400):
        uc = int(rng.integers(0, 2 ** 32)) if i % 2 else int(rng.integers(0, 2 ** 22))
        px, py = [int(v) for v in rng.integers(-2000, 2000, 2)]
        sc = int(rng.integers(0, 3))
        x, y = [int(v) for v in rng.integers(-500, 500, 2)]
        rows.append((uc, px, py, sc, x, y, R.ref_mv_bits(px, py, sc, x, y), R.ref_mv_cost(uc, px, py, sc, x, y)))
    lam = rng.uniform(0.1, 5000, 64)
    lc = np.array([R.ref_lambda_to_cost(float(v)) for v in lam], np.int64)
    return np.array(rows, np.int64), lam, lc


def golden_clip(rng):
    R = B.ref()
    rows = []
    for i in range(300):
        pw, ph = [(416, 240), (1920, 1080), (3840, 2160)][i % 3]
        cx = int(rng.integers(0, pw // 8)) * 8
        cy = int(rng.integers(0, ph // 8)) * 8
        p = [int(v) for v in rng.integers(-9000, 9000, 2)]
        sr = int(rng.choice([4, 64, 128]))
        out = np.zeros(4, np.int32)
        R.ref_set_search_range(pw, ph, cx, cy, p[0], p[1], sr, out)
        mv = np.array(p, np.int32)
        R.ref_clip_mv(pw, ph, cx, cy, mv)
        rows.append([pw, ph, cx, cy, p[0], p[1], sr] + out.tolist() + mv.tolist())
    return np.array(rows, np.int64)


def golden_filters(rng, bit_depth):
    R = B.ref()
    mx = (1 << bit_depth) - 1
    src = rng.integers(0, mx + 1, (40, 64)).astype(np.int16)
    mid = rng.integers(-8192, 8192, (40, 64)).astype(np.int16)
    outs = {}
    for chroma in (0, 1):
        for frac in range(8 if chroma else 4):
            for last in (0, 1):
                d = np.zeros((17, 33), np.int16)
                R.ref_filter_hor(chroma, B.ptr(src, 8 * 64 + 8), 64, B.ptr(d), 33, 33, 17, frac, last, bit_depth)
                outs["h_%d_%d_%d" % (chroma, frac, last)] = d
                for first in (0, 1):
                    s = src if first else mid
                    d = np.zeros((17, 33), np.int16)
                    R.ref_filter_ver(chroma, B.ptr(s, 8 * 64 + 8), 64, B.ptr(d), 33, 33, 17, frac, first, last, bit_depth)
                    outs["v_%d_%d_%d_%d" % (chroma, frac, first, last)] = d
    return src, mid, outs


def golden_transform(rng, bit_depth):
    R = B.ref()
    outs = {}
    for n in (4, 8, 16, 32):
        blk = rng.integers(-(1 << bit_depth) + 1, 1 << bit_depth, (3, n, n)).astype(np.int32)
        blk[0] = (1 << bit_depth) - 1
        for dst in ((0, 1) if n == 4 else (0,)):
            c = np.zeros_like(blk)
            for t in range(3):
                R.ref_fwd_transform(bit_depth, np.ascontiguousarray(blk[t]), c[t], n, n, dst)
            outs["blk_%d" % n] = blk
            outs["coef_%d_%d" % (n, dst)] = c
    return outs


def golden_search(rng, bit_depth, n, mode):
    """jobs in the hmgpu_me_job layout + the reference's answers (ref_me_batch)"""
    w_, h_ = 416, 240
    fr = synth.luma_frames(w_, h_, 4, bit_depth).astype(np.int16)
    jobs = np.zeros(n, hmgpu.ME_JOB)
    for i in range(n):
        w, h = SHAPES[i % len(SHAPES)]
        cus = 8 if max(w, h) <= 8 else 16 if max(w, h) <= 16 else 32 if max(w, h) <= 32 else 64
        cx = int(rng.integers(0, w_ // cus)) * cus
        cy = int(rng.integers(0, h_ // cus)) * cus
        j = jobs[i]
        j["pu_x"] = cx + int(rng.integers(0, (cus - w) // 4 + 1)) * 4
        j["pu_y"] = cy + int(rng.integers(0, (cus - h) // 4 + 1)) * 4
        j["pu_w"], j["pu_h"], j["ref_slot"] = w, h, int(rng.integers(0, 3))
        pred = (int(rng.integers(-60, 60)), int(rng.integers(-60, 60))) if i % 7 else (int(rng.integers(-2500, 2500)), int(rng.integers(-1500, 1500)))
        j["pred_x"], j["pred_y"], j["start_x"], j["start_y"] = pred[0], pred[1], pred[0], pred[1]
        bd = np.zeros(4, np.int32)
        B.oracle().hmo_clip_bounds(w_, h_, cx, cy, bd)
        j["clip_hmin"], j["clip_hmax"], j["clip_vmin"], j["clip_vmax"] = bd
        sr = 64 if mode == "tz" else int(rng.choice([4, 8, 32]))
        ltrb = np.zeros(4, np.int32)
        B.ref().ref_set_search_range(w_, h_, cx, cy, pred[0], pred[1], sr, ltrb)
        j["win_l"], j["win_t"], j["win_r"], j["win_b"] = ltrb
        j["search_range"] = sr
        j["ui_cost"] = B.ref().ref_lambda_to_cost(float(rng.uniform(4, 200)))
        fl = hmgpu.F_INTEGER | hmgpu.F_FRAC | (hmgpu.F_FEN if i % 5 else 0) | (hmgpu.F_HADME if i % 6 else 0)
        if mode == "fs":
            fl |= hmgpu.F_FULL
        elif i % 2:
            fl |= hmgpu.F_HAS_2NX2N
            j["i2n_x"], j["i2n_y"] = int(rng.integers(-20, 20)), int(rng.integers(-20, 20))
        j["flags"] = fl
    pads = [padded_ref(fr[k]) for k in range(3)]
    res, _ = B.me_batch(B.ref().ref_me_batch, jobs, pads, fr[3], bit_depth)
    return jobs, res


def golden_selective(rng, bit_depth, n):
    """xTZSearchSelective (FastSearch=2): jobs with kind = KIND_SELECTIVE, their MV predictors (side array) and the reference's
    integer MV / SAD (ref_tz_selective)"""
    w_, h_ = 416, 240
    fr = synth.luma_frames(w_, h_, 4, bit_depth).astype(np.int16)
    jobs, _ = golden_search(rng, bit_depth, n, "tz")
    side = rng.integers(-80, 81, 6 * n).astype(np.int16)
    pads = [padded_ref(fr[k]) for k in range(3)]
    pw = pads[0].shape[1]
    rows = np.zeros((n, 3), np.int64)
    for i in range(n):
        j = jobs[i]
        j["kind"] = hmgpu.KIND_SELECTIVE
        j["org_offset"] = 6 * i
        j["search_range"] = int(rng.choice([8, 16, 64]))
        cu_x, cu_y = -(int(j["clip_hmin"]) // 4) - 71, -(int(j["clip_vmin"]) // 4) - 71
        ltrb = np.zeros(4, np.int32)
        B.ref().ref_set_search_range(w_, h_, cu_x, cu_y, int(j["pred_x"]), int(j["pred_y"]), int(j["search_range"]), ltrb)
        j["win_l"], j["win_t"], j["win_r"], j["win_b"] = ltrb
        if i % 4 == 0:
            side[6 * i:6 * i + 2] = (int(j["pred_x"]) + 12, int(j["pred_y"]) - 8)
        w, h, x0, y0 = int(j["pu_w"]), int(j["pu_h"]), int(j["pu_x"]), int(j["pu_y"])
        blk = np.ascontiguousarray(fr[3][y0:y0 + h, x0:x0 + w])
        pad = pads[int(j["ref_slot"])]
        mv = np.array([int(j["start_x"]), int(j["start_y"])], np.int32)
        sad = np.zeros(1, np.uint32)
        B.ref().ref_tz_selective(B.ptr(blk), w, w, h, B.ptr(pad, (y0 + M) * pw + x0 + M), pw, int(ltrb[0]), int(ltrb[1]), int(ltrb[2]), int(ltrb[3]),
                                 int(j["ui_cost"]), int(j["pred_x"]), int(j["pred_y"]), bit_depth, w_, h_, cu_x, cu_y, int(j["search_range"]),
                                 int(bool(int(j["flags"]) & hmgpu.F_HAS_2NX2N)), int(j["i2n_x"]), int(j["i2n_y"]),
                                 np.ascontiguousarray(side[6 * i:6 * i + 6].astype(np.int32)), mv, sad)
        rows[i] = (int(mv[0]), int(mv[1]), int(sad[0]))
    return jobs, side, rows


def main():
    assert B.have_ref(), "build oracle/_ref/libhmref.so first (make -C oracle ref)"
    rng = np.random.default_rng(20261018)
    out = {}
    for bd in (8, 10):
        org, cur, rows = golden_dist(rng, bd)
        out.update({"dist_org_%d" % bd: org, "dist_cur_%d" % bd: cur, "dist_rows_%d" % bd: rows})
        src, mid, f = golden_filters(rng, bd)
        out.update({"filt_src_%d" % bd: src, "filt_mid_%d" % bd: mid})
        out.update({"filt_%d_%s" % (bd, k): v for k, v in f.items()})
        out.update({"tr_%d_%s" % (bd, k): v for k, v in golden_transform(rng, bd).items()})
        for mode, n in (("tz", 144), ("fs", 48)):
            jobs, res = golden_search(rng, bd, n, mode)
            out["search_%s_jobs_%d" % (mode, bd)] = jobs.view(np.uint8).reshape(n, -1)
            out["search_%s_res_%d" % (mode, bd)] = res
    rows, lam, lc = golden_cost(rng)
    out.update({"cost_rows": rows, "cost_lambda": lam, "cost_lambda_ui": lc, "clip_rows": golden_clip(rng)})
    for bd in (8, 10):                                   # appended last: the arrays above keep their values
        jobs, side, rows = golden_selective(rng, bd, 48)
        out["sel_jobs_%d" % bd] = jobs.view(np.uint8).reshape(len(jobs), -1)
        out["sel_side_%d" % bd] = side
        out["sel_rows_%d" % bd] = rows
    path = os.path.join(HERE, "hm162_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
