"""Synthetic ME work-lists with the shape of HM's own call stream.

TEncCu::xCompressCU (TEncCu.cpp:466-1040) visits every CU of the CTU quadtree (64 -> 8) and
xCheckRDCostInter (:1532) runs predInterSearch on its partitions: 2Nx2N, 2NxN, Nx2N at every
depth, the four AMP shapes for 16- and 32-wide CUs (64-wide AMP is merge-only, TEncCu.cpp:431-435;
inter NxN never occurs at these cfgs, :674).  predInterSearch (TEncSearch.cpp:3075) then calls
xMotionEstimation once per PU x reference picture.  frame_jobs() enumerates exactly that PU set
for one P picture and attaches a plausible predictor to each job (true global motion of the
synthetic clip + jitter), the reference's clipMv bounds and xSetSearchRange window.

In the real encoder the predictor of a PU depends on already-coded neighbours; a work-list is
what several encoder instances sharing one GPU submit concurrently, and what bench.py times.
"""
import numpy as np

from hmgpu import (F_FEN, F_FRAC, F_FULL, F_HADME, F_HAS_2NX2N, F_INTEGER, ME_JOB)

# partitions as (x, y, w, h) in units of CU size / 4
_PARTS_SYM = [
    [(0, 0, 4, 4)],                       # 2Nx2N
    [(0, 0, 4, 2), (0, 2, 4, 2)],         # 2NxN
    [(0, 0, 2, 4), (2, 0, 2, 4)],         # Nx2N
]
_PARTS_AMP = [
    [(0, 0, 4, 1), (0, 1, 4, 3)],         # 2NxnU
    [(0, 0, 4, 3), (0, 3, 4, 1)],         # 2NxnD
    [(0, 0, 1, 4), (1, 0, 3, 4)],         # nLx2N
    [(0, 0, 3, 4), (3, 0, 1, 4)],         # nRx2N
]


def lambda_to_cost(lam):
    """m_uiCost = floor(65536*sqrt(lambda)) (TComRdCost.cpp:196-209)"""
    return int(np.floor(65536.0 * np.sqrt(lam)))


def pu_list(pic_w, pic_h, amp=True):
    """-> int array [n, 7]: cu_x, cu_y, pu_x, pu_y, pu_w, pu_h, is_first_2Nx2N_at_depth0"""
    out = []
    for depth in range(4):
        s = 64 >> depth
        q = s // 4
        modes = list(_PARTS_SYM)
        if amp and s in (16, 32):
            modes += _PARTS_AMP
        for cy in range(0, pic_h - s + 1, s):
            for cx in range(0, pic_w - s + 1, s):
                for mi, mode in enumerate(modes):
                    for (px, py, pw, ph) in mode:
                        out.append((cx, cy, cx + px * q, cy + py * q, pw * q, ph * q,
                                    1 if (depth == 0 and mi == 0) else 0))
    return np.array(out, dtype=np.int32)


def clip_bounds_np(pic_w, pic_h, cu_x, cu_y):
    """TComDataCU::clipMv bounds (TComDataCU.cpp:2917-2929), vectorised"""
    return ((-64 - 8 - cu_x + 1) * 4, (pic_w + 8 - cu_x - 1) * 4,
            (-64 - 8 - cu_y + 1) * 4, (pic_h + 8 - cu_y - 1) * 4)


def search_range_np(bd, pred_x, pred_y, sr):
    """xSetSearchRange (TEncSearch.cpp:3911-3927), vectorised (int16 wrap never triggers here)"""
    hmin, hmax, vmin, vmax = bd
    px = np.clip(pred_x, hmin, hmax)
    py = np.clip(pred_y, vmin, vmax)
    r4 = sr << 2
    left = np.clip(px - r4, hmin, hmax) >> 2
    top = np.clip(py - r4, vmin, vmax) >> 2
    right = np.clip(px + r4, hmin, hmax) >> 2
    bottom = np.clip(py + r4, vmin, vmax) >> 2
    return left, top, right, bottom


def frame_jobs(pic_w, pic_h, n_refs=4, seed=7, full_search=False, search_range=64, fen=True, hadme=True,
               lam=57.9, motion_qpel=(12, -8), jitter=6, amp=True, frac=True, pus=None, ref_dist=None):
    """ME jobs of one P picture: every PU of every CU of the quadtree x every reference.
    motion_qpel: true global motion per frame (synth.py translates by (+3,-2) px/frame);
    ref_dist[slot]: temporal distance of the reference in that slot (default slot+1)."""
    rng = np.random.default_rng(seed)
    if pus is None:
        pus = pu_list(pic_w, pic_h, amp)
    n = len(pus) * n_refs
    jobs = np.zeros(n, ME_JOB)
    rep = np.repeat(pus, n_refs, axis=0)
    ref = np.tile(np.arange(n_refs), len(pus))
    cu_x, cu_y = rep[:, 0], rep[:, 1]
    jobs["pu_x"], jobs["pu_y"], jobs["pu_w"], jobs["pu_h"] = rep[:, 2], rep[:, 3], rep[:, 4], rep[:, 5]
    jobs["ref_slot"] = ref
    # the further the reference, the larger the displacement
    dist = (ref + 1) if ref_dist is None else np.asarray(ref_dist)[ref]
    mvx = motion_qpel[0] * dist + rng.integers(-jitter, jitter + 1, n)
    mvy = motion_qpel[1] * dist + rng.integers(-jitter, jitter + 1, n)
    jobs["pred_x"], jobs["pred_y"] = mvx, mvy
    jobs["start_x"], jobs["start_y"] = mvx, mvy
    bd = clip_bounds_np(pic_w, pic_h, cu_x, cu_y)
    jobs["clip_hmin"], jobs["clip_hmax"], jobs["clip_vmin"], jobs["clip_vmax"] = bd
    left, top, right, bottom = search_range_np(bd, mvx, mvy, search_range)
    jobs["win_l"], jobs["win_t"], jobs["win_r"], jobs["win_b"] = left, top, right, bottom
    jobs["i2n_x"] = (motion_qpel[0] * dist) // 4 + rng.integers(-1, 2, n)
    jobs["i2n_y"] = (motion_qpel[1] * dist) // 4 + rng.integers(-1, 2, n)
    jobs["search_range"] = search_range
    jobs["ui_cost"] = lambda_to_cost(lam)
    flags = np.full(n, F_INTEGER | (F_FRAC if frac else 0) | (F_FEN if fen else 0) | (F_HADME if hadme else 0), np.uint8)
    if full_search:
        flags |= F_FULL
    else:
        flags[rep[:, 6] == 0] |= F_HAS_2NX2N
    jobs["flags"] = flags
    return jobs
