"""Shared helpers of the test-suite: the oracle (oracle/hm_oracle.c) is the checker, the
product (libhmgpu via hm-16.2_b200/hmgpu.py) is the thing checked."""
import ctypes as C

import numpy as np

import hmgpu
from oracle import binding as B

M = 80  # luma padding


def padded_ref(luma):
    """int16 padded plane (replicate, 80) of an unpadded luma frame, via the oracle"""
    h, w = luma.shape
    out = np.zeros((h + 2 * M, w + 2 * M), np.int16)
    B.oracle().hmo_extend_border(np.ascontiguousarray(luma, np.int16), w, h, M, out)
    return out


def oracle_me(jobs, refs_padded, org, bit_depth, org_blocks=None):
    """expected hmgpu_me_result for every job, computed by the CPU oracle.
    refs_padded: list of padded int16 planes by slot; org: unpadded int16 source picture."""
    O = B.oracle()
    out = np.zeros(len(jobs), hmgpu.ME_RESULT)
    pic_h, pic_w = org.shape
    for i, j in enumerate(jobs):
        w, h = int(j["pu_w"]), int(j["pu_h"])
        fl = int(j["flags"])
        if fl & hmgpu.F_ORG_BLOCK:
            blk = np.ascontiguousarray(org_blocks[int(j["org_offset"]):int(j["org_offset"]) + w * h]).reshape(h, w)
        else:
            blk = np.ascontiguousarray(org[int(j["pu_y"]):int(j["pu_y"]) + h, int(j["pu_x"]):int(j["pu_x"]) + w])
        ref = refs_padded[int(j["ref_slot"])]
        pw = ref.shape[1]
        s = B.SearchT()
        s.org = B.ptr(blk); s.org_stride = w; s.w = w; s.h = h
        s.ref = B.ptr(ref, (int(j["pu_y"]) + M) * pw + int(j["pu_x"]) + M); s.ref_stride = pw
        s.l, s.t, s.r, s.b = int(j["win_l"]), int(j["win_t"]), int(j["win_r"]), int(j["win_b"])
        s.ui_cost = int(j["ui_cost"]); s.pred_x = int(j["pred_x"]); s.pred_y = int(j["pred_y"])
        s.fen = int(bool(fl & hmgpu.F_FEN)); s.hadme = int(bool(fl & hmgpu.F_HADME))
        s.lossless = int(bool(fl & hmgpu.F_LOSSLESS)); s.bit_depth = bit_depth
        # the oracle derives clipMv from (pic, cu); the job carries the bounds: invert them
        s.cu_x = -(int(j["clip_hmin"]) // 4) - 71
        s.cu_y = -(int(j["clip_vmin"]) // 4) - 71
        s.pic_w = int(j["clip_hmax"]) // 4 - 7 + s.cu_x
        s.pic_h = int(j["clip_vmax"]) // 4 - 7 + s.cu_y
        s.search_range = int(j["search_range"])
        s.start_x, s.start_y = int(j["start_x"]), int(j["start_y"])
        s.has_2nx2n = int(bool(fl & hmgpu.F_HAS_2NX2N)); s.i2n_x = int(j["i2n_x"]); s.i2n_y = int(j["i2n_y"])
        n = 0
        if fl & hmgpu.F_INTEGER:
            if fl & hmgpu.F_FULL:
                O.hmo_pattern_search(C.byref(s))
            elif int(j["kind"]) == hmgpu.KIND_SELECTIVE:
                off = int(j["org_offset"])
                for k in range(3):
                    s.sel_pred[k][0] = int(org_blocks[off + 2 * k]); s.sel_pred[k][1] = int(org_blocks[off + 2 * k + 1])
                O.hmo_tz_selective(C.byref(s))
            else:
                O.hmo_tz_search(C.byref(s))
            n = s.n_cand
            out[i]["int_sad"] = s.sad
        else:
            s.mv_x, s.mv_y = int(j["start_x"]), int(j["start_y"])
        out[i]["int_x"], out[i]["int_y"] = s.mv_x, s.mv_y
        if fl & hmgpu.F_FRAC:
            s.n_cand = 0
            O.hmo_frac_search(C.byref(s))
            n += s.n_cand
            out[i]["half_x"], out[i]["half_y"] = s.half_x, s.half_y
            out[i]["qter_x"], out[i]["qter_y"] = s.qter_x, s.qter_y
            out[i]["frac_cost"] = s.frac_cost
        out[i]["n_cand"] = n
    return out


def assert_results_equal(got, exp, jobs=None):
    for f in hmgpu.ME_RESULT.names:
        bad = np.nonzero(got[f] != exp[f])[0]
        if len(bad):
            i = int(bad[0])
            raise AssertionError("field %s differs at job %d (%d of %d jobs): got %s expected %s%s" % (
                f, i, len(bad), len(got), got[i], exp[i], "" if jobs is None else " job=%s" % (jobs[i],)))
