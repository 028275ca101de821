"""ctypes bindings for the CPU checker libraries -- TEST INFRASTRUCTURE ONLY.

  * ``oracle()``  -> oracle/_ref/liboracle.so : our plain-C restatement (oracle/hm_oracle.c)
  * ``ref()``     -> oracle/_ref/libhmref.so  : the UNMODIFIED HM-16.2 reference compiled from
                     /root/reference plus the extern "C" shim oracle/ref_harness.cpp

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (hm-16.2_b200/, libhmgpu.so) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
vp = C.c_void_p
ci = C.c_int
cu = C.c_uint


class SearchT(C.Structure):
    """mirror of hmo_search_t (oracle/hm_oracle.h)"""
    _fields_ = [
        ("org", vp), ("org_stride", ci), ("w", ci), ("h", ci),
        ("ref", vp), ("ref_stride", ci),
        ("l", ci), ("t", ci), ("r", ci), ("b", ci),
        ("ui_cost", cu), ("pred_x", ci), ("pred_y", ci),
        ("fen", ci), ("hadme", ci), ("lossless", ci), ("bit_depth", ci),
        ("pic_w", ci), ("pic_h", ci), ("cu_x", ci), ("cu_y", ci), ("search_range", ci),
        ("start_x", ci), ("start_y", ci),
        ("has_2nx2n", ci), ("i2n_x", ci), ("i2n_y", ci),
        ("mv_x", ci), ("mv_y", ci), ("sad", cu),
        ("half_x", ci), ("half_y", ci), ("qter_x", ci), ("qter_y", ci), ("frac_cost", cu),
        ("n_cand", cu),
        ("sel_pred", ci * 2 * 3),
    ]


def build_oracle():
    """compile oracle/hm_oracle.c (always possible; gcc only)"""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def build_ref():
    """compile the reference from /root/reference (only where it exists)"""
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    path = os.path.join(OUT, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.hmo_sad.restype = cu
    L.hmo_sad.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]
    L.hmo_hads.restype = cu
    L.hmo_hads.argtypes = [vp, ci, vp, ci, ci, ci, ci]
    L.hmo_sse.restype = cu
    L.hmo_sse.argtypes = [vp, ci, vp, ci, ci, ci, ci]
    L.hmo_component_bits.restype = cu
    L.hmo_component_bits.argtypes = [ci]
    L.hmo_mv_bits.restype = cu
    L.hmo_mv_bits.argtypes = [ci] * 5
    L.hmo_mv_cost.restype = cu
    L.hmo_mv_cost.argtypes = [cu] + [ci] * 5
    L.hmo_bits_cost.restype = cu
    L.hmo_bits_cost.argtypes = [cu, cu]
    L.hmo_lambda_to_cost.restype = cu
    L.hmo_lambda_to_cost.argtypes = [C.c_double]
    L.hmo_calc_rd_cost_sad.restype = C.c_double
    L.hmo_calc_rd_cost_sad.argtypes = [C.c_double, cu, cu]
    L.hmo_clip_bounds.argtypes = [ci] * 4 + [i32p]
    L.hmo_clip_mv.argtypes = [ci] * 4 + [i32p]
    L.hmo_set_search_range.argtypes = [ci] * 7 + [i32p]
    L.hmo_filter_hor.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, ci]
    L.hmo_filter_ver.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci]
    L.hmo_extend_border.argtypes = [i16p, ci, ci, ci, i16p]
    L.hmo_phase_planes.argtypes = [i16p, ci, ci, ci, i16p]
    L.hmo_pred_inter_blk.argtypes = [ci, vp, ci, ci, ci, ci, ci, ci, ci, vp, ci]
    L.hmo_add_avg.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, ci]
    L.hmo_bipred_key.argtypes = [vp, ci, vp, ci, ci, ci, vp, ci]
    for f in (L.hmo_pattern_search, L.hmo_tz_search, L.hmo_frac_search):
        f.argtypes = [C.POINTER(SearchT)]
        f.restype = None
    L.hmo_motion_estimation.argtypes = [C.POINTER(SearchT), ci, ci, C.POINTER(cu), i32p, C.POINTER(cu)]
    L.hmo_fwd_transform.argtypes = [ci, i32p, i32p, ci, ci, ci]
    L.hmo_transform_matrix.argtypes = [ci, i32p]
    L.hmo_inv_transform.argtypes = [ci, i32p, i32p, ci, ci]
    L.hmo_quant.restype = cu
    L.hmo_quant.argtypes = [i32p, ci, ci, ci, ci, ci, i32p, i32p]
    L.hmo_me_batch.restype = C.c_double
    L.hmo_me_batch.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, vp]
    L.hmo_intra_pred.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp]
    L.hmo_intra_use_filtered.restype = ci
    L.hmo_intra_use_filtered.argtypes = [ci, ci, ci]
    L.hmo_intra_costs.argtypes = [vp, vp, vp, ci, ci, ci, u32p]
    L.hmo_sao_blk_stats.argtypes = [vp, ci, vp, ci, ci, ci, ci, i32p, i32p, ci, i64p, i64p]
    L.hmo_sao_offset_block.argtypes = [ci, i32p, vp, ci, vp, ci, ci, ci, ci, ci]
    L.hmo_deblock_picture.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, ci, ci, ci, ci]
    L.hmo_rdoq.restype = ci
    L.hmo_rdoq.argtypes = [vp, vp, i32p, i32p]
    L.hmo_scan_order.argtypes = [ci, ci, vp, vp]
    L.hmo_dequant.argtypes = [i32p, ci, ci, ci, ci, ci, i32p]
    L.hmo_rdoq_batch.restype = C.c_longlong
    L.hmo_rdoq_batch.argtypes = [vp, i32p, vp, ci, vp, i32p, i32p]
    _oracle = L
    return L


# hmo_rdoq_bits / hmo_rdoq_tu of oracle/hm_oracle.h
RDOQ_BITS = np.dtype([("sig_group", np.int32, (2, 2)), ("sig", np.int32, (44, 2)), ("last_x", np.int32, (2, 10)),
                      ("last_y", np.int32, (2, 10)), ("greater_one", np.int32, (24, 2)), ("level_abs", np.int32, (6, 2))])
RDOQ_TU = np.dtype([("log2_size", np.int32), ("channel", np.int32), ("scan", np.int32), ("sign_hide", np.int32), ("qbits", np.int32),
                    ("qp_per", np.int32), ("qp_rem", np.int32), ("go_rice_init", np.int32), ("cbf_bits", np.int32, 2),
                    ("bit_depth", np.int32), ("pad", np.int32), ("err_scale", np.float64), ("lambda", np.float64)], align=True)


def rdoq(tu, bits, coef):
    """hmo_rdoq on one TU: tu = RDOQ_TU record, bits = RDOQ_BITS record, coef = int32 raster -> (levels, abs_sum)"""
    coef = np.ascontiguousarray(coef, np.int32).ravel()
    level = np.zeros_like(coef)
    tu = np.ascontiguousarray(tu)
    bits = np.ascontiguousarray(bits)
    s = oracle().hmo_rdoq(tu.ctypes.data, bits.ctypes.data, coef, level)
    return level, int(s)


def have_ref():
    return os.path.exists(os.path.join(OUT, "libhmref.so"))


def ref():
    global _ref
    if _ref is not None:
        return _ref
    L = C.CDLL(os.path.join(OUT, "libhmref.so"))
    L.ref_sad_me.restype = cu
    L.ref_sad_me.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci]
    L.ref_dist_subpel.restype = cu
    L.ref_dist_subpel.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci]
    L.ref_dist_generic.restype = cu
    L.ref_dist_generic.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]
    L.ref_calc_had.restype = cu
    L.ref_calc_had.argtypes = [vp, ci, vp, ci, ci, ci, ci]
    L.ref_sse.restype = cu
    L.ref_sse.argtypes = [vp, ci, vp, ci, ci, ci, ci]
    L.ref_lambda_to_cost.restype = cu
    L.ref_lambda_to_cost.argtypes = [C.c_double]
    L.ref_mv_cost.restype = cu
    L.ref_mv_cost.argtypes = [cu] + [ci] * 5
    L.ref_mv_bits.restype = cu
    L.ref_mv_bits.argtypes = [ci] * 5
    L.ref_bits_cost.restype = cu
    L.ref_bits_cost.argtypes = [cu, cu]
    L.ref_calc_rd_cost_sad.restype = C.c_double
    L.ref_calc_rd_cost_sad.argtypes = [C.c_double, cu, cu]
    L.ref_clip_mv.argtypes = [ci] * 4 + [i32p]
    L.ref_set_search_range.argtypes = [ci] * 7 + [i32p]
    L.ref_filter_hor.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, ci]
    L.ref_filter_ver.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci]
    L.ref_extend_border.restype = ci
    L.ref_extend_border.argtypes = [i16p, ci, ci, i16p, ci]
    L.ref_pattern_search.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, cu, ci, ci, ci, ci, i32p, u32p]
    L.ref_tz_search.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, cu, ci, ci, ci, ci,
                                ci, ci, ci, ci, ci, ci, ci, ci, i32p, u32p]
    L.ref_tz_selective.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, cu, ci, ci, ci,
                                   ci, ci, ci, ci, ci, ci, ci, ci, i32p, i32p, u32p]
    L.ref_frac_search.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, cu, ci, ci, ci, ci, ci, i32p, i32p, u32p]
    L.ref_fwd_transform.argtypes = [ci, i32p, i32p, ci, ci, ci]
    L.ref_partial_butterfly.argtypes = [ci, i32p, i32p, ci, ci]
    L.ref_inv_transform.argtypes = [ci, i32p, i32p, ci, ci]
    L.ref_quant_scale.restype = ci
    L.ref_quant_scale.argtypes = [ci]
    L.ref_pred_inter_blk.argtypes = [ci, vp, ci, ci, ci, ci, ci, ci, ci, vp, ci]
    L.ref_me_batch.restype = C.c_double
    L.ref_me_batch.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, vp]
    L.ref_intra_pred.argtypes = [ci, vp, ci, ci, ci, ci, ci, vp]
    L.ref_intra_use_filtered.restype = ci
    L.ref_intra_use_filtered.argtypes = [ci, ci, ci]
    L.ref_sao_blk_stats.argtypes = [vp, ci, vp, ci, ci, ci, ci, i32p, i32p, ci, i64p, i64p]
    L.ref_sao_offset_block.argtypes = [ci, i32p, vp, ci, vp, ci, ci, ci, ci, ci]
    _ref = L
    return L


def me_batch(lib_fn, jobs, refs_padded, org, bit_depth, org_blocks=None, margin=80):
    """run a job batch (numpy structured array with the hmgpu_me_job layout) through
    hmo_me_batch / ref_me_batch.  refs_padded: list of padded int16 planes by slot.
    -> (results as raw bytes-compatible uint8 array [n, 24], cpu_seconds)"""
    jobs = np.ascontiguousarray(jobs)
    n = len(jobs)
    res = np.zeros((n, 24), np.uint8)
    if n == 0:
        return res, 0.0
    pw = refs_padded[0].shape[1]
    ptrs = (C.c_void_p * len(refs_padded))(*[r.ctypes.data + (margin * pw + margin) * 2 for r in refs_padded])
    org = np.ascontiguousarray(org, np.int16)
    ob = None if org_blocks is None else np.ascontiguousarray(org_blocks, np.int16)
    secs = lib_fn(jobs.ctypes.data, n, C.cast(ptrs, C.c_void_p), pw, org.ctypes.data, org.shape[1],
                  None if ob is None else ob.ctypes.data, bit_depth, res.ctypes.data)
    return res, secs


def ptr(a, off=0):
    """address of element `off` of a contiguous int16/int32 numpy array"""
    return a.ctypes.data + off * a.itemsize


def dequant(level, log2_size, qp_per, qp_rem, bit_depth):
    """hmo_dequant on one TU: levels (int32 raster) -> transform coefficients"""
    level = np.ascontiguousarray(level, np.int32).ravel()
    coef = np.zeros_like(level)
    oracle().hmo_dequant(level, level.size, log2_size, qp_per, qp_rem, bit_depth, coef)
    return coef


def rdoq_batch(tus, bits_index, coef_offset, bits, coef):
    """hmo_rdoq_batch: every TU of `tus` (RDOQ_TU records) in one C call -> (levels laid out like coef, sum of the uiAbsSum values)"""
    tus = np.ascontiguousarray(tus, RDOQ_TU)
    bits = np.ascontiguousarray(bits, RDOQ_BITS)
    bits_index = np.ascontiguousarray(bits_index, np.int32)
    coef_offset = np.ascontiguousarray(coef_offset, np.uint32)
    coef = np.ascontiguousarray(coef, np.int32)
    level = np.zeros_like(coef)
    total = oracle().hmo_rdoq_batch(tus.ctypes.data, bits_index, coef_offset.ctypes.data, len(tus), bits.ctypes.data, coef, level)
    return level, int(total)
