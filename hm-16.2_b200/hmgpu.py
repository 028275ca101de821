"""ctypes binding of libhmgpu.so -- the host-side mirror used by tests/ and bench.py.

Every call goes through the C ABI declared in include/hmgpu.h; there is no Python or CPU
implementation behind it.  If the shared library is missing or no CUDA device is present the
import / Context() fails loudly (no fallback).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhmgpu.so")

# flags (include/hmgpu.h)
F_FEN, F_HADME, F_LOSSLESS, F_HAS_2NX2N, F_FULL, F_INTEGER, F_FRAC, F_ORG_BLOCK = (1 << i for i in range(8))
DF_SAD, DF_SAD_GENERIC, DF_HADS, DF_SSE = range(4)
KIND_DEFAULT, KIND_SELECTIVE = 0, 1

ME_JOB = np.dtype([
    ("pu_x", "<i2"), ("pu_y", "<i2"), ("pu_w", "u1"), ("pu_h", "u1"), ("ref_slot", "u1"), ("flags", "u1"),
    ("pred_x", "<i2"), ("pred_y", "<i2"), ("start_x", "<i2"), ("start_y", "<i2"),
    ("win_l", "<i2"), ("win_t", "<i2"), ("win_r", "<i2"), ("win_b", "<i2"),
    ("i2n_x", "<i2"), ("i2n_y", "<i2"),
    ("clip_hmin", "<i2"), ("clip_hmax", "<i2"), ("clip_vmin", "<i2"), ("clip_vmax", "<i2"),
    ("search_range", "<i2"), ("kind", "<i2"),
    ("ui_cost", "<u4"), ("org_offset", "<u4")], align=True)
ME_RESULT = np.dtype([
    ("int_x", "<i2"), ("int_y", "<i2"), ("int_sad", "<u4"),
    ("half_x", "<i2"), ("half_y", "<i2"), ("qter_x", "<i2"), ("qter_y", "<i2"),
    ("frac_cost", "<u4"), ("n_cand", "<u4")], align=True)
DIST_ITEM = np.dtype([
    ("org_offset", "<u4"), ("cur_offset", "<u4"), ("org_stride", "<i4"), ("cur_stride", "<i4"),
    ("w", "u1"), ("h", "u1"), ("func", "u1"), ("sub_shift", "u1")], align=True)
MC_JOB = np.dtype([
    ("pu_x", "<i2"), ("pu_y", "<i2"), ("pu_w", "u1"), ("pu_h", "u1"), ("ref_slot", "u1"), ("reserved", "u1"),
    ("mv_x", "<i2"), ("mv_y", "<i2"), ("dst_offset", "<u4")], align=True)
PRED_JOB = np.dtype([
    ("pu_x", "<i2"), ("pu_y", "<i2"), ("pu_w", "u1"), ("pu_h", "u1"), ("ref_slot", "i1", (2,)),
    ("mv_x", "<i2", (2,)), ("mv_y", "<i2", (2,)), ("dst_offset", "<u4")], align=True)
assert ME_JOB.itemsize == 48 and ME_RESULT.itemsize == 24 and DIST_ITEM.itemsize == 20 and MC_JOB.itemsize == 16
assert PRED_JOB.itemsize == 20
INTRA_JOB = np.dtype([("org_offset", "<u4"), ("ref_offset", "<u4"), ("size", "u1"), ("flags", "u1"), ("reserved", "<u2")], align=True)
assert INTRA_JOB.itemsize == 12
# hmgpu_rdoq_bits / hmgpu_rdoq_job of include/hmgpu.h
RDOQ_BITS = np.dtype([("sig_group", "<i4", (2, 2)), ("sig", "<i4", (44, 2)), ("last_x", "<i4", (2, 10)), ("last_y", "<i4", (2, 10)),
                      ("greater_one", "<i4", (24, 2)), ("level_abs", "<i4", (6, 2))])
RDOQ_JOB = np.dtype([("log2_size", "<i4"), ("channel", "<i4"), ("scan", "<i4"), ("flags", "<u4"), ("qbits", "<i4"), ("qp_per", "<i4"),
                     ("qp_rem", "<i4"), ("go_rice_init", "<i4"), ("cbf_bits", "<i4", (2,)), ("bit_depth", "<i4"), ("bits_index", "<i4"),
                     ("coef_offset", "<u4"), ("reserved", "<u4"), ("err_scale", "<f8"), ("lambda", "<f8")], align=True)
assert RDOQ_BITS.itemsize == 768 and RDOQ_JOB.itemsize == 72
RDOQ_SIGN_HIDE = 1
IF_ABOVE, IF_LEFT, IF_EDGE_FILTERS, IF_SATD, IF_NO_SMOOTH = 1, 2, 4, 8, 16

EXPORTS = [
    "hmgpu_create", "hmgpu_destroy", "hmgpu_last_error", "hmgpu_abi_version", "hmgpu_launch_count",
    "hmgpu_stream", "hmgpu_synchronize", "hmgpu_set_option", "hmgpu_clip_bounds_ctu", "hmgpu_host_alloc", "hmgpu_host_free", "hmgpu_struct_sizes", "hmgpu_ref_upload", "hmgpu_ref_release",
    "hmgpu_ref_download_plane", "hmgpu_ref_upload_device", "hmgpu_org_upload", "hmgpu_org_upload_device",
    "hmgpu_me_search", "hmgpu_me_submit", "hmgpu_me_wait", "hmgpu_pu_submit", "hmgpu_pu_wait", "hmgpu_me_search_device", "hmgpu_clip_bounds", "hmgpu_search_range",
    "hmgpu_dist_batch", "hmgpu_intra_costs", "hmgpu_sao_stats", "hmgpu_sao_apply", "hmgpu_deblock", "hmgpu_mv_bits", "hmgpu_mv_cost", "hmgpu_mc_luma", "hmgpu_predict", "hmgpu_pred_error", "hmgpu_merge_skip_dist", "hmgpu_fwd_transform", "hmgpu_inv_transform",
    "hmgpu_quant", "hmgpu_rdoq", "hmgpu_dequant", "hmgpu_residual_tus", "hmgpu_profile_enable", "hmgpu_profile_stage_count", "hmgpu_profile_stage_name",
    "hmgpu_profile_read", "hmgpu_microbench"]

_lib = None


class HmGpuError(RuntimeError):
    pass


def lib():
    """load libhmgpu.so (fails loudly when it has not been built)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HmGpuError("libhmgpu.so not built: run `make -C hm-16.2_b200` (or __graft_entry__.build())")
    L = C.CDLL(LIB_PATH)
    vp, ci, cu = C.c_void_p, C.c_int, C.c_uint
    L.hmgpu_create.argtypes = [ci, ci, ci, ci, ci, C.POINTER(vp)]
    L.hmgpu_destroy.argtypes = [vp]
    L.hmgpu_destroy.restype = None
    L.hmgpu_last_error.argtypes = [vp]
    L.hmgpu_last_error.restype = C.c_char_p
    L.hmgpu_launch_count.argtypes = [vp]
    L.hmgpu_launch_count.restype = C.c_uint64
    L.hmgpu_stream.argtypes = [vp]
    L.hmgpu_stream.restype = vp
    L.hmgpu_synchronize.argtypes = [vp]
    L.hmgpu_set_option.argtypes = [vp, C.c_char_p, ci]
    L.hmgpu_clip_bounds_ctu.argtypes = [ci, ci, ci, ci, ci, ci, vp]
    L.hmgpu_clip_bounds_ctu.restype = None
    L.hmgpu_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.hmgpu_host_free.argtypes = [vp, vp]
    L.hmgpu_struct_sizes.argtypes = [vp]
    L.hmgpu_struct_sizes.restype = None
    L.hmgpu_ref_upload.argtypes = [vp, ci, vp, ci, vp, vp, ci]
    L.hmgpu_ref_release.argtypes = [vp, ci]
    L.hmgpu_ref_download_plane.argtypes = [vp, ci, ci, ci, vp]
    L.hmgpu_ref_upload_device.argtypes = [vp, ci, vp, ci]
    L.hmgpu_org_upload.argtypes = [vp, vp, ci]
    L.hmgpu_org_upload_device.argtypes = [vp, vp, ci]
    L.hmgpu_me_search.argtypes = [vp, vp, ci, vp, ci, vp]
    L.hmgpu_me_search_device.argtypes = [vp, vp, ci, vp, vp, ci]
    L.hmgpu_me_submit.argtypes = [vp, vp, ci, vp, ci]
    L.hmgpu_me_wait.argtypes = [vp, vp]
    L.hmgpu_pu_submit.argtypes = [vp, vp, ci, vp, ci, vp, vp, ci]
    L.hmgpu_pu_wait.argtypes = [vp, vp, vp]
    L.hmgpu_clip_bounds.argtypes = [ci, ci, ci, ci, vp]
    L.hmgpu_clip_bounds.restype = None
    L.hmgpu_search_range.argtypes = [vp, ci, ci, ci, vp]
    L.hmgpu_search_range.restype = None
    L.hmgpu_dist_batch.argtypes = [vp, vp, ci, vp, ci, vp, ci, vp]
    L.hmgpu_intra_costs.argtypes = [vp, vp, ci, vp, ci, vp, ci, vp]
    L.hmgpu_sao_stats.argtypes = [vp, vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp]
    L.hmgpu_sao_apply.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp]
    L.hmgpu_deblock.argtypes = [vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, ci, ci, ci, ci]
    L.hmgpu_mv_bits.argtypes = [ci] * 5
    L.hmgpu_mv_bits.restype = cu
    L.hmgpu_mv_cost.argtypes = [cu] + [ci] * 5
    L.hmgpu_mv_cost.restype = cu
    L.hmgpu_mc_luma.argtypes = [vp, vp, ci, vp, ci]
    L.hmgpu_predict.argtypes = [vp, vp, ci, ci, vp, ci]
    L.hmgpu_pred_error.argtypes = [vp, vp, ci, ci, vp]
    L.hmgpu_merge_skip_dist.argtypes = [vp, vp, ci, vp, vp, ci, vp, ci, vp]
    L.hmgpu_fwd_transform.argtypes = [vp, vp, ci, ci, ci, vp]
    L.hmgpu_inv_transform.argtypes = [vp, vp, ci, ci, ci, vp]
    L.hmgpu_quant.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, vp]
    L.hmgpu_rdoq.argtypes = [vp, vp, ci, vp, ci, vp, ci, vp, vp]
    L.hmgpu_dequant.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.hmgpu_residual_tus.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci, vp, vp, vp, vp]
    L.hmgpu_profile_enable.argtypes = [vp, ci]
    L.hmgpu_profile_stage_name.argtypes = [ci]
    L.hmgpu_profile_stage_name.restype = C.c_char_p
    L.hmgpu_profile_read.argtypes = [vp, vp, vp, ci]
    L.hmgpu_microbench.argtypes = [vp, ci, C.POINTER(C.c_double)]
    _lib = L
    return L


def clip_bounds(pic_w, pic_h, cu_x, cu_y, max_cu=None):
    b = np.zeros(4, np.int16)
    if max_cu is None:
        lib().hmgpu_clip_bounds(pic_w, pic_h, cu_x, cu_y, b.ctypes.data)
    else:
        lib().hmgpu_clip_bounds_ctu(pic_w, pic_h, cu_x, cu_y, max_cu, max_cu, b.ctypes.data)
    return b


def mv_bits(pred_x, pred_y, scale, x, y):
    return int(lib().hmgpu_mv_bits(int(pred_x), int(pred_y), int(scale), int(x), int(y)))


def mv_cost(ui_cost, pred_x, pred_y, scale, x, y):
    return int(lib().hmgpu_mv_cost(int(ui_cost) & 0xffffffff, int(pred_x), int(pred_y), int(scale), int(x), int(y)))


def search_range(bounds, pred_x, pred_y, srch_rng):
    b = np.ascontiguousarray(bounds, np.int16)
    out = np.zeros(4, np.int16)
    lib().hmgpu_search_range(b.ctypes.data, int(pred_x), int(pred_y), int(srch_rng), out.ctypes.data)
    return out


class Context:
    """one encoder instance bound to one CUDA device (hmgpu_create / hmgpu_destroy)"""

    def __init__(self, pic_w, pic_h, bit_depth=8, max_refs=4, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.hmgpu_create(device, pic_w, pic_h, bit_depth, max_refs, C.byref(self.h))
        if rc != 0:
            raise HmGpuError("hmgpu_create failed (%d): %s" % (rc, self.L.hmgpu_last_error(None).decode()))
        self.pic_w, self.pic_h, self.bit_depth = pic_w, pic_h, bit_depth

    def close(self):
        if self.h:
            for p in getattr(self, "_pinned", []):
                self.L.hmgpu_host_free(self.h, p)
            self._pinned = []
            self.L.hmgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise HmGpuError("libhmgpu error %d: %s" % (rc, self.L.hmgpu_last_error(self.h).decode()))

    @property
    def launches(self):
        return int(self.L.hmgpu_launch_count(self.h))

    @property
    def stream(self):
        return self.L.hmgpu_stream(self.h)

    def profile_enable(self, on=True):
        self._check(self.L.hmgpu_profile_enable(self.h, int(on)))

    def profile_read(self, reset=True):
        """-> {stage: (device_ms, launches)} accumulated since the last reset"""
        n = self.L.hmgpu_profile_stage_count()
        ms = np.zeros(n, np.float64)
        cnt = np.zeros(n, np.uint64)
        self._check(self.L.hmgpu_profile_read(self.h, ms.ctypes.data, cnt.ctypes.data, int(reset)))
        return {self.L.hmgpu_profile_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def microbench(self, which):
        g = C.c_double()
        self._check(self.L.hmgpu_microbench(self.h, which, C.byref(g)))
        return g.value

    def host_array(self, shape, dtype):
        """numpy array backed by page-locked memory (hmgpu_host_alloc); freed with the context"""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self.L.hmgpu_host_alloc(self.h, max(n, 1), C.byref(p)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def synchronize(self):
        self._check(self.L.hmgpu_synchronize(self.h))

    def set_option(self, name, value):
        """tuning knob of this context (hmgpu_set_option): kernel mapping switches, the resident server, ..."""
        self._check(self.L.hmgpu_set_option(self.h, name.encode(), int(value)))

    def ref_upload(self, slot, luma, cb=None, cr=None):
        luma = np.ascontiguousarray(luma, np.int16)
        assert luma.shape == (self.pic_h, self.pic_w)
        if cb is not None:
            cb = np.ascontiguousarray(cb, np.int16)
            cr = np.ascontiguousarray(cr, np.int16)
        self._check(self.L.hmgpu_ref_upload(self.h, slot, luma.ctypes.data, self.pic_w,
                                            cb.ctypes.data if cb is not None else None,
                                            cr.ctypes.data if cr is not None else None, self.pic_w // 2))

    def ref_upload_device(self, slot, d_ptr, stride):
        self._check(self.L.hmgpu_ref_upload_device(self.h, slot, d_ptr, stride))

    def ref_release(self, slot):
        self._check(self.L.hmgpu_ref_release(self.h, slot))

    def ref_plane(self, slot, fx, fy):
        out = np.zeros((self.pic_h + 160, self.pic_w + 160), np.int16)
        self._check(self.L.hmgpu_ref_download_plane(self.h, slot, fx, fy, out.ctypes.data))
        return out

    def org_upload(self, luma):
        luma = np.ascontiguousarray(luma, np.int16)
        assert luma.shape == (self.pic_h, self.pic_w)
        self._check(self.L.hmgpu_org_upload(self.h, luma.ctypes.data, self.pic_w))

    def org_upload_device(self, d_ptr, stride):
        self._check(self.L.hmgpu_org_upload_device(self.h, d_ptr, stride))

    def me_search(self, jobs, org_blocks=None, out=None):
        jobs = np.ascontiguousarray(jobs, ME_JOB)
        res = np.zeros(len(jobs), ME_RESULT) if out is None else out
        if org_blocks is not None:
            org_blocks = np.ascontiguousarray(org_blocks, np.int16)
        self._check(self.L.hmgpu_me_search(self.h, jobs.ctypes.data, len(jobs),
                                           org_blocks.ctypes.data if org_blocks is not None else None,
                                           org_blocks.size if org_blocks is not None else 0, res.ctypes.data))
        return res

    def me_submit(self, jobs, org_blocks=None):
        """asynchronous half of me_search for small batches; collect with me_wait()"""
        jobs = np.ascontiguousarray(jobs, ME_JOB)
        if org_blocks is not None:
            org_blocks = np.ascontiguousarray(org_blocks, np.int16)
        self._check(self.L.hmgpu_me_submit(self.h, jobs.ctypes.data, len(jobs),
                                           org_blocks.ctypes.data if org_blocks is not None else None,
                                           org_blocks.size if org_blocks is not None else 0))
        self._pending = len(jobs)

    def me_wait(self):
        res = np.zeros(self._pending, ME_RESULT)
        self._check(self.L.hmgpu_me_wait(self.h, res.ctypes.data))
        self._pending = 0
        return res

    def pu_submit(self, jobs, pred_jobs, pred_funcs, org_blocks=None):
        """searches + prediction-error jobs of one PU in one mailbox round trip; collect with pu_wait()"""
        jobs = np.ascontiguousarray(jobs, ME_JOB)
        pred_jobs = np.ascontiguousarray(pred_jobs, PRED_JOB)
        pred_funcs = np.ascontiguousarray(pred_funcs, np.uint8)
        assert len(pred_funcs) == len(pred_jobs)
        if org_blocks is not None:
            org_blocks = np.ascontiguousarray(org_blocks, np.int16)
        self._check(self.L.hmgpu_pu_submit(self.h, jobs.ctypes.data if len(jobs) else None, len(jobs),
                                           org_blocks.ctypes.data if org_blocks is not None else None,
                                           org_blocks.size if org_blocks is not None else 0,
                                           pred_jobs.ctypes.data if len(pred_jobs) else None,
                                           pred_funcs.ctypes.data if len(pred_jobs) else None, len(pred_jobs)))
        self._pending, self._pending_pred = len(jobs), len(pred_jobs)

    def pu_wait(self):
        res = np.zeros(self._pending, ME_RESULT)
        out = np.zeros(self._pending_pred, np.uint32)
        self._check(self.L.hmgpu_pu_wait(self.h, res.ctypes.data if len(res) else None, out.ctypes.data if len(out) else None))
        self._pending = self._pending_pred = 0
        return res, out

    def me_search_device(self, d_jobs, n_jobs, d_org_blocks, d_results, flags_any):
        self._check(self.L.hmgpu_me_search_device(self.h, d_jobs, n_jobs, d_org_blocks, d_results, int(flags_any)))

    def dist_batch(self, org, cur, items):
        org = np.ascontiguousarray(org, np.int16)
        cur = np.ascontiguousarray(cur, np.int16)
        items = np.ascontiguousarray(items, DIST_ITEM)
        out = np.zeros(len(items), np.uint32)
        self._check(self.L.hmgpu_dist_batch(self.h, org.ctypes.data, org.size, cur.ctypes.data, cur.size,
                                            items.ctypes.data, len(items), out.ctypes.data))
        return out

    def sao_stats(self, rec, org, ctu_w, ctu_h, skip_r, skip_b, ctu_flags=None):
        """SAO statistics of one picture component (hmgpu_sao_stats): -> int64 [n_ctus, 5 types, 2 (diff, count), 32 classes]"""
        rec = np.ascontiguousarray(rec, np.int16)
        org = np.ascontiguousarray(org, np.int16)
        h, w = rec.shape
        n_ctus = ((w + ctu_w - 1) // ctu_w) * ((h + ctu_h - 1) // ctu_h)
        skip_r = np.ascontiguousarray(skip_r, np.int32)
        skip_b = np.ascontiguousarray(skip_b, np.int32)
        flags = None if ctu_flags is None else np.ascontiguousarray(ctu_flags, np.uint8)
        out = np.zeros((n_ctus, 5, 2, 32), np.int64)
        self._check(self.L.hmgpu_sao_stats(self.h, rec.ctypes.data, w, org.ctypes.data, w, w, h, ctu_w, ctu_h,
                                           None if flags is None else flags.ctypes.data, skip_r.ctypes.data, skip_b.ctypes.data, out.ctypes.data))
        return out

    def sao_apply(self, rec, ctu_w, ctu_h, types, offsets, ctu_flags=None):
        """SAO applied to one picture component (hmgpu_sao_apply): -> int16 picture"""
        rec = np.ascontiguousarray(rec, np.int16)
        h, w = rec.shape
        types = np.ascontiguousarray(types, np.int8)
        offsets = np.ascontiguousarray(offsets, np.int32)
        flags = None if ctu_flags is None else np.ascontiguousarray(ctu_flags, np.uint8)
        out = np.zeros_like(rec)
        self._check(self.L.hmgpu_sao_apply(self.h, rec.ctypes.data, w, w, h, ctu_w, ctu_h, None if flags is None else flags.ctypes.data,
                                           types.ctypes.data, offsets.ctypes.data, out.ctypes.data))
        return out

    def deblock(self, y, cb, cr, bs_ver, bs_hor, qp, nofilter, beta_offset_div2=0, tc_offset_div2=0, cb_qp_offset=0, cr_qp_offset=0):
        """deblocking of one picture (hmgpu_deblock): -> filtered copies (y, cb, cr)"""
        y, cb, cr = [np.array(a, np.int16, order="C", copy=True) for a in (y, cb, cr)]
        h, w = y.shape
        bs_ver, bs_hor, nofilter = [np.ascontiguousarray(a, np.uint8) for a in (bs_ver, bs_hor, nofilter)]
        qp = np.ascontiguousarray(qp, np.int8)
        self._check(self.L.hmgpu_deblock(self.h, y.ctypes.data, cb.ctypes.data, cr.ctypes.data, w, h, bs_ver.ctypes.data, bs_hor.ctypes.data,
                                         qp.ctypes.data, nofilter.ctypes.data, beta_offset_div2, tc_offset_div2, cb_qp_offset, cr_qp_offset))
        return y, cb, cr

    def intra_costs(self, jobs, org_blocks, ref_lines):
        """distortion of the 35 luma intra modes of every job (hmgpu_intra_costs): -> uint32 [n_jobs, 35]"""
        jobs = np.ascontiguousarray(jobs, INTRA_JOB)
        org_blocks = np.ascontiguousarray(org_blocks, np.int16)
        ref_lines = np.ascontiguousarray(ref_lines, np.int16)
        out = np.zeros((len(jobs), 35), np.uint32)
        self._check(self.L.hmgpu_intra_costs(self.h, jobs.ctypes.data, len(jobs), org_blocks.ctypes.data, org_blocks.size,
                                             ref_lines.ctypes.data, ref_lines.size, out.ctypes.data))
        return out

    def mc_luma(self, jobs, n_dst):
        jobs = np.ascontiguousarray(jobs, MC_JOB)
        dst = np.zeros(n_dst, np.int16)
        self._check(self.L.hmgpu_mc_luma(self.h, jobs.ctypes.data, len(jobs), dst.ctypes.data, n_dst))
        return dst

    def predict(self, jobs, n_dst, with_chroma=True):
        """motionCompensation of whole PUs (Y [+ Cb, Cr], uni / bi): -> int16 array of n_dst elements"""
        jobs = np.ascontiguousarray(jobs, PRED_JOB)
        dst = np.zeros(n_dst, np.int16)
        self._check(self.L.hmgpu_predict(self.h, jobs.ctypes.data, len(jobs), int(with_chroma), dst.ctypes.data, n_dst))
        return dst

    def merge_skip_dist(self, jobs, org_offset, org_blocks, n_pred):
        """MC of every merge candidate (Y, Cb, Cr) + SSE of the skip reconstruction per component: -> (pred int16[n_pred], sse uint32[n, 3])"""
        jobs = np.ascontiguousarray(jobs, PRED_JOB)
        org_offset = np.ascontiguousarray(org_offset, np.uint32)
        org_blocks = np.ascontiguousarray(org_blocks, np.int16)
        pred = np.zeros(n_pred, np.int16)
        sse = np.zeros((len(jobs), 3), np.uint32)
        self._check(self.L.hmgpu_merge_skip_dist(self.h, jobs.ctypes.data, len(jobs), org_offset.ctypes.data, org_blocks.ctypes.data,
                                                 org_blocks.size, pred.ctypes.data, n_pred, sse.ctypes.data))
        return pred, sse

    def pred_error(self, jobs, func):
        """luma prediction error (DF_SAD / DF_HADS) of every job against the source picture"""
        jobs = np.ascontiguousarray(jobs, PRED_JOB)
        out = np.zeros(len(jobs), np.uint32)
        self._check(self.L.hmgpu_pred_error(self.h, jobs.ctypes.data, len(jobs), int(func), out.ctypes.data))
        return out

    def fwd_transform(self, resi, n, use_dst=False):
        resi = np.ascontiguousarray(resi, np.int16).reshape(-1, n, n)
        out = np.zeros(resi.shape, np.int32)
        self._check(self.L.hmgpu_fwd_transform(self.h, resi.ctypes.data, resi.shape[0], n, int(use_dst), out.ctypes.data))
        return out

    def inv_transform(self, coeff, n, use_dst=False):
        coeff = np.ascontiguousarray(coeff, np.int32).reshape(-1, n, n)
        out = np.zeros(coeff.shape, np.int16)
        self._check(self.L.hmgpu_inv_transform(self.h, coeff.ctypes.data, coeff.shape[0], n, int(use_dst), out.ctypes.data))
        return out

    def quant(self, coeff, n, qp_per, qp_rem, is_intra):
        coeff = np.ascontiguousarray(coeff, np.int32).reshape(-1, n, n)
        level = np.zeros(coeff.shape, np.int32)
        delta = np.zeros(coeff.shape, np.int32)
        abs_sum = np.zeros(coeff.shape[0], np.uint32)
        self._check(self.L.hmgpu_quant(self.h, coeff.ctypes.data, coeff.shape[0], n, qp_per, qp_rem, int(is_intra),
                                       level.ctypes.data, delta.ctypes.data, abs_sum.ctypes.data))
        return level, delta, abs_sum

    def rdoq(self, jobs, bits, coef, out=None):
        """rate-distortion optimised quantisation of every TU of `jobs` (hmgpu_rdoq): -> (levels laid out like coef, uiAbsSum per TU).
        coef / out from host_array() (page-locked) are copied without the staging pass"""
        jobs = np.ascontiguousarray(jobs, RDOQ_JOB).ravel()
        bits = np.ascontiguousarray(bits, RDOQ_BITS).ravel()
        coef = np.ascontiguousarray(coef, np.int32).ravel()
        level = np.zeros_like(coef) if out is None else out
        assert level.dtype == np.int32 and level.size == coef.size and level.flags.c_contiguous
        abs_sum = np.zeros(len(jobs), np.int32)
        self._check(self.L.hmgpu_rdoq(self.h, jobs.ctypes.data, len(jobs), bits.ctypes.data, len(bits), coef.ctypes.data, coef.size,
                                      level.ctypes.data, abs_sum.ctypes.data))
        return level, abs_sum

    def dequant(self, level, n, qp_per, qp_rem):
        """hmgpu_dequant: [n_tus, n, n] levels -> transform coefficients (xDeQuant, flat quantiser)"""
        level = np.ascontiguousarray(level, np.int32).reshape(-1, n, n)
        coef = np.zeros_like(level)
        self._check(self.L.hmgpu_dequant(self.h, level.ctypes.data, level.shape[0], n, qp_per, qp_rem, coef.ctypes.data))
        return coef

    def residual_tus(self, resi, n, jobs, bits, use_dst=False):
        """hmgpu_residual_tus: transform + RDOQ + dequantisation + inverse transform + distortions of [n_tus, n, n] residual blocks
        -> (levels, uiAbsSum, reconstructed residuals, distortions [n_tus, 2] = (coded, nothing coded))"""
        resi = np.ascontiguousarray(resi, np.int16).reshape(-1, n, n)
        jobs = np.ascontiguousarray(jobs, RDOQ_JOB).ravel()
        bits = np.ascontiguousarray(bits, RDOQ_BITS).ravel()
        assert len(jobs) == resi.shape[0]
        level = np.zeros(resi.shape, np.int32)
        abs_sum = np.zeros(len(jobs), np.int32)
        rec = np.zeros(resi.shape, np.int16)
        dist = np.zeros((len(jobs), 2), np.uint32)
        self._check(self.L.hmgpu_residual_tus(self.h, resi.ctypes.data, resi.shape[0], n, int(use_dst), jobs.ctypes.data, bits.ctypes.data, len(bits),
                                              level.ctypes.data, abs_sum.ctypes.data, rec.ctypes.data, dist.ctypes.data))
        return level, abs_sum, rec, dist
