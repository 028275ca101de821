"""The kernels outside the ME step at production sizes, for ncu: phase_planes / pad_convert (one 1080p
reference upload), predict (luma + chroma, uni / bi), pred_error (SATD), fwd_transform<32/16/8/4>, quant, dist_batch.
usage: python profiles/prof_misc.py [reps]   (prints device ms per call measured with CUDA events)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
import hmgpu  # noqa: E402
import synth  # noqa: E402

W, H = 1920, 1080
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(3)
frames = synth.luma_frames(W, H, 3, 8).astype(np.int16)
cb = (frames[:, ::2, ::2] // 2 + 64).astype(np.int16)
ctx = hmgpu.Context(W, H, 8, 2)
stream = torch.cuda.ExternalStream(ctx.stream)
d_frames = torch.from_numpy(frames).cuda()


def timed(name, fn, n=reps):
    fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    print("%-28s %.3f ms/call" % (name, e0.elapsed_time(e1) / n))


# reference upload (device-resident source): pad_convert + phase_planes
timed("ref_upload_device (planes)", lambda: ctx.ref_upload_device(0, d_frames[0].data_ptr(), W))
ctx.ref_upload(0, frames[0], cb[0], cb[0])
ctx.ref_upload(1, frames[1], cb[1], cb[1])
ctx.org_upload(frames[2])
# PU prediction jobs over the picture: 16x16 PUs, a third of them bi-predicted
n = 20000
pj = np.zeros(n, hmgpu.PRED_JOB)
pj["pu_w"] = pj["pu_h"] = 16
pj["pu_x"] = rng.integers(0, (W - 16) // 4, n) * 4
pj["pu_y"] = rng.integers(0, (H - 16) // 4, n) * 4
pj["ref_slot"] = np.stack([rng.integers(0, 2, n), np.where(np.arange(n) % 3 == 0, 1, -1)], 1)
pj["mv_x"] = rng.integers(-60, 60, (n, 2))
pj["mv_y"] = rng.integers(-60, 60, (n, 2))
pj["dst_offset"] = np.arange(n) * 384
timed("predict 20000 x 16x16 Y+C", lambda: ctx.predict(pj, n * 384, True), 1)
timed("pred_error HADS 20000", lambda: ctx.pred_error(pj, hmgpu.DF_HADS), 1)
for tn, cnt in ((32, 8000), (16, 30000), (8, 100000), (4, 300000)):
    resi = rng.integers(-255, 256, (cnt, tn, tn)).astype(np.int16)
    timed("fwd_transform %dx%d x %d" % (tn, tn, cnt), lambda: ctx.fwd_transform(resi, tn), 1)
coef = rng.integers(-30000, 30000, (30000, 16, 16)).astype(np.int32)
timed("quant 16x16 x 30000", lambda: ctx.quant(coef, 16, 4, 2, False), 1)
m = 50000
org = rng.integers(0, 256, 1 << 20).astype(np.int16)
cur = rng.integers(0, 256, 1 << 20).astype(np.int16)
it = np.zeros(m, hmgpu.DIST_ITEM)
it["w"] = it["h"] = 16
it["org_stride"] = it["cur_stride"] = 64
it["org_offset"] = rng.integers(0, (1 << 20) - 64 * 16, m)
it["cur_offset"] = rng.integers(0, (1 << 20) - 64 * 16, m)
it["func"] = np.arange(m) % 4
timed("dist_batch 50000 x 16x16", lambda: ctx.dist_batch(org, cur, it), 1)
# f4: intra mode pre-selection, 8 000 PUs of 8x8 / 16x16 / 32x32
ni = 8000
ij = np.zeros(ni, hmgpu.INTRA_JOB)
oo = lo = 0
for i in range(ni):
    nn = (8, 16, 32)[i % 3]
    ij[i] = (oo, lo, nn, hmgpu.IF_ABOVE | hmgpu.IF_LEFT | hmgpu.IF_EDGE_FILTERS | hmgpu.IF_SATD, 0)
    oo += nn * nn; lo += 2 * (4 * nn + 1)
iorg = rng.integers(0, 256, oo).astype(np.int16)
ilin = rng.integers(0, 256, lo).astype(np.int16)
timed("intra_costs 8000 PUs x 35", lambda: ctx.intra_costs(ij, iorg, ilin), 1)
# f3: SAO statistics and application of the 1080p luma component
rec = frames[0]
orgp = np.clip(frames[0] + rng.integers(-4, 5, frames[0].shape), 0, 255).astype(np.int16)
sk_r, sk_b = np.full(5, 5, np.int32), np.full(5, 4, np.int32)
timed("sao_stats 1080p luma", lambda: ctx.sao_stats(rec, orgp, 64, 64, sk_r, sk_b), 1)
n_ctu = ((W + 63) // 64) * ((H + 63) // 64)
stype = (np.arange(n_ctu) % 6 - 1).astype(np.int8)
soff = np.zeros((n_ctu, 32), np.int32); soff[:, :5] = (3, 1, 0, -1, -3); soff[:, 12:16] = (2, 1, -1, -2)
timed("sao_apply 1080p luma", lambda: ctx.sao_apply(rec, 64, 64, stype, soff), 1)
# f2: merge candidates of 4 000 CUs of 16x16, five each
nm = 20000
mj = pj[:nm].copy()
mj["pu_x"] &= ~7; mj["pu_y"] &= ~7
moff = (np.arange(nm) // 5 * 384).astype(np.uint32)
morg = rng.integers(0, 256, nm // 5 * 384).astype(np.int16)
timed("merge_skip_dist 20000 cands", lambda: ctx.merge_skip_dist(mj, moff, morg, nm * 384), 1)
ctx.close()
