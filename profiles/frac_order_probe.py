"""Does the order of the jobs matter to the fractional stage?  The bench work-list is ordered by CU depth (every PU shape sweeps
the picture on its own); here the same jobs are also submitted sorted by (64-row band, shape).  Prints the per-stage device times
of both orders for the kernel selected by HMGPU_FRAC_TMA.  usage: [HMGPU_FRAC_TMA=0|1] python profiles/frac_order_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
import hmgpu  # noqa: E402
import synth  # noqa: E402
import worklist  # noqa: E402

W, H, NREF = 1920, 1080, 4
frames = synth.luma_frames(W, H, NREF + 2, 8).astype(np.int16)
jobs = worklist.frame_jobs(W, H, n_refs=NREF, ref_dist=[NREF + 1 - k for k in range(NREF)])
key = (jobs["pu_y"].astype(np.int64) >> 6) * 256 + (jobs["pu_w"].astype(np.int64) // 4 - 1) * 16 + (jobs["pu_h"].astype(np.int64) // 4 - 1)
orders = {"bench order": jobs, "band/shape order": jobs[np.argsort(key, kind="stable")]}
flags_any = int(np.bitwise_or.reduce(jobs["flags"]))
ctx = hmgpu.Context(W, H, 8, NREF)
d_frames = torch.from_numpy(frames).cuda()
for s in range(NREF):
    ctx.ref_upload_device(s, d_frames[s].data_ptr(), W)
ctx.org_upload_device(d_frames[NREF + 1].data_ptr(), W)
md5 = {}
for name, jl in orders.items():
    d_jobs = torch.from_numpy(jl.view(np.uint8).reshape(len(jl), -1).copy()).cuda()
    d_res = torch.zeros((len(jl), hmgpu.ME_RESULT.itemsize), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.me_search_device(d_jobs.data_ptr(), len(jl), None, d_res.data_ptr(), flags_any)
    ctx.synchronize()
    ctx.profile_read(reset=True)
    ctx.profile_enable(True)
    for _ in range(5):
        ctx.me_search_device(d_jobs.data_ptr(), len(jl), None, d_res.data_ptr(), flags_any)
    ctx.synchronize()
    prof = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    print("%-18s" % name, {k: round(v[0] / 5, 3) for k, v in prof.items() if v[1]})
    res = d_res.cpu().numpy().view(hmgpu.ME_RESULT).reshape(-1)
    md5[name] = int(res["frac_cost"].astype(np.int64).sum()), int(res["n_cand"].astype(np.int64).sum())
print(md5)
ctx.close()
