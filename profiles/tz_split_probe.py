"""Where the TZ stage's time goes: the stage timed (CUDA events of the library's own stage profile) on subsets of the bench
workload -- every job, only the PU shapes of the one-thread-per-job kernels, only the larger PUs.
usage: python profiles/tz_split_probe.py"""
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
jobs_all = worklist.frame_jobs(W, H, n_refs=NREF, ref_dist=[NREF + 1 - k for k in range(NREF)])
thread_shapes = {(4, 8), (4, 16), (8, 4), (8, 8), (8, 16), (12, 16), (16, 4), (16, 8), (16, 12), (16, 16), (32, 8), (8, 32), (16, 32), (32, 16)}
is_thread = np.array([(int(w), int(h)) in thread_shapes for w, h in zip(jobs_all["pu_w"], jobs_all["pu_h"])])
ctx = hmgpu.Context(W, H, 8, NREF)
d_frames = torch.from_numpy(frames).cuda()
for s in range(NREF):
    ctx.ref_upload_device(s, d_frames[s].data_ptr(), W)
ctx.org_upload_device(d_frames[NREF + 1].data_ptr(), W)
for name, sel in (("all jobs", np.ones(len(jobs_all), bool)), ("thread-kernel shapes", is_thread), ("larger PUs", ~is_thread)):
    jobs = np.ascontiguousarray(jobs_all[sel])
    flags_any = int(np.bitwise_or.reduce(jobs["flags"]))
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(len(jobs), -1).copy()).cuda()
    d_res = torch.zeros((len(jobs), hmgpu.ME_RESULT.itemsize), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for i in range(3):
        ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
    ctx.synchronize()
    ctx.profile_read(reset=True)
    ctx.profile_enable(True)
    n = 5
    for i in range(n):
        ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
    ctx.synchronize()
    prof = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    print("%-22s %8d jobs  " % (name, len(jobs)) + "  ".join("%s %.3f ms" % (k, v[0] / n) for k, v in prof.items() if v[1]))
ctx.close()
