"""One device-resident step of the bench workload (1080p, 4 refs, TZ + fractional), for ncu.
usage: python profiles/prof_step.py [n_steps]"""
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = synth.luma_frames(W, H, NREF + 2, 8).astype(np.int16)
jobs = worklist.frame_jobs(W, H, n_refs=NREF, ref_dist=[NREF + 1 - k for k in range(NREF)])
flags_any = int(np.bitwise_or.reduce(jobs["flags"]))
ctx = hmgpu.Context(W, H, 8, NREF)
d_frames = torch.from_numpy(frames).cuda()
d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(len(jobs), -1).copy()).cuda()
d_res = torch.zeros((len(jobs), hmgpu.ME_RESULT.itemsize), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for s in range(NREF):
    ctx.ref_upload_device(s, d_frames[s].data_ptr(), W)
ctx.org_upload_device(d_frames[NREF + 1].data_ptr(), W)
for i in range(steps):
    ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
ctx.synchronize()
res = d_res.cpu().numpy().view(hmgpu.ME_RESULT).reshape(-1)
print("jobs", len(jobs), "cands", int(res["n_cand"].astype(np.int64).sum()))
ctx.close()
