"""A/B of the TZ batch mappings on the bench workload (1080p, 4 refs): per-stage device times and an MD5 of the
result array, so that two runs with different HMGPU_TZ_SPLIT values can be compared byte for byte.
usage: HMGPU_TZ_SPLIT=k python profiles/tz_ab.py [n_steps]"""
import hashlib
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
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
for i in range(3):
    ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
ctx.synchronize()
ctx.profile_read(reset=True)
ctx.profile_enable(True)
for i in range(steps):
    ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
ctx.synchronize()
prof = ctx.profile_read(reset=True)
res = d_res.cpu().numpy()
r = res.view(hmgpu.ME_RESULT).reshape(-1)
print("split", os.environ.get("HMGPU_TZ_SPLIT", "default"), "jobs", len(jobs), "cands", int(r["n_cand"].astype(np.int64).sum()),
      "md5", hashlib.md5(res.tobytes()).hexdigest())
print("stage ms/step:", {k: round(v[0] / steps, 4) if isinstance(v, (tuple, list)) else v for k, v in prof.items()} if isinstance(prof, dict) else prof)
ctx.close()
