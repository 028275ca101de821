"""Per-call latency of hmgpu_me_search for single jobs (the HM drop-in calling pattern)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
import hmgpu  # noqa: E402
import synth  # noqa: E402
import worklist  # noqa: E402

W, H = 416, 240
fr = synth.luma_frames(W, H, 3, 8).astype(np.int16)
ctx = hmgpu.Context(W, H, 8, 2)
ctx.ref_upload(0, fr[0]); ctx.ref_upload(1, fr[1]); ctx.org_upload(fr[2])
tz = worklist.frame_jobs(W, H, n_refs=2, ref_dist=[2, 1])
fs = worklist.frame_jobs(W, H, n_refs=2, ref_dist=[2, 1], full_search=True)


def probe(name, jobs, n_calls=300, per_call=1, gap_us=0):
    sel = jobs[np.random.default_rng(0).choice(len(jobs), n_calls * per_call, replace=True)]
    for i in range(20):
        ctx.me_search(sel[i * per_call:(i + 1) * per_call])
    ctx.profile_read(reset=True); ctx.profile_enable(True)
    dt = 0.0
    for i in range(n_calls):
        t0 = time.perf_counter()
        ctx.me_search(sel[i * per_call:(i + 1) * per_call])
        t1 = time.perf_counter()
        dt += t1 - t0
        while (time.perf_counter() - t1) * 1e6 < gap_us:   # emulate the host work between two calls
            pass
    p = ctx.profile_read(reset=True); ctx.profile_enable(False)
    dev = {k: round(v[0] / max(1, v[1]) * 1e3, 1) for k, v in p.items() if v[1]}
    print("%-28s %7.1f us/call wall   device us/launch %s" % (name, dt / n_calls * 1e6, dev))


for shape in [(8, 8), (16, 16), (32, 32), (64, 64)]:
    m = (tz["pu_w"] == shape[0]) & (tz["pu_h"] == shape[1])
    probe("tz %dx%d x1" % shape, tz[m])
    probe("tz %dx%d x4" % shape, tz[m], per_call=4)
for shape in [(8, 8), (64, 64)]:
    m = (fs["pu_w"] == shape[0]) & (fs["pu_h"] == shape[1])
    probe("fs sr64 %dx%d x1" % shape, fs[m], n_calls=min(100, int(m.sum()) - 25))
m = (tz["pu_w"] == 16) & (tz["pu_h"] == 16)
for gap in (20, 50, 200, 1000, 5000):
    probe("tz 16x16 x1 gap %d us" % gap, tz[m], n_calls=200, gap_us=gap)
os.environ["HMGPU_NO_FASTPATH"] = "1"
m = (tz["pu_w"] == 16) & (tz["pu_h"] == 16)
probe("tz 16x16 x1 (batch pipeline)", tz[m])
ctx.close()
