"""RDOQ (f1, hmgpu_rdoq) at production batch size: the calls of the reference encoder's own xRateDistOptQuant
(tests/golden/rdoq_golden.npz: TUs of every size, luma + chroma, many coder states) repeated REP times in one batch -- about what
the residual quadtrees of a few 1080p CTU rows ask for.  Prints device ms of the four launches (stage timers), the wall time of
the blocking C-ABI call with host buffers, and -- as the reported CPU baseline -- the oracle (oracle/hm_rdoq.c, pinned to the same
dumped calls) on one host core over the un-repeated calls.  The batch's levels are checked against the dumped ones before timing.
usage: python profiles/prof_rdoq.py [rep] [calls]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import hmgpu  # noqa: E402
import rdoqdump  # noqa: E402
from test_golden import rdoq_golden_calls  # noqa: E402
from test_gpu_rdoq import batch_of  # noqa: E402

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps_timed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
calls = rdoq_golden_calls()
jobs1, bits, coef1 = batch_of(calls)
want = np.concatenate([c["level"] for c in calls])
n1 = coef1.size
jobs = np.tile(jobs1, rep)
coef = np.tile(coef1, rep)
jobs["coef_offset"] = (jobs["coef_offset"].astype(np.int64) + np.repeat(np.arange(rep, dtype=np.int64) * n1, len(jobs1))).astype(np.uint32)
sizes = {lg: int((jobs["log2_size"] == lg).sum()) for lg in (2, 3, 4, 5)}
print("batch: %d TUs (%s), %d coefficients, %d sets of bit estimates" % (len(jobs), sizes, coef.size, len(bits)))

modes = [int(m) for m in os.environ.get("RDOQ_MODES", "1,0").split(",")]
with hmgpu.Context(64, 64, 8, 1) as ctx:
    for mode in modes:
        ctx.set_option("rdoq_tu", mode)
        print("--- %s" % ("one thread per TU (rdoq_tu_kernel, one launch)" if mode else "one lane group per TU (rdoq_kernel<log2, lanes>, one launch per size)"))
        level, abs_sum = ctx.rdoq(jobs, bits, coef)                         # warm-up + check
        assert np.array_equal(level.reshape(rep, n1), np.tile(want, (rep, 1))), "levels differ from the reference encoder's"
        ctx.profile_enable(True)
        ctx.profile_read(True)
        t0 = time.perf_counter()
        for _ in range(reps_timed):
            ctx.rdoq(jobs, bits, coef)
        wall = (time.perf_counter() - t0) / reps_timed
        st = ctx.profile_read(True)
        dev_ms = st["quant"][0] / reps_timed
        print("device: %.3f ms per batch (%d launches)  -> %.2f M TU/s, %.1f M coef/s" % (dev_ms, st["quant"][1] // reps_timed, len(jobs) / dev_ms / 1e3, coef.size / dev_ms / 1e3))
        print("C-ABI call with host buffers: %.3f ms per batch -> %.2f M TU/s, %.1f M coef/s" % (wall * 1e3, len(jobs) / wall / 1e6, coef.size / wall / 1e6))
        # each size class alone, and the un-repeated calls (a small batch: the latency of the longest chain)
        for lg in (5, 4, 3, 2):
            sel = jobs["log2_size"] == lg
            ctx.rdoq(jobs[sel], bits, coef)
            ctx.profile_read(True)
            ctx.rdoq(jobs[sel], bits, coef)
            ms = ctx.profile_read(True)["quant"][0]
            print("  %2dx%-2d: %6d TUs  %.3f ms  (%.3f us per TU across the machine)" % (1 << lg, 1 << lg, int(sel.sum()), ms, ms * 1e3 / max(int(sel.sum()), 1)))
        ctx.rdoq(jobs1, bits, coef1)
        ctx.profile_read(True)
        ctx.rdoq(jobs1, bits, coef1)
        print("  the %d calls once: %.3f ms" % (len(jobs1), ctx.profile_read(True)["quant"][0]))
        ctx.profile_enable(False)

if os.environ.get("RDOQ_NO_CPU"):
    sys.exit(0)
try:
    from oracle import binding as B
    prepared = [(rdoqdump.to_tu_and_bits(c, B.RDOQ_TU, B.RDOQ_BITS), c["coef"]) for c in calls]
    t0 = time.perf_counter()
    loops = 0
    while time.perf_counter() - t0 < 10.0:
        for (tu, ob), cf in prepared:
            B.rdoq(tu, ob, cf)
        loops += 1
    cpu = (time.perf_counter() - t0) / loops
    print("oracle on 1 host core (incl. ctypes call overhead): %.3f ms per %d TUs -> %.3f M TU/s, %.2f M coef/s" % (cpu * 1e3, len(calls), len(calls) / cpu / 1e6, n1 / cpu / 1e6))
except Exception as e:  # the oracle is the checker; its timing is optional here
    print("oracle timing skipped:", e)
