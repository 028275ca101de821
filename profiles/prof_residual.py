"""hmgpu_residual_tus (forward transform, RDOQ, dequantiser, inverse transform, distortions of a batch of TUs in
one call) at production batch sizes: device ms per stage (stage timers = CUDA events on the context's stream) and the wall time of
the blocking C-ABI call with host buffers.  Residuals are synthetic; the quantiser parameters and bit estimates are those of the
reference encoder's own calls (tests/golden/rdoq_golden.npz).
usage: python profiles/prof_residual.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import hmgpu  # noqa: E402
from test_gpu_residual import calls_of, jobs_for, residual_blocks  # noqa: E402

rng = np.random.default_rng(1)
with hmgpu.Context(64, 64, 8, 1) as ctx:
    for n, count in ((32, 4000), (16, 16000), (8, 64000), (4, 128000)):
        base = residual_blocks(rng, 64, n, 8)
        resi = np.tile(base, (count // 64, 1, 1))
        count = len(resi)
        jobs, bits = jobs_for(calls_of(n, 8), count, n)
        ctx.residual_tus(resi, n, jobs, bits)
        ctx.profile_enable(True)
        ctx.profile_read(True)
        t0 = time.perf_counter()
        level, abs_sum, rec, dist = ctx.residual_tus(resi, n, jobs, bits)
        wall = time.perf_counter() - t0
        st = {k: v for k, v in ctx.profile_read(True).items() if v[1]}
        ctx.profile_enable(False)
        dev = sum(v[0] for v in st.values())
        print("%2dx%-2d x %6d TUs (%d coded): device %.3f ms (%s) -> %.1f M samples/s;  C-ABI call %.2f ms"
              % (n, n, count, int((abs_sum > 0).sum()), dev, ", ".join("%s %.3f" % (k, v[0]) for k, v in st.items()), count * n * n / dev / 1e3, wall * 1e3))
