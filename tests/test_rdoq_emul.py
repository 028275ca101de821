"""f1 without a GPU: the __host__ __device__ body of the RDOQ kernel (hm-16.2_b200/csrc/rdoq_impl.cuh), run lane by lane by
tests/rdoq_emul.cpp in the phase order of rdoq.cu, against the calls of the reference's xRateDistOptQuant dumped by the
instrumented reference encoder (tests/golden/rdoq_golden.npz) and against the oracle on random TUs.  This checks the kernel's
LOGIC where there is no GPU; the kernel itself is checked by tests/test_gpu_rdoq.py.  Also: the scan tables of the kernel
against the oracle's (which follow TComRom.cpp's ScanGenerator)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import hmgpu
import rdoqdump
from oracle import binding as B
from test_golden import rdoq_golden_calls

HERE = os.path.dirname(os.path.abspath(__file__))
LANES = {2: 1, 3: 2, 4: 8, 5: 32}          # lanes per TU of rdoq.cu's launch classes


@pytest.fixture(scope="module")
def emul():
    tmp = tempfile.mkdtemp(prefix="rdoq_emul_")
    so = os.path.join(tmp, "librdoq_emul.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, os.path.join(HERE, "rdoq_emul.cpp")])
    L = C.CDLL(so)
    L.rdoq_emul.restype = C.c_int
    L.rdoq_emul.argtypes = [C.c_void_p] * 4 + [C.c_int]
    L.rdoq_emul_scan_table.argtypes = [C.c_void_p]
    L.rdoq_emul_tu.restype = C.c_int
    L.rdoq_emul_tu.argtypes = [C.c_void_p] * 4 + [C.c_int]
    return L


def run_emul(L, job, bits, coef, lanes):
    coef = np.ascontiguousarray(coef, np.int32)
    level = np.full_like(coef, 12345)
    job = np.ascontiguousarray(job)
    bits = np.ascontiguousarray(bits)
    s = L.rdoq_emul(job.ctypes.data, bits.ctypes.data, coef.ctypes.data, level.ctypes.data, lanes)
    # the one-thread-per-TU body must say the same, alone in its warp and beside a longer neighbour
    for ghost in (0, 1):
        level2 = np.full_like(coef, 54321)
        s2 = L.rdoq_emul_tu(job.ctypes.data, bits.ctypes.data, coef.ctypes.data, level2.ctypes.data, ghost)
        assert s2 == s and np.array_equal(level2, level), ("thread-per-TU body differs from the lane-group body", ghost, s2, s)
    return level, s


def test_scan_tables_match_oracle(emul):
    tab = np.zeros(4335, np.uint16)
    emul.rdoq_emul_scan_table(tab.ctypes.data)
    scan_at, cg_at = [0, 16, 80, 336], [0, 1, 5, 21]
    for t in range(3):
        for s in range(4):
            n = 16 << (2 * s)
            scan = np.zeros(n, np.uint16)
            scan_cg = np.zeros(max(n // 16, 1), np.uint16)
            B.oracle().hmo_scan_order(s + 2, t, scan.ctypes.data, scan_cg.ctypes.data)
            assert np.array_equal(tab[t * 1360 + scan_at[s]:t * 1360 + scan_at[s] + n], scan), (t, s)
            assert np.array_equal(tab[4080 + t * 85 + cg_at[s]:4080 + t * 85 + cg_at[s] + n // 16], scan_cg), (t, s)


def test_kernel_body_on_the_reference_encoders_calls(emul):
    calls = rdoq_golden_calls()
    assert len(calls) >= 1000
    for i, c in enumerate(calls):
        job, bits = rdoqdump.to_tu_and_bits(c, hmgpu.RDOQ_JOB, hmgpu.RDOQ_BITS)
        for lanes in {LANES[c["log2"]], 1, 32}:
            level, s = run_emul(emul, job, bits, c["coef"], lanes)
            assert s == c["abs_sum"] and np.array_equal(level, c["level"]), (i, lanes, {k: c[k] for k in rdoqdump.HDR})


def random_tus(rng, n, calls):
    """random TUs in the statistics RDOQ sees (Laplacian coefficients decaying with frequency, now and then dense or huge),
    with the bit estimates / lambda / scale of a dumped call of the same size and channel"""
    out = []
    for k in range(n):
        c = calls[int(rng.integers(len(calls)))]
        nn = c["w"]
        yy, xx = np.mgrid[0:nn, 0:nn]
        mode = k % 4
        scale = [400.0, 3000.0, 60.0, 20000.0][mode] / (1.0 + (xx + yy) * [0.8, 0.15, 1.5, 0.02][mode])
        coef = np.round(rng.laplace(0.0, 1.0, (nn, nn)) * scale).astype(np.int64)
        if mode == 3:
            coef[rng.integers(nn), rng.integers(nn)] = int(rng.choice([32767, -32768]))
        coef = np.clip(coef, -32768, 32767).astype(np.int32).ravel()
        d = dict(c)
        d["coef"] = coef
        d["sign_hide"] = int(rng.integers(2))
        d["scan"] = int(rng.integers(3)) if nn <= 8 else 0
        out.append(d)
    return out


def test_kernel_body_on_random_tus_against_oracle(emul):
    rng = np.random.default_rng(2024)
    calls = rdoq_golden_calls()
    changed = 0
    for i, c in enumerate(random_tus(rng, 1500, calls)):
        tu, obits = rdoqdump.to_tu_and_bits(c, B.RDOQ_TU, B.RDOQ_BITS)
        want, want_sum = B.rdoq(tu, obits, c["coef"])
        job, bits = rdoqdump.to_tu_and_bits(c, hmgpu.RDOQ_JOB, hmgpu.RDOQ_BITS)
        level, s = run_emul(emul, job, bits, c["coef"], LANES[c["log2"]])
        assert s == want_sum and np.array_equal(level, want), (i, {k: c[k] for k in rdoqdump.HDR})
        changed += int(np.abs(want).sum() != want_sum)
    assert changed > 50


def test_golden_batch_is_the_dumped_calls():
    """hm-16.2_b200/rdoq_batch.py (bench.py's rdoq leg, smoke()) builds the same jobs / coefficients / levels as the dump reader"""
    import rdoq_batch
    from test_gpu_rdoq import batch_of
    calls = rdoq_golden_calls()
    jobs1, bits1, coef1 = batch_of(calls)
    jobs, bits, coef, level, abs_sum = rdoq_batch.golden_batch(2)
    n, n1 = len(jobs1), coef1.size
    assert len(jobs) == 2 * n and coef.size == 2 * n1 and np.array_equal(coef[:n1], coef1) and np.array_equal(coef[n1:], coef1)
    for f in jobs1.dtype.names:
        if f not in ("bits_index", "coef_offset"):
            assert np.array_equal(jobs[f][:n], jobs1[f]) and np.array_equal(jobs[f][n:], jobs1[f]), f
    assert np.array_equal(jobs["coef_offset"][:n], jobs1["coef_offset"]) and np.array_equal(jobs["coef_offset"][n:].astype(np.int64), jobs1["coef_offset"].astype(np.int64) + n1)
    for i in range(n):
        assert bits[jobs["bits_index"][i]].tobytes() == bits1[jobs1["bits_index"][i]].tobytes(), i
    assert np.array_equal(level[:n1], np.concatenate([c["level"] for c in calls])) and np.array_equal(abs_sum[:n], [c["abs_sum"] for c in calls])


def test_a_warp_of_different_tus_in_lockstep():
    """rdoq_tu_kernel's warp on the CPU (tests/rdoq_emul_warp.cpp): 32 threads = 32 lanes, real collectives, the device's strides.
    The reference encoder's calls of one size, shuffled, 32 at a time (the last warp not full) -- and every TU comes out as the
    reference returned it, whatever its neighbours were."""
    from test_gpu_rdoq import batch_of
    tmp = tempfile.mkdtemp(prefix="rdoq_warp_")
    so = os.path.join(tmp, "librdoq_emul_warp.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-std=c++17", "-o", so, os.path.join(HERE, "rdoq_emul_warp.cpp")])
    L = C.CDLL(so)
    L.rdoq_emul_warp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    calls = rdoq_golden_calls()
    jobs, bits, coef = batch_of(calls)
    want = np.concatenate([c["level"] for c in calls])
    want_sum = np.array([c["abs_sum"] for c in calls], np.int32)
    level = np.zeros_like(coef)                       # (the kernel's level buffer starts out as zeros; only non-zero levels are stored)
    abs_sum = np.full(len(jobs), -1, np.int32)
    rng = np.random.default_rng(7)
    warps = 0
    for lg in (5, 4, 3, 2):
        idx = rng.permutation(np.flatnonzero(jobs["log2_size"] == lg))
        if lg == 2:
            idx = idx[:200]                           # (enough of the smallest size: a warp of threads costs more than its 32 TUs)
        for a in range(0, len(idx), 32):
            sel = idx[a:a + 32]
            wj = np.ascontiguousarray(jobs[sel])
            ws = np.zeros(32, np.int32)
            L.rdoq_emul_warp(wj.ctypes.data, len(sel), lg, bits.ctypes.data, coef.ctypes.data, level.ctypes.data, ws.ctypes.data)
            abs_sum[sel] = ws[:len(sel)]
            warps += 1
        done = idx
        assert np.array_equal(abs_sum[done], want_sum[done]), lg
        for i in done:
            o, n = int(jobs["coef_offset"][i]), 1 << (2 * lg)
            assert np.array_equal(level[o:o + n], want[o:o + n]), (lg, int(i))
    assert warps >= 20
