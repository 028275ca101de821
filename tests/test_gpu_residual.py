"""The reconstruction side of the residual-costing loop: hmgpu_dequant against the calls of the reference
encoder's own xDeQuant (tests/golden/dequant_golden.npz, dumped by the instrumented reference encoder), and hmgpu_residual_tus --
forward transform, RDOQ, dequantisation, inverse transform and the two distortions in one call -- against the same chain made of
the oracle's functions, each of which is pinned to the reference on its own (test_oracle_vs_ref.py, test_golden.py).  Bit-exact."""
import numpy as np
import pytest

import hmgpu
import rdoqdump
from oracle import binding as B
from test_golden import dequant_golden_calls, rdoq_golden_calls


def residual_blocks(rng, count, n, bit_depth):
    """residuals as a prediction leaves them: a smooth error, an edge, noise -- in varying strength"""
    yy, xx = np.mgrid[0:n, 0:n]
    out = np.zeros((count, n, n), np.int64)
    top = (1 << bit_depth) - 1
    for t in range(count):
        amp = [2, 8, 40, 200][t % 4] << (bit_depth - 8)
        out[t] = (amp * np.sin(xx * rng.uniform(0.05, 1.5) + yy * rng.uniform(0.05, 1.5) + rng.uniform(0, 6))
                  + (amp // 2) * (xx > rng.integers(n)) + rng.normal(0, amp / 4 + 0.5, (n, n))).round()
    return np.clip(out, -top, top).astype(np.int16)


def jobs_for(calls, count, n):
    """RDOQ jobs + bit-estimate sets for `count` TUs of n x n out of dumped calls of that size"""
    jobs = np.zeros(count, hmgpu.RDOQ_JOB)
    sets, index = [], {}
    for i in range(count):
        c = calls[i % len(calls)]
        job, bits = rdoqdump.to_tu_and_bits(c, hmgpu.RDOQ_JOB, hmgpu.RDOQ_BITS)
        key = bits.tobytes()
        if key not in index:
            index[key] = len(sets)
            sets.append(bits)
        jobs[i] = job
        jobs[i]["bits_index"], jobs[i]["coef_offset"] = index[key], i * n * n
    return jobs, np.array(sets, hmgpu.RDOQ_BITS)


def oracle_chain(resi, n, jobs, bits, bit_depth, use_dst):
    """transformNxN -> invTransformNxN -> distortions with the oracle's functions, one TU at a time"""
    O = B.oracle()
    log2 = n.bit_length() - 1
    level = np.zeros(resi.shape, np.int32)
    abs_sum = np.zeros(len(resi), np.int32)
    rec = np.zeros(resi.shape, np.int16)
    dist = np.zeros((len(resi), 2), np.uint32)
    zero = np.zeros((n, n), np.int16)
    for t in range(len(resi)):
        coef = np.zeros((n, n), np.int32)
        O.hmo_fwd_transform(bit_depth, np.ascontiguousarray(resi[t], np.int32), coef, n, n, int(use_dst))
        tu = np.zeros(1, B.RDOQ_TU)
        for f in ("log2_size", "channel", "scan", "qbits", "qp_per", "qp_rem", "go_rice_init", "cbf_bits", "bit_depth", "err_scale", "lambda"):
            tu[f] = jobs[f][t]
        tu["sign_hide"] = jobs["flags"][t]
        ob = np.zeros(1, B.RDOQ_BITS)
        for f in ob.dtype.names:
            ob[f] = bits[f][jobs["bits_index"][t]]
        lv, s = B.rdoq(tu, ob, coef)
        level[t], abs_sum[t] = lv.reshape(n, n), s
        deq = B.dequant(lv, log2, int(jobs["qp_per"][t]), int(jobs["qp_rem"][t]), bit_depth)
        back = np.zeros((n, n), np.int32)
        O.hmo_inv_transform(bit_depth, np.ascontiguousarray(deq.reshape(n, n)), back, n, int(use_dst))
        rec[t] = back.astype(np.int16)
        org = np.ascontiguousarray(resi[t])
        dist[t, 0] = O.hmo_sse(B.ptr(org), n, B.ptr(rec[t]), n, n, n, bit_depth)
        dist[t, 1] = O.hmo_sse(B.ptr(org), n, B.ptr(zero), n, n, n, bit_depth)
    return level, abs_sum, rec, dist


def calls_of(n, bit_depth):
    return [c for c in rdoq_golden_calls() if c["w"] == n and c["bit_depth"] == bit_depth]


def test_oracle_chain_reconstructs_what_it_codes():
    """(no GPU) the chain the GPU test compares with: nothing coded -> zero reconstruction and equal distortions; coding pays"""
    rng = np.random.default_rng(5)
    for n in (4, 16):
        calls = calls_of(n, 8)
        resi = residual_blocks(rng, 24, n, 8)
        jobs, bits = jobs_for(calls, len(resi), n)
        level, abs_sum, rec, dist = oracle_chain(resi, n, jobs, bits, 8, n == 4)
        empty = abs_sum == 0
        assert empty.any() and (~empty).any()
        assert not rec[empty].any() and np.array_equal(dist[empty, 0], dist[empty, 1])
        assert np.array_equal(np.abs(level).sum(axis=(1, 2)) > 0, ~empty)
        assert dist[~empty, 0].astype(np.int64).sum() < dist[~empty, 1].astype(np.int64).sum()


@pytest.mark.gpu
def test_dequant_matches_the_reference_encoders_calls():
    calls = dequant_golden_calls()
    assert len(calls) >= 300 and {c["log2"] for c in calls} == {2, 3, 4, 5} and {c["bit_depth"] for c in calls} == {8, 10}
    groups = {}
    for c in calls:
        groups.setdefault((c["bit_depth"], c["w"], c["per"], c["rem"]), []).append(c)
    for bit_depth in (8, 10):
        with hmgpu.Context(64, 64, bit_depth, 1) as ctx:
            for (bd, n, per, rem), g in sorted(groups.items()):
                if bd != bit_depth:
                    continue
                got = ctx.dequant(np.stack([c["level"].reshape(n, n) for c in g]), n, per, rem)
                for k, c in enumerate(g):
                    assert np.array_equal(got[k].ravel(), c["coef"]), (bd, n, per, rem, k)
            # the clips at both ends, against the oracle (the encoder's own levels never reach them)
            lv = np.array([[32767, -32768, 40000, -40000, 1, -1, 0, 12345] * 2], np.int32).reshape(1, 4, 4)
            for per in (0, 3, 9, 12):
                assert np.array_equal(ctx.dequant(lv, 4, per, 5).ravel(), B.dequant(lv, 2, per, 5, bit_depth)), (bit_depth, per)
            with pytest.raises(hmgpu.HmGpuError):
                ctx.dequant(lv, 4, 13, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("bit_depth", [8, 10])
def test_residual_tus_matches_the_oracle_chain(bit_depth):
    rng = np.random.default_rng(40 + bit_depth)
    with hmgpu.Context(64, 64, bit_depth, 1) as ctx:
        for n in (4, 8, 16, 32):
            calls = calls_of(n, bit_depth)
            assert calls, (n, bit_depth)
            for use_dst in ((False, True) if n == 4 else (False,)):
                resi = residual_blocks(rng, 70 if n < 32 else 37, n, bit_depth)
                jobs, bits = jobs_for(calls, len(resi), n)
                for mapping in (1, 0):
                    ctx.set_option("rdoq_tu", mapping)
                    n0 = ctx.launches
                    level, abs_sum, rec, dist = ctx.residual_tus(resi, n, jobs, bits, use_dst)
                    assert ctx.launches - n0 == 5                  # transform, RDOQ, dequantiser, inverse transform, distortions
                    e_level, e_sum, e_rec, e_dist = oracle_chain(resi, n, jobs, bits, bit_depth, use_dst)
                    assert np.array_equal(abs_sum, e_sum) and np.array_equal(level, e_level), (n, use_dst, mapping)
                    assert np.array_equal(rec, e_rec) and np.array_equal(dist, e_dist), (n, use_dst, mapping)
                    assert (abs_sum > 0).any()
        # a job that does not describe its TU is refused
        resi = residual_blocks(rng, 3, 8, bit_depth)
        jobs, bits = jobs_for(calls_of(8, bit_depth), 3, 8)
        for field, value in (("coef_offset", 1), ("log2_size", 4), ("bit_depth", bit_depth + 2)):
            broken = jobs.copy()
            broken[field][2] = value
            with pytest.raises(hmgpu.HmGpuError):
                ctx.residual_tus(resi, 8, broken, bits)
