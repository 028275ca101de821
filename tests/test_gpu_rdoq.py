"""f1: rate-distortion optimised quantisation on the device, hmgpu_rdoq, against (1) the calls of the reference's
own TComTrQuant::xRateDistOptQuant dumped by the instrumented reference encoder (tests/golden/rdoq_golden.npz: inputs, CABAC bit
estimates, returned levels) -- all of them in ONE batch, so TUs of every size, both channels, many coder states and QPs share
the launches -- and (2) the oracle (pinned to the same calls) on random TUs.  Levels and uiAbsSum are bit-exact."""
import numpy as np
import pytest

import hmgpu
import rdoqdump
from oracle import binding as B
from test_golden import rdoq_golden_calls
from test_rdoq_emul import random_tus

pytestmark = pytest.mark.gpu


def batch_of(calls):
    """dumped calls -> (jobs, bit-estimate sets, coefficients): distinct coder states become distinct sets"""
    jobs = np.zeros(len(calls), hmgpu.RDOQ_JOB)
    sets, index = [], {}
    at = 0
    for i, c in enumerate(calls):
        job, bits = rdoqdump.to_tu_and_bits(c, hmgpu.RDOQ_JOB, hmgpu.RDOQ_BITS)
        key = bits.tobytes()
        if key not in index:
            index[key] = len(sets)
            sets.append(bits)
        jobs[i] = job
        jobs[i]["bits_index"], jobs[i]["coef_offset"] = index[key], at
        at += c["coef"].size
    return jobs, np.array(sets, hmgpu.RDOQ_BITS), np.concatenate([c["coef"] for c in calls]).astype(np.int32)


def check(ctx, calls, want_level, want_sum):
    jobs, bits, coef = batch_of(calls)
    level, abs_sum = ctx.rdoq(jobs, bits, coef)
    bad = np.flatnonzero(abs_sum != want_sum)
    assert bad.size == 0, "TU %d: uiAbsSum %d vs %d (%s)" % (bad[0], abs_sum[bad[0]], want_sum[bad[0]], {k: calls[bad[0]][k] for k in rdoqdump.HDR})
    for i, c in enumerate(calls):
        a = int(jobs[i]["coef_offset"])
        assert np.array_equal(level[a:a + c["coef"].size], want_level[i]), (i, {k: c[k] for k in rdoqdump.HDR})
    return jobs, bits, coef


# both mappings of the kernel: one thread per TU (32 TUs of a size per warp, one launch; the default) and one lane group per TU
MAPPINGS = pytest.mark.parametrize("per_thread", [1, 0])


@MAPPINGS
def test_rdoq_matches_the_reference_encoders_calls(per_thread):
    calls = rdoq_golden_calls()
    assert len(calls) >= 1000 and {c["log2"] for c in calls} == {2, 3, 4, 5}
    with hmgpu.Context(64, 64, 8, 1) as ctx:
        ctx.set_option("rdoq_tu", per_thread)
        n0 = ctx.launches
        jobs, bits, coef = check(ctx, calls, [c["level"] for c in calls], np.array([c["abs_sum"] for c in calls]))
        assert ctx.launches - n0 == (1 if per_thread else 4)          # one launch for all sizes / one per TU size
        assert len(bits) > 10
        # one TU at a time, and each size class alone (ragged batches: the last warp / lane group is partly empty)
        for n in (1, 3, 33):
            check(ctx, calls[:n], [c["level"] for c in calls[:n]], np.array([c["abs_sum"] for c in calls[:n]]))
        for lg in (2, 3, 4, 5):
            sub = [c for c in calls if c["log2"] == lg][:37]
            check(ctx, sub, [c["level"] for c in sub], np.array([c["abs_sum"] for c in sub]))
        # errors are reported, not thrown away
        for field, value in (("log2_size", 6), ("scan", 3), ("bits_index", len(bits)), ("coef_offset", coef.size), ("qp_rem", 6), ("lambda", 0.0)):
            broken = jobs[:2].copy()
            broken[field][1] = value
            with pytest.raises(hmgpu.HmGpuError):
                ctx.rdoq(broken, bits, coef)


@MAPPINGS
def test_rdoq_matches_oracle_on_random_tus(per_thread):
    rng = np.random.default_rng(99)
    calls = random_tus(rng, 3000, rdoq_golden_calls())
    want_level, want_sum = [], []
    for c in calls:
        tu, obits = rdoqdump.to_tu_and_bits(c, B.RDOQ_TU, B.RDOQ_BITS)
        lv, s = B.rdoq(tu, obits, c["coef"])
        want_level.append(lv); want_sum.append(s)
    with hmgpu.Context(64, 64, 8, 1) as ctx:
        ctx.set_option("rdoq_tu", per_thread)
        check(ctx, calls, want_level, np.array(want_sum))
        # the same TUs in another order: other neighbours in every warp, the same levels
        order = rng.permutation(len(calls))
        check(ctx, [calls[i] for i in order], [want_level[i] for i in order], np.array(want_sum)[order])


@MAPPINGS
def test_rdoq_all_zero_and_uncovered_coefficients(per_thread):
    """TUs that quantise to nothing return zero levels and uiAbsSum 0; coefficients no job covers come back as zeros"""
    calls = [dict(c) for c in rdoq_golden_calls()[:40]]
    for c in calls:
        c["coef"] = (c["coef"] % 3 - 1).astype(np.int32)       # magnitudes far below one quantiser step
    jobs, bits, coef = batch_of(calls)
    coef = np.concatenate([coef, np.full(100, 7777, np.int32)])
    with hmgpu.Context(64, 64, 8, 1) as ctx:
        ctx.set_option("rdoq_tu", per_thread)
        level, abs_sum = ctx.rdoq(jobs, bits, coef)
    assert not level.any() and not abs_sum.any()
