"""f4: intra mode pre-selection on the device, hmgpu_intra_costs, against the oracle (which
tests/test_oracle_vs_ref.py pins to the reference's xPredIntraAng / xPredIntraPlanar / xDCPredFiltering)."""
import numpy as np
import pytest

import hmgpu
from oracle import binding as B

pytestmark = pytest.mark.gpu


def _smooth(line):
    """the [1 2 1] smoothing of initAdiPatternChType (TComPattern.cpp:296-325), ends copied"""
    f = line.astype(np.int32).copy()
    f[1:-1] = (line[:-2].astype(np.int32) + 2 * line[1:-1] + line[2:] + 2) >> 2
    return f.astype(np.int16)


def _jobs(rng, bd, count):
    mx = (1 << bd) - 1
    jobs = np.zeros(count, hmgpu.INTRA_JOB)
    orgs, lines = [], []
    o_off = l_off = 0
    for i in range(count):
        n = int(rng.choice([4, 8, 16, 32, 64], p=[0.3, 0.3, 0.2, 0.15, 0.05]))
        kind = i % 3
        if kind == 0:
            line = rng.integers(0, mx + 1, 4 * n + 1)
        elif kind == 1:
            line = np.clip(np.linspace(mx // 5, 4 * mx // 5, 4 * n + 1) + rng.integers(-4, 5, 4 * n + 1), 0, mx)
        else:
            line = rng.choice(np.array([0, mx]), 4 * n + 1)
        line = line.astype(np.int16)
        org = np.clip(rng.integers(0, mx + 1, (n, n)) if kind != 1 else line[2 * n + 1:3 * n + 1][None, :] + rng.integers(-6, 7, (n, n)), 0, mx).astype(np.int16)
        flags = int(rng.integers(0, 32))
        if i % 4:
            flags |= hmgpu.IF_ABOVE | hmgpu.IF_LEFT | hmgpu.IF_EDGE_FILTERS | hmgpu.IF_SATD    # the common case
            flags &= ~hmgpu.IF_NO_SMOOTH
        jobs[i] = (o_off, l_off, n, flags, 0)
        orgs.append(org.reshape(-1)); lines += [line, _smooth(line)]
        o_off += n * n; l_off += 2 * (4 * n + 1)
    return jobs, np.concatenate(orgs), np.concatenate(lines)


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_intra_costs_match_oracle(bit_depth):
    rng = np.random.default_rng(70 + bit_depth)
    jobs, orgs, lines = _jobs(rng, bit_depth, 400)
    O = B.oracle()
    exp = np.zeros((len(jobs), 35), np.uint32)
    for i, j in enumerate(jobs):
        n = int(j["size"])
        lo = int(j["ref_offset"])
        row = np.zeros(35, np.uint32)
        O.hmo_intra_costs(B.ptr(lines, lo), B.ptr(lines, lo + 4 * n + 1), B.ptr(orgs, int(j["org_offset"])), n, bit_depth, int(j["flags"]), row)
        exp[i] = row
    with hmgpu.Context(64, 64, bit_depth, 1) as ctx:
        got = ctx.intra_costs(jobs, orgs, lines)
        bad = np.argwhere(got != exp)
        assert bad.size == 0, "job %d mode %d: %d vs %d (size %d flags %d)" % (bad[0][0], bad[0][1], got[tuple(bad[0])], exp[tuple(bad[0])],
                                                                            jobs[bad[0][0]]["size"], jobs[bad[0][0]]["flags"])
        # errors are reported, not thrown away
        broken = jobs[:1].copy(); broken["size"] = 12
        with pytest.raises(hmgpu.HmGpuError):
            ctx.intra_costs(broken, orgs, lines)
        broken = jobs[:1].copy(); broken["ref_offset"] = lines.size
        with pytest.raises(hmgpu.HmGpuError):
            ctx.intra_costs(broken, orgs, lines)
