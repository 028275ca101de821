"""A batch of TUs for hmgpu_rdoq out of the calls of the reference encoder's xRateDistOptQuant that tests/golden/rdoq_golden.npz
holds (dumped by the instrumented reference encoder, tests/golden/make_rdoq_golden.py): TUs of every size, luma and chroma, many
coder states, with the levels the reference returned.  Used by bench.py's rdoq leg and by __graft_entry__.smoke()."""
import os

import numpy as np

import hmgpu

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rdoq_golden.npz")
# columns of the dump's header (tests/rdoqdump.py HDR)
_COL = {"log2": 2, "channel": 3, "scan": 5, "qbits": 7, "per": 8, "rem": 9, "cbf_bits0": 10, "cbf_bits1": 11, "sign_hide": 12, "go_rice_init": 13, "bit_depth": 15}
# estBitsSbacStruct words -> hmgpu_rdoq_bits fields (tests/rdoqdump.py BITS_LAYOUT)
_BITS = {"sig_group": (0, (2, 2)), "sig": (4, (44, 2)), "last_x": (92, (2, 10)), "last_y": (112, (2, 10)), "greater_one": (132, (24, 2)), "level_abs": (180, (6, 2))}


def golden_batch(rep=1, path=GOLDEN):
    """-> (jobs, bits, coef, levels the reference returned, uiAbsSum the reference returned), the dumped calls `rep` times over"""
    z = np.load(path)
    hdr, n = z["hdr"], len(z["abs_sum"])
    jobs = np.zeros(n, hmgpu.RDOQ_JOB)
    jobs["log2_size"], jobs["channel"], jobs["scan"] = hdr[:, _COL["log2"]], hdr[:, _COL["channel"]], hdr[:, _COL["scan"]]
    jobs["flags"], jobs["qbits"], jobs["qp_per"], jobs["qp_rem"] = hdr[:, _COL["sign_hide"]], hdr[:, _COL["qbits"]], hdr[:, _COL["per"]], hdr[:, _COL["rem"]]
    jobs["go_rice_init"], jobs["bit_depth"] = hdr[:, _COL["go_rice_init"]], hdr[:, _COL["bit_depth"]]
    jobs["cbf_bits"] = hdr[:, [_COL["cbf_bits0"], _COL["cbf_bits1"]]]
    jobs["err_scale"], jobs["lambda"] = z["scale_lambda"][:, 0], z["scale_lambda"][:, 1]
    jobs["bits_index"], jobs["coef_offset"] = z["bits_index"], z["offset"][:-1]
    bits = np.zeros(len(z["bits"]), hmgpu.RDOQ_BITS)
    for name, (first, shape) in _BITS.items():
        bits[name] = z["bits"][:, first:first + int(np.prod(shape))].reshape((-1,) + shape)
    coef, level, n1 = z["coef"].astype(np.int32), z["level"].astype(np.int32), int(z["offset"][-1])
    if rep > 1:
        jobs = np.tile(jobs, rep)
        jobs["coef_offset"] = (jobs["coef_offset"].astype(np.int64) + np.repeat(np.arange(rep, dtype=np.int64) * n1, n)).astype(np.uint32)
        coef, level = np.tile(coef, rep), np.tile(level, rep)
    return jobs, bits, coef, level, np.tile(z["abs_sum"].astype(np.int32), rep)
