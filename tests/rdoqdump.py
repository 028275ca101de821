"""Reader of the RDOQ dumps written by the instrumented reference encoder (oracle/rdoq_dump.inc; test infrastructure).
-> list of calls of TComTrQuant::xRateDistOptQuant, each a dict: the TU's geometry / scan / quantiser parameters ('hdr' fields by
name), err_scale, lambda, the CABAC bit estimates ('bits', the raw estBitsSbacStruct as int32), the coefficients it was given
('coef') and the levels / absolute sum it returned ('level', 'abs_sum')."""
import numpy as np

HDR = ("w", "h", "log2", "channel", "comp", "scan", "first_sig_ctx", "qbits", "per", "rem", "cbf_bits0", "cbf_bits1", "sign_hide",
       "go_rice_init", "go_rice_adapt", "bit_depth", "ext_precision", "scaling_lists", "tskip", "intra", "root_cbf", "bits_bytes",
       "max_dyn_range", "adapt_qp_select")

# estBitsSbacStruct (TComTrQuant.h:59-73) as int32 words: name -> (first word, shape)
BITS_LAYOUT = {"sig_group": (0, (2, 2)), "sig": (4, (44, 2)), "last_x": (92, (2, 10)), "last_y": (112, (2, 10)),
               "greater_one": (132, (24, 2)), "level_abs": (180, (6, 2)), "cbf": (192, (10, 2)), "root_cbf": (212, (4, 2)),
               "go_rice_stats": (220, (4,))}
BITS_WORDS = 224


def read(path):
    data = open(path, "rb").read()
    pos, calls = 0, []
    while pos < len(data):
        if data[pos:pos + 1] != b"R":
            raise ValueError("bad record tag at %d" % pos)
        pos += 1
        hdr = np.frombuffer(data, np.int32, 24, pos)
        pos += 96
        c = {k: int(v) for k, v in zip(HDR, hdr)}
        dd = np.frombuffer(data, np.float64, 2, pos)
        pos += 16
        c["err_scale"], c["lambda"] = float(dd[0]), float(dd[1])
        assert c["bits_bytes"] == 4 * BITS_WORDS, c["bits_bytes"]
        c["bits"] = np.frombuffer(data, np.int32, BITS_WORDS, pos).copy()
        pos += c["bits_bytes"]
        n = c["w"] * c["h"]
        c["coef"] = np.frombuffer(data, np.int32, n, pos).copy()
        pos += 4 * n
        c["level"] = np.frombuffer(data, np.int32, n, pos).copy()
        pos += 4 * n
        c["abs_sum"] = int(np.frombuffer(data, np.int32, 1, pos)[0])
        pos += 4
        calls.append(c)
    return calls


def bits_table(words):
    """the estBitsSbacStruct words -> dict of the tables RDOQ reads"""
    return {k: np.asarray(words[o:o + int(np.prod(sh))]).reshape(sh) for k, (o, sh) in BITS_LAYOUT.items()}


def to_tu_and_bits(c, tu_dtype, bits_dtype):
    """a dumped call -> (TU record, bit-estimate record) in the layouts of oracle/hm_oracle.h (binding.RDOQ_TU / RDOQ_BITS) or of
    include/hmgpu.h (hmgpu.RDOQ_JOB / RDOQ_BITS): both name their fields alike"""
    t = bits_table(c["bits"])
    bits = np.zeros((), bits_dtype)
    for k in ("sig_group", "sig", "last_x", "last_y", "greater_one", "level_abs"):
        bits[k] = t[k]
    tu = np.zeros((), tu_dtype)
    tu["log2_size"], tu["channel"], tu["scan"] = c["log2"], c["channel"], c["scan"]
    tu["qbits"], tu["qp_per"], tu["qp_rem"], tu["go_rice_init"] = c["qbits"], c["per"], c["rem"], c["go_rice_init"]
    tu["cbf_bits"] = (c["cbf_bits0"], c["cbf_bits1"])
    tu["err_scale"], tu["lambda"] = c["err_scale"], c["lambda"]
    if "sign_hide" in tu_dtype.names:
        tu["sign_hide"], tu["bit_depth"] = c["sign_hide"], c["bit_depth"]
    else:
        tu["flags"], tu["bit_depth"] = c["sign_hide"], c["bit_depth"]
    return tu, bits


def supported(c):
    """the configurations hmo_rdoq / hmgpu_rdoq restate (every call of the BASELINE cfgs)"""
    return (c["w"] == c["h"] and not c["scaling_lists"] and not c["ext_precision"] and not c["go_rice_adapt"]
            and c["max_dyn_range"] == 15 and not c["adapt_qp_select"])


# ---- xDeQuant dumps (oracle/rdoq_dump.inc, hm_deq_after) ----------------------------------------------------------------------
DEQ_HDR = ("w", "h", "log2", "channel", "per", "rem", "bit_depth", "max_dyn_range", "scaling_lists", "tskip", "ext_precision", "spare")


def read_dequant(path):
    """-> list of calls of TComTrQuant::xDeQuant: header fields by name, 'level' (what it was given), 'coef' (what it produced)"""
    data = open(path, "rb").read()
    pos, calls = 0, []
    while pos < len(data):
        if data[pos:pos + 1] != b"D":
            raise ValueError("bad record tag at %d" % pos)
        pos += 1
        c = {k: int(v) for k, v in zip(DEQ_HDR, np.frombuffer(data, np.int32, 12, pos))}
        pos += 48
        n = c["w"] * c["h"]
        c["level"] = np.frombuffer(data, np.int32, n, pos).copy()
        pos += 4 * n
        c["coef"] = np.frombuffer(data, np.int32, n, pos).copy()
        pos += 4 * n
        calls.append(c)
    return calls


def dequant_supported(c):
    """what hmo_dequant / hmgpu_dequant restate: square TUs, flat quantiser, no extended precision (transform-skipped TUs are
    dequantised by the same arithmetic as long as extended precision is off)"""
    return c["w"] == c["h"] and not c["scaling_lists"] and not c["ext_precision"] and c["max_dyn_range"] == 15
