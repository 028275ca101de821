#!/usr/bin/env python
"""Golden vectors for the rate-distortion optimised quantiser: calls of TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1974) made
by the INSTRUMENTED REFERENCE ENCODER oracle/_ref/TAppEncoderRdoq (oracle/Makefile target `rdoq`, hooks in oracle/rdoq_dump.inc)
while it encodes short synthetic clips: per call the TU parameters, lambda, the CABAC bit estimates, the coefficients it was
given and the levels it returned.

Run where /root/reference exists:  python tests/golden/make_rdoq_golden.py   -> tests/golden/rdoq_golden.npz
"""
import collections
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rdoqdump  # noqa: E402
import synth  # noqa: E402

ENC = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRdoq")
CFG = os.path.join(ROOT, "oracle", "_ref", "cfg")

# (cfg, width, height, frames, qp, bit depth, extra encoder options, keep every k-th call while dumping)
CASES = [
    ("encoder_lowdelay_P_main.cfg", 176, 144, 3, 30, 8, [], 3),
    ("encoder_lowdelay_P_main.cfg", 176, 144, 2, 18, 8, ["--SignHideFlag=0"], 5),
    ("encoder_randomaccess_main.cfg", 176, 144, 5, 37, 8, ["--DecodingRefreshType=2", "--IntraPeriod=16"], 5),
    ("encoder_randomaccess_main10.cfg", 176, 144, 3, 24, 10, ["--DecodingRefreshType=2", "--IntraPeriod=16"], 5),
    ("encoder_intra_main.cfg", 136, 72, 1, 12, 8, [], 3),
]
PER_CLASS = 6          # calls kept per (case, size, channel, scan, root cbf, transform skip, something coded?)


def dump_case(cfg, w, h, frames, qp, bd, extra, every, tmp, tag):
    yuv = synth.write_yuv(os.path.join(tmp, tag + ".yuv"), w, h, frames, bd, seed=31 + len(tag))
    dump = os.path.join(tmp, tag + ".dump")
    cmd = [ENC, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-q", str(qp),
           "-b", os.path.join(tmp, tag + ".bin"), "-o", os.path.join(tmp, tag + ".rec.yuv")]
    if bd != 8:
        cmd += ["--InputBitDepth=%d" % bd]
    subprocess.run(cmd + extra, check=True, capture_output=True, env=dict(os.environ, HM_RDOQ_DUMP=dump, HM_RDOQ_EVERY=str(every)))
    return rdoqdump.read(dump)


def select(calls):
    """a few calls of every class, the ones with the most coded coefficients first (they exercise the group decisions, the
    last-position search and the sign hiding), then the first ones met"""
    groups = collections.defaultdict(list)
    for c in calls:
        if rdoqdump.supported(c):
            groups[(c["log2"], c["channel"], c["scan"], c["root_cbf"], c["tskip"], c["abs_sum"] > 0)].append(c)
    out = []
    for key in sorted(groups):
        g = groups[key]
        busy = sorted(g, key=lambda c: -int(np.count_nonzero(c["level"])))[:PER_CLASS // 2]
        ids = {id(c) for c in busy}
        out += busy + [c for c in g if id(c) not in ids][:PER_CLASS - len(busy)]
    return out


def pack(calls):
    tables, index = [], {}
    hdr, dd, which, offs, coef, level, sums = [], [], [], [0], [], [], []
    for c in calls:
        key = c["bits"].tobytes()
        if key not in index:
            index[key] = len(tables)
            tables.append(c["bits"])
        which.append(index[key])
        hdr.append([c[k] for k in rdoqdump.HDR])
        dd.append([c["err_scale"], c["lambda"]])
        coef.append(c["coef"])
        level.append(c["level"])
        sums.append(c["abs_sum"])
        offs.append(offs[-1] + len(c["coef"]))
    return {"hdr": np.array(hdr, np.int32), "scale_lambda": np.array(dd, np.float64), "bits_index": np.array(which, np.int32),
            "bits": np.array(tables, np.int32), "offset": np.array(offs, np.int64), "coef": np.concatenate(coef).astype(np.int32),
            "level": np.concatenate(level).astype(np.int32), "abs_sum": np.array(sums, np.int32)}


def unpack(z):
    """the fixture -> list of calls in rdoqdump.read()'s form"""
    calls = []
    for i in range(len(z["abs_sum"])):
        c = {k: int(v) for k, v in zip(rdoqdump.HDR, z["hdr"][i])}
        c["err_scale"], c["lambda"] = float(z["scale_lambda"][i, 0]), float(z["scale_lambda"][i, 1])
        c["bits"] = z["bits"][z["bits_index"][i]]
        a, b = int(z["offset"][i]), int(z["offset"][i + 1])
        c["coef"], c["level"], c["abs_sum"] = z["coef"][a:b], z["level"][a:b], int(z["abs_sum"][i])
        calls.append(c)
    return calls


def main():
    kept = []
    with tempfile.TemporaryDirectory(prefix="hmrdoq_") as tmp:
        for i, (cfg, w, h, frames, qp, bd, extra, every) in enumerate(CASES):
            calls = dump_case(cfg, w, h, frames, qp, bd, extra, every, tmp, "case%d" % i)
            sel = select(calls)
            print("%-34s %dx%d qp %d: %d calls dumped, %d kept" % (cfg, w, h, qp, len(calls), len(sel)))
            kept += sel
    out = os.path.join(ROOT, "tests", "golden", "rdoq_golden.npz")
    np.savez_compressed(out, **pack(kept))
    print("%d calls, %d coded coefficients -> %s (%.0f KB)" % (len(kept), sum(int(np.count_nonzero(c["level"])) for c in kept), out,
                                                                  os.path.getsize(out) / 1024))


if __name__ == "__main__":
    main()
