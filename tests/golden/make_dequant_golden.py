#!/usr/bin/env python
"""Golden vectors for the dequantiser: calls of TComTrQuant::xDeQuant (TComTrQuant.cpp:1203-1313) made by the INSTRUMENTED
REFERENCE ENCODER oracle/_ref/TAppEncoderRdoq (oracle/Makefile target `rdoq`, hook hm_deq_after in oracle/rdoq_dump.inc) while it
encodes short synthetic clips: per call the TU size, QP per / rem, bit depth, the levels it was given and the coefficients it
produced.

Run where /root/reference exists:  python tests/golden/make_dequant_golden.py   -> tests/golden/dequant_golden.npz
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
# (cfg, width, height, frames, qp, bit depth, keep every k-th call): QPs chosen so that both the right-shift and the left-shift
# branch of the dequantiser are met at every TU size
CASES = [
    ("encoder_lowdelay_P_main.cfg", 176, 144, 2, 4, 8, 7),
    ("encoder_lowdelay_P_main.cfg", 176, 144, 2, 22, 8, 7),
    ("encoder_randomaccess_main.cfg", 176, 144, 3, 37, 8, 5),
    ("encoder_randomaccess_main10.cfg", 176, 144, 2, 10, 10, 7),
    ("encoder_randomaccess_main10.cfg", 176, 144, 2, 45, 10, 3),
    ("encoder_intra_main.cfg", 136, 72, 1, 51, 8, 1),
]
PER_CLASS = 4          # calls kept per (case, size, channel, per, rem, transform skip)


def dump_case(cfg, w, h, frames, qp, bd, every, tmp, tag):
    yuv = synth.write_yuv(os.path.join(tmp, tag + ".yuv"), w, h, frames, bd, seed=77 + len(tag))
    dump = os.path.join(tmp, tag + ".dump")
    cmd = [ENC, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-q", str(qp),
           "-b", os.path.join(tmp, tag + ".bin"), "-o", os.path.join(tmp, tag + ".rec.yuv")]
    if bd != 8:
        cmd += ["--InputBitDepth=%d" % bd]
    subprocess.run(cmd, check=True, capture_output=True, env=dict(os.environ, HM_DEQ_DUMP=dump, HM_DEQ_EVERY=str(every)))
    return rdoqdump.read_dequant(dump)


def main():
    kept = []
    with tempfile.TemporaryDirectory(prefix="hmdeq_") as tmp:
        for i, (cfg, w, h, frames, qp, bd, every) in enumerate(CASES):
            calls = [c for c in dump_case(cfg, w, h, frames, qp, bd, every, tmp, "case%d" % i) if rdoqdump.dequant_supported(c)]
            groups = collections.defaultdict(list)
            for c in calls:
                groups[(c["log2"], c["channel"], c["per"], c["rem"], c["tskip"])].append(c)
            sel = []
            for key in sorted(groups):       # the busiest calls of every class first
                sel += sorted(groups[key], key=lambda c: -int(np.abs(c["level"]).sum()))[:PER_CLASS]
            print("%-34s qp %d %d-bit: %d calls dumped, %d kept" % (cfg, qp, bd, len(calls), len(sel)))
            kept += sel
    offs = np.concatenate([[0], np.cumsum([c["level"].size for c in kept])]).astype(np.int64)
    out = os.path.join(ROOT, "tests", "golden", "dequant_golden.npz")
    np.savez_compressed(out, hdr=np.array([[c[k] for k in rdoqdump.DEQ_HDR] for c in kept], np.int32), offset=offs,
                        level=np.concatenate([c["level"] for c in kept]).astype(np.int32),
                        coef=np.concatenate([c["coef"] for c in kept]).astype(np.int32))
    print("%d calls -> %s (%.0f KB)" % (len(kept), out, os.path.getsize(out) / 1024))


if __name__ == "__main__":
    main()
