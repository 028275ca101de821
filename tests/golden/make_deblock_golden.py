#!/usr/bin/env python
"""Golden vectors for the deblocking filter: pictures before / after TComLoopFilter::loopFilterPic, with the per-unit data the
filter saw (boundary strengths, QPs, no-filter flags), dumped by the INSTRUMENTED REFERENCE DECODER oracle/_ref/TAppDecoderDbk
(oracle/Makefile target `dbk`, hooks in oracle/dbk_dump.inc) while it decodes streams made by the unmodified reference encoder.

Run where /root/reference exists:  python tests/golden/make_deblock_golden.py   -> tests/golden/deblock_golden.npz
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dbkdump  # noqa: E402
import synth  # noqa: E402

ENC = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRef")
DEC = os.path.join(ROOT, "oracle", "_ref", "TAppDecoderDbk")
CFG = os.path.join(ROOT, "oracle", "_ref", "cfg")

# (cfg, width, height, frames, qp, bit depth, extra encoder options, pictures kept)
CASES = [
    ("encoder_lowdelay_P_main.cfg", 176, 144, 3, 37, 8, [], [0, 2]),
    ("encoder_lowdelay_P_main.cfg", 200, 120, 3, 24, 8, ["--LoopFilterOffsetInPPS=1", "--LoopFilterBetaOffset_div2=2", "--LoopFilterTcOffset_div2=-1",
                                                          "--CbQpOffset=4", "--CrQpOffset=-3"], [1, 2]),
    ("encoder_randomaccess_main.cfg", 176, 144, 9, 32, 8, ["--DecodingRefreshType=2", "--IntraPeriod=16"], [4, 8]),
    ("encoder_randomaccess_main10.cfg", 176, 144, 9, 30, 10, ["--DecodingRefreshType=2", "--IntraPeriod=16"], [0, 3]),
    ("encoder_intra_main.cfg", 136, 72, 1, 45, 8, [], [0]),
]


def dump_case(cfg, w, h, frames, qp, bd, extra, tmp, tag):
    yuv = synth.write_yuv(os.path.join(tmp, tag + ".yuv"), w, h, frames, bd, seed=77 + len(tag))
    bits = os.path.join(tmp, tag + ".bin")
    cmd = [ENC, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-q", str(qp), "-b", bits]
    if bd != 8:
        cmd += ["--InputBitDepth=%d" % bd]
    subprocess.run(cmd + extra, check=True, capture_output=True)
    dump = os.path.join(tmp, tag + ".dump")
    subprocess.run([DEC, "-b", bits], check=True, capture_output=True, env=dict(os.environ, HM_DBK_DUMP=dump))
    return dbkdump.read(dump)


def pack(pic):
    d = {k: np.asarray(pic[k]) for k in ("bs_ver", "bs_hor", "qp", "nofilter")}
    for i, c in enumerate(("y", "cb", "cr")):
        d["pre_" + c] = pic["pre"][i]
        d["post_" + c] = pic["post"][i]
    d["params"] = np.array([pic[k] for k in ("w", "h", "bd_luma", "bd_chroma", "beta_offset_div2", "tc_offset_div2", "cb_qp_offset", "cr_qp_offset", "poc")], np.int32)
    return d


def main():
    out = {}
    n = 0
    with tempfile.TemporaryDirectory(prefix="hmdbk_") as tmp:
        for ci, (cfg, w, h, frames, qp, bd, extra, keep) in enumerate(CASES):
            pics = dump_case(cfg, w, h, frames, qp, bd, extra, tmp, "c%d" % ci)
            for k in keep:                      # decoding order
                for name, v in pack(pics[k]).items():
                    out["p%d_%s" % (n, name)] = v
                n += 1
    out["n_pictures"] = np.array([n], np.int32)
    path = os.path.join(ROOT, "tests", "golden", "deblock_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", n, "pictures")


if __name__ == "__main__":
    main()
