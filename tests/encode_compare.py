#!/usr/bin/env python
"""Encode the same synthetic clip with CPU HM (oracle/_ref/TAppEncoderRef, the unmodified
reference) and with the GPUME encoder (hm-16.2_b200/host/build/TAppEncoderGpu --GPUME=1|2),
compare bitstream and reconstruction MD5, print timings.

usage: encode_compare.py [--cfg lowdelay_P_main] [--size 416x240] [--frames 4] [--qp 32]
                         [--gpume 1] [--skip-cpu] [--bit-depth 8] [-- extra HM options...]
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "hm-16.2_b200")
sys.path.insert(0, HERE)
import synth  # noqa: E402

REF_ENC = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRef")
GPU_ENC = os.path.join(HERE, "host", "build", "TAppEncoderGpu")
CFG_DIR = os.path.join(ROOT, "oracle", "_ref", "cfg")


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run(enc, cfg, yuv, w, h, frames, qp, out_prefix, extra, bit_depth=8, env=None):
    cmd = [enc, "-c", cfg, "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-q", str(qp),
           "-b", out_prefix + ".bin", "-o", out_prefix + ".yuv"]
    if bit_depth != 8:
        cmd += ["--InputBitDepth=%d" % bit_depth]
    cmd += extra
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        sys.stderr.write(p.stdout[-2000:] + p.stderr[-2000:])
        raise SystemExit("encoder failed: %s (rc %d)" % (" ".join(cmd), p.returncode))
    m = re.search(r"Total Time:\s+([0-9.]+) sec", p.stdout)
    stats = [ln for ln in p.stderr.splitlines() if ln.startswith("[GPUME]")]
    return {"wall_s": wall, "cpu_total_time_s": float(m.group(1)) if m else None, "bitstream_md5": md5(out_prefix + ".bin"),
            "recon_md5": md5(out_prefix + ".yuv"), "bytes": os.path.getsize(out_prefix + ".bin"), "gpume": stats}


def available():
    return all(os.path.exists(p) for p in (REF_ENC, GPU_ENC, CFG_DIR))


def compare(cfg_name, size, frames, qp, gpume=1, extra=(), bit_depth=8, skip_cpu=False):
    """encode the synthetic clip with CPU HM and with the GPUME encoder; -> dict with MD5s, times, fps"""
    extra = list(extra)
    w, h = [int(v) for v in size.split("x")]
    cfg = os.path.join(CFG_DIR, "encoder_%s.cfg" % cfg_name)
    tmp = tempfile.mkdtemp(prefix="hmenc_")
    yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), w, h, frames, bit_depth)
    out = {"cfg": cfg_name, "size": size, "frames": frames, "qp": qp, "extra": extra}
    # the two encoders are single-threaded: they run side by side on different cores
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as pool:
        fc = None if skip_cpu else pool.submit(run, REF_ENC, cfg, yuv, w, h, frames, qp, os.path.join(tmp, "cpu"), extra, bit_depth)
        fg = pool.submit(run, GPU_ENC, cfg, yuv, w, h, frames, qp, os.path.join(tmp, "gpu"), extra + ["--GPUME=%d" % gpume], bit_depth)
        if fc is not None:
            out["cpu"] = fc.result()
            out["cpu"]["fps"] = frames / out["cpu"]["wall_s"]
        out["gpu"] = fg.result()
    out["gpu"]["fps"] = frames / out["gpu"]["wall_s"]
    if "cpu" in out:
        out["bitstream_identical"] = out["cpu"]["bitstream_md5"] == out["gpu"]["bitstream_md5"]
        out["recon_identical"] = out["cpu"]["recon_md5"] == out["gpu"]["recon_md5"]
        out["speedup_wall"] = out["cpu"]["wall_s"] / out["gpu"]["wall_s"]
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="lowdelay_P_main")
    ap.add_argument("--size", default="416x240")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--qp", type=int, default=32)
    ap.add_argument("--gpume", type=int, default=1)
    ap.add_argument("--bit-depth", type=int, default=8)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    out = compare(a.cfg, a.size, a.frames, a.qp, a.gpume, a.extra, a.bit_depth, a.skip_cpu)
    print(json.dumps(out, indent=1))
    if "cpu" in out and not (out["bitstream_identical"] and out["recon_identical"]):
        raise SystemExit(3)


if __name__ == "__main__":
    main()
