"""Reader of the job streams the HM binding writes in capture mode (HmGpuHost.h: HMGPU_CAPTURE=<file> with --GPUME=2).

A stream is what a real encode asks of libhmgpu, in order: reference-picture uploads ('R'), source pictures ('O') and
xMotionEstimation jobs ('J') -- each job with the answer the CPU search of the encoder gave (integer MV, half / quarter
offsets, cost after the fractional search).  replay() runs a stream through a hmgpu.Context: every maximal run of jobs
between two uploads goes to the device as ONE batch and the results are compared with the captured CPU answers."""
import os
import struct
import subprocess

import numpy as np

from hmgpu import ME_JOB, ME_RESULT

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_ENC = os.path.join(HERE, "host", "build", "TAppEncoderGpu")


def capture_encode(cfg, yuv, w, h, frames, qp, out_file, extra=(), bit_depth=8):
    """run the patched encoder in capture mode (no GPU involved); returns its stderr summary"""
    cmd = [GPU_ENC, "-c", cfg, "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-q", str(qp),
           "-b", out_file + ".bin", "-o", "/dev/null", "--GPUME=2"] + list(extra)
    if bit_depth != 8:
        cmd += ["--InputBitDepth=%d" % bit_depth]
    p = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, HMGPU_CAPTURE=out_file))
    if p.returncode != 0:
        raise RuntimeError("capture encode failed: %s\n%s" % (" ".join(cmd), p.stderr[-2000:]))
    return [ln for ln in p.stderr.splitlines() if ln.startswith("[GPUME]")]


def read_stream(path):
    """-> (pic_w, pic_h, bit_depth, events); events: ('R', slot, poc, luma) | ('O', poc, luma) | ('J', jobs, cpu_results, key_blocks)
    with consecutive jobs merged into one 'J' event (key-block offsets rebased into the event's key array)"""
    data = np.fromfile(path, np.uint8)
    buf = data.tobytes()
    pos, n = 0, len(buf)
    assert buf[0:1] == b"H"
    magic, w, h, bd = struct.unpack_from("<4i", buf, 1)
    assert magic == 0x50414348
    pos = 17
    events = []
    jobs, res, keys, key_elems = [], [], [], 0
    js, rs = ME_JOB.itemsize, ME_RESULT.itemsize

    def flush():
        nonlocal jobs, res, keys, key_elems
        if jobs:
            events.append(("J", np.frombuffer(b"".join(jobs), ME_JOB).copy(), np.frombuffer(b"".join(res), ME_RESULT).copy(),
                           np.concatenate(keys) if keys else None))
            jobs, res, keys, key_elems = [], [], [], 0
    while pos < n:
        tag = buf[pos:pos + 1]
        pos += 1
        if tag in (b"R", b"O"):
            flush()
            slot, poc = struct.unpack_from("<2i", buf, pos)
            pos += 8
            luma = np.frombuffer(buf, np.int16, w * h, pos).reshape(h, w)
            pos += 2 * w * h
            events.append(("R", slot, poc, luma) if tag == b"R" else ("O", poc, luma))
        elif tag == b"J":
            j = np.frombuffer(buf, ME_JOB, 1, pos).copy()
            pos += js
            res.append(buf[pos:pos + rs])
            pos += rs
            (nk,) = struct.unpack_from("<i", buf, pos)
            pos += 4
            if nk:
                j["org_offset"] = key_elems
                keys.append(np.frombuffer(buf, np.int16, nk, pos))
                key_elems += nk
                pos += 2 * nk
            jobs.append(j.tobytes())
        else:
            raise ValueError("bad record tag %r at %d" % (tag, pos - 1))
    flush()
    return w, h, bd, events


FIELDS = ("int_x", "int_y", "half_x", "half_y", "qter_x", "qter_y", "frac_cost")


def replay(ctx, events, timer=None):
    """feed a stream to a hmgpu.Context; -> dict(jobs, batches, mismatches, candidates).  timer(fn) may wrap the search calls."""
    n_jobs = n_bad = n_batches = n_cand = 0
    for ev in events:
        if ev[0] == "R":
            ctx.ref_upload(ev[1], ev[3])
        elif ev[0] == "O":
            ctx.org_upload(ev[2])
        else:
            _, jobs, cpu, keys = ev
            got = timer(lambda: ctx.me_search(jobs, keys)) if timer else ctx.me_search(jobs, keys)
            bad = np.zeros(len(jobs), bool)
            for f in FIELDS:
                bad |= got[f] != cpu[f]
            n_bad += int(bad.sum())
            n_jobs += len(jobs)
            n_batches += 1
            n_cand += int(got["n_cand"].astype(np.int64).sum())
    return {"jobs": n_jobs, "batches": n_batches, "mismatches": n_bad, "candidates": n_cand}
