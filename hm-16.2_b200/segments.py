"""Segment sharding: the only way this path scales across GPUs.

An intra period that starts with an IDR (--DecodingRefreshType=2) is a closed segment: no
reference crosses its boundary, so segments are independent encodes (`-fs start -f length`,
TAppEncTop.cpp:369) and need no exchange step -- no collective, only a final concatenation of the
Annex-B streams.  Segment k goes to rank k mod N; several segments (encoder processes) share one
GPU because a single sequential encoder cannot fill it.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time


def plan_segments(n_frames, intra_period, n_ranks):
    """-> list of dicts {segment, rank, frame_start, n_frames}; covers [0, n_frames) exactly once"""
    assert intra_period > 0 and n_ranks > 0
    segs = []
    k = 0
    for start in range(0, n_frames, intra_period):
        segs.append({"segment": k, "rank": k % n_ranks, "frame_start": start,
                     "n_frames": min(intra_period, n_frames - start)})
        k += 1
    return segs


def segments_of_rank(plan, rank):
    return [s for s in plan if s["rank"] == rank]


def encoder_cmd(encoder, cfg, yuv, width, height, qp, seg, out_prefix, extra=()):
    """command line of one segment encode (the reference's own options)"""
    return [encoder, "-c", cfg, "-i", yuv, "-wdt", str(width), "-hgt", str(height), "-fr", "30",
            "-fs", str(seg["frame_start"]), "-f", str(seg["n_frames"]), "-q", str(qp),
            "-b", "%s_seg%03d.bin" % (out_prefix, seg["segment"]), "-o", "%s_seg%03d.yuv" % (out_prefix, seg["segment"])] + list(extra)


def run_rank(encoder, cfg, yuv, width, height, qp, plan, rank, out_prefix, extra=(), device=None, max_parallel=4):
    """encode this rank's segments, up to max_parallel encoder processes at a time on its GPU"""
    env = dict(os.environ)
    if device is not None:
        env["HMGPU_DEVICE"] = str(device)
    todo = segments_of_rank(plan, rank)
    running, done = [], []
    while todo or running:
        while todo and len(running) < max_parallel:
            seg = todo.pop(0)
            p = subprocess.Popen(encoder_cmd(encoder, cfg, yuv, width, height, qp, seg, out_prefix, extra),
                                 stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
            running.append((seg, p))
        seg, p = running.pop(0)
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("segment %d failed: %s" % (seg["segment"], err.decode()[-500:]))
        done.append(seg["segment"])
    return done


class MpsDaemon:
    """CUDA MPS for the duration of a `with` block.

    Several encoder processes on one GPU are the deployment model of this path (a single sequential encoder cannot fill a
    B200), but without MPS the kernels of different processes time-slice the whole GPU: four 832x480 GPUME encoders took
    25-40 s each instead of 7 s (profiles/r1l_mps_sharing.log).  Under MPS they run concurrently at single-process speed,
    resident mailbox servers included.  `ok` is False when the control daemon could not be started (the block still runs)."""

    def __init__(self):
        self.dir = tempfile.mkdtemp(prefix="hmgpu_mps_")
        self.env = dict(os.environ, CUDA_MPS_PIPE_DIRECTORY=os.path.join(self.dir, "pipe"),
                        CUDA_MPS_LOG_DIRECTORY=os.path.join(self.dir, "log"))
        self.ok = False

    def __enter__(self):
        os.makedirs(self.env["CUDA_MPS_PIPE_DIRECTORY"], exist_ok=True)
        os.makedirs(self.env["CUDA_MPS_LOG_DIRECTORY"], exist_ok=True)
        exe = shutil.which("nvidia-cuda-mps-control")
        if exe:
            try:
                self.ok = subprocess.run([exe, "-d"], env=self.env, timeout=30).returncode == 0
                time.sleep(0.5)
                if self.ok:
                    # the MPS server itself starts with the first client (seconds): do that before anybody is timed
                    here = os.path.dirname(os.path.abspath(__file__))
                    subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import hmgpu; hmgpu.Context(64, 64, 8, 1).close()" % here],
                                   env=self.env, timeout=120, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            except (OSError, subprocess.TimeoutExpired):
                self.ok = False
        return self

    def __exit__(self, *exc):
        if self.ok:
            try:
                subprocess.run(["nvidia-cuda-mps-control"], input=b"quit\n", env=self.env, timeout=60)
            except (OSError, subprocess.TimeoutExpired):
                pass
        shutil.rmtree(self.dir, ignore_errors=True)
        return False
