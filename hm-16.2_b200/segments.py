"""Segment sharding: the only way this path scales across GPUs.

An intra period that starts with an IDR (--DecodingRefreshType=2) is a closed segment: no
reference crosses its boundary, so segments are independent encodes (`-fs start -f length`,
TAppEncTop.cpp:369) and need no exchange step -- no collective, only a final concatenation of the
Annex-B streams.  Segment k goes to rank k mod N; several segments (encoder processes) share one
GPU because a single sequential encoder cannot fill it -- through the per-GPU broker daemon (BrokerDaemon).
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


def run_rank(encoder, cfg, yuv, width, height, qp, plan, rank, out_prefix, extra=(), device=None, max_parallel=4, env=None):
    """encode this rank's segments, up to max_parallel encoder processes at a time on its GPU (env: e.g. BrokerDaemon.env)"""
    env = dict(os.environ if env is None else env)
    if device is not None:
        env["HMGPU_DEVICE"] = str(device)
    todo = segments_of_rank(plan, rank)
    running, done = [], []
    while todo or running:
        while todo and len(running) < max_parallel:
            seg = todo.pop(0)
            p = subprocess.Popen(encoder_cmd(encoder, cfg, yuv, width, height, qp, seg, out_prefix, extra),
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            running.append((seg, p))
        seg, p = running.pop(0)
        out, err = p.communicate()
        if p.returncode != 0:     # HM prints its configuration errors on stdout
            raise RuntimeError("segment %d failed (rc %d): %s %s" % (seg["segment"], p.returncode, out.decode()[-300:], err.decode()[-300:]))
        done.append(seg["segment"])
    return done


class BrokerDaemon:
    """The per-GPU broker daemon (hm-16.2_b200/hmgpud, csrc/hmgpud.cu) for the duration of a `with` block.

    Several encoder processes on one GPU are the deployment model of this path (a single sequential encoder cannot fill a
    B200).  The daemon owns the only CUDA context of the GPU; encoders started with `env` (HMGPU_BROKER=<socket>) attach to it
    through a shared-memory mailbox instead of creating a context each (2-4 s per process, and kernels of different processes
    time-slice the GPU unless MPS runs).  Raises when the daemon does not come up: there is no fallback."""

    def __init__(self, device=0, ctas=4, idle_us=2000, socket_path=None):
        self.device = device
        self.dir = tempfile.mkdtemp(prefix="hmgpud_")
        self.socket = socket_path or os.path.join(self.dir, "hmgpud.%d.sock" % device)
        self.cmd = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "hmgpud"), "--device", str(device),
                    "--socket", self.socket, "--ctas", str(ctas), "--idle-us", str(idle_us)]
        self.env = dict(os.environ, HMGPU_BROKER=self.socket)
        self.proc = None
        self.startup_s = None

    def __enter__(self):
        t0 = time.perf_counter()
        self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        line = self.proc.stdout.readline()                  # "hmgpud ready: ..." once the context exists and the socket listens
        if not line.startswith("hmgpud ready"):
            rest = line + (self.proc.stdout.read() if self.proc.poll() is not None else "")
            self.__exit__(None, None, None)
            raise RuntimeError("hmgpud did not start: %s" % rest.strip())
        self.banner = line.strip()
        self.startup_s = time.perf_counter() - t0
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            if self.proc.poll() is None:
                self.proc.terminate()
                try:
                    self.proc.wait(timeout=20)
                except subprocess.TimeoutExpired:
                    self.proc.kill()
            self.proc.stdout.close()
            self.proc = None
        shutil.rmtree(self.dir, ignore_errors=True)
        return False
