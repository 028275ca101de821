"""Multi-GPU host logic on CPU: two gloo ranks shard closed intra-period segments with no
data-path collective; only the bookkeeping (which rank did what, max-over-ranks time) is reduced."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import segments
import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, intra, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = segments.plan_segments(n_frames, intra, world)
    mine = segments.segments_of_rank(plan, rank)
    # "encode": a checksum of the frames of my segments (independent inputs, no exchange)
    frames = synth.luma_frames(64, 48, n_frames, 8)
    rec = torch.zeros(len(plan), 3, dtype=torch.int64)
    for s in mine:
        chunk = frames[s["frame_start"]:s["frame_start"] + s["n_frames"]]
        rec[s["segment"]] = torch.tensor([1, int(chunk.astype(np.int64).sum()), s["n_frames"]])
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)       # pretend device time of this rank
    dist.barrier()
    dist.all_reduce(rec, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((rec.numpy().copy(), float(t[0])))
    dist.destroy_process_group()


def test_two_ranks_cover_every_segment_once():
    n_frames, intra, world = 44, 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, intra, q)) for r in range(world)]
    for p in procs:
        p.start()
    rec, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plan = segments.plan_segments(n_frames, intra, world)
    assert (rec[:, 0] == 1).all()                       # every segment done by exactly one rank
    assert int(rec[:, 2].sum()) == n_frames             # frames covered exactly once
    frames = synth.luma_frames(64, 48, n_frames, 8)
    for s in plan:
        exp = int(frames[s["frame_start"]:s["frame_start"] + s["n_frames"]].astype(np.int64).sum())
        assert int(rec[s["segment"], 1]) == exp         # same result as a single-process run
    assert tmax == 11.0                                 # max over ranks


def test_plan_properties():
    for n, ip, w in [(256, 32, 8), (64, 32, 4), (33, 8, 3), (7, 32, 2)]:
        plan = segments.plan_segments(n, ip, w)
        assert sum(s["n_frames"] for s in plan) == n
        assert [s["frame_start"] for s in plan] == list(range(0, n, ip))
        assert all(s["rank"] == s["segment"] % w for s in plan)
        cmd = segments.encoder_cmd("enc", "c.cfg", "in.yuv", 1920, 1080, 32, plan[-1], "out")
        assert cmd[cmd.index("-fs") + 1] == str(plan[-1]["frame_start"])


# ---- the real thing: segments.run_rank drives encoder processes over closed segments -----------------------------------------
import hashlib
import shutil
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ENC = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRef")
GPU_ENC = os.path.join(ROOT, "hm-16.2_b200", "host", "build", "TAppEncoderGpu")
RA_CFG = os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_randomaccess_main.cfg")
RA_EXTRA = ["--DecodingRefreshType=2", "--IntraPeriod=16"]     # periodic IDR needs IntraPeriod > GOPSize (8)


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _rank_worker(rank, world, port, yuv, w, h, n_frames, prefix, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = segments.plan_segments(n_frames, 16, world)
    done = segments.run_rank(REF_ENC, RA_CFG, yuv, w, h, 32, plan, rank, prefix, RA_EXTRA, max_parallel=2)
    mask = torch.zeros(len(plan), dtype=torch.int64)
    mask[done] = 1
    dist.all_reduce(mask, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(mask.numpy().copy())
    dist.destroy_process_group()


@pytest.mark.skipif(not (os.path.exists(REF_ENC) and os.path.exists(RA_CFG)), reason="reference encoder not built (needs /root/reference)")
def test_run_rank_encodes_real_segments_on_two_ranks():
    """two gloo ranks encode the closed segments of one clip with segments.run_rank (the CPU reference encoder stands in for the
    GPU encoder: same command line, same `-fs/-f` sharding): every segment is encoded exactly once, and a segment's streams are
    byte-identical to those of a single process that encodes the same plan alone -- segments are independent encodes."""
    w, h, n_frames = 64, 64, 48
    tmp = tempfile.mkdtemp(prefix="hmseg_")
    try:
        yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), w, h, n_frames, 8)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, yuv, w, h, n_frames, os.path.join(tmp, "two"), q)) for r in range(2)]
        for p in procs:
            p.start()
        mask = q.get(timeout=240)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert (mask == 1).all()
        plan1 = segments.plan_segments(n_frames, 16, 1)
        assert segments.run_rank(REF_ENC, RA_CFG, yuv, w, h, 32, plan1, 0, os.path.join(tmp, "one"), RA_EXTRA, max_parallel=3) == [0, 1, 2]
        for s in plan1:
            for ext in ("bin", "yuv"):
                a = os.path.join(tmp, "two_seg%03d.%s" % (s["segment"], ext))
                b = os.path.join(tmp, "one_seg%03d.%s" % (s["segment"], ext))
                assert os.path.getsize(a) > 0 and _md5(a) == _md5(b)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF_ENC) and os.path.exists(GPU_ENC) and os.path.exists(RA_CFG)), reason="encoder binaries not built (need /root/reference at build time)")
def test_run_rank_gpu_segments_match_cpu_hm():
    """segments.run_rank with the GPUME encoder: three closed segments, two encoder processes at a time on ONE GPU through the
    broker daemon; bitstream and reconstruction of every segment identical to CPU HM on the same segment."""
    w, h, n_frames = 416, 240, 48
    tmp = tempfile.mkdtemp(prefix="hmseg_")
    try:
        yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), w, h, n_frames, 8)
        plan = segments.plan_segments(n_frames, 16, 1)
        with segments.BrokerDaemon(device=0) as broker:
            done = segments.run_rank(GPU_ENC, RA_CFG, yuv, w, h, 32, plan, 0, os.path.join(tmp, "gpu"), RA_EXTRA + ["--GPUME=1"],
                                     max_parallel=2, env=broker.env)
        assert sorted(done) == [0, 1, 2]
        segments.run_rank(REF_ENC, RA_CFG, yuv, w, h, 32, plan, 0, os.path.join(tmp, "cpu"), RA_EXTRA, max_parallel=3)
        for s in plan:
            for ext in ("bin", "yuv"):
                assert _md5(os.path.join(tmp, "gpu_seg%03d.%s" % (s["segment"], ext))) == _md5(os.path.join(tmp, "cpu_seg%03d.%s" % (s["segment"], ext)))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
