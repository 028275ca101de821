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
