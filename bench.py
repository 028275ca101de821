#!/usr/bin/env python
"""bench.py -- throughput of the HM-16.2 inter-search hot path on B200 (libhmgpu) beside the
reference's own CPU implementation.

Workload (BASELINE.json configs[1]): encoder_lowdelay_P_main.cfg, 1920x1080 8-bit, TZSearch
(FastSearch=1, SearchRange=64), FEN=1, HadamardME=1, 4 reference pictures, QP 32 lambda.
One "step" = the whole motion-estimation work of ONE P picture: the newest reference is turned
into its 16 quarter-pel planes, the source picture is installed, and every PU of every CU of
the CTU quadtree is searched against all 4 references (worklist.frame_jobs: ~760 k
xMotionEstimation bodies = TZ integer search + half/quarter-pel SATD refinement each).

  value : ME Gcandidates/s, inputs resident in HBM, CUDA events on the library's own stream.
  e2e   : the same step through the C ABI with HOST buffers (reference upload, source upload,
          job list in, results out), wall clock around blocking calls, copies included.
  --impl reference : the reference's own xTZSearch + xPatternSearchFracDIF (oracle/_ref/libhmref.so,
          compiled from the unmodified HM sources) over a bounded sample of the same job list
          on all host cores.

Run:  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hm-16.2_b200"))
sys.path.insert(0, ROOT)

PIC_W, PIC_H, BIT_DEPTH, N_REFS, SEARCH_RANGE, LAMBDA = 1920, 1080, 8, 4, 64, 57.9
METRIC = "ME Gcand/s (1080p lowdelay_P TZSearch, bit-exact vs CPU HM)"
UNIT = "Gcand/s"


def workload_config(n_jobs):
    return {"workload": "encoder_lowdelay_P_main.cfg 1920x1080 8-bit TZSearch SR64 FEN1 HADME1 4 refs QP32: "
                        "all PUs of one P picture x 4 references per step",
            "jobs_per_step": int(n_jobs), "pic": "%dx%d" % (PIC_W, PIC_H), "bit_depth": BIT_DEPTH,
            "n_refs": N_REFS, "search_range": SEARCH_RANGE,
            "l2": "inputs larger than L2 (4 refs x 16 planes = 165 MB + 36 MB jobs + work lists; rotating reference slot)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_work(jobs, results):
    """SURVEY.md 8(d) per-unit figures x the units of one step (independent of GPU tricks).
    -> dict stage -> (int_ops, bytes)"""
    import hmgpu
    w = jobs["pu_w"].astype(np.int64)
    h = jobs["pu_h"].astype(np.int64)
    fen = (jobs["flags"] & hmgpu.F_FEN) != 0
    rows = np.where(fen & (h > 8), h // 2, h)
    int_cand = results["n_cand"].astype(np.int64) - 18
    # integer SAD candidate: W*(H>>s) abs-diff-accumulates (+1 MV cost).  Compulsory bytes: the search window is
    # read ONCE per search, (W + R - L) x (H + B - T) reference samples, plus the W x H source block (1 B/sample at 8 bit)
    tz_ops = int((int_cand * (w * rows + 1)).sum())
    win_w = w + jobs["win_r"].astype(np.int64) - jobs["win_l"].astype(np.int64)
    win_h = h + jobs["win_b"].astype(np.int64) - jobs["win_t"].astype(np.int64)
    tz_bytes = int((win_w * win_h + w * h).sum())
    # SATD candidate: 576 int ops per 8x8 tile, 112 per 4x4 tile; 18 candidates share one (W+8)x(H+8) footprint
    t8 = ((w % 8) == 0) & ((h % 8) == 0)
    tiles = np.where(t8, (w // 8) * (h // 8), (w // 4) * (h // 4))
    per_tile = np.where(t8, 576, 112)
    frac_ops = int((18 * tiles * per_tile).sum())
    frac_bytes = int(((w + 8) * (h + 8) + w * h).sum())
    plane_bytes = (PIC_W + 160) * (PIC_H + 160)
    return {"tz": (tz_ops, tz_bytes), "frac_dist": (frac_ops, frac_bytes),
            "planes": (192 * plane_bytes, 17 * plane_bytes)}


def ncu_traffic(kernel_csv):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary under profiles/ (written by profiles/summarize_ncu.py); None when absent."""
    import csv
    try:
        rows = {r[0]: r for r in csv.reader(open(os.path.join(ROOT, "profiles", kernel_csv)))}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
        # a stage is several launches (one column each): their sum is the traffic of the stage
        return (sum(float(v) for v in rd[2:] if v) * scale[rd[1]] + sum(float(v) for v in wr[2:] if v) * scale[wr[1]])
    except (OSError, KeyError, ValueError, IndexError):
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import hmgpu
    import synth
    import worklist

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass

    # every rank codes its own picture (segment sharding: independent inputs, no exchange)
    frames = synth.luma_frames(PIC_W, PIC_H, N_REFS + 2, BIT_DEPTH, seed=1234 + rank).astype(np.int16)
    # slot k holds frame k, the picture being coded is frame N_REFS + 1
    jobs = worklist.frame_jobs(PIC_W, PIC_H, n_refs=N_REFS, seed=7 + rank, search_range=SEARCH_RANGE, lam=LAMBDA,
                               ref_dist=[N_REFS + 1 - k for k in range(N_REFS)])
    n_jobs = len(jobs)
    flags_any = int(np.bitwise_or.reduce(jobs["flags"]))
    ctx = hmgpu.Context(PIC_W, PIC_H, BIT_DEPTH, N_REFS, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    d_frames = torch.from_numpy(frames).cuda()
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(n_jobs, -1).copy()).cuda()
    d_res = torch.zeros((n_jobs, hmgpu.ME_RESULT.itemsize), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for s in range(N_REFS):
        ctx.ref_upload_device(s, d_frames[s].data_ptr(), PIC_W)
    ctx.synchronize()

    def step_device(i):
        # the newest reconstructed picture replaces the oldest reference slot, then one picture of ME
        ctx.ref_upload_device(i % N_REFS, d_frames[i % N_REFS].data_ptr(), PIC_W)
        ctx.org_upload_device(d_frames[N_REFS + 1].data_ptr(), PIC_W)
        ctx.me_search_device(d_jobs.data_ptr(), n_jobs, None, d_res.data_ptr(), flags_any)

    # end-to-end inputs live in page-locked host memory (hmgpu_host_alloc): the library copies them directly
    h_frames = ctx.host_array(frames.shape, np.int16)
    h_frames[...] = frames
    h_jobs = ctx.host_array((n_jobs,), hmgpu.ME_JOB)
    h_jobs[...] = jobs
    h_res = ctx.host_array((n_jobs,), hmgpu.ME_RESULT)

    def step_host(i):
        ctx.ref_upload(i % N_REFS, h_frames[i % N_REFS])
        ctx.org_upload(h_frames[N_REFS + 1])
        return ctx.me_search(h_jobs, out=h_res)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # ---- kernel-resident timing --------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    ctx.profile_read(reset=True)
    ctx.profile_enable(True)
    launches0 = ctx.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    prof = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    res = d_res.cpu().numpy().view(hmgpu.ME_RESULT).reshape(-1)
    cand_per_step = int(res["n_cand"].astype(np.int64).sum())

    # ---- end-to-end through the C ABI with host buffers ----------------------------------------
    for i in range(min(2, args.warmup)):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        res_h = step_host(args.warmup + i)
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    assert len(res_h) == n_jobs

    int_peak = ctx.microbench(0) if rank == 0 else None
    sad4_peak = ctx.microbench(1) if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(cand_per_step)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_cand = float(c[0])

    ctx.close()                                            # the encode legs run through broker daemons: free this rank's context first
    seg = None if args.no_encode else encode_segment_leg(rank, world, local_rank, dist, args.full)

    if rank == 0:
        value = total_cand * args.steps / (dev_ms * 1e-3) / 1e9
        e2e_value = total_cand * args.steps / (e2e_ms * 1e-3) / 1e9
        frame_bytes = PIC_W * PIC_H * 2
        work = algorithmic_work(jobs, res)
        stage_ms = {k: v[0] / args.steps for k, v in prof.items() if v[1]}
        # roofline = the dominant KERNEL: the stage with the longest average launch (the fractional stage is ONE launch of
        # fracw_group_kernel; the TZ stage is 17 launches on side streams, none longer than half of it -- the launch list
        # profiles/r2t_ncu_launches_bench.csv shows both).  The longest STAGE is named next to it (roofline.longest_stage) and every
        # stage has its own entry in roofline_int32.per_stage.
        per_launch = {k: stage_ms[k] / max(1.0, prof[k][1] / args.steps) for k in stage_ms if k in work}
        dom = max(per_launch, key=lambda k: per_launch[k])
        longest = max((k for k in stage_ms if k in work), key=lambda k: stage_ms[k])
        ops, byts = work[dom]
        dur = stage_ms[dom] * 1e-3
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic_csv = {"tz": "r1n_ncu_tz_stage.csv", "frac_dist": "r2k_ncu_fracw_group.csv"}.get(dom)
        notes = {
            "tz": "the TZ stage (classify + 14 one-thread-per-job shape kernels + 2 warp-per-job launches) is a chain of dependent "
                  "points per job: bound by latency / instruction issue, not by HBM (DRAM traffic << algorithmic bytes: the planes "
                  "are served by L2); see roofline_int32 and full_search.roofline for the integer-pipe view",
            "frac_dist": "the fractional stage (one CTA per (reference, CTU) group of jobs, windows of the phase planes staged by TMA, "
                         "me_fracw.cu) is bound by integer issue (ncu: profiles/r2k_ncu_fracw_group.csv: 62 % issue active, ALU pipe "
                         "52 %, FMA-heavy pipe 47 %), not by HBM: DRAM traffic 0.47 GB per step against 0.60 GB algorithmic; see roofline_int32"}
        # The dominant stage is integer-pipe work out of L1 / shared memory (the north star's "INT32-pipe roofline"; the task's
        # schema only names hbm | tensor, neither of which bounds it), so the headline roofline is the measured INT32 lane-op
        # rate; the same stage in HBM terms is kept as a sub-object.
        kernel_names = {"frac_dist": "fracw_group_kernel (me_fracw.cu)", "tz": "TZ stage: tzt_search_kernel<...> x 14 + tz_search_kernel x 2",
                        "planes": "phase_planes_kernel"}
        roofline = {"kernel": kernel_names.get(dom, dom), "stage": dom, "bound": "int32", "achieved": ops / dur / 1e9, "peak": int_peak, "unit": "G int-op/s",
                    "frac": ops / dur / 1e9 / int_peak,
                    "peak_source": "hmgpu_microbench(0): LOP3 + IMAD.IADD lane-ops/s on both integer pipes, measured in this run",
                    "algorithmic_ops_per_step": ops, "launches_per_step": int(prof[dom][1] // args.steps),
                    "traffic": ncu_traffic(traffic_csv) if traffic_csv else None,
                    "traffic_source": ("profiles/" + traffic_csv) if traffic_csv else None,
                    "hbm": {"achieved": byts / dur / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": byts / dur / 1e9 / hbm_peak,
                            "algorithmic_bytes_per_step": byts,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
                    "longest_stage": {"stage": longest, "ms": stage_ms[longest], "launches_per_step": int(prof[longest][1] // args.steps),
                                      "frac_int32": work[longest][0] / (stage_ms[longest] * 1e-3) / 1e9 / int_peak,
                                      "note": notes.get(longest, "") if longest != dom else ""},
                    "note": notes.get(dom, "")}
        roofline_int32 = {"peak": int_peak, "unit": "Gop/s", "vabsdiff4_peak_glaneops": sad4_peak,
                          "per_stage": {k: {"ms": stage_ms[k], "gops": work[k][0] / (stage_ms[k] * 1e-3) / 1e9,
                                            "frac": work[k][0] / (stage_ms[k] * 1e-3) / 1e9 / int_peak,
                                            "gbs": work[k][1] / (stage_ms[k] * 1e-3) / 1e9} for k in stage_ms if k in work}}
        me_ops = sum(work[k][0] for k in ("tz", "frac_dist") if k in stage_ms)
        me_ms = sum(stage_ms[k] for k in ("tz", "frac_dist", "frac_expand", "frac_select") if k in stage_ms)
        if me_ms > 0:
            # the whole ME step (integer + fractional search of every job) against the INT32 peak, in algorithmic operations
            roofline_int32["me_step"] = {"ms": me_ms, "gops": me_ops / (me_ms * 1e-3) / 1e9, "frac": me_ops / (me_ms * 1e-3) / 1e9 / int_peak}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8", "data": "synthetic", "config": workload_config(n_jobs),
               "candidates_per_step": total_cand, "jobs_per_s": n_jobs * world * args.steps / (dev_ms * 1e-3),
               "pictures_per_s_me_only": world * args.steps / (dev_ms * 1e-3),
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                       "h2d_bytes_per_step": int(2 * frame_bytes + jobs.nbytes),
                       "d2h_bytes_per_step": int(res.nbytes), "timing": "wall clock around the blocking C-ABI calls"},
               "stage_ms_per_step": stage_ms, "roofline": roofline, "roofline_int32": roofline_int32}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference_run(jobs, frames, sample_jobs=args.cpu_sample, procs=1)
        if seg is not None:
            out["encode_segments"] = seg
        if world == 1 and not args.no_full_search:
            out["full_search"] = full_search_leg(local_rank, max(2, args.steps // 2), sad4_peak, int_peak)
        if world == 1 and not args.no_rdoq:
            out["rdoq"] = rdoq_leg(local_rank, max(2, args.steps))
        if world == 1 and not args.no_encode:
            out["real_stream"] = real_stream_leg(local_rank)
            out["encode"] = encode_runs(local_rank, args.full)
            out["encode_shared_gpu"] = encode_shared_gpu_leg(local_rank, args.full)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def full_search_leg(local_rank, steps, sad4_peak, int_peak=None):
    """BASELINE.json configs[0] as a batch: encoder_lowdelay_P_main.cfg 416x240, FastSearch=0, SearchRange=64, 4 references --
    every PU of one P picture against its full +-64 window (xPatternSearch), device-resident.  This is the stage whose
    roofline is the integer pipe: algorithmic work = W*(H>>s) abs-diff-accumulates per candidate."""
    import torch
    import hmgpu
    import synth
    import worklist
    w, h, n_refs = 416, 240, 4
    frames = synth.luma_frames(w, h, n_refs + 2, 8, seed=99).astype(np.int16)
    jobs = worklist.frame_jobs(w, h, n_refs=n_refs, full_search=True, search_range=64, lam=LAMBDA, frac=False,
                               ref_dist=[n_refs + 1 - k for k in range(n_refs)])
    ctx = hmgpu.Context(w, h, 8, n_refs, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    d_frames = torch.from_numpy(frames).cuda()
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(len(jobs), -1).copy()).cuda()
    d_res = torch.zeros((len(jobs), hmgpu.ME_RESULT.itemsize), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for s in range(n_refs):
        ctx.ref_upload_device(s, d_frames[s].data_ptr(), w)
    ctx.org_upload_device(d_frames[n_refs + 1].data_ptr(), w)
    flags_any = int(np.bitwise_or.reduce(jobs["flags"]))
    for _ in range(2):
        ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = d_res.cpu().numpy().view(hmgpu.ME_RESULT).reshape(-1)
    cand = res["n_cand"].astype(np.int64)
    pw, ph = jobs["pu_w"].astype(np.int64), jobs["pu_h"].astype(np.int64)
    rows = np.where(ph > 8, ph // 2, ph)                       # FEN
    px_sads = int((cand * pw * rows).sum())
    ctx.close()
    achieved = px_sads / 4 / (ms * 1e-3) / 1e9                 # giga VABSDIFF4-lane-ops/s that the algorithm needs
    return {"workload": "encoder_lowdelay_P_main.cfg 416x240 FastSearch=0 SearchRange=64 FEN1, all PUs of one P picture x 4 references",
            "jobs_per_step": int(len(jobs)), "candidates_per_step": int(cand.sum()), "ms_per_step": ms,
            "gcand_per_s": float(cand.sum()) / (ms * 1e-3) / 1e9, "pixel_sads_per_s": px_sads / (ms * 1e-3),
            "roofline": {"kernel": "full_search_packed_kernel", "bound": "int32",
                         # one abs-diff-accumulate per pixel-candidate, against the measured INT32 lane-op rate
                         "achieved": px_sads / (ms * 1e-3) / 1e9, "peak": int_peak, "unit": "G int-op/s",
                         "frac": (px_sads / (ms * 1e-3) / 1e9 / int_peak) if int_peak else None,
                         "peak_source": "hmgpu_microbench(0): LOP3 + IMAD.IADD lane-ops/s on both integer pipes, measured in this run",
                         # the same work as the kernel issues it: 4 pixel-candidates per VABSDIFF4.U8.ACC lane-op (ALU pipe only)
                         "packed": {"achieved": achieved, "peak": sad4_peak, "unit": "G VABSDIFF4 lane-op/s", "frac": achieved / sad4_peak,
                                    "peak_source": "hmgpu_microbench(1) measured in this run"}}}


def rdoq_leg(local_rank, steps, rep=64):
    """rate-distortion optimised quantisation (TComTrQuant::xRateDistOptQuant) of a batch of TUs through hmgpu_rdoq.
    The batch is the reference encoder's own calls (tests/golden/rdoq_golden.npz: 1067 TUs of every size, luma + chroma, 686 coder
    states) `rep` times over; the levels are compared with the ones the reference returned before anything is timed.  Device time =
    the stage timers around the launch (CUDA events on the context's stream); e2e = the blocking C-ABI call with page-locked host
    buffers, copies inside.  cpu_baseline: the oracle's restatement (oracle/hm_rdoq.c, pinned to the same calls) on one host core."""
    import hmgpu
    import rdoq_batch
    jobs, bits, coef, want, want_sum = rdoq_batch.golden_batch(rep)
    n_coef = int(coef.size)
    with hmgpu.Context(64, 64, 8, 1, device=local_rank) as ctx:
        h_coef, h_level = ctx.host_array(coef.shape, np.int32), ctx.host_array(coef.shape, np.int32)
        h_coef[:] = coef
        for _ in range(2):
            level, abs_sum = ctx.rdoq(jobs, bits, h_coef, out=h_level)
        identical = bool(np.array_equal(level, want) and np.array_equal(abs_sum, want_sum))
        ctx.profile_enable(True)
        ctx.profile_read(True)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.rdoq(jobs, bits, h_coef, out=h_level)
        wall = (time.perf_counter() - t0) / steps
        ms, launches = ctx.profile_read(True)["quant"]
        ctx.profile_enable(False)
    ms /= steps
    out = {"workload": "the reference encoder's %d xRateDistOptQuant calls x %d: %d TUs (4x4 %d, 8x8 %d, 16x16 %d, 32x32 %d), %d coefficients, %d sets of bit estimates"
                       % (len(jobs) // rep, rep, len(jobs), *[int((jobs["log2_size"] == lg).sum()) for lg in (2, 3, 4, 5)], n_coef, len(bits)),
           "levels_identical_to_reference": identical, "kernel": "rdoq_tu_kernel (rdoq.cu): one thread per TU, one launch", "launches_per_batch": launches // steps,
           "ms_per_batch": ms, "mtu_per_s": len(jobs) / ms / 1e3, "mcoef_per_s": n_coef / ms / 1e3,
           "e2e": {"ms_per_batch": wall * 1e3, "mcoef_per_s": n_coef / wall / 1e6, "h2d_bytes": int(coef.nbytes + jobs.nbytes + bits.nbytes),
                   "d2h_bytes": int(coef.nbytes + 4 * len(jobs)), "timing": "wall clock around the blocking C-ABI call, page-locked host buffers"},
           # 4 bytes in + 4 bytes out per coefficient are compulsory; what bounds the kernel is the chain of a TU (one level
           # decision after the other: the level coder's state and the running cost pass from coefficient to coefficient)
           "roofline": {"bound": "latency", "hbm": {"achieved": 8.0 * n_coef / (ms * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes_per_batch": 8 * n_coef},
                        "ns_per_chain_step": ms * 1e6 / 1024.0,
                        "note": "the batch cannot finish before its longest chain: 1024 dependent steps of a 32x32 TU; "
                                "throughput grows with the batch until every scheduler holds several warps (DESIGN.md, RDOQ)"}}
    try:
        from oracle import binding as B
        tus = np.zeros(len(jobs) // rep, B.RDOQ_TU)
        for f in ("log2_size", "channel", "scan", "qbits", "qp_per", "qp_rem", "go_rice_init", "cbf_bits", "bit_depth", "err_scale", "lambda"):
            tus[f] = jobs[f][:len(tus)]
        tus["sign_hide"] = jobs["flags"][:len(tus)]
        obits = np.zeros(len(bits), B.RDOQ_BITS)
        for f in obits.dtype.names:
            obits[f] = bits[f]
        n1, m = n_coef // rep, len(tus)
        t0, loops, bad = time.perf_counter(), 0, 0
        while time.perf_counter() - t0 < 3.0:
            lv, total = B.rdoq_batch(tus, jobs["bits_index"][:m], jobs["coef_offset"][:m], obits, coef[:n1])
            bad += int(not np.array_equal(lv, want[:n1])) + int(total != int(want_sum[:m].sum()))
            loops += 1
        cpu = (time.perf_counter() - t0) / loops
        out["cpu_baseline"] = {"value": n1 / cpu / 1e6, "unit": "Mcoef/s", "cores": 1, "kind": "port", "mismatches": bad,
                               "sample": "the %d calls once per pass, %d passes, oracle/hm_rdoq.c (hmo_rdoq_batch: one C call per pass)" % (m, loops)}
    except Exception as e:                                          # (the oracle is the checker: its timing is a reported extra)
        out["cpu_baseline"] = {"unavailable": repr(e)}
    out["cpu_reference_encoder"] = rdoq_reference_speed()
    return out


def rdoq_reference_speed():
    """the reference's own xRateDistOptQuant on this host: a short encode by the instrumented reference encoder
    (oracle/_ref/TAppEncoderRdoq, HM_RDOQ_TIME: seconds between the two hooks around the call, summed over the encode).  Its mix of
    TUs is the encode's own (many TUs quantise to nothing and return early), not the bench batch: a second reported baseline."""
    import synth
    import tempfile
    enc = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRdoq")
    cfg = os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_lowdelay_P_main.cfg")
    if not (os.path.exists(enc) and os.path.exists(cfg)):
        return {"unavailable": "oracle/_ref/TAppEncoderRdoq not built"}
    try:
        with tempfile.TemporaryDirectory(prefix="hmrdoq_") as tmp:
            yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), 416, 240, 3, 8, seed=5)
            tf = os.path.join(tmp, "time.txt")
            t0 = time.perf_counter()
            subprocess.run([enc, "-c", cfg, "-i", yuv, "-wdt", "416", "-hgt", "240", "-fr", "30", "-f", "3", "-q", "32", "-b", os.path.join(tmp, "s.bin"),
                            "-o", os.path.join(tmp, "rec.yuv")], check=True, capture_output=True, env=dict(os.environ, HM_RDOQ_TIME=tf), timeout=120)
            wall = time.perf_counter() - t0
            calls, coefs, secs = open(tf).read().split()
        return {"value": int(coefs) / float(secs) / 1e6, "unit": "Mcoef/s", "cores": 1, "kind": "reference", "calls": int(calls), "coefficients": int(coefs),
                "seconds_in_rdoq": float(secs), "encode_s": wall,
                "sample": "encoder_lowdelay_P_main.cfg 416x240 3 frames QP32: every xRateDistOptQuant call of the encode, timed inside the reference encoder"}
    except Exception as e:
        return {"unavailable": repr(e)}


def real_stream_leg(local_rank, frames=3):
    """ME throughput on the encoder's REAL call stream: the patched HM encoder in capture mode (HMGPU_CAPTURE, CPU search, no
    GPU) codes `frames` pictures of the 1080p lowdelay-P clip and writes every xMotionEstimation call as the job the binding
    sends, with the CPU search's answer; the stream is then replayed through libhmgpu, one batch per picture, and every result
    is compared with the CPU's.  Unlike the synthetic work-list of the headline (every PU shape of every CU, predictors near
    the true motion) this is exactly what an encode asks for, in its data-dependent mix of shapes and predictors."""
    ec = _enc()
    if ec is None:
        return {"unavailable": "encoder binaries not built (need /root/reference at build time)"}
    import tempfile
    import torch
    import capture
    import hmgpu
    import synth
    tmp = tempfile.mkdtemp(prefix="hmcap_")
    yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), PIC_W, PIC_H, frames, 8)
    path = os.path.join(tmp, "stream.bin")
    t0 = time.perf_counter()
    capture.capture_encode(os.path.join(ec.CFG_DIR, "encoder_lowdelay_P_main.cfg"), yuv, PIC_W, PIC_H, frames, 32, path)
    cap_s = time.perf_counter() - t0
    w, h, bd, events = capture.read_stream(path)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    walls = []

    def timer(fn):
        t = time.perf_counter()
        r = fn()
        walls.append(time.perf_counter() - t)
        return r
    with hmgpu.Context(w, h, bd, 16, device=local_rank) as ctx:
        capture.replay(ctx, events)                           # warm-up pass (allocations, first launches)
        walls.clear()
        out = capture.replay(ctx, events, timer)
        # the largest batch (one whole P picture) device-resident, CUDA events on the library's stream
        big = max((e for e in events if e[0] == "J"), key=lambda e: len(e[1]))
        jobs = big[1]
        stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
        d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(len(jobs), -1).copy()).cuda()
        d_res = torch.zeros((len(jobs), hmgpu.ME_RESULT.itemsize), dtype=torch.uint8, device="cuda")
        flags_any = int(np.bitwise_or.reduce(jobs["flags"]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(4):
            if rep == 1:
                e0.record(stream)
            ctx.me_search_device(d_jobs.data_ptr(), len(jobs), None, d_res.data_ptr(), flags_any)
        e1.record(stream)
        ctx.synchronize()
        dev_ms = e0.elapsed_time(e1) / 3
        res = d_res.cpu().numpy().view(hmgpu.ME_RESULT).reshape(-1)
        big_cand = int(res["n_cand"].astype(np.int64).sum())
        same = all((res[f] == big[2][f]).all() for f in capture.FIELDS)
    wall = float(sum(walls))
    return {"workload": "encoder_lowdelay_P_main.cfg 1920x1080 QP32 TZSearch, %d frames: the xMotionEstimation calls of a real CPU HM encode "
                        "(captured by the binding in capture mode), replayed one batch per picture" % frames,
            "capture_encode_s": cap_s, "jobs": out["jobs"], "batches": out["batches"], "candidates": out["candidates"],
            "mismatches_vs_cpu_search": out["mismatches"],
            "e2e_host_buffers": {"seconds": wall, "gcand_per_s": out["candidates"] / wall / 1e9, "jobs_per_s": out["jobs"] / wall},
            "largest_batch_device_resident": {"jobs": int(len(jobs)), "candidates": big_cand, "ms": dev_ms,
                                              "gcand_per_s": big_cand / (dev_ms * 1e-3) / 1e9, "identical_to_cpu_search": bool(same)}}


def _enc():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import encode_compare
    return encode_compare if encode_compare.available() else None


def encode_runs(local_rank, full=False):
    """whole-encoder numbers (BASELINE.json: encode fps, bitstream MD5-identical): ONE HM encoder with GPUME=1, attached to the
    per-GPU broker daemon, beside the unmodified CPU encoder (oracle/_ref, baseline leg only) on the same clip; the two run side
    by side on different host cores.  Default: bounded clips; --full: BASELINE cfg 2 at its stated size (32 frames, QP 27/32/37)."""
    ec = _enc()
    if ec is None:
        return {"unavailable": "encoder binaries not built (need /root/reference at build time)"}
    import segments
    import synth
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    cases = [("cfg1_lowdelay_P_fullsearch_SR64_416x240_3f", "lowdelay_P_main", 416, 240, 3, 32, ["--FastSearch=0", "--SearchRange=64"]),
             ("cfg2_lowdelay_P_TZ_1920x1080_4f_qp32", "lowdelay_P_main", 1920, 1080, 4, 32, [])]
    if full:
        cases = [("cfg2_lowdelay_P_TZ_1920x1080_32f_qp%d" % qp, "lowdelay_P_main", 1920, 1080, 32, qp, []) for qp in (27, 32, 37)]
    runs = {}
    tmp = tempfile.mkdtemp(prefix="hmenc_")
    with segments.BrokerDaemon(device=local_rank) as brk:
        env = dict(brk.env, HMGPU_SERVER_STATS="1")

        def one(case):
            name, cfg_name, w, h, n, qp, extra = case
            cfg = os.path.join(ec.CFG_DIR, "encoder_%s.cfg" % cfg_name)
            yuv = os.path.join(tmp, "in_%dx%d_%d.yuv" % (w, h, n))
            with ThreadPoolExecutor(2) as pool:
                fc = pool.submit(ec.run, ec.REF_ENC, cfg, yuv, w, h, n, qp, os.path.join(tmp, name + "_cpu"), extra)
                fg = pool.submit(ec.run, ec.GPU_ENC, cfg, yuv, w, h, n, qp, os.path.join(tmp, name + "_gpu"), extra + ["--GPUME=1"], 8, env)
                c, g = fc.result(), fg.result()
            return name, {"cpu_fps": n / c["wall_s"], "gpu_fps": n / g["wall_s"], "cpu_s": c["wall_s"], "gpu_s": g["wall_s"],
                          "speedup": c["wall_s"] / g["wall_s"],
                          "bitstream_md5_identical": c["bitstream_md5"] == g["bitstream_md5"],
                          "recon_md5_identical": c["recon_md5"] == g["recon_md5"], "gpume": g["gpume"]}
        for w, h, n in sorted({(c[2], c[3], c[4]) for c in cases}):
            synth.write_yuv(os.path.join(tmp, "in_%dx%d_%d.yuv" % (w, h, n)), w, h, n, 8)
        try:
            if full:                                          # the three QPs side by side (6 processes, 6 host cores)
                with ThreadPoolExecutor(len(cases)) as pool:
                    for name, r in pool.map(one, cases):
                        runs[name] = r
            else:
                for case in cases:
                    name, r = one(case)
                    runs[name] = r
        except SystemExit as e:
            runs["error"] = str(e)
        runs["broker"] = {"banner": brk.banner, "startup_s": brk.startup_s}
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return runs


def encode_segment_leg(rank, world, local_rank, dist, full=False):
    """Multi-GPU encode throughput over the natural shard (BASELINE cfg 4): closed intra-period segments of
    encoder_randomaccess_main.cfg (--DecodingRefreshType=2) at 1920x1080, S = host cores / 8 segments per GPU (so that 8 GPUs
    use every host core: each encoder needs a core for its serial part), one encoder process per segment, all processes of a
    GPU attached to that GPU's broker daemon.  No exchange step.  fps = all frames / slowest rank.  Beside it the unmodified CPU
    encoder over the SAME segments with the same number of processes (at N = 8: all host cores).  MD5 compared per segment.
    Default: 9-frame segments (one I + one hierarchical-B GOP of 8); --full: 32-frame segments (cfg 4's size)."""
    ec = _enc()
    import torch
    if ec is None:
        return {"unavailable": "encoder binaries not built (need /root/reference at build time)"}
    import segments
    import synth
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    w, h = 1920, 1080
    n = 32 if full else 9
    cores = os.cpu_count() or 8
    per_gpu = max(1, cores // 8)
    tmp = tempfile.mkdtemp(prefix="hmseg_")
    cfg = os.path.join(ec.CFG_DIR, "encoder_randomaccess_main.cfg")
    extra = ["--DecodingRefreshType=2", "--IntraPeriod=%d" % (32 if full else 16)]     # HM wants IntraPeriod > GOP size for IDR periods
    yuvs = [synth.write_yuv(os.path.join(tmp, "seg%d.yuv" % k), w, h, n, 8, seed=1234 + rank * per_gpu + k) for k in range(per_gpu)]

    def many(enc, tag, more, env):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        with ThreadPoolExecutor(per_gpu) as pool:
            outs = list(pool.map(lambda k: ec.run(enc, cfg, yuvs[k], w, h, n, 32, os.path.join(tmp, "%s%d" % (tag, k)), extra + more, 8, env), range(per_gpu)))
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), outs

    with segments.BrokerDaemon(device=local_rank) as brk:
        g_s, g = many(ec.GPU_ENC, "g", ["--GPUME=1"], brk.env)
        startup = brk.startup_s
    c_s, c = many(ec.REF_ENC, "c", [], None)
    same = torch.tensor([int(all(a["bitstream_md5"] == b["bitstream_md5"] and a["recon_md5"] == b["recon_md5"] for a, b in zip(g, c)))], device="cuda")
    if dist is not None:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out = None
    if rank == 0:
        frames = world * per_gpu * n
        out = {"workload": "encoder_randomaccess_main.cfg --DecodingRefreshType=2, 1920x1080 QP32, %d closed %d-frame segment(s) "
                           "per GPU, one encoder process per segment, attached to the GPU's broker daemon" % (per_gpu, n),
               "segments": world * per_gpu, "frames": frames, "host_cores": cores, "encoder_processes": world * per_gpu,
               "gpu_fps_total": frames / g_s, "slowest_rank_s": g_s,
               "cpu_fps_total_same_processes": frames / c_s, "cpu_s": c_s, "gpu_over_cpu": c_s / g_s,
               "all_md5_identical": bool(int(same[0])), "broker_startup_s_not_timed": startup, "gpume_rank0": g[0]["gpume"]}
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


def encode_shared_gpu_leg(local_rank, full=False):
    """Several encoder instances on ONE GPU through the broker daemon, beside the same number of CPU HM processes: one process
    per host core, each coding its own closed segment (encoder_randomaccess_main.cfg, --DecodingRefreshType=2), all started
    together.  Default: 832x480, 16-frame segments; --full: 1920x1080, 32-frame segments (cfg 4)."""
    ec = _enc()
    if ec is None:
        return {"unavailable": "encoder binaries not built (need /root/reference at build time)"}
    import re
    import segments
    import synth
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    w, h, n = (1920, 1080, 32) if full else (832, 480, 16)
    n_procs = os.cpu_count() or 8
    tmp = tempfile.mkdtemp(prefix="hmshare_")
    cfg = os.path.join(ec.CFG_DIR, "encoder_randomaccess_main.cfg")
    extra = ["--DecodingRefreshType=2", "--IntraPeriod=%d" % n]
    yuvs = [synth.write_yuv(os.path.join(tmp, "s%d.yuv" % k), w, h, n, 8, seed=4321 + k) for k in range(n_procs)]

    def many(enc, tag, more, env):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(n_procs) as pool:
            outs = list(pool.map(lambda k: ec.run(enc, cfg, yuvs[k], w, h, n, 32, os.path.join(tmp, "%s%d" % (tag, k)),
                                                  extra + more, 8, env), range(n_procs)))
        return time.perf_counter() - t0, outs

    with segments.BrokerDaemon(device=local_rank) as brk:
        g_s, g = many(ec.GPU_ENC, "g", ["--GPUME=1"], dict(brk.env, HMGPU_SERVER_STATS="1"))
        startup = brk.startup_s
    c_s, c = many(ec.REF_ENC, "c", [], None)
    same = all(a["bitstream_md5"] == b["bitstream_md5"] and a["recon_md5"] == b["recon_md5"] for a, b in zip(g, c))
    attach = [float(m.group(1)) for o in g for m in [re.search(r"([0-9.]+) s waiting for the one-time", " ".join(o["gpume"]))] if m]
    search = [float(m.group(1)) for o in g for m in [re.search(r"([0-9.]+) s in motionSearch overall", " ".join(o["gpume"]))] if m]
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return {"workload": "%d encoder processes (= host cores) on ONE GPU through the broker daemon, one closed %d-frame segment each "
                        "(randomaccess_main, %dx%d, QP32), started together; CPU arm: the same %d segments on the same cores" % (n_procs, n, w, h, n_procs),
            "host_cores": os.cpu_count(), "encoder_processes": n_procs, "gpu_procs_on_one_gpu_fps_total": n_procs * n / g_s, "gpu_s": g_s,
            "cpu_procs_fps_total": n_procs * n / c_s, "cpu_s": c_s, "gpu_over_cpu": c_s / g_s, "all_md5_identical": same,
            "broker_startup_s_not_timed": startup, "gpu_proc_attach_s_mean": float(np.mean(attach)) if attach else None,
            "gpu_proc_binding_s_mean": float(np.mean(search)) if search else None, "gpume_proc0": g[0]["gpume"]}


# ---- the reference on the host cores ------------------------------------------------------------

def _cpu_worker(a):
    kind, jobs_bytes, n, frames, bit_depth, want_count = a
    import hmgpu
    from oracle import binding as B

    def padded_ref(luma):
        # the reference reads its padded reconstruction (TComPicYuv::extendPicBorder, margin 80)
        out = np.zeros((PIC_H + 160, PIC_W + 160), np.int16)
        B.oracle().hmo_extend_border(np.ascontiguousarray(luma, np.int16), PIC_W, PIC_H, 80, out)
        return out

    jobs = np.frombuffer(jobs_bytes, dtype=hmgpu.ME_JOB)
    pads = [padded_ref(frames[k]) for k in range(N_REFS)]
    fn = B.ref().ref_me_batch if kind == "reference" else B.oracle().hmo_me_batch
    t0 = time.perf_counter()
    _, cpu_s = B.me_batch(fn, jobs, pads, frames[N_REFS + 1], bit_depth)
    wall = time.perf_counter() - t0
    cand = None
    if want_count:                                           # untimed: the candidate count (the reference does not count them)
        cnt, _ = B.me_batch(B.oracle().hmo_me_batch, jobs, pads, frames[N_REFS + 1], bit_depth)
        cand = int(cnt.view(hmgpu.ME_RESULT)["n_cand"].astype(np.int64).sum())
    return cpu_s, wall, cand


def cpu_reference_run(jobs, frames, sample_jobs, procs, cand_known=None):
    """time the reference's xTZSearch + xPatternSearchFracDIF on a bounded, evenly spaced sample
    of the step's job list, `procs` processes (the reference is single-threaded per process)."""
    from oracle import binding as B
    kind = "reference" if B.have_ref() else "port"
    stride = max(1, len(jobs) // max(1, sample_jobs))
    sample = np.ascontiguousarray(jobs[::stride][:sample_jobs])
    # dealt round-robin: the job list is ordered by CU depth (64x64 PUs first, 8x4 / 4x8 last), so contiguous chunks would give
    # the first worker several times the work of the last and the wall clock (max over workers) would flatter the GPU
    chunks = [sample[k::procs] for k in range(procs)]
    argsl = [(kind, np.ascontiguousarray(c).tobytes(), len(c), frames, BIT_DEPTH, cand_known is None) for c in chunks if len(c)]
    t0 = time.perf_counter()
    if procs == 1:
        outs = [_cpu_worker(argsl[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            outs = pool.map(_cpu_worker, argsl)
    wall = max(o[1] for o in outs)
    cand = cand_known if cand_known is not None else sum(o[2] for o in outs)
    return {"value": cand / wall / 1e9, "candidates": cand, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": "%d of %d jobs of one step (every %dth), %s xTZSearch+xPatternSearchFracDIF, %.1f s wall"
                      % (len(sample), len(jobs), stride, "libhmref.so" if kind == "reference" else "oracle/hm_oracle.c", wall),
            "jobs_per_s": len(sample) / wall, "seconds": wall, "total_s_incl_setup": time.perf_counter() - t0}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import synth
    import worklist
    frames = synth.luma_frames(PIC_W, PIC_H, N_REFS + 2, BIT_DEPTH, seed=1234).astype(np.int16)
    jobs = worklist.frame_jobs(PIC_W, PIC_H, n_refs=N_REFS, seed=7, search_range=SEARCH_RANGE, lam=LAMBDA,
                               ref_dist=[N_REFS + 1 - k for k in range(N_REFS)])
    procs = os.cpu_count() or 1
    vals = []
    cand = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(jobs, frames, sample_jobs=args.cpu_sample, procs=procs, cand_known=cand)
        cand = r["candidates"]                               # counted once (an untimed oracle pass), the same for every step
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    secs = float(np.mean([r["seconds"] for r in vals]))
    last = vals[-1]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(len(jobs)),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1200000,
                    help="jobs of the step in the CPU baseline sample (default: the whole step, ~10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encode", action="store_true", help="skip the whole-encoder CPU vs GPUME runs")
    ap.add_argument("--no-full-search", action="store_true", help="skip the full-search (BASELINE configs[0]) leg")
    ap.add_argument("--no-rdoq", action="store_true", help="skip the RDOQ (SURVEY 8 f1) leg")
    ap.add_argument("--full", action="store_true", help="encode legs at BASELINE size (cfg 2: 32 frames at QP 27/32/37; cfg 4: 32-frame "
                                                        "1080p segments) -- tens of minutes; logs of such runs are under profiles/")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
