"""The per-GPU broker daemon (hm-16.2_b200/hmgpud) with real clients: libhmgpu in client mode (HMGPU_BROKER) must give the
same bytes as the oracle for everything an encoder asks of it -- per-PU calls through the shared-memory mailbox (the
daemon's resident server kernel, more job lines than server CTAs, bi-pred key patterns, prediction-error lines), picture
uploads, large batches, prediction -- for several clients at once, and the daemon must survive a client that dies."""
import os
import signal
import subprocess
import sys
import time

import numpy as np
import pytest

import hmgpu
import segments
import synth
import worklist
from util import assert_results_equal, oracle_me, padded_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 416, 240


@pytest.fixture(scope="module")
def daemon():
    with segments.BrokerDaemon(device=0, ctas=4, idle_us=2000) as d:
        yield d


def _remote_context(daemon, monkeypatch, *a, **k):
    monkeypatch.setenv("HMGPU_BROKER", daemon.socket)
    ctx = hmgpu.Context(*a, **k)
    monkeypatch.delenv("HMGPU_BROKER")
    return ctx


def _frames(n, seed=5):
    return synth.luma_frames(W, H, n, 8, seed=seed).astype(np.int16)


def test_mailbox_calls_and_batches_match_oracle(daemon, monkeypatch):
    fr = _frames(4)
    pads = [padded_ref(fr[k]) for k in range(3)]
    jobs = worklist.frame_jobs(W, H, n_refs=3, seed=3)
    rng = np.random.default_rng(1)
    jobs = jobs[rng.permutation(len(jobs))[:1500]]
    exp = oracle_me(jobs, pads, fr[3], 8)
    with _remote_context(daemon, monkeypatch, W, H, 8, 3) as ctx:
        for k in range(3):
            ctx.ref_upload(k, fr[k])
        ctx.org_upload(fr[3])
        # per-PU calls of 1..13 jobs: more lines than the 4 server CTAs of a client, so CTAs walk several lines per call
        got = np.zeros(600, hmgpu.ME_RESULT)
        i = 0
        while i < 600:
            n = min(600 - i, 1 + (i * 7) % 13)
            got[i:i + n] = ctx.me_search(jobs[i:i + n])
            i += n
        assert_results_equal(got, exp[:600], jobs[:600])
        # asynchronous pair
        ctx.me_submit(jobs[600:604])
        assert_results_equal(ctx.me_wait(), exp[600:604], jobs[600:604])
        # the whole list as one batch (batch area of the segment, batch kernels in the daemon)
        assert_results_equal(ctx.me_search(jobs), exp, jobs)
        # a new picture: the server is drained, the slot re-uploaded, searches go on
        ctx.ref_upload(1, fr[0])
        e2 = oracle_me(jobs[:8], [pads[0], pads[0], pads[2]], fr[3], 8)
        assert_results_equal(ctx.me_search(jobs[:8]), e2, jobs[:8])
        assert ctx.launches > 0
        # idle for longer than the server's idle exit: the next call restarts it through the control socket
        time.sleep(0.05)
        assert_results_equal(ctx.me_search(jobs[:3]), e2[:3], jobs[:3])


def _pu_case(bit_depth=8):
    """pictures, searches and prediction-error jobs of made-up PUs, with the oracle's answers"""
    import test_gpu_predict as T
    from oracle import binding as B
    O = B.oracle()
    rng = np.random.default_rng(4)
    pics = T._pictures(bit_depth, 2, 31)
    org = T._pictures(bit_depth, 1, 41)[0][0]
    pads = [(T._pad(y, T.M), None, None) for (y, cb, cr) in pics]
    pj, _ = T._jobs(rng, 40, 2, bi_every=4)
    funcs = np.where(np.arange(40) % 3 == 0, hmgpu.DF_SAD, hmgpu.DF_HADS).astype(np.uint8)
    exp = np.zeros(40, np.uint32)
    for i, j in enumerate(pj):
        w, h = int(j["pu_w"]), int(j["pu_h"])
        pred = np.ascontiguousarray(T._oracle_component(O, pads, j, 0, bit_depth))
        blk = np.ascontiguousarray(org[int(j["pu_y"]):int(j["pu_y"]) + h, int(j["pu_x"]):int(j["pu_x"]) + w])
        exp[i] = (O.hmo_sad(B.ptr(blk), w, B.ptr(pred), w, w, h, 0, bit_depth, 0) if funcs[i] == hmgpu.DF_SAD
                  else O.hmo_hads(B.ptr(blk), w, B.ptr(pred), w, w, h, bit_depth))
    jobs = worklist.frame_jobs(W, H, n_refs=2, seed=8)[:2000:400]
    exp_me = oracle_me(jobs, [padded_ref(pics[0][0]), padded_ref(pics[1][0])], org, bit_depth)
    return pics, org, pj, funcs, exp, jobs, exp_me


def _check_pu_calls(ctx, pics, org, pj, funcs, exp, jobs, exp_me):
    for s, (y, cb, cr) in enumerate(pics):
        ctx.ref_upload(s, y)
    ctx.org_upload(org)
    # searches and prediction-error jobs of a "PU" in one mailbox round trip
    for i in range(0, 40, 10):
        ctx.pu_submit(jobs, pj[i:i + 10], funcs[i:i + 10])
        res, out = ctx.pu_wait()
        assert_results_equal(res, exp_me, jobs)
        assert out.tolist() == exp[i:i + 10].tolist(), i
    # prediction-error jobs alone, and more lines than the mailbox holds (runs inside the wait as batched calls)
    ctx.pu_submit(jobs[:0], pj[:27], funcs[:27])
    res, out = ctx.pu_wait()
    assert len(res) == 0 and out.tolist() == exp[:27].tolist()
    ctx.pu_submit(jobs, pj[:30], funcs[:30])
    res, out = ctx.pu_wait()
    assert_results_equal(res, exp_me, jobs)
    assert out.tolist() == exp[:30].tolist()
    # the batched entry point
    for f in (hmgpu.DF_SAD, hmgpu.DF_HADS):
        sel = funcs == f
        assert ctx.pred_error(pj[sel], f).tolist() == exp[sel].tolist()


def test_prediction_error_lines_through_the_broker(daemon, monkeypatch):
    case = _pu_case()
    with _remote_context(daemon, monkeypatch, W, H, 8, 2) as ctx:
        _check_pu_calls(ctx, *case)


@pytest.mark.parametrize("ctas", [16, 3])
def test_prediction_error_lines_in_process(ctas, monkeypatch):
    """the same mixed calls through the mailbox of an in-process context (own server kernel)"""
    monkeypatch.setenv("HMGPU_SERVER_CTAS", str(ctas))
    case = _pu_case()
    with hmgpu.Context(W, H, 8, 2) as ctx:
        _check_pu_calls(ctx, *case)


CLIENT = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, %(root)r + "/hm-16.2_b200"); sys.path.insert(0, %(root)r + "/tests"); sys.path.insert(0, %(root)r)
import hmgpu, synth, worklist
from util import assert_results_equal, oracle_me, padded_ref
seed, rounds = int(sys.argv[1]), int(sys.argv[2])
W, H = 416, 240
fr = synth.luma_frames(W, H, 3, 8, seed=seed).astype(np.int16)
jobs = worklist.frame_jobs(W, H, n_refs=2, seed=seed)
jobs = jobs[np.random.default_rng(seed).permutation(len(jobs))[:240]]
exp = oracle_me(jobs, [padded_ref(fr[0]), padded_ref(fr[1])], fr[2], 8)
with hmgpu.Context(W, H, 8, 2) as ctx:
    for r in range(rounds):
        ctx.ref_upload(0, fr[0]); ctx.ref_upload(1, fr[1]); ctx.org_upload(fr[2])
        for i in range(0, len(jobs), 4):
            assert_results_equal(ctx.me_search(jobs[i:i + 4]), exp[i:i + 4], jobs[i:i + 4])
        print("round", r, flush=True)
print("client", seed, "ok", flush=True)
"""


def test_many_clients_and_a_dying_one(daemon):
    env = dict(daemon.env)
    code = CLIENT % {"root": ROOT}
    procs = [subprocess.Popen([sys.executable, "-c", code, str(100 + k), "3"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for k in range(6)]
    victim = subprocess.Popen([sys.executable, "-c", code, "77", "1000"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert victim.stdout.readline().startswith("round")       # it is in the middle of its calls, server kernel resident
    victim.send_signal(signal.SIGKILL)
    victim.wait()
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0 and "ok" in out, err[-2000:]
    # the daemon is still there and takes new clients (the victim's server kernel was made to leave, its segment is reused)
    p = subprocess.run([sys.executable, "-c", code, "5", "1"], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]
    assert daemon.proc.poll() is None


def test_encoder_through_the_broker_md5(daemon):
    """the patched HM encoder attached to the daemon (no CUDA context of its own) codes the same bytes as CPU HM"""
    import encode_compare
    if not encode_compare.available():
        pytest.skip("encoder binaries not built (need /root/reference at build time)")
    import tempfile
    tmp = tempfile.mkdtemp(prefix="hmbrk_")
    yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), W, H, 6, 8)
    for cfg_name, extra in (("lowdelay_P_main", []), ("randomaccess_main", ["--DecodingRefreshType=2", "--IntraPeriod=16"])):
        cfg = os.path.join(encode_compare.CFG_DIR, "encoder_%s.cfg" % cfg_name)
        c = encode_compare.run(encode_compare.REF_ENC, cfg, yuv, W, H, 6, 32, os.path.join(tmp, "cpu"), extra)
        g = encode_compare.run(encode_compare.GPU_ENC, cfg, yuv, W, H, 6, 32, os.path.join(tmp, "gpu"), extra + ["--GPUME=1"], env=daemon.env)
        assert c["bitstream_md5"] == g["bitstream_md5"] and c["recon_md5"] == g["recon_md5"], (cfg_name, g["gpume"])
        assert "through the broker daemon" in " ".join(g["gpume"])
