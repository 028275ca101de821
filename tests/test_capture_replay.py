"""The encoder's REAL call stream ("a job list captured from a real CPU encode"): the patched HM encoder in
capture mode (HMGPU_CAPTURE, no GPU involved) writes every xMotionEstimation call as the hmgpu_me_job the binding would send,
with the answer of the encoder's own CPU search.  CPU test: the oracle reproduces those answers (pins the oracle AND the
job marshalling of HmGpuHost.cpp to the real encoder); GPU test: so does libhmgpu, batch by batch."""
import os
import tempfile

import numpy as np
import pytest

import capture
import hmgpu
import synth
from util import oracle_me, padded_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG_DIR = os.path.join(ROOT, "oracle", "_ref", "cfg")
pytestmark = pytest.mark.skipif(not (os.path.exists(capture.GPU_ENC) and os.path.isdir(CFG_DIR)),
                                reason="encoder binaries not built (need /root/reference at build time)")


def _capture(cfg_name, w, h, frames, extra=()):
    tmp = tempfile.mkdtemp(prefix="hmcap_")
    yuv = synth.write_yuv(os.path.join(tmp, "in.yuv"), w, h, frames, 8)
    path = os.path.join(tmp, "stream.bin")
    capture.capture_encode(os.path.join(CFG_DIR, "encoder_%s.cfg" % cfg_name), yuv, w, h, frames, 32, path, extra)
    out = capture.read_stream(path)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


@pytest.mark.parametrize("cfg_name,frames,extra", [("lowdelay_P_main", 3, []), ("lowdelay_main", 3, []),
                                                   ("lowdelay_P_main", 2, ["--FastSearch=0", "--SearchRange=16"])])
def test_oracle_reproduces_the_encoders_searches(cfg_name, frames, extra):
    w, h, bd, events = _capture(cfg_name, 416, 240, frames, extra)
    assert (w, h, bd) == (416, 240, 8)
    refs, org, n_checked = {}, None, 0
    rng = np.random.default_rng(5)
    for ev in events:
        if ev[0] == "R":
            refs[ev[1]] = padded_ref(ev[3])
        elif ev[0] == "O":
            org = ev[2]
        else:
            _, jobs, cpu, keys = ev
            pick = np.sort(rng.choice(len(jobs), min(len(jobs), 1200), replace=False))      # a sample keeps the CPU suite short
            pads = [refs.get(s) for s in range(16)]
            exp = oracle_me(jobs[pick], pads, org, 8, keys)
            for f in capture.FIELDS:
                bad = np.nonzero(exp[f] != cpu[f][pick])[0]
                assert len(bad) == 0, (f, jobs[pick][bad[0]], exp[bad[0]], cpu[pick][bad[0]])
            n_checked += len(pick)
    assert n_checked >= 1000


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_name,w,h,frames", [("lowdelay_P_main", 832, 480, 3), ("randomaccess_main", 416, 240, 9)])
def test_device_reproduces_the_encoders_searches(cfg_name, w, h, frames):
    extra = ["--DecodingRefreshType=2", "--IntraPeriod=16"] if cfg_name.startswith("random") else []
    w, h, bd, events = _capture(cfg_name, w, h, frames, extra)
    with hmgpu.Context(w, h, bd, 16) as ctx:
        out = capture.replay(ctx, events)
    assert out["jobs"] > 10000 and out["mismatches"] == 0, out
