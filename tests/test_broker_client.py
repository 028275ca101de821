"""libhmgpu as the client of a broker daemon (csrc/remote.cu), exercised on the CPU against a MOCK daemon written here:
the wire format of csrc/broker_proto.h, the shared-memory segment layout, picture uploads through the upload area, chunking
of large batches through the batch area, refusal of the entry points a client does not have, and the loud failure when no
daemon is reachable (no private CUDA context, no CPU fallback).  The real daemon (hm-16.2_b200/hmgpud) needs a GPU: its
parity tests are in tests/test_gpu_broker.py."""
import mmap
import os
import socket
import struct
import threading

import numpy as np
import pytest

import hmgpu

MAGIC = 0x484D4742
OPS = dict(CREATE=1, SERVER_START=2, SERVER_SYNC=3, SERVER_QUERY=4, REF_UPLOAD=5, ORG_UPLOAD=6, REF_RELEASE=7, ME_BATCH=8,
           PRED_ERROR=9, PREDICT=10, SET_OPTION=11, LAUNCH_COUNT=12, DESTROY=13)
MSG = struct.Struct("<II6i32s")          # BrokerMsg, 64 bytes
REPLY = struct.Struct("<i3iQ232s")       # BrokerReply, 256 bytes
HDR = struct.Struct("<IIQQQQQQQ")        # BrokerShmHeader (first 64 of 128 bytes)
assert MSG.size == 64 and REPLY.size == 256


class MockDaemon(threading.Thread):
    """speaks the protocol for ONE client; remembers what it was asked"""

    def __init__(self, path, batch_bytes=1 << 16):
        super().__init__(daemon=True)
        self.path, self.batch_bytes = path, batch_bytes
        self.log, self.uploads, self.batches = [], [], []
        self.srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        if os.path.exists(path):
            os.unlink(path)
        self.srv.bind(path)
        self.srv.listen(1)
        self.shm_name = "/hmgpu_mock.%d" % os.getpid()

    def run(self):
        c, _ = self.srv.accept()
        shm = None
        try:
            while True:
                buf = b""
                while len(buf) < MSG.size:
                    part = c.recv(MSG.size - len(buf))
                    if not part:
                        return
                    buf += part
                magic, op, *rest = MSG.unpack(buf)
                a, text = rest[:6], rest[6].split(b"\0")[0].decode()
                assert magic == MAGIC
                self.log.append((op, a, text))
                rc, v, v64, out = 0, [0, 0, 0], 0, b""
                if op == OPS["CREATE"]:
                    self.w, self.h = a[0], a[1]
                    mail = 1 << 19
                    self.up_off, self.up_bytes = 4096 + mail, (a[0] * a[1] * 3 + 4095) // 4096 * 4096
                    self.ba_off = self.up_off + self.up_bytes
                    total = self.ba_off + self.batch_bytes
                    fd = os.open("/dev/shm" + self.shm_name, os.O_CREAT | os.O_RDWR, 0o600)
                    os.ftruncate(fd, total)
                    shm = mmap.mmap(fd, total)
                    os.close(fd)
                    shm[:HDR.size] = HDR.pack(MAGIC, 1, total, 4096, mail, self.up_off, self.up_bytes, self.ba_off, self.batch_bytes)
                    v, v64, out = [4, 2000, 0], total, self.shm_name.encode()
                elif op in (OPS["REF_UPLOAD"], OPS["ORG_UPLOAD"]):
                    n = self.w * self.h * (3 if (op == OPS["REF_UPLOAD"] and a[1]) else 2)
                    self.uploads.append((op, a[0], bytes(shm[self.up_off:self.up_off + n])))
                elif op == OPS["ME_BATCH"]:
                    n, n_org = a[0], a[1]
                    ob = (2 * n_org + 255) // 256 * 256
                    jb = (48 * n + 255) // 256 * 256
                    jobs = np.frombuffer(bytes(shm[self.ba_off + ob:self.ba_off + ob + 48 * n]), hmgpu.ME_JOB)
                    self.batches.append((n, n_org, a[2]))
                    res = np.zeros(n, hmgpu.ME_RESULT)
                    res["int_x"], res["int_y"], res["n_cand"] = jobs["pu_x"], jobs["pu_y"], a[2] + np.arange(n)   # an echo the test can check
                    shm[self.ba_off + ob + jb:self.ba_off + ob + jb + 24 * n] = res.tobytes()
                elif op == OPS["PRED_ERROR"]:
                    rc, out = -4, b"slot 3 was uploaded without chroma"      # an error text must reach hmgpu_last_error
                elif op == OPS["LAUNCH_COUNT"]:
                    v64 = 4242
                c.sendall(REPLY.pack(rc, *v, v64, out))
                if op == OPS["DESTROY"]:
                    return
        finally:
            c.close()
            self.srv.close()
            if shm is not None:
                shm.close()
                os.unlink("/dev/shm" + self.shm_name)
            os.unlink(self.path)


@pytest.fixture
def mock(tmp_path, monkeypatch):
    d = MockDaemon(str(tmp_path / "mock.sock"))
    d.start()
    monkeypatch.setenv("HMGPU_BROKER", d.path)
    yield d
    d.join(timeout=5)


def test_client_speaks_the_protocol(mock):
    import worklist
    w, h = 64, 32
    with hmgpu.Context(w, h, 8, 4) as ctx:
        assert [m[0] for m in mock.log] == [OPS["CREATE"]] and mock.log[0][1][:5] == [w, h, 8, 4, 1]
        # uploads: strided rows arrive tight in the upload area; the slot becomes valid on the client side
        luma = (np.arange(w * h, dtype=np.int16) % 251).reshape(h, w)
        cb = np.full((h // 2, w // 2), 7, np.int16)
        cr = np.full((h // 2, w // 2), 9, np.int16)
        ctx.ref_upload(2, luma, cb, cr)
        ctx.org_upload(luma[::-1])
        (op0, slot0, b0), (op1, _, b1) = mock.uploads
        assert op0 == OPS["REF_UPLOAD"] and slot0 == 2 and b0 == luma.tobytes() + cb.tobytes() + cr.tobytes()
        assert op1 == OPS["ORG_UPLOAD"] and b1 == np.ascontiguousarray(luma[::-1]).tobytes()
        # a batch larger than the mailbox goes through the batch area, cut to what the area holds
        jobs = np.zeros(2000, hmgpu.ME_JOB)
        jobs["pu_w"] = jobs["pu_h"] = 8
        jobs["pu_x"] = (np.arange(2000) % 7) * 8
        jobs["pu_y"] = (np.arange(2000) % 3) * 8
        jobs["ref_slot"] = 2
        jobs["flags"] = hmgpu.F_INTEGER | hmgpu.F_FRAC
        jobs["search_range"] = 8
        bd = hmgpu.clip_bounds(w, h, 0, 0)
        jobs["clip_hmin"], jobs["clip_hmax"], jobs["clip_vmin"], jobs["clip_vmax"] = bd
        jobs["win_l"], jobs["win_t"], jobs["win_r"], jobs["win_b"] = hmgpu.search_range(bd, 0, 0, 8)
        jobs["ui_cost"] = 1000
        res = ctx.me_search(jobs)
        assert len(mock.batches) > 1 and sum(b[0] for b in mock.batches) == 2000
        assert (res["int_x"] == jobs["pu_x"]).all() and (res["int_y"] == jobs["pu_y"]).all()
        assert (res["n_cand"] == np.arange(2000)).all()
        # jobs are validated on the client before anything travels (the slot must have been uploaded through THIS context)
        bad = jobs[:40].copy()
        bad["ref_slot"] = 1
        with pytest.raises(hmgpu.HmGpuError, match="not uploaded"):
            ctx.me_search(bad)
        # the daemon's error text reaches hmgpu_last_error
        pj = np.zeros(1, hmgpu.PRED_JOB)
        pj["pu_w"] = pj["pu_h"] = 8
        pj["ref_slot"] = [[2, -1]]
        with pytest.raises(hmgpu.HmGpuError, match="uploaded without chroma"):
            ctx.pred_error(pj, hmgpu.DF_HADS)
        assert ctx.launches == 4242
        # measurement / device-pointer entry points do not exist for a client
        with pytest.raises(hmgpu.HmGpuError, match="not available through the broker"):
            ctx.ref_upload_device(0, 1234, w)
        with pytest.raises(hmgpu.HmGpuError, match="not available through the broker"):
            ctx.microbench(0)
        assert ctx.stream is None
    assert mock.log[-1][0] == OPS["DESTROY"]


def test_no_daemon_no_fallback(tmp_path, monkeypatch):
    monkeypatch.setenv("HMGPU_BROKER", str(tmp_path / "nobody.sock"))
    monkeypatch.setenv("HMGPU_BROKER_WAIT_MS", "50")
    with pytest.raises(hmgpu.HmGpuError, match="cannot reach the broker daemon.*no CPU fallback"):
        hmgpu.Context(416, 240)
