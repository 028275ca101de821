"""System-level parity: the HM encoder with GPUME=1/2 (motion search in libhmgpu)
produces the same bitstream and reconstruction, byte for byte, as the unmodified CPU encoder.
GPUME=2 additionally cross-checks every xMotionEstimation call against the CPU search inside
the encoder and aborts on the first difference."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "encode_compare.py")
NEEDED = [os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRef"),
          os.path.join(ROOT, "hm-16.2_b200", "host", "build", "TAppEncoderGpu"),
          os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_lowdelay_P_main.cfg")]

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not all(os.path.exists(p) for p in NEEDED),
                                 reason="encoder binaries not built (need /root/reference at build time)")]


def _compare(*args):
    p = subprocess.run([sys.executable, SCRIPT] + list(args), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    out = json.loads(p.stdout)
    assert out["bitstream_identical"] and out["recon_identical"], out
    return out


def test_lowdelay_p_tz_crosscheck():
    out = _compare("--cfg", "lowdelay_P_main", "--frames", "3", "--gpume", "2")
    assert "calls cross-checked" in out["gpu"]["gpume"][0]
    assert " 0 calls cross-checked" not in out["gpu"]["gpume"][0]


def test_lowdelay_p_full_search():
    # BASELINE cfg 1 semantics (FastSearch=0) with a smaller range to keep the CPU run short
    _compare("--cfg", "lowdelay_P_main", "--frames", "3", "--gpume", "1", "--", "--FastSearch=0", "--SearchRange=24")


def test_lowdelay_b_bipred():
    # encoder_lowdelay_main.cfg: B slices, bi-pred refinement (2*org - pred key patterns, 9x9 full search)
    _compare("--cfg", "lowdelay_main", "--frames", "4", "--gpume", "2")


def test_randomaccess_closed_gop():
    _compare("--cfg", "randomaccess_main", "--frames", "18", "--gpume", "1", "--", "--DecodingRefreshType=2", "--IntraPeriod=16")


def test_randomaccess_main10_closed_gop():
    # BASELINE cfg 5 semantics at a small size: 10-bit pictures (uint16 planes, >>2 distortion shift), SearchRange 128
    _compare("--cfg", "randomaccess_main10", "--frames", "18", "--gpume", "2", "--bit-depth", "10", "--",
             "--DecodingRefreshType=2", "--IntraPeriod=16", "--SearchRange=128")


def test_lowdelay_p_selective_search():
    # FastSearch=2 (xTZSearchSelective): GPUME=2 cross-checks every call, GPUME=1 exercises the batched path
    _compare("--cfg", "lowdelay_P_main", "--frames", "3", "--gpume", "2", "--", "--FastSearch=2")
    _compare("--cfg", "lowdelay_P_main", "--frames", "3", "--gpume", "1", "--", "--FastSearch=2")


@pytest.mark.slow
def test_lowdelay_p_tz_1080p():
    # BASELINE cfg 2 at its picture size (1920x1080, TZSearch, 4 references), 8 frames
    out = _compare("--cfg", "lowdelay_P_main", "--size", "1920x1080", "--frames", "8", "--gpume", "1")
    assert "xMotionEstimation calls on libhmgpu" in out["gpu"]["gpume"][0]


@pytest.mark.slow
def test_randomaccess_main10_4k():
    # BASELINE cfg 5 at its picture size: 3840x2160 10-bit, SearchRange 128, closed GOP (uint16 planes of 4160 x 2320 samples)
    _compare("--cfg", "randomaccess_main10", "--size", "3840x2160", "--frames", "3", "--gpume", "1", "--bit-depth", "10", "--",
             "--DecodingRefreshType=2", "--IntraPeriod=16", "--SearchRange=128")
