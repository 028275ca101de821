"""The reference arm of bench.py runs without a GPU: check its JSON line against the contract the driver reads.

(The GPU arm needs a B200 and is run by the driver itself; here the keys and the bookkeeping of `--impl reference`, which times
oracle/_ref/libhmref.so -- the compiled reference -- on the host cores.)
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhmref.so")), reason="oracle/_ref not built")
def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "60000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "Gcand/s" and line["higher_is_better"] is True and line["dtype"] == "u8"
    assert line["steps"] == 1 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert 0 < line["value"] < 1.0          # a CPU: tens of Mcand/s, not Gcand/s
