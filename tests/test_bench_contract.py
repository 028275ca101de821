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


def test_rdoq_leg_bookkeeping_with_a_stand_in_context(monkeypatch):
    """bench.py's rdoq leg needs a B200 for its numbers; its bookkeeping -- the batch, the level check, the two CPU
    baselines (the oracle in one C call per pass, the reference's own xRateDistOptQuant timed inside the instrumented encoder)
    and the keys of its record -- runs here with a stand-in for the context that answers with the reference encoder's levels."""
    import importlib.util
    import numpy as np
    import hmgpu
    import rdoq_batch

    class StandIn:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def host_array(self, shape, dtype):
            return np.zeros(shape, dtype)

        def rdoq(self, jobs, bits, coef, out=None):
            _, _, _, level, abs_sum = rdoq_batch.golden_batch(len(jobs) // 1067)
            out[:] = level
            return out, abs_sum

        def profile_enable(self, on=True):
            pass

        def profile_read(self, reset=True):
            return {"quant": (4.0, 2)}

    monkeypatch.setattr(hmgpu, "Context", StandIn)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    leg = bench.rdoq_leg(0, 2, rep=2)
    assert leg["levels_identical_to_reference"] is True and leg["launches_per_batch"] == 1 and leg["ms_per_batch"] == 2.0
    assert leg["e2e"]["h2d_bytes"] > 2 * 4 * 157664 and leg["e2e"]["d2h_bytes"] > 2 * 4 * 157664
    assert leg["roofline"]["bound"] == "latency" and leg["roofline"]["hbm"]["algorithmic_bytes_per_batch"] == 8 * 2 * 157664
    cb = leg["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["mismatches"] == 0 and 1.0 < cb["value"] < 500.0
    ref = leg["cpu_reference_encoder"]
    if "unavailable" not in ref:
        assert ref["kind"] == "reference" and ref["calls"] > 10000 and 1.0 < ref["value"] < 1000.0
