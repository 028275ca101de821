"""The C-ABI library loads and exports every symbol include/hmgpu.h declares; struct layouts of
the ctypes binding match; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hmgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "hmgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = hmgpu.lib()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libhmgpu.so does not export %s" % n
    assert set(hmgpu.EXPORTS) <= set(names)
    assert L.hmgpu_abi_version() == 3


def test_struct_layouts_match_binding():
    out = np.zeros(5, np.int32)
    hmgpu.lib().hmgpu_struct_sizes(out.ctypes.data)
    assert out.tolist() == [hmgpu.ME_JOB.itemsize, hmgpu.ME_RESULT.itemsize, hmgpu.DIST_ITEM.itemsize, hmgpu.MC_JOB.itemsize,
                            hmgpu.PRED_JOB.itemsize]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hmgpu.HmGpuError, match="no CPU fallback"):
        hmgpu.Context(416, 240)


def test_product_does_not_import_oracle():
    """the product path must never route through oracle/ (test infrastructure)"""
    pkg = os.path.join(ROOT, "hm-16.2_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle/", "liboracle", "hm_oracle", "libhmref", "hmo_"):
                    assert needle not in txt, "%s mentions %r" % (os.path.join(dirpath, f), needle)


def test_worklist_shape():
    import worklist
    jobs = worklist.frame_jobs(416, 240, n_refs=2)
    assert jobs.dtype == hmgpu.ME_JOB and len(jobs) > 1000
    assert set(np.unique(jobs["pu_w"])) <= {4, 8, 12, 16, 24, 32, 64}
    # every PU inside the picture, windows inside the clip bounds
    assert (jobs["pu_x"] + jobs["pu_w"] <= 416).all() and (jobs["pu_y"] + jobs["pu_h"] <= 240).all()
    assert ((jobs["win_r"].astype(int) << 2) <= jobs["clip_hmax"]).all()
    assert ((jobs["win_l"].astype(int) << 2) >= jobs["clip_hmin"] - 3).all()
