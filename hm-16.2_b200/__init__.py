"""hm-16.2_b200 -- B200-native inter-search hot path of the HM-16.2 HEVC reference encoder.

  csrc/        hand-written sm_100a CUDA kernels + the extern "C" ABI (include/hmgpu.h)
  libhmgpu.so  built in-tree by `make -C hm-16.2_b200`
  hmgpu.py     ctypes binding (tests, bench)
  worklist.py  HM-shaped ME job lists (CTU quadtree x partitions x references)
  synth.py     deterministic synthetic 4:2:0 video
  rdoq_batch.py  batches of TUs for hmgpu_rdoq out of the reference encoder's dumped calls
  host/        the C++ host side that plugs into HM behind the GPUME cfg switch

The directory name is not a Python identifier; import the modules with this directory on
sys.path (tests/conftest.py, bench.py) or through __graft_entry__._import_package().
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)

import hmgpu  # noqa: E402
import synth  # noqa: E402
import worklist  # noqa: E402
import rdoq_batch  # noqa: E402

__all__ = ["hmgpu", "synth", "worklist", "rdoq_batch"]
