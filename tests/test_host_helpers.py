"""The scalar host-side helpers of the C ABI against the oracle (CPU only: they need no GPU).

hmgpu_clip_bounds / hmgpu_clip_bounds_ctu  = TComDataCU::clipMv bounds (TComDataCU.cpp:2917-2929)
hmgpu_search_range                         = TEncSearch::xSetSearchRange (TEncSearch.cpp:3911-3927)
hmgpu_mv_bits / hmgpu_mv_cost              = TComRdCost::getBits / getCost(x, y) (TComRdCost.h:171-188)

The GPU parity tests build their jobs WITH these helpers, so they only prove the device and the oracle agree on whatever
window the helper produced; here the helpers themselves are pinned.
"""
import numpy as np

import hmgpu
from oracle import binding as B


def test_clip_bounds_match_oracle():
    O = B.oracle()
    rng = np.random.default_rng(11)
    for _ in range(5000):
        pic_w = int(rng.integers(2, 2046)) * 4
        pic_h = int(rng.integers(2, 2046)) * 4
        cu_x = int(rng.integers(0, pic_w // 8 + 1)) * 8
        cu_y = int(rng.integers(0, pic_h // 8 + 1)) * 8
        exp = np.zeros(4, np.int32)
        O.hmo_clip_bounds(pic_w, pic_h, cu_x, cu_y, exp)
        got = hmgpu.clip_bounds(pic_w, pic_h, cu_x, cu_y)
        assert got.tolist() == exp.tolist(), (pic_w, pic_h, cu_x, cu_y)
        assert hmgpu.clip_bounds(pic_w, pic_h, cu_x, cu_y, max_cu=64).tolist() == exp.tolist()
    # the widest picture hmgpu_create admits still fits int16 quarter-pel bounds
    b = hmgpu.clip_bounds(8184, 8184, 0, 0)
    assert b[1] == (8184 + 7) * 4 and b[3] == (8184 + 7) * 4
    # smaller CTUs move the low bounds only (g_uiMaxCUWidth in clipMv)
    b64, b32 = hmgpu.clip_bounds(416, 240, 64, 32, max_cu=64), hmgpu.clip_bounds(416, 240, 64, 32, max_cu=32)
    assert b32[0] - b64[0] == 32 * 4 and b32[2] - b64[2] == 32 * 4 and b32[1] == b64[1] and b32[3] == b64[3]


def test_search_range_matches_oracle():
    O = B.oracle()
    rng = np.random.default_rng(12)
    for k in range(20000):
        pic_w = int(rng.integers(4, 1024)) * 4
        pic_h = int(rng.integers(4, 600)) * 4
        cu_x = int(rng.integers(0, pic_w // 8)) * 8
        cu_y = int(rng.integers(0, pic_h // 8)) * 8
        # predictors well inside, near and far outside the clip bounds (clipMv runs twice in xSetSearchRange)
        spread = [64, 1024, 20000][k % 3]
        pred_x = int(rng.integers(-spread, spread + 1))
        pred_y = int(rng.integers(-spread, spread + 1))
        sr = int(rng.choice([1, 4, 8, 16, 64, 128, 256]))
        exp = np.zeros(4, np.int32)
        O.hmo_set_search_range(pic_w, pic_h, cu_x, cu_y, pred_x, pred_y, sr, exp)
        got = hmgpu.search_range(hmgpu.clip_bounds(pic_w, pic_h, cu_x, cu_y), pred_x, pred_y, sr)
        assert got.tolist() == exp.tolist(), (pic_w, pic_h, cu_x, cu_y, pred_x, pred_y, sr)


def test_mv_bits_and_cost_match_oracle():
    O = B.oracle()
    rng = np.random.default_rng(13)
    for k in range(20000):
        scale = int(rng.integers(0, 3))
        pred_x, pred_y = int(rng.integers(-2048, 2048)), int(rng.integers(-2048, 2048))
        lim = 2048 >> scale
        x, y = int(rng.integers(-lim, lim)), int(rng.integers(-lim, lim))
        # m_uiCost from small lambdas up to values whose product with the bit count wraps around 32 bits
        ui_cost = int(rng.integers(1, 1 << [12, 20, 28, 32][k % 4]))
        assert hmgpu.mv_bits(pred_x, pred_y, scale, x, y) == O.hmo_mv_bits(pred_x, pred_y, scale, x, y)
        assert hmgpu.mv_cost(ui_cost, pred_x, pred_y, scale, x, y) == O.hmo_mv_cost(ui_cost, pred_x, pred_y, scale, x, y)


def test_set_option_needs_a_context():
    # the entry point exists and rejects a NULL context without touching CUDA
    assert hmgpu.lib().hmgpu_set_option(None, b"tz_thread", 0) == -1
