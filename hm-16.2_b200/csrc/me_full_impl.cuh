// me_full_impl.cuh -- exhaustive integer search by one CTA per job (device code shared by the
// batch kernels in me_full.cu and the fused low-latency kernel in me_single.cu).
//
// Replaces TEncSearch::xPatternSearch (TEncSearch.cpp:3932-3989): every candidate of the
// window [L..R] x [T..B] is costed with SAD (sub-sampled rows under FEN, :3950-3956) plus
// TComRdCost::getCost(x, y) at scale 2, and the FIRST strict minimum in raster order
// (y outer, x inner, :3959-3983) wins.  The reduction therefore carries the 64-bit key
// (cost << 32 | raster index) and takes its minimum.
//
// The reference window (W + R - L) x (H' rows, all of them needed) is staged once in shared
// memory with aligned 128-bit loads; the PU block is staged next to it.  8-bit pictures use
// packed bytes: a thread owns the four candidates that share one aligned 32-bit column of the
// window, builds their byte-shifted operands with one funnel shift each and accumulates with
// VABSDIFF4.U8.ACC.  >8-bit pictures and int16 key patterns (bi-pred) take a scalar path.
#pragma once
#include "hmgpu_internal.cuh"

#define FS_THREADS 256

__device__ __forceinline__ unsigned long long fs_block_min(unsigned long long key, unsigned long long* s_red)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other < key ? other : key;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = key;
  __syncthreads();
  if (warp == 0)
  {
    key = lane < (FS_THREADS / 32) ? s_red[lane] : ~0ull;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
    {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other < key ? other : key;
    }
  }
  return key; // valid in warp 0
}

// ---- packed 8-bit path -------------------------------------------------------------------------
// Executed by one CTA of FS_THREADS threads; smem: dynamic shared memory sized by
// fs_packed_smem_bytes(); thread 0 writes *out.
__device__ __forceinline__ void full_search_block_packed(const hmgpu_me_job& jb, const RefTable& refs, const OrgView& org,
                                                         unsigned char* smem, unsigned long long* s_red, hmgpu_me_result* out)
{
  const int W = jb.pu_w, H = jb.pu_h;
  const int sub = ((jb.flags & HMGPU_F_FEN) && H > 8) ? 1 : 0;
  const int rows = H >> sub, rmul = 1 << sub;
  const int L = jb.win_l, T = jb.win_t, R = jb.win_r, B = jb.win_b;
  const int nx = R - L + 1, ny = B - T + 1;
  if (nx <= 0 || ny <= 0)
  {
    if (threadIdx.x == 0)
    {
      hmgpu_me_result r; memset(&r, 0, sizeof r);
      r.int_sad = 0xffffffffu - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, 0, 0);
      *out = r;
    }
    return;
  }
  const int pitch = refs.pitch;
  const uint8_t* ref00 = (const uint8_t*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;
  // window origin, aligned down to 16 bytes in memory
  const uint8_t* wl = ref00 + (ptrdiff_t)T * pitch + L;
  const int mis = (int)((uintptr_t)wl & 15);
  const uint8_t* wbase = wl - mis;
  const int win_w = mis + nx - 1 + W + 4;                 // bytes needed per row (+4 for the funnel look-ahead)
  const int rw16 = (win_w + 15) >> 4;                     // 16-byte chunks per row
  const int spitch = rw16 * 16 + 16;                      // odd multiple of 16 B keeps rows on different banks
  const int win_h = ny - 1 + H;
  uint32_t* s_org = (uint32_t*)smem;                      // rows * W bytes (visited rows only)
  unsigned char* s_win = smem + ((rows * W + 15) & ~15);

  // stage PU block (visited rows only) and window
  {
    const int wq = W >> 2;
    const uint8_t* o = (const uint8_t*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
    for (int i = threadIdx.x; i < rows * wq; i += FS_THREADS)
    {
      const int r = i / wq, k = i - r * wq;
      s_org[i] = __ldg((const uint32_t*)(o + (size_t)(r * rmul) * org.pitch) + k);
    }
    for (int i = threadIdx.x; i < win_h * rw16; i += FS_THREADS)
    {
      const int r = i / rw16, k = i - r * rw16;
      const uint4 v = __ldg((const uint4*)(wbase + (size_t)r * pitch) + k);
      *(uint4*)(s_win + (size_t)r * spitch + k * 16) = v;
    }
  }
  __syncthreads();

  // work items: (candidate row y, aligned 4-byte column group g). group g covers window byte
  // offsets 4g .. 4g+3, i.e. candidates x = L + 4g - mis + s, s = 0..3.
  const int wq = W >> 2;
  const int ng = (mis + nx + 3) >> 2;
  unsigned long long best = ~0ull;
  for (int it = threadIdx.x; it < ny * ng; it += FS_THREADS)
  {
    const int cy = it / ng, g = it - cy * ng;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int r = 0; r < rows; r++)
    {
      const uint32_t* wrow = (const uint32_t*)(s_win + (size_t)(cy + r * rmul) * spitch) + g;
      const uint32_t* orow = s_org + r * wq;
      uint32_t lo = wrow[0];
      for (int k = 0; k < wq; k++)
      {
        const uint32_t hi = wrow[k + 1];
        const uint32_t o = orow[k];
        a0 = vabsdiff4_acc(lo, o, a0);
        a1 = vabsdiff4_acc(__funnelshift_r(lo, hi, 8), o, a1);
        a2 = vabsdiff4_acc(__funnelshift_r(lo, hi, 16), o, a2);
        a3 = vabsdiff4_acc(__funnelshift_r(lo, hi, 24), o, a3);
        lo = hi;
      }
    }
    const int y = T + cy;
    const uint32_t acc[4] = { a0, a1, a2, a3 };
#pragma unroll
    for (int s = 0; s < 4; s++)
    {
      const int xi = 4 * g - mis + s;                      // candidate index along x
      if (xi < 0 || xi >= nx) continue;
      const int x = L + xi;
      const uint32_t cost = hm_sad_norm(acc[s], sub, 8) + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, x, y);
      const unsigned long long key = ((unsigned long long)cost << 32) | (uint32_t)(cy * nx + xi);
      best = key < best ? key : best;
    }
  }
  best = fs_block_min(best, s_red);
  if (threadIdx.x == 0)
  {
    const uint32_t idx = (uint32_t)best, cost = (uint32_t)(best >> 32);
    const int by = T + (int)(idx / nx), bx = L + (int)(idx % nx);
    hmgpu_me_result r; memset(&r, 0, sizeof r);
    r.int_x = (int16_t)bx; r.int_y = (int16_t)by;
    r.int_sad = cost - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, bx, by);
    r.n_cand = (uint32_t)(nx * ny);
    *out = r;
  }
}

// ---- generic path: Px elements, int16 key pattern ------------------------------------------------
// Executed by one CTA of FS_THREADS threads; s_org: 64*64 int16 of shared memory; thread 0 writes *out.
template <typename Px>
__device__ __forceinline__ void full_search_block_generic(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks,
                                                          const RefTable& refs, const OrgView& org, int16_t* s_org,
                                                          unsigned long long* s_red, hmgpu_me_result* out)
{
  const int W = jb.pu_w, H = jb.pu_h;
  const int sub = ((jb.flags & HMGPU_F_FEN) && H > 8) ? 1 : 0;
  const int rows = H >> sub, rmul = 1 << sub;
  const int L = jb.win_l, T = jb.win_t, R = jb.win_r, B = jb.win_b;
  const int nx = R - L + 1, ny = B - T + 1;
  if (nx <= 0 || ny <= 0)
  {
    if (threadIdx.x == 0)
    {
      hmgpu_me_result r; memset(&r, 0, sizeof r);
      r.int_sad = 0xffffffffu - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, 0, 0);
      *out = r;
    }
    return;
  }
  const int pitch = refs.pitch;
  const Px* ref00 = (const Px*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;
  if (jb.flags & HMGPU_F_ORG_BLOCK)
  {
    const int16_t* o = org_blocks + jb.org_offset;
    for (int i = threadIdx.x; i < W * H; i += FS_THREADS) s_org[i] = o[i];
  }
  else
  {
    const Px* o = (const Px*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
    for (int i = threadIdx.x; i < W * H; i += FS_THREADS)
    {
      const int r = i / W, k = i - r * W;
      s_org[i] = (int16_t)o[(size_t)r * org.pitch + k];
    }
  }
  __syncthreads();
  unsigned long long best = ~0ull;
  for (int it = threadIdx.x; it < nx * ny; it += FS_THREADS)
  {
    const int cy = it / nx, cx = it - cy * nx;
    const Px* p = ref00 + (ptrdiff_t)(T + cy) * pitch + (L + cx);
    uint32_t acc = 0;
    for (int r = 0; r < rows; r++)
    {
      const Px* pr = p + (size_t)(r * rmul) * pitch;
      const int16_t* o = s_org + (r * rmul) * W;
      for (int k = 0; k < W; k++) acc += (uint32_t)hm_abs((int)o[k] - (int)__ldg(pr + k));
    }
    const uint32_t cost = hm_sad_norm(acc, sub, refs.bit_depth) + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, L + cx, T + cy);
    const unsigned long long key = ((unsigned long long)cost << 32) | (uint32_t)it;
    best = key < best ? key : best;
  }
  best = fs_block_min(best, s_red);
  if (threadIdx.x == 0)
  {
    const uint32_t idx = (uint32_t)best, cost = (uint32_t)(best >> 32);
    const int by = T + (int)(idx / nx), bx = L + (int)(idx % nx);
    hmgpu_me_result r; memset(&r, 0, sizeof r);
    r.int_x = (int16_t)bx; r.int_y = (int16_t)by;
    r.int_sad = cost - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, bx, by);
    r.n_cand = (uint32_t)(nx * ny);
    *out = r;
  }
}

