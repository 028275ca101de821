// me_frac_impl.cuh -- device code of the fractional refinement shared by the batch kernels
// (me_frac.cu) and the fused low-latency kernel (me_single.cu): candidate tables of
// s_acMvRefineH / s_acMvRefineQ (TEncSearch.cpp:51-75), row loaders and the distortion of one
// 8x8 / 4x4 tile (SATD per xCalcHADs8x8/4x4, TComRdCost.cpp:1343-1534, or SAD).
#pragma once
#include "hmgpu_internal.cuh"

static __constant__ int8_t c_refine_h[9][2] = { {0,0},{0,-1},{0,1},{-1,0},{1,0},{-1,-1},{1,-1},{-1,1},{1,1} };
static __constant__ int8_t c_refine_q[9][2] = { {0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1} };

// work item: job index (26 bits) | tile index (6 bits)
#define WORK_TILE_BITS 6

__device__ __forceinline__ int job_tile_size(const hmgpu_me_job& j) { return ((j.pu_w & 7) == 0 && (j.pu_h & 7) == 0) ? 8 : 4; }

// load n (4 or 8) consecutive pixels of a row at an arbitrary address into ints
template <int N>
__device__ __forceinline__ void load_row(const uint8_t* p, int* out)
{
  const uintptr_t a = (uintptr_t)p;
  const int sh = (int)(a & 3) * 8;
  const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
  uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
  const uint32_t v0 = __funnelshift_r(w0, w1, sh);
  out[0] = v0 & 0xff; out[1] = (v0 >> 8) & 0xff; out[2] = (v0 >> 16) & 0xff; out[3] = v0 >> 24;
  if (N == 8)
  {
    const uint32_t w2 = __ldg(q + 2);
    const uint32_t v1 = __funnelshift_r(w1, w2, sh);
    out[4] = v1 & 0xff; out[5] = (v1 >> 8) & 0xff; out[6] = (v1 >> 16) & 0xff; out[7] = v1 >> 24;
  }
}
template <int N>
__device__ __forceinline__ void load_row(const uint16_t* p, int* out)
{
#pragma unroll
  for (int k = 0; k < N; k++) out[k] = (int)__ldg(p + k);
}
template <int N>
__device__ __forceinline__ void load_row(const int16_t* p, int* out)
{
#pragma unroll
  for (int k = 0; k < N; k++) out[k] = (int)p[k];
}

// distortion of one TS x TS tile: d = org - ref, SATD or SAD
template <typename Px, int TS>
__device__ __forceinline__ uint32_t tile_dist(const hmgpu_me_job& jb, const int16_t* org_blocks, const OrgView& org,
                                              const Px* ref, int pitch, int tx, int ty, bool satd)
{
  int d[TS * TS];
  if (jb.flags & HMGPU_F_ORG_BLOCK)
  {
    const int16_t* o = org_blocks + jb.org_offset + (size_t)ty * jb.pu_w + tx;
#pragma unroll
    for (int r = 0; r < TS; r++) load_row<TS>(o + (size_t)r * jb.pu_w, d + r * TS);
  }
  else
  {
    const Px* o = (const Px*)org.base + (size_t)(jb.pu_y + ty) * org.pitch + jb.pu_x + tx;
#pragma unroll
    for (int r = 0; r < TS; r++) load_row<TS>(o + (size_t)r * org.pitch, d + r * TS);
  }
#pragma unroll
  for (int r = 0; r < TS; r++)
  {
    int v[TS];
    load_row<TS>(ref + (ptrdiff_t)(ty + r) * pitch + tx, v);
#pragma unroll
    for (int k = 0; k < TS; k++) d[r * TS + k] -= v[k];
  }
  if (satd) return TS == 8 ? hm_satd8x8(d) : hm_satd4x4(d);
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < TS * TS; i++) s += (uint32_t)hm_abs(d[i]);
  return s;
}

// ---- packed-byte SATD of one tile (shared by me_frac2.cu and me_fracw.cu) ---------------------------------------------
// sign patterns of the 8-point Hadamard (Sylvester order): s_kj = (-1)^popc(k & j); +1 -> 0x01, -1 -> 0xff
__host__ __device__ constexpr uint32_t had_pat4(int k, bool neg)
{
  uint32_t w = 0;
  for (int j = 0; j < 4; j++)
  {
    int s = 0;
    for (int b = 0; b < 3; b++) s ^= ((k >> b) & (j >> b) & 1);
    const bool minus = (s != 0) != neg;
    w |= (minus ? 0xffu : 0x01u) << (8 * j);
  }
  return w;
}
template <int K, bool NEG> struct HadPat
{
  static constexpr uint32_t lo = had_pat4(K, NEG);
  static constexpr uint32_t hi = had_pat4(K, NEG != ((K & 4) != 0));
};

// out[k] = init[k] + (NEG ? -1 : 1) * sum_j s_kj p_j  for one row of 8 packed pixels (w0 = p0..3, w1 = p4..7)
template <bool NEG>
__device__ __forceinline__ void had_row8(uint32_t w0, uint32_t w1, const int* init, int* out)
{
  out[0] = hm_dp4a_us(w1, HadPat<0, NEG>::hi, hm_dp4a_us(w0, HadPat<0, NEG>::lo, init[0]));
  out[1] = hm_dp4a_us(w1, HadPat<1, NEG>::hi, hm_dp4a_us(w0, HadPat<1, NEG>::lo, init[1]));
  out[2] = hm_dp4a_us(w1, HadPat<2, NEG>::hi, hm_dp4a_us(w0, HadPat<2, NEG>::lo, init[2]));
  out[3] = hm_dp4a_us(w1, HadPat<3, NEG>::hi, hm_dp4a_us(w0, HadPat<3, NEG>::lo, init[3]));
  out[4] = hm_dp4a_us(w1, HadPat<4, NEG>::hi, hm_dp4a_us(w0, HadPat<4, NEG>::lo, init[4]));
  out[5] = hm_dp4a_us(w1, HadPat<5, NEG>::hi, hm_dp4a_us(w0, HadPat<5, NEG>::lo, init[5]));
  out[6] = hm_dp4a_us(w1, HadPat<6, NEG>::hi, hm_dp4a_us(w0, HadPat<6, NEG>::lo, init[6]));
  out[7] = hm_dp4a_us(w1, HadPat<7, NEG>::hi, hm_dp4a_us(w0, HadPat<7, NEG>::lo, init[7]));
}
template <bool NEG>
__device__ __forceinline__ void had_row4(uint32_t w0, const int* init, int* out)
{
  out[0] = hm_dp4a_us(w0, HadPat<0, NEG>::lo, init[0]);
  out[1] = hm_dp4a_us(w0, HadPat<1, NEG>::lo, init[1]);
  out[2] = hm_dp4a_us(w0, HadPat<2, NEG>::lo, init[2]);
  out[3] = hm_dp4a_us(w0, HadPat<3, NEG>::lo, init[3]);
}

// vertical 8-point pass over the 8 columns of d[64] + sum of absolute values, rounding (s+2)>>2
__device__ __forceinline__ uint32_t had_cols8_abs(int* d)
{
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++)
  {
    int* v = d + c;
#pragma unroll
    for (int i = 0; i < 4; i++) { const int a = v[i * 8], b = v[(i + 4) * 8]; v[i * 8] = a + b; v[(i + 4) * 8] = a - b; }
#pragma unroll
    for (int i = 0; i < 8; i += 4)
#pragma unroll
      for (int j = i; j < i + 2; j++) { const int a = v[j * 8], b = v[(j + 2) * 8]; v[j * 8] = a + b; v[(j + 2) * 8] = a - b; }
#pragma unroll
    for (int i = 0; i < 8; i += 2) s += 2u * (uint32_t)max(hm_abs(v[i * 8]), hm_abs(v[(i + 1) * 8]));
  }
  return (s + 2) >> 2;
}
__device__ __forceinline__ uint32_t had_cols4_abs(int* d)
{
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
  {
    int* v = d + c;
    const int a0 = v[0] + v[8], a1 = v[4] + v[12], a2 = v[0] - v[8], a3 = v[4] - v[12];
    s += 2u * (uint32_t)max(hm_abs(a0), hm_abs(a1)) + 2u * (uint32_t)max(hm_abs(a2), hm_abs(a3));
  }
  return (s + 1) >> 1;
}

