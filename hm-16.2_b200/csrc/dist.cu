// dist.cu -- the TComRdCost distortion family on caller-supplied Pel blocks, one warp per item.
//
// One item = one DistParam::DistFunc call of the reference (TComRdCost.h:60-101): the function
// table m_afpDistortFunc (TComRdCost.cpp:223-276) maps to `func`:
//   SAD          xGetSAD4/8/16/32/64/12/24/48 (:493-964)  rows 0,2^s,.. then <<s, >>(bd-8)
//   SAD_GENERIC  xGetSAD (:465-491)                        every row, iSubShift ignored
//   HADS         xGetHADs (:1537-1604)                     8x8 / 4x4 / 2x2 tiles, per-tile rounding
//   SSE          xGetSSE* (:970-1315)                      each term >> 2(bd-8)
// This is the parity surface for the per-candidate values; the search kernels use their own
// fused versions of the same arithmetic.
#include "hmgpu_internal.cuh"

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int TS>
__device__ __forceinline__ uint32_t had_tile_i16(const int16_t* o, int os, const int16_t* c, int cs)
{
  int d[TS * TS];
#pragma unroll
  for (int r = 0; r < TS; r++)
#pragma unroll
    for (int k = 0; k < TS; k++) d[r * TS + k] = (int)o[r * os + k] - (int)c[r * cs + k];
  if (TS == 8) return hm_satd8x8(d);
  if (TS == 4) return hm_satd4x4(d);
  // 2x2 (xCalcHADs2x2, TComRdCost.cpp:1321-1341): no rounding
  const int m0 = d[0] + d[2], m1 = d[1] + d[3], m2 = d[0] - d[2], m3 = d[1] - d[3];
  return (uint32_t)(hm_abs(m0 + m1) + hm_abs(m0 - m1) + hm_abs(m2 + m3) + hm_abs(m2 - m3));
}

__global__ void __launch_bounds__(128)
dist_batch_kernel(const int16_t* __restrict__ org, const int16_t* __restrict__ cur,
                  const hmgpu_dist_item* __restrict__ items, int n_items, int bit_depth, uint32_t* __restrict__ out)
{
  const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= n_items) return;
  const hmgpu_dist_item it = items[item];
  const int16_t* o = org + it.org_offset;
  const int16_t* c = cur + it.cur_offset;
  const int w = it.w, h = it.h;
  uint32_t acc = 0;
  if (it.func == HMGPU_DF_SAD || it.func == HMGPU_DF_SAD_GENERIC)
  {
    const int s = it.func == HMGPU_DF_SAD ? it.sub_shift : 0;
    const int rows = (h + (1 << s) - 1) >> s;
    for (int i = lane; i < rows * w; i += 32)
    {
      const int r = (i / w) << s, k = i % w;
      acc += (uint32_t)hm_abs((int)o[r * it.org_stride + k] - (int)c[r * it.cur_stride + k]);
    }
    acc = warp_sum(acc);
    acc = (acc << s) >> (bit_depth - 8);
  }
  else if (it.func == HMGPU_DF_HADS)
  {
    const int ts = ((w & 7) == 0 && (h & 7) == 0) ? 8 : ((w & 3) == 0 && (h & 3) == 0) ? 4 : 2;
    const int tw = w / ts, nt = tw * (h / ts);
    for (int t = lane; t < nt; t += 32)
    {
      const int ty = (t / tw) * ts, tx = (t % tw) * ts;
      const int16_t* po = o + ty * it.org_stride + tx;
      const int16_t* pc = c + ty * it.cur_stride + tx;
      if (ts == 8) acc += had_tile_i16<8>(po, it.org_stride, pc, it.cur_stride);
      else if (ts == 4) acc += had_tile_i16<4>(po, it.org_stride, pc, it.cur_stride);
      else acc += had_tile_i16<2>(po, it.org_stride, pc, it.cur_stride);
    }
    acc = warp_sum(acc) >> (bit_depth - 8);
  }
  else
  {
    const int sh = (bit_depth - 8) << 1;
    for (int i = lane; i < h * w; i += 32)
    {
      const int r = i / w, k = i % w;
      const int t = (int)o[r * it.org_stride + k] - (int)c[r * it.cur_stride + k];
      acc += (uint32_t)((t * t) >> sh);
    }
    acc = warp_sum(acc);
  }
  if (lane == 0) out[item] = acc;
}

int hmgpu_launch_dist(hmgpu_ctx* ctx, const int16_t* d_org, const int16_t* d_cur,
                      const hmgpu_dist_item* d_items, int n_items, uint32_t* d_out)
{
  const int warps_per_block = 4;
  const int blocks = (n_items + warps_per_block - 1) / warps_per_block;
  HmgpuStage st(ctx, HMGPU_ST_DIST, 1);
  dist_batch_kernel<<<blocks, warps_per_block * 32, 0, ctx->stream>>>(d_org, d_cur, d_items, n_items, ctx->bit_depth, d_out);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
