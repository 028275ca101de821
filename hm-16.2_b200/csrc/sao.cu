// sao.cu -- sample adaptive offset, encoder side: the statistics of one picture component (first half).
//
// Replaces TEncSampleAdaptiveOffset::getStatistics -> getBlkStats (TEncSampleAdaptiveOffset.cpp:312-363, 910-1340) without the
// pre-deblock sample mode (SAOLcuBoundary 0, the setting of every BASELINE cfg): for every CTU and each of the five SAO types
// (EO 0 / 90 / 135 / 45 degrees, band offset) the sum of (source - reconstruction) and the number of samples per class.  The
// decision that follows (deriveModeNewRDO / deriveModeMergeRDO: rate-distortion with CABAC estimates) stays on the host.
//
// The reference walks every line with running sign buffers; the class of a sample depends only on the sample and its two
// neighbours along the direction of the type, edgeType = sgn(c - a) + sgn(c - b), so one thread classifies one sample for all
// five types at once (nine loads of the 3x3 neighbourhood, served by L1).  Which samples a type visits depends on the
// availability of the neighbouring CTUs and on the lines skipped at the right / bottom CTU boundary; the first line of the
// diagonal types has its own range (:1129-1141, :1226-1245).  One CTA per CTU; edge classes are accumulated in registers and
// reduced with shuffles, bands with shared-memory atomics.
#include "hmgpu_internal.cuh"

#define SAO_THREADS 256

__device__ __forceinline__ int sao_sgn(int v) { return (v > 0) - (v < 0); }

__global__ void __launch_bounds__(SAO_THREADS)
sao_stats_kernel(const int16_t* __restrict__ rec, int rec_stride, const int16_t* __restrict__ org, int org_stride, int width, int height,
                 int ctu_w, int ctu_h, int ctus_x, const uint8_t* __restrict__ ctu_flags, int4 skr, int skr4, int4 skb, int skb4,
                 int bit_depth, long long* __restrict__ stats)
{
  __shared__ int s_diff[5][32];
  __shared__ int s_cnt[5][32];
  const int ctu = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int x0 = (ctu % ctus_x) * ctu_w, y0 = (ctu / ctus_x) * ctu_h;
  const int w = min(ctu_w, width - x0), h = min(ctu_h, height - y0);
  for (int i = tid; i < 5 * 32; i += SAO_THREADS) { (&s_diff[0][0])[i] = 0; (&s_cnt[0][0])[i] = 0; }
  __syncthreads();
  // left / above / above-left: the caller's (slice and tile boundaries, deriveLoopFilterBoundaryAvailibility) or the picture's;
  // right / below / above-right always from the picture geometry, as getStatistics sets them (:334-338)
  bool L = x0 > 0, A = y0 > 0, AL = x0 > 0 && y0 > 0;
  if (ctu_flags) { const int f = ctu_flags[ctu]; L = f & 1; A = f & 4; AL = f & 16; }
  const bool R = x0 + ctu_w < width, B = y0 + ctu_h < height, AR = y0 > 0 && R;
  const int sx = L ? 0 : 1;
  const int ex0 = R ? w - skr.x : w - 1, ey0 = B ? h - skb.x : h;
  const int ex1 = R ? w - skr.y : w, sy1 = A ? 0 : 1, ey1 = B ? h - skb.y : h - 1;
  const int ex2 = R ? w - skr.z : w - 1, ey2 = B ? h - skb.z : h - 1;
  const int ex3 = R ? w - skr.w : w - 1, ey3 = B ? h - skb.w : h - 1;
  const int ex4 = R ? w - skr4 : w, ey4 = B ? h - skb4 : h;
  int ed[4][5], ec[4][5];                          // edge types: per-thread sums per class
#pragma unroll
  for (int t = 0; t < 4; t++)
#pragma unroll
    for (int c = 0; c < 5; c++) { ed[t][c] = 0; ec[t][c] = 0; }
  const int16_t* rblk = rec + (size_t)y0 * rec_stride + x0;
  const int16_t* oblk = org + (size_t)y0 * org_stride + x0;
  for (int i = tid; i < w * h; i += SAO_THREADS)
  {
    const int y = i / w, x = i - y * w;
    const int16_t* p = rblk + (ptrdiff_t)y * rec_stride + x;
    const int c = p[0], d = (int)oblk[(ptrdiff_t)y * org_stride + x] - c;
    const bool in0 = y < ey0 && x >= sx && x < ex0;
    const bool in1 = y >= sy1 && y < ey1 && x < ex1;
    const bool in2 = y == 0 ? (x >= (AL ? 0 : 1) && x < (A ? ex2 : 1)) : (y < ey2 && x >= sx && x < ex2);
    const bool in3 = y == 0 ? (x >= (A ? sx : ex3) && x < ((!R && AR) ? w : ex3)) : (y < ey3 && x >= sx && x < ex3);
    // a neighbour is read only where its type visits the sample (outside, it may lie beyond the picture)
    int e[4];
    e[0] = in0 ? 2 + sao_sgn(c - p[-1]) + sao_sgn(c - p[1]) : -1;
    e[1] = in1 ? 2 + sao_sgn(c - p[-rec_stride]) + sao_sgn(c - p[rec_stride]) : -1;
    e[2] = in2 ? 2 + sao_sgn(c - p[-rec_stride - 1]) + sao_sgn(c - p[rec_stride + 1]) : -1;
    e[3] = in3 ? 2 + sao_sgn(c - p[-rec_stride + 1]) + sao_sgn(c - p[rec_stride - 1]) : -1;
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
      for (int k = 0; k < 5; k++) { const bool hit = e[t] == k; ed[t][k] += hit ? d : 0; ec[t][k] += hit ? 1 : 0; }
    if (y < ey4 && x < ex4)
    {
      const int b = c >> (bit_depth - 5);
      atomicAdd(&s_diff[4][b], d);
      atomicAdd(&s_cnt[4][b], 1);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; t++)
#pragma unroll
    for (int k = 0; k < 5; k++)
    {
      int a = ed[t][k], n = ec[t][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
      if (lane == 0) { atomicAdd(&s_diff[t][k], a); atomicAdd(&s_cnt[t][k], n); }
    }
  __syncthreads();
  // stats[ctu][type][0 = diff, 1 = count][class]
  for (int i = tid; i < 5 * 32; i += SAO_THREADS)
  {
    const int t = i >> 5, k = i & 31;
    stats[((size_t)ctu * 5 + t) * 64 + k] = (long long)s_diff[t][k];
    stats[((size_t)ctu * 5 + t) * 64 + 32 + k] = (long long)s_cnt[t][k];
  }
}

// SAO applied to a picture component: TComSampleAdaptiveOffset::offsetCTU -> offsetBlock (TComSampleAdaptiveOffset.cpp:309-620) for
// every CTU with its own type and offsets (SAOBlkParam after reconstructBlkSAOParams), one thread per sample.  A sample the
// type does not visit keeps its value (the reference copies the picture first, SAOProcess :660-680).
__global__ void __launch_bounds__(SAO_THREADS)
sao_apply_kernel(const int16_t* __restrict__ src, int stride, int width, int height, int ctu_w, int ctu_h, int ctus_x,
                 const uint8_t* __restrict__ ctu_flags, const int8_t* __restrict__ types, const int32_t* __restrict__ offsets,
                 int bit_depth, int16_t* __restrict__ dst)
{
  __shared__ int s_off[32];
  const int ctu = blockIdx.x, tid = threadIdx.x;
  const int x0 = (ctu % ctus_x) * ctu_w, y0 = (ctu / ctus_x) * ctu_h;
  const int w = min(ctu_w, width - x0), h = min(ctu_h, height - y0);
  const int type = types[ctu];
  if (tid < 32) s_off[tid] = offsets[(size_t)ctu * 32 + tid];
  __syncthreads();
  const bool cl = x0 > 0, cr = x0 + ctu_w < width, ca = y0 > 0, cb = y0 + ctu_h < height;
  int f = (cl ? 1 : 0) | (cr ? 2 : 0) | (ca ? 4 : 0) | (cb ? 8 : 0) | (cl && ca ? 16 : 0) | (cr && ca ? 32 : 0) | (cl && cb ? 64 : 0) | (cr && cb ? 128 : 0);
  if (ctu_flags) f = ctu_flags[ctu];
  const bool L = f & 1, R = f & 2, A = f & 4, B = f & 8, AL = f & 16, AR = f & 32, BL = f & 64, BR = f & 128;
  const int sx = L ? 0 : 1, ex = R ? w : w - 1, maxv = (1 << bit_depth) - 1;
  const int16_t* sblk = src + (size_t)y0 * stride + x0;
  int16_t* dblk = dst + (size_t)y0 * stride + x0;
  for (int i = tid; i < w * h; i += SAO_THREADS)
  {
    const int y = i / w, x = i - y * w;
    const int16_t* p = sblk + (ptrdiff_t)y * stride + x;
    const int c = p[0];
    bool in = false;
    int cls = 0;
    if (type == 0) { in = x >= sx && x < ex; if (in) cls = 2 + sao_sgn(c - p[-1]) + sao_sgn(c - p[1]); }
    else if (type == 1) { in = y >= (A ? 0 : 1) && y < (B ? h : h - 1); if (in) cls = 2 + sao_sgn(c - p[-stride]) + sao_sgn(c - p[stride]); }
    else if (type == 2)
    {
      if (y == 0) in = x >= (AL ? 0 : 1) && x < (A ? ex : 1);
      else if (y == h - 1) in = x >= (B ? sx : w - 1) && x < (BR ? w : w - 1);
      else in = x >= sx && x < ex;
      if (in) cls = 2 + sao_sgn(c - p[-stride - 1]) + sao_sgn(c - p[stride + 1]);
    }
    else if (type == 3)
    {
      if (y == 0) in = x >= (A ? sx : w - 1) && x < (AR ? w : w - 1);
      else if (y == h - 1) in = x >= (BL ? 0 : 1) && x < (B ? ex : 1);
      else in = x >= sx && x < ex;
      if (in) cls = 2 + sao_sgn(c - p[-stride + 1]) + sao_sgn(c - p[stride - 1]);
    }
    else if (type == 4) { in = true; cls = c >> (bit_depth - 5); }
    dblk[(ptrdiff_t)y * stride + x] = (int16_t)(in ? min(maxv, max(0, c + s_off[cls])) : c);
  }
}

int hmgpu_launch_sao_apply(hmgpu_ctx* ctx, const int16_t* d_src, int stride, int width, int height, int ctu_w, int ctu_h, const uint8_t* d_flags,
                           const int8_t* d_types, const int32_t* d_offsets, int16_t* d_dst)
{
  const int ctus_x = (width + ctu_w - 1) / ctu_w, n_ctus = ctus_x * ((height + ctu_h - 1) / ctu_h);
  HmgpuStage st(ctx, HMGPU_ST_DIST, 1);
  sao_apply_kernel<<<n_ctus, SAO_THREADS, 0, ctx->stream>>>(d_src, stride, width, height, ctu_w, ctu_h, ctus_x, d_flags, d_types, d_offsets,
                                                          ctx->bit_depth, d_dst);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

int hmgpu_launch_sao_stats(hmgpu_ctx* ctx, const int16_t* d_rec, int rec_stride, const int16_t* d_org, int org_stride, int width, int height,
                           int ctu_w, int ctu_h, const uint8_t* d_flags, const int32_t* skip_r, const int32_t* skip_b, long long* d_stats)
{
  const int ctus_x = (width + ctu_w - 1) / ctu_w, n_ctus = ctus_x * ((height + ctu_h - 1) / ctu_h);
  HmgpuStage st(ctx, HMGPU_ST_DIST, 1);
  sao_stats_kernel<<<n_ctus, SAO_THREADS, 0, ctx->stream>>>(d_rec, rec_stride, d_org, org_stride, width, height, ctu_w, ctu_h, ctus_x, d_flags,
                                                          make_int4(skip_r[0], skip_r[1], skip_r[2], skip_r[3]), skip_r[4],
                                                          make_int4(skip_b[0], skip_b[1], skip_b[2], skip_b[3]), skip_b[4], ctx->bit_depth, d_stats);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
