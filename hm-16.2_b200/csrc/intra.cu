// intra.cu -- intra mode pre-selection of a luma PU: the prediction of all 35 intra modes and their
// distortion against the source block, batched over PUs.
//
// Replaces the first-pass loop of TEncSearch::estIntraPredQT (TEncSearch.cpp:2352-2395: predIntraAng + distParam.DistFunc per
// mode; the mode bits and the candidate list stay on the host) with TComPrediction::predIntraAng for blocks without DPCM
// (TComPrediction.cpp:407-492): xPredIntraPlanar (:746-803), predIntraGetPredValDC (:182-225), xPredIntraAng (:245-405),
// xDCPredFiltering (:808-835) and the choice between filtered and unfiltered reference samples
// (filteringIntraReferenceSamples, TComPattern.cpp:529-554).  The reference samples themselves -- availability, substitution,
// [1 2 1] / strong smoothing, initAdiPatternChType (TComPattern.cpp:225-330) -- depend on the reconstruction of the neighbouring
// CUs and arrive from the host, both versions, as one line of 4N+1 samples each (bottom-left -> top-left -> above-right).
//
// Mapping: one CTA per PU; a work item = (mode, 8x8 tile) -- 4x4 for a 4x4 PU, the SATD tiling of xGetHADs -- whose thread
// predicts its 64 samples straight from the line in shared memory with the closed form of the mode (no intermediate block, no
// transposition for the horizontal modes), subtracts them from the source tile and runs the Hadamard transform in registers;
// one shared-memory atomic per item.  The line is addressed as m[-2N .. 2N] with m[0] = top-left corner, so that the
// reference's refAbove[k] = m[k] and refLeft[k] = m[-k].
#include "hmgpu_internal.cuh"

#define INTRA_THREADS 128

__device__ __constant__ int8_t c_intra_ang[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
__device__ __constant__ int16_t c_intra_inv[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };
__device__ __constant__ int8_t c_intra_thr[5] = { 10, 7, 1, 0, 10 };    // m_aucIntraFilter, luma, 4x4 .. 64x64

struct IntraMode
{
  const int* m;          // the line this mode reads (filtered or not), pointing at the top-left corner
  int mode, s, angle, inv;
  bool ver, edge;
};

__device__ __forceinline__ int intra_main(const IntraMode& M, int k)
{
  // main reference at index k: the line itself, or (negative angles) the side reference projected with the inverse angle:
  // refMain[k] = refSide[(128 + |k| * invAngle) >> 8]  (TComPrediction.cpp:300-305)
  return k >= 0 ? M.m[M.s * k] : M.m[-M.s * ((128 + (-k) * M.inv) >> 8)];
}

// sample (x, y) of the prediction, picture orientation
__device__ __forceinline__ int intra_sample(const IntraMode& M, int n, int lg, int x, int y, int dc, bool dc_filter, int maxv)
{
  const int* m = M.m;
  if (M.mode == 0)
    return ((n - 1 - x) * m[-(y + 1)] + (x + 1) * m[n + 1] + (n - 1 - y) * m[x + 1] + (y + 1) * m[-(n + 1)] + n) >> (lg + 1);
  if (M.mode == 1)
  {
    if (dc_filter)
    {
      if (x == 0 && y == 0) return (m[1] + m[-1] + 2 * dc + 2) >> 2;
      if (y == 0) return (m[x + 1] + 3 * dc + 2) >> 2;
      if (x == 0) return (m[-(y + 1)] + 3 * dc + 2) >> 2;
    }
    return dc;
  }
  const int u = M.ver ? x : y, v = M.ver ? y : x;     // position along / across the main reference
  if (M.angle == 0)
  {
    int p = m[M.s * (u + 1)];
    if (M.edge && u == 0) p = min(maxv, max(0, p + ((m[-M.s * (v + 1)] - m[0]) >> 1)));
    return p;
  }
  const int pos = (v + 1) * M.angle, di = pos >> 5, df = pos & 31;
  const int a = intra_main(M, u + di + 1);
  if (!df) return a;
  return ((32 - df) * a + df * intra_main(M, u + di + 2) + 16) >> 5;
}

__global__ void __launch_bounds__(INTRA_THREADS)
intra_costs_kernel(const hmgpu_intra_job* __restrict__ jobs, const int16_t* __restrict__ org_blocks, const int16_t* __restrict__ ref_lines,
                   int bit_depth, uint32_t* __restrict__ dist)
{
  __shared__ int s_line[2][4 * 64 + 1];
  __shared__ int16_t s_org[64 * 64];
  __shared__ uint32_t s_acc[35];
  __shared__ int s_dc;
  const hmgpu_intra_job jb = jobs[blockIdx.x];
  const int n = jb.size, tid = threadIdx.x;
  const int lg = 31 - __clz(n);
  for (int i = tid; i < 2 * (4 * n + 1); i += INTRA_THREADS)
    s_line[i / (4 * n + 1)][i % (4 * n + 1)] = (int)ref_lines[jb.ref_offset + i];
  for (int i = tid; i < n * n; i += INTRA_THREADS) s_org[i] = org_blocks[jb.org_offset + i];
  if (tid < 35) s_acc[tid] = 0;
  __syncthreads();
  const bool above = jb.flags & HMGPU_IF_ABOVE, left = jb.flags & HMGPU_IF_LEFT;
  if (tid < 32)
  {
    // predIntraGetPredValDC on the unfiltered samples
    const int* m = s_line[0] + 2 * n;
    int sum = 0;
    for (int i = tid; i < n; i += 32) sum += (above ? m[i + 1] : 0) + (left ? m[-(i + 1)] : 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (tid == 0) s_dc = (above && left) ? (sum + n) / (2 * n) : ((above || left) ? (sum + n / 2) / n : m[-1]);
  }
  __syncthreads();
  const int dc = s_dc;
  const bool dc_filter = above && left && n <= 16;
  const bool satd = jb.flags & HMGPU_IF_SATD;
  const int ts = n >= 8 ? 8 : 4, tpr = n / ts, tiles = tpr * tpr;
  const int maxv = (1 << bit_depth) - 1;
  for (int it = tid; it < 35 * tiles; it += INTRA_THREADS)
  {
    const int mode = it / tiles, t = it - mode * tiles;
    const int ty = (t / tpr) * ts, tx = (t - (t / tpr) * tpr) * ts;
    IntraMode M;
    M.mode = mode;
    {
      // filteringIntraReferenceSamples: never for DC, otherwise by the distance of the mode from pure horizontal / vertical
      const int diff = min(abs(mode - 10), abs(mode - 26));
      const bool filtered = !(jb.flags & HMGPU_IF_NO_SMOOTH) && mode != 1 && diff > c_intra_thr[lg - 2];
      M.m = s_line[filtered ? 1 : 0] + 2 * n;
    }
    M.ver = mode >= 18;
    const int am = M.ver ? mode - 26 : -(mode - 10);
    M.angle = mode < 2 ? 0 : (am < 0 ? -c_intra_ang[-am] : c_intra_ang[am]);
    M.inv = mode < 2 ? 0 : c_intra_inv[abs(am)];
    M.s = M.ver ? 1 : -1;
    M.edge = (jb.flags & HMGPU_IF_EDGE_FILTERS) && n <= 16;
    uint32_t v;
    if (ts == 8)
    {
      int d[64];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++)
          d[r * 8 + c] = (int)s_org[(ty + r) * n + tx + c] - intra_sample(M, n, lg, tx + c, ty + r, dc, dc_filter, maxv);
      if (satd) v = hm_satd8x8(d);
      else { v = 0; for (int i = 0; i < 64; i++) v += (uint32_t)hm_abs(d[i]); }
    }
    else
    {
      int d[16];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++)
          d[r * 4 + c] = (int)s_org[(ty + r) * n + tx + c] - intra_sample(M, n, lg, tx + c, ty + r, dc, dc_filter, maxv);
      if (satd) v = hm_satd4x4(d);
      else { v = 0; for (int i = 0; i < 16; i++) v += (uint32_t)hm_abs(d[i]); }
    }
    atomicAdd(&s_acc[mode], v);
  }
  __syncthreads();
  // xGetHADs / xGetSAD: the total is shifted by DISTORTION_PRECISION_ADJUSTMENT(bitDepth - 8)
  if (tid < 35) dist[(size_t)blockIdx.x * 35 + tid] = s_acc[tid] >> (bit_depth - 8);
}

int hmgpu_launch_intra_costs(hmgpu_ctx* ctx, const hmgpu_intra_job* d_jobs, int n_jobs, const int16_t* d_org, const int16_t* d_lines, uint32_t* d_dist)
{
  HmgpuStage st(ctx, HMGPU_ST_DIST, 1);
  intra_costs_kernel<<<n_jobs, INTRA_THREADS, 0, ctx->stream>>>(d_jobs, d_org, d_lines, ctx->bit_depth, d_dist);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
