// transform.cu -- forward core transform and scalar quantiser, batched over TUs; luma MC.
//
// hmgpu_launch_fwd_transform replaces TComTrQuant::xT -> xTrMxN (TComTrQuant.cpp:1805-1827,
// 836-885): rows then columns, shift1 = log2(N) + bitDepth + 6 - 15, shift2 = log2(N) + 6,
// round-to-nearest after each stage.  partialButterfly4/8/16/32 (:387-758) and fastForwardDst
// (:413-435) factor an exact integer matrix product with no intermediate rounding, so the
// kernel evaluates the product directly (int32, results fit 16 bits -- which is also why an
// int8 tensor-core MMA cannot reproduce it).
// hmgpu_launch_quant replaces the scalar branch of TComTrQuant::xQuant (:1120-1199).
// hmgpu_launch_mc_luma replaces the luma part of TComPrediction::xPredInterBlk (:660-698) for
// uni-prediction: a phase-plane gather.
#include "hmgpu_internal.cuh"

// magnitudes of the HEVC core transform, C[j] ~ 64*sqrt(2)*cos(j*pi/64) (H.265 8.6.4.2)
__constant__ int c_dct_mag[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                   61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
__constant__ int c_dst4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };
__constant__ int c_quant_scales[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };

__device__ __forceinline__ int dct_coef(int n, int k, int i)
{
  if (k == 0) return 64;
  int m = ((k * (32 / n)) * (2 * i + 1)) & 127;
  if (m > 64) m = 128 - m;
  return (m > 32) ? -c_dct_mag[64 - m] : c_dct_mag[m];
}

template <int N>
__global__ void __launch_bounds__(256)
fwd_transform_kernel(const int16_t* __restrict__ resi, int n_tus, int use_dst, int bit_depth, int32_t* __restrict__ coeff)
{
  constexpr int TPB = (N * N >= 256) ? 1 : 256 / (N * N);   // TUs per block
  constexpr int LOG2N = N == 4 ? 2 : N == 8 ? 3 : N == 16 ? 4 : 5;
  // rows padded by one word: the lanes of a warp read column x of 32 different rows (j = lane), which with a pitch of N words
  // is ONE bank for N = 32 (ncu, profiles/r2l_ncu_misc_kernels.csv: L1 97 % busy, 4 % issue active, 1.22 ms per 8 000 TUs)
  constexpr int P = N + 1;
  __shared__ int s_m[N * N];
  __shared__ int s_a[TPB][N * P];
  __shared__ int s_b[TPB][N * P];
  const int tid = threadIdx.x;
  for (int i = tid; i < N * N; i += 256)
    s_m[i] = (use_dst && N == 4) ? c_dst4[i / N][i % N] : dct_coef(N, i / N, i % N);
  const int tu0 = blockIdx.x * TPB;
  for (int i = tid; i < TPB * N * N; i += 256)
  {
    const int t = i / (N * N);
    const int e = i % (N * N);
    if (tu0 + t < n_tus) s_a[t][(e / N) * P + (e % N)] = (int)resi[(size_t)(tu0 + t) * N * N + e];
  }
  __syncthreads();
  const int shift1 = LOG2N + bit_depth + 6 - 15, shift2 = LOG2N + 6;
  const int rnd1 = shift1 > 0 ? (1 << (shift1 - 1)) : 0, rnd2 = 1 << (shift2 - 1);
  // stage 1: b[k*N + j] = (sum_i M[k][i] * a[j*N + i] + rnd) >> shift1
  for (int i = tid; i < TPB * N * N; i += 256)
  {
    const int t = i / (N * N), e = i % (N * N), k = e / N, j = e % N;
    int acc = 0;
#pragma unroll
    for (int x = 0; x < N; x++) acc += s_m[k * N + x] * s_a[t][j * P + x];
    s_b[t][k * P + j] = (acc + rnd1) >> shift1;
  }
  __syncthreads();
  for (int i = tid; i < TPB * N * N; i += 256)
  {
    const int t = i / (N * N), e = i % (N * N), k = e / N, j = e % N;
    if (tu0 + t >= n_tus) continue;
    int acc = 0;
#pragma unroll
    for (int x = 0; x < N; x++) acc += s_m[k * N + x] * s_b[t][j * P + x];
    coeff[(size_t)(tu0 + t) * N * N + k * N + j] = (acc + rnd2) >> shift2;
  }
}

// Inverse core transform: TComTrQuant::xIT -> xITrMxN -> partialButterflyInverse4/8/16/32 / fastInverseDst (TComTrQuant.cpp:894-960,
// 437-810), square TUs.  The partial butterflies factor the transposed integer matrix product exactly, so the kernel evaluates it
// directly: stage 1 (shift 7) clipped to the 16-bit transform dynamic range, stage 2 (shift 20 - bitDepth) clipped to the Pel range.
template <int N>
__global__ void __launch_bounds__(256)
inv_transform_kernel(const int32_t* __restrict__ coeff, int n_tus, int use_dst, int bit_depth, int16_t* __restrict__ resi)
{
  constexpr int TPB = (N * N >= 256) ? 1 : 256 / (N * N);   // TUs per block
  constexpr int P = N + 1;
  __shared__ int s_m[N * P];                                  // M[i][k], rows padded: stage loops read column k of 32 rows
  __shared__ int s_a[TPB][N * P];
  __shared__ int s_b[TPB][N * P];
  const int tid = threadIdx.x;
  for (int i = tid; i < N * N; i += 256)
    s_m[(i / N) * P + (i % N)] = (use_dst && N == 4) ? c_dst4[i / N][i % N] : dct_coef(N, i / N, i % N);
  const int tu0 = blockIdx.x * TPB;
  for (int i = tid; i < TPB * N * N; i += 256)
  {
    const int t = i / (N * N), e = i % (N * N);
    if (tu0 + t < n_tus) s_a[t][(e / N) * P + (e % N)] = coeff[(size_t)(tu0 + t) * N * N + e];
  }
  __syncthreads();
  const int shift2 = 20 - bit_depth;
  // stage 1: b[j][k] = clip((sum_i M[i][k] * a[i][j] + 64) >> 7)
  for (int i = tid; i < TPB * N * N; i += 256)
  {
    const int t = i / (N * N), e = i % (N * N), j = e / N, k = e % N;
    int acc = 0;
#pragma unroll
    for (int x = 0; x < N; x++) acc += s_m[x * P + k] * s_a[t][x * P + j];
    s_b[t][j * P + k] = min(32767, max(-32768, (acc + 64) >> 7));
  }
  __syncthreads();
  for (int i = tid; i < TPB * N * N; i += 256)
  {
    const int t = i / (N * N), e = i % (N * N), j = e / N, k = e % N;
    if (tu0 + t >= n_tus) continue;
    int acc = 0;
#pragma unroll
    for (int x = 0; x < N; x++) acc += s_m[x * P + k] * s_b[t][x * P + j];
    resi[(size_t)(tu0 + t) * N * N + j * N + k] = (int16_t)min(32767, max(-32768, (acc + (1 << (shift2 - 1))) >> shift2));
  }
}

int hmgpu_launch_inv_transform(hmgpu_ctx* ctx, const int32_t* d_coeff, int n_tus, int n, int use_dst, int16_t* d_resi)
{
  HmgpuStage st(ctx, HMGPU_ST_TRANSFORM, 1);
  switch (n)
  {
    case 4:  inv_transform_kernel<4><<<(n_tus + 15) / 16, 256, 0, ctx->stream>>>(d_coeff, n_tus, use_dst, ctx->bit_depth, d_resi); break;
    case 8:  inv_transform_kernel<8><<<(n_tus + 3) / 4, 256, 0, ctx->stream>>>(d_coeff, n_tus, 0, ctx->bit_depth, d_resi); break;
    case 16: inv_transform_kernel<16><<<n_tus, 256, 0, ctx->stream>>>(d_coeff, n_tus, 0, ctx->bit_depth, d_resi); break;
    case 32: inv_transform_kernel<32><<<n_tus, 256, 0, ctx->stream>>>(d_coeff, n_tus, 0, ctx->bit_depth, d_resi); break;
    default: return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

int hmgpu_launch_fwd_transform(hmgpu_ctx* ctx, const int16_t* d_resi, int n_tus, int n, int use_dst, int32_t* d_coeff)
{
  HmgpuStage st(ctx, HMGPU_ST_TRANSFORM, 1);
  switch (n)
  {
    case 4:  fwd_transform_kernel<4><<<(n_tus + 15) / 16, 256, 0, ctx->stream>>>(d_resi, n_tus, use_dst, ctx->bit_depth, d_coeff); break;
    case 8:  fwd_transform_kernel<8><<<(n_tus + 3) / 4, 256, 0, ctx->stream>>>(d_resi, n_tus, 0, ctx->bit_depth, d_coeff); break;
    case 16: fwd_transform_kernel<16><<<n_tus, 256, 0, ctx->stream>>>(d_resi, n_tus, 0, ctx->bit_depth, d_coeff); break;
    case 32: fwd_transform_kernel<32><<<n_tus, 256, 0, ctx->stream>>>(d_resi, n_tus, 0, ctx->bit_depth, d_coeff); break;
    default: return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

__global__ void quant_kernel(const int32_t* __restrict__ coeff, size_t total, int nn, int qbits, long long add,
                             int scale, int32_t* __restrict__ level, int32_t* __restrict__ delta_u,
                             uint32_t* __restrict__ abs_sum)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t q = 0;
  if (i < total)
  {
    const int c = coeff[i];
    const long long t = (long long)hm_abs(c) * scale;
    q = (uint32_t)((t + add) >> qbits);
    if (delta_u) delta_u[i] = (int)((t - ((long long)q << qbits)) >> (qbits - 8));
    int v = c < 0 ? -(int)q : (int)q;
    v = min(32767, max(-32768, v));
    level[i] = v;
  }
  // uiAcSum of the TU: the lanes of a warp belong to one TU (nn >= 32: a multiple of 32 coefficients per TU, blocks of 256
  // threads start on a TU boundary) or to two (4x4 TUs: the halves of the warp), so the warp adds up before the atomic -- one
  // atomic per nonzero coefficient made this kernel 421 us per 7.7 M coefficients at 4 % issue active (profiles/r2s_ncu_misc_kernels.csv)
  if (nn >= 32)
  {
    const uint32_t s = __reduce_add_sync(0xffffffffu, q);
    if (lane == 0 && s) atomicAdd(&abs_sum[i / nn], s);
  }
  else
  {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((lane & 15) == 0 && q) atomicAdd(&abs_sum[i / nn], q);
  }
}

int hmgpu_launch_quant(hmgpu_ctx* ctx, const int32_t* d_coeff, int n_tus, int n, int qp_per, int qp_rem,
                       int is_intra, int32_t* d_level, int32_t* d_delta, uint32_t* d_abs_sum)
{
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  const int transform_shift = 15 - ctx->bit_depth - log2n;   // getTransformShift, TComTrQuant.h
  const int qbits = 14 + qp_per + transform_shift;
  const long long add = (long long)(is_intra ? 171 : 85) << (qbits - 9);
  static const int scales[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };
  const size_t total = (size_t)n_tus * n * n;
  HMGPU_CUDA(ctx, cudaMemsetAsync(d_abs_sum, 0, sizeof(uint32_t) * n_tus, ctx->stream));
  HmgpuStage st(ctx, HMGPU_ST_QUANT, 1);
  quant_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_coeff, total, n * n, qbits, add, scales[qp_rem],
                                                                         d_level, d_delta, d_abs_sum);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

// xDeQuant (TComTrQuant.cpp:1203-1313), the branch without scaling lists (:1276-1311): level x g_invQuantScales[rem], rounded and
// shifted right by IQUANT_SHIFT - (transform shift + per) -- or left when that is not positive -- between the clip of the input
// to what the 32-bit intermediate carries and the clip of the output to the 16-bit transform range.  One thread per coefficient;
// the QP of a TU comes from its RDOQ job (hmgpu_residual_tus) or is the same for the whole call (hmgpu_dequant).
__global__ void dequant_kernel(const int32_t* __restrict__ level, size_t total, int nn, int transform_shift, const hmgpu_rdoq_job* __restrict__ jobs,
                               int qp_per, int qp_rem, int32_t* __restrict__ coef)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (jobs) { const hmgpu_rdoq_job* j = jobs + i / nn; qp_per = j->qp_per; qp_rem = j->qp_rem; }
  const int scale = qp_rem == 0 ? 40 : qp_rem == 1 ? 45 : qp_rem == 2 ? 51 : qp_rem == 3 ? 57 : qp_rem == 4 ? 64 : 72;     // g_invQuantScales
  const int right_shift = 6 - (transform_shift + qp_per);
  const int target = min(16, 32 + right_shift - 7);
  const int q = min((1 << (target - 1)) - 1, max(-(1 << (target - 1)), level[i]));
  int v;
  if (right_shift > 0) v = (q * scale + (1 << (right_shift - 1))) >> right_shift;
  else v = (int)((unsigned)(q * scale) << -right_shift);
  coef[i] = min(32767, max(-32768, v));
}

int hmgpu_launch_dequant(hmgpu_ctx* ctx, const int32_t* d_level, int n_tus, int n, const hmgpu_rdoq_job* d_jobs, int qp_per, int qp_rem, int32_t* d_coef)
{
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  const size_t total = (size_t)n_tus * n * n;
  HmgpuStage st(ctx, HMGPU_ST_QUANT, 1);
  dequant_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_level, total, n * n, 15 - ctx->bit_depth - log2n, d_jobs, qp_per, qp_rem, d_coef);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

template <typename Px>
__global__ void mc_luma_kernel(const hmgpu_mc_job* __restrict__ jobs, RefTable refs, int16_t* __restrict__ dst)
{
  const hmgpu_mc_job jb = jobs[blockIdx.x];
  const int ph = (jb.mv_y & 3) * 4 + (jb.mv_x & 3);
  const Px* p = (const Px*)refs.base[jb.ref_slot] + (size_t)ph * refs.plane_elems
              + (ptrdiff_t)(jb.pu_y + (jb.mv_y >> 2)) * refs.pitch + (jb.pu_x + (jb.mv_x >> 2));
  int16_t* d = dst + jb.dst_offset;
  for (int i = threadIdx.x; i < jb.pu_w * jb.pu_h; i += blockDim.x)
  {
    const int r = i / jb.pu_w, k = i - r * jb.pu_w;
    d[i] = (int16_t)p[(ptrdiff_t)r * refs.pitch + k];
  }
}

int hmgpu_launch_mc_luma(hmgpu_ctx* ctx, const hmgpu_mc_job* d_jobs, int n_jobs, int16_t* d_dst)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  HmgpuStage st(ctx, HMGPU_ST_MC, 1);
  if (ctx->px_bytes == 1) mc_luma_kernel<uint8_t><<<n_jobs, 128, 0, ctx->stream>>>(d_jobs, rt, d_dst);
  else mc_luma_kernel<uint16_t><<<n_jobs, 128, 0, ctx->stream>>>(d_jobs, rt, d_dst);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
