// planes.cu -- reference picture preparation on the device.
//
// Replaces TComPicYuv::extendPicBorder (TComPicYuv.cpp:171-215) and the per-search
// interpolation of TEncSearch::xExtDIFUpSamplingH/Q (TEncSearch.cpp:5565-5766):
//   1. pad_convert_kernel   : int16 recon -> padded (80 px replicate) integer plane in the
//                             context pixel type (uint8 at 8 bit, uint16 above)
//   2. phase_planes_kernel  : the 15 fractional planes P[fy][fx] = clip(V_fy(H_fx(ref)))
//                             with the two-pass semantics of TComInterpolationFilter.cpp:166-251
//                             (horizontal pass first, 14-bit intermediate with -8192 offset,
//                             vertical pass last with rounding and clip).
// HBM-bound: reads one plane, writes 15 (+1): algorithmic bytes = 17 * plane bytes.
#include "hmgpu_internal.cuh"

// luma taps, H.265 table 8-11 (same numbers as TComInterpolationFilter.cpp:57-63)
__constant__ int c_luma_taps[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 } };

template <typename Px>
__global__ void pad_convert_kernel(const int16_t* __restrict__ src, int src_stride, int w, int h,
                                   Px* __restrict__ dst, int pitch, int pw, int ph)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= pw || y >= ph) return;
  const int sx = min(w - 1, max(0, x - HMGPU_MARGIN));
  const int sy = min(h - 1, max(0, y - HMGPU_MARGIN));
  dst[(size_t)y * pitch + x] = (Px)src[(size_t)sy * src_stride + sx];
}

template <typename Px>
__global__ void copy_convert_kernel(const int16_t* __restrict__ src, int src_stride, int w, int h,
                                    Px* __restrict__ dst, int pitch)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(size_t)y * pitch + x] = (Px)src[(size_t)y * src_stride + x];
}

// padded chroma plane kept as int16 (used by motion compensation only)
__global__ void pad_chroma_kernel(const int16_t* __restrict__ src, int src_stride, int w, int h,
                                  int16_t* __restrict__ dst, int pitch, int pw, int ph)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= pw || y >= ph) return;
  const int sx = min(w - 1, max(0, x - HMGPU_CMARGIN));
  const int sy = min(h - 1, max(0, y - HMGPU_CMARGIN));
  dst[(size_t)y * pitch + x] = src[(size_t)sy * src_stride + sx];
}

#define PP_TX 64
#define PP_TY 16
#define PP_SW (PP_TX + 8)   // source tile width  (cols -3 .. +4, one spare)
#define PP_SH (PP_TY + 7)   // source tile height (rows -3 .. +4)

// One block = one 64x16 tile of all 16 phase planes.
// smem: source tile (int16) + 4 horizontal intermediates (int16, exact reference values).
template <typename Px>
__global__ void __launch_bounds__(256)
phase_planes_kernel(Px* __restrict__ planes, size_t plane_elems, int pitch, int pw, int ph, int bit_depth)
{
  __shared__ int16_t s_src[PP_SH][PP_SW];
  __shared__ int16_t s_h[4][PP_SH][PP_TX];

  const int x0 = blockIdx.x * PP_TX, y0 = blockIdx.y * PP_TY;
  const int tid = threadIdx.x;
  const Px* p00 = planes;

  for (int i = tid; i < PP_SH * PP_SW; i += 256)
  {
    const int r = i / PP_SW, c = i % PP_SW;
    const int sx = min(pw - 1, max(0, x0 + c - 3));
    const int sy = min(ph - 1, max(0, y0 + r - 3));
    s_src[r][c] = (int16_t)p00[(size_t)sy * pitch + sx];
  }
  __syncthreads();

  const int head = max(2, 14 - bit_depth);          // headRoom (TComInterpolationFilter.cpp:192)
  const int shift1 = 6 - head;                      // first, not last (:207-210)
  const int off1 = -(8192 << shift1);
  for (int i = tid; i < PP_SH * PP_TX; i += 256)
  {
    const int r = i / PP_TX, c = i % PP_TX;
    const int16_t* s = &s_src[r][c];                // s[0] is column x-3
    // frac 0: filterCopy(isFirst=true, isLast=false) (:111-126)
    s_h[0][r][c] = (int16_t)((int16_t)(s[3] << head) - 8192);
#pragma unroll
    for (int f = 1; f < 4; f++)
    {
      int sum = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) sum += s[k] * c_luma_taps[f][k];
      s_h[f][r][c] = (int16_t)((sum + off1) >> shift1);
    }
  }
  __syncthreads();

  const int shift2 = 6 + head;                      // not first, last (:200-205)
  const int off2 = (1 << (shift2 - 1)) + (8192 << 6);
  const int max_val = (1 << bit_depth) - 1;
  // vertical pass: one thread = FOUR horizontally adjacent samples of all 15 fractional planes, stored as one packed word
  // (uint8) or two (uint16) per plane -- byte-wide stores made this kernel LSU-bound at 14 % of the HBM rate (round 1)
  for (int i = tid; i < PP_TY * (PP_TX / 4); i += 256)
  {
    const int r = i / (PP_TX / 4), c = (i - r * (PP_TX / 4)) * 4;
    const int x = x0 + c, y = y0 + r;
    if (x >= pw || y >= ph) continue;               // pitch is a multiple of 64: a group of four never straddles the row end
#pragma unroll
    for (int fx = 0; fx < 4; fx++)
    {
      int col[8][4];
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        const short4 q = *(const short4*)&s_h[fx][r + k][c];   // rows y-3 .. y+4, 8-byte aligned (c % 4 == 0)
        col[k][0] = q.x; col[k][1] = q.y; col[k][2] = q.z; col[k][3] = q.w;
      }
#pragma unroll
      for (int fy = 0; fy < 4; fy++)
      {
        if (fx == 0 && fy == 0) continue;           // integer plane already in place
        int v[4];
#pragma unroll
        for (int e = 0; e < 4; e++)
        {
          if (fy == 0)
          {
            // filterCopy(isFirst=false, isLast=true) (:127-147)
            v[e] = (int16_t)((col[3][e] + 8192 + (1 << (head - 1))) >> head);
          }
          else
          {
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += col[k][e] * c_luma_taps[fy][k];
            v[e] = (int16_t)((sum + off2) >> shift2);
          }
          v[e] = min(max_val, max(0, v[e]));
        }
        Px* out = planes + (size_t)(fy * 4 + fx) * plane_elems + (size_t)y * pitch + x;
        if (sizeof(Px) == 1) *(uint32_t*)out = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
        else *(uint2*)out = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
      }
    }
  }
}

// ---- 8-bit pictures: the same two passes on packed operands ---------------------------------------------------------------
// The horizontal pass reads the source tile as packed bytes: the eight taps of an output are two IDP.4A (unsigned pixels x
// signed taps, the taps of a phase as two constant words), the byte alignment of the four outputs of a thread comes from funnel
// shifts of three aligned words.  The vertical pass reads the 14-bit intermediates as PAIRS of vertically adjacent rows in one
// word (the horizontal pass stores them that way): an even output row is four IDP.2A, the odd row below it five (its first and
// last tap stand alone in their pair), instead of eight IMAD each.  Same integers, same rounding: ncu had shown the kernel
// bound by instruction issue (76 % issue active, profiles/r2l_ncu_misc_kernels.csv), not by the 41 MB it writes.
__host__ __device__ constexpr uint32_t pp_pack4(int a, int b, int c, int d)
{
  return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}
__host__ __device__ constexpr int pp_tap(int f, int k)
{
  return f == 1 ? (k == 0 ? -1 : k == 1 ? 4 : k == 2 ? -10 : k == 3 ? 58 : k == 4 ? 17 : k == 5 ? -5 : k == 6 ? 1 : 0)
       : f == 2 ? (k == 0 ? -1 : k == 1 ? 4 : k == 2 ? -11 : k == 3 ? 40 : k == 4 ? 40 : k == 5 ? -11 : k == 6 ? 4 : -1)
       : f == 3 ? (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? -5 : k == 3 ? 17 : k == 4 ? 58 : k == 5 ? -10 : k == 6 ? 4 : -1)
       : (k == 3 ? 64 : 0);
}
// taps k0..k0+1 of phase F as the two low bytes of a word (IDP.2A.LO multiplies them with the low / high half of its first operand)
template <int F, int K0> struct PpTap2 { static constexpr uint32_t w = pp_pack4(K0 < 0 ? 0 : pp_tap(F, K0), K0 + 1 > 7 ? 0 : pp_tap(F, K0 + 1), 0, 0); };
template <int F> struct PpTap4 { static constexpr uint32_t lo = pp_pack4(pp_tap(F, 0), pp_tap(F, 1), pp_tap(F, 2), pp_tap(F, 3)),
                                                           hi = pp_pack4(pp_tap(F, 4), pp_tap(F, 5), pp_tap(F, 6), pp_tap(F, 7)); };

__device__ __forceinline__ int pp_dp2a(uint32_t pair, uint32_t taps, int acc)
{
  int r;
  asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(pair), "r"(taps), "r"(acc));
  return r;
}

#define PP8_SWW ((PP_TX + 12) / 4)   // source tile words per row: columns -4 .. +7 around the tile (aligned words)

template <int FY>
__device__ __forceinline__ void pp8_vertical(const uint32_t (&p)[5][4], int off2, int shift2, int (&even)[4], int (&odd)[4])
{
#pragma unroll
  for (int e = 0; e < 4; e++)
  {
    int a = off2, b = off2;
    a = pp_dp2a(p[0][e], PpTap2<FY, 0>::w, a); a = pp_dp2a(p[1][e], PpTap2<FY, 2>::w, a);
    a = pp_dp2a(p[2][e], PpTap2<FY, 4>::w, a); a = pp_dp2a(p[3][e], PpTap2<FY, 6>::w, a);
    b = pp_dp2a(p[0][e], PpTap2<FY, -1>::w, b); b = pp_dp2a(p[1][e], PpTap2<FY, 1>::w, b); b = pp_dp2a(p[2][e], PpTap2<FY, 3>::w, b);
    b = pp_dp2a(p[3][e], PpTap2<FY, 5>::w, b); b = pp_dp2a(p[4][e], PpTap2<FY, 7>::w, b);
    even[e] = min(255, max(0, a >> shift2));
    odd[e] = min(255, max(0, b >> shift2));
  }
}

__global__ void __launch_bounds__(256)
phase_planes8_kernel(uint8_t* __restrict__ planes, size_t plane_elems, int pitch, int pw, int ph)
{
  __shared__ uint32_t s_src[PP_SH][PP8_SWW];                // packed bytes, word 0 = columns x0-4 .. x0-1
  __shared__ __align__(16) uint32_t s_hp[4][(PP_SH + 1) / 2][PP_TX];      // [phase][row pair][column] = (row 2j | row 2j+1 << 16); read as uint4
  const int x0 = blockIdx.x * PP_TX, y0 = blockIdx.y * PP_TY;
  const int tid = threadIdx.x;
  for (int i = tid; i < PP_SH * PP8_SWW; i += 256)
  {
    const int r = i / PP8_SWW, wq = i - r * PP8_SWW;
    const int sy = min(ph - 1, max(0, y0 + r - 3));
    const int sx = x0 - 4 + wq * 4;
    const uint8_t* row = planes + (size_t)sy * pitch;
    uint32_t v;
    if (sx >= 0 && sx + 3 < pw) v = *(const uint32_t*)(row + sx);     // pitch and x0 are multiples of 4: aligned
    else
    {
      v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) v |= (uint32_t)row[min(pw - 1, max(0, sx + b))] << (8 * b);
    }
    s_src[r][wq] = v;
  }
  __syncthreads();
  // horizontal pass (8-bit: headRoom 6, shift 0, offset -8192): four adjacent outputs of one row per thread
  int16_t* s_h16 = (int16_t*)s_hp;
  for (int i = tid; i < PP_SH * (PP_TX / 4); i += 256)
  {
    const int r = i / (PP_TX / 4), c = (i - r * (PP_TX / 4)) * 4;
    const uint32_t w0 = s_src[r][c / 4], w1 = s_src[r][c / 4 + 1], w2 = s_src[r][c / 4 + 2];
    // output column c + j reads bytes (c + j - 3) .. (c + j + 4) = byte 1 + j of w0 onwards
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      const uint32_t lo = j < 3 ? __funnelshift_r(w0, w1, 8 * (j + 1)) : w1;
      const uint32_t hi = j < 3 ? __funnelshift_r(w1, w2, 8 * (j + 1)) : w2;
      int16_t* dst = s_h16 + ((size_t)(r >> 1) * PP_TX + c + j) * 2 + (r & 1);
      const size_t ph_stride = (size_t)((PP_SH + 1) / 2) * PP_TX * 2;
      dst[0] = (int16_t)((int)((lo >> 24) << 6) - 8192);                        // phase 0: filterCopy, sample (c + j) << 6
      dst[ph_stride] = (int16_t)(hm_dp4a_us(hi, PpTap4<1>::hi, hm_dp4a_us(lo, PpTap4<1>::lo, -8192)));
      dst[2 * ph_stride] = (int16_t)(hm_dp4a_us(hi, PpTap4<2>::hi, hm_dp4a_us(lo, PpTap4<2>::lo, -8192)));
      dst[3 * ph_stride] = (int16_t)(hm_dp4a_us(hi, PpTap4<3>::hi, hm_dp4a_us(lo, PpTap4<3>::lo, -8192)));
    }
  }
  __syncthreads();
  // vertical pass (shift 12, offset 2048 + 8192 * 64): four adjacent columns of TWO rows (even, odd) per thread
  const int shift2 = 12, off2 = (1 << 11) + (8192 << 6);
  for (int i = tid; i < (PP_TY / 2) * (PP_TX / 4); i += 256)
  {
    const int rp = i / (PP_TX / 4), c = (i - rp * (PP_TX / 4)) * 4;
    const int x = x0 + c, y = y0 + 2 * rp;
    if (x >= pw || y >= ph) continue;
    const bool odd_row = y + 1 < ph;
#pragma unroll
    for (int fx = 0; fx < 4; fx++)
    {
      uint32_t p[5][4];                             // row pairs rp .. rp + 4 (rows y-3 .. y+6)
#pragma unroll
      for (int k = 0; k < 5; k++)
      {
        const uint4 q = *(const uint4*)&s_hp[fx][rp + k][c];
        p[k][0] = q.x; p[k][1] = q.y; p[k][2] = q.z; p[k][3] = q.w;
      }
#pragma unroll
      for (int fy = 0; fy < 4; fy++)
      {
        if (fx == 0 && fy == 0) continue;           // integer plane already in place
        int ev[4], od[4];
        if (fy == 0)
        {
          // filterCopy(isFirst=false, isLast=true): rows y and y+1 are the high half of pair rp+1 and the low half of pair rp+2
#pragma unroll
          for (int e = 0; e < 4; e++)
          {
            ev[e] = min(255, max(0, ((int)(int16_t)(p[1][e] >> 16) + 8192 + 32) >> 6));
            od[e] = min(255, max(0, ((int)(int16_t)(p[2][e] & 0xffffu) + 8192 + 32) >> 6));
          }
        }
        else if (fy == 1) pp8_vertical<1>(p, off2, shift2, ev, od);
        else if (fy == 2) pp8_vertical<2>(p, off2, shift2, ev, od);
        else pp8_vertical<3>(p, off2, shift2, ev, od);
        uint8_t* out = planes + (size_t)(fy * 4 + fx) * plane_elems + (size_t)y * pitch + x;
        *(uint32_t*)out = (uint32_t)ev[0] | ((uint32_t)ev[1] << 8) | ((uint32_t)ev[2] << 16) | ((uint32_t)ev[3] << 24);
        if (odd_row) *(uint32_t*)(out + pitch) = (uint32_t)od[0] | ((uint32_t)od[1] << 8) | ((uint32_t)od[2] << 16) | ((uint32_t)od[3] << 24);
      }
    }
  }
}

template <typename Px>
static int launch_planes_t(hmgpu_ctx* ctx, int slot, const int16_t* d_src, int src_stride)
{
  Px* planes = (Px*)ctx->refs[slot].planes;
  dim3 b(32, 8), g((ctx->pw + 31) / 32, (ctx->ph + 7) / 8);
  HmgpuStage st(ctx, HMGPU_ST_PLANES, 2);
  pad_convert_kernel<Px><<<g, b, 0, ctx->stream>>>(d_src, src_stride, ctx->pic_w, ctx->pic_h, planes, ctx->pitch, ctx->pw, ctx->ph);
  dim3 g2((ctx->pw + PP_TX - 1) / PP_TX, (ctx->ph + PP_TY - 1) / PP_TY);
  if (sizeof(Px) == 1 && ctx->bit_depth == 8)
    phase_planes8_kernel<<<g2, 256, 0, ctx->stream>>>((uint8_t*)planes, ctx->plane_elems, ctx->pitch, ctx->pw, ctx->ph);
  else
    phase_planes_kernel<Px><<<g2, 256, 0, ctx->stream>>>(planes, ctx->plane_elems, ctx->pitch, ctx->pw, ctx->ph, ctx->bit_depth);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

int hmgpu_launch_planes(hmgpu_ctx* ctx, int slot, const int16_t* d_src, int src_stride)
{
  return ctx->px_bytes == 1 ? launch_planes_t<uint8_t>(ctx, slot, d_src, src_stride)
                            : launch_planes_t<uint16_t>(ctx, slot, d_src, src_stride);
}

int hmgpu_launch_chroma(hmgpu_ctx* ctx, int16_t* d_dst, const int16_t* d_src, int src_stride)
{
  dim3 b(32, 8), g((ctx->cpw + 31) / 32, (ctx->cph + 7) / 8);
  HmgpuStage st(ctx, HMGPU_ST_PLANES, 1);
  pad_chroma_kernel<<<g, b, 0, ctx->stream>>>(d_src, src_stride, ctx->pic_w / 2, ctx->pic_h / 2, d_dst, ctx->cpitch, ctx->cpw, ctx->cph);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

int hmgpu_launch_org(hmgpu_ctx* ctx, const int16_t* d_src, int src_stride)
{
  dim3 b(32, 8), g((ctx->pic_w + 31) / 32, (ctx->pic_h + 7) / 8);
  HmgpuStage st(ctx, HMGPU_ST_ORG, 1);
  if (ctx->px_bytes == 1)
    copy_convert_kernel<uint8_t><<<g, b, 0, ctx->stream>>>(d_src, src_stride, ctx->pic_w, ctx->pic_h, (uint8_t*)ctx->d_org, ctx->org_pitch);
  else
    copy_convert_kernel<uint16_t><<<g, b, 0, ctx->stream>>>(d_src, src_stride, ctx->pic_w, ctx->pic_h, (uint16_t*)ctx->d_org, ctx->org_pitch);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
