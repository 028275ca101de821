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

template <typename Px>
static int launch_planes_t(hmgpu_ctx* ctx, int slot, const int16_t* d_src, int src_stride)
{
  Px* planes = (Px*)ctx->refs[slot].planes;
  dim3 b(32, 8), g((ctx->pw + 31) / 32, (ctx->ph + 7) / 8);
  HmgpuStage st(ctx, HMGPU_ST_PLANES, 2);
  pad_convert_kernel<Px><<<g, b, 0, ctx->stream>>>(d_src, src_stride, ctx->pic_w, ctx->pic_h, planes, ctx->pitch, ctx->pw, ctx->ph);
  dim3 g2((ctx->pw + PP_TX - 1) / PP_TX, (ctx->ph + PP_TY - 1) / PP_TY);
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
