// predict.cu -- motion compensation of a PU (luma + 4:2:0 chroma, uni- and bi-directional) and the
// prediction-error costs built on it.
//
// Replaces, for one PU:
//   TComPrediction::motionCompensation -> xPredInterUni / xPredInterBi -> xPredInterBlk
//     (TComPrediction.cpp:514-698): luma at quarter-pel through the 8-tap filters, chroma at
//     eighth-pel through the 4-tap filters (TComInterpolationFilter.cpp:57-75, 166-251), horizontal
//     pass first, vertical pass last, filterCopy semantics for integer phases;
//   bi-prediction: both lists kept at 14 bits (isLast = !bi) and combined by TComYuv::addAvg
//     (TComYuv.cpp:336-392);
//   TEncSearch::xGetInterPredictionError (TEncSearch.cpp:2952-2972, merge estimation) and
//   xGetTemplateCost's distortion (:3771-3811, AMVP): luma prediction + SAD or Hadamard SATD against
//   the source picture.
//
// The uni-directional luma result equals a phase-plane sample, but the 14-bit
// intermediates of bi-prediction and the chroma samples are not stored anywhere, so this file runs
// the filters themselves: one CTA per (PU, component), the two separable passes through shared memory.
#include "hmgpu_internal.cuh"

#define PR_THREADS 128
#define IF_PREC 14
#define IF_FILT 6
#define IF_OFFS (1 << (IF_PREC - 1))

static __constant__ int c_luma_taps[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 } };
static __constant__ int c_chroma_taps[8][4] = {
  {  0, 64,  0,  0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
  { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

// One separable pass of TComInterpolationFilter::filter<N, isVertical, isFirst, isLast> / filterCopy
// (TComInterpolationFilter.cpp:94-251), including the int16 store of the un-clipped value.
// src points at output sample (0,0); S is the element type of the source.
template <typename S, int NT>
__device__ __forceinline__ void mc_pass(const S* src, int sstride, int16_t* dst, int dstride, int w, int h,
                                        int frac, bool vertical, bool is_first, bool is_last, int bit_depth)
{
  const int head = max(2, IF_PREC - bit_depth);
  const int max_val = (1 << bit_depth) - 1;
  if (frac == 0)
  {
    for (int i = threadIdx.x; i < w * h; i += PR_THREADS)
    {
      const int y = i / w, x = i - y * w;
      const int v = (int)src[(ptrdiff_t)y * sstride + x];
      int o;
      if (is_first == is_last) o = v;
      else if (is_first) o = (int)(int16_t)(v << head) - IF_OFFS;
      else
      {
        o = (int)(int16_t)((v + IF_OFFS + (1 << (head - 1))) >> head);
        o = min(max_val, max(0, o));
      }
      dst[y * dstride + x] = (int16_t)o;
    }
    return;
  }
  const int* c = NT == 8 ? c_luma_taps[frac] : c_chroma_taps[frac];
  const int cs = vertical ? sstride : 1;
  int shift = IF_FILT, offset;
  if (is_last)
  {
    shift += is_first ? 0 : head;
    offset = 1 << (shift - 1);
    offset += is_first ? 0 : (IF_OFFS << IF_FILT);
  }
  else
  {
    shift -= is_first ? head : 0;
    offset = is_first ? -(IF_OFFS << shift) : 0;
  }
  for (int i = threadIdx.x; i < w * h; i += PR_THREADS)
  {
    const int y = i / w, x = i - y * w;
    const S* p = src + (ptrdiff_t)y * sstride + x - (NT / 2 - 1) * cs;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < NT; k++) sum += (int)p[(ptrdiff_t)k * cs] * c[k];
    int val = (int)(int16_t)((sum + offset) >> shift);
    if (is_last) val = min(max_val, max(0, val));
    dst[y * dstride + x] = (int16_t)val;
  }
}

// xPredInterBlk of one component of one list into `out` (stride w): clipped samples (bi == false) or the
// 14-bit intermediate (bi == true).  tmp: (h + NT - 1) * w int16 of shared memory.
template <typename S, int NT>
__device__ __forceinline__ void mc_block(const S* ref, int rstride, int mvx, int mvy, int w, int h, bool bi, int bit_depth,
                                         int16_t* tmp, int16_t* out)
{
  constexpr int SH = NT == 8 ? 2 : 3, HALF = NT / 2;
  const S* r = ref + (mvx >> SH) + (ptrdiff_t)(mvy >> SH) * rstride;
  const int fx = mvx & ((1 << SH) - 1), fy = mvy & ((1 << SH) - 1);
  if (fy == 0) mc_pass<S, NT>(r, rstride, out, w, w, h, fx, false, true, !bi, bit_depth);
  else if (fx == 0) mc_pass<S, NT>(r, rstride, out, w, w, h, fy, true, true, !bi, bit_depth);
  else
  {
    mc_pass<S, NT>(r - (ptrdiff_t)(HALF - 1) * rstride, rstride, tmp, w, w, h + NT - 1, fx, false, true, false, bit_depth);
    __syncthreads();
    mc_pass<int16_t, NT>(tmp + (HALF - 1) * w, w, out, w, w, h, fy, true, false, !bi, bit_depth);
  }
  __syncthreads();
}

struct PredPlanes
{
  const void* luma[HMGPU_MAX_REFS];     // sample (0,0) of the integer luma plane (Px)
  const int16_t* cb[HMGPU_MAX_REFS];    // sample (0,0) of the padded chroma planes
  const int16_t* cr[HMGPU_MAX_REFS];
  int pitch, cpitch, bit_depth;
};

// prediction of component comp (0 Y, 1 Cb, 2 Cr) of job jb into s_out (stride = component width)
template <typename Px>
__device__ __forceinline__ void predict_component(const hmgpu_pred_job& jb, int comp, const PredPlanes& pl,
                                                  int16_t* s_tmp, int16_t* s_l0, int16_t* s_out)
{
  const int w = comp ? jb.pu_w >> 1 : jb.pu_w, h = comp ? jb.pu_h >> 1 : jb.pu_h;
  const bool bi = jb.ref_slot[0] >= 0 && jb.ref_slot[1] >= 0;
  for (int l = 0; l < 2; l++)
  {
    if (jb.ref_slot[l] < 0) continue;
    int16_t* out = (bi && l == 0) ? s_l0 : s_out;
    if (comp == 0)
    {
      const Px* ref = (const Px*)pl.luma[jb.ref_slot[l]] + (ptrdiff_t)jb.pu_y * pl.pitch + jb.pu_x;
      mc_block<Px, 8>(ref, pl.pitch, jb.mv_x[l], jb.mv_y[l], w, h, bi, pl.bit_depth, s_tmp, out);
    }
    else
    {
      const int16_t* plane = comp == 1 ? pl.cb[jb.ref_slot[l]] : pl.cr[jb.ref_slot[l]];
      const int16_t* ref = plane + (ptrdiff_t)(jb.pu_y >> 1) * pl.cpitch + (jb.pu_x >> 1);
      mc_block<int16_t, 4>(ref, pl.cpitch, jb.mv_x[l], jb.mv_y[l], w, h, bi, pl.bit_depth, s_tmp, out);
    }
  }
  if (bi)
  {
    // TComYuv::addAvg (TComYuv.cpp:336-392)
    const int shift = max(2, IF_PREC - pl.bit_depth) + 1;
    const int offset = (1 << (shift - 1)) + 2 * IF_OFFS;
    const int max_val = (1 << pl.bit_depth) - 1;
    for (int i = threadIdx.x; i < w * h; i += PR_THREADS)
      s_out[i] = (int16_t)min(max_val, max(0, ((int)s_l0[i] + (int)s_out[i] + offset) >> shift));
    __syncthreads();
  }
}

template <typename Px>
__global__ void __launch_bounds__(PR_THREADS)
predict_kernel(const hmgpu_pred_job* __restrict__ jobs, PredPlanes pl, int16_t* __restrict__ dst)
{
  __shared__ int16_t s_tmp[(64 + 7) * 64];
  __shared__ int16_t s_l0[64 * 64];
  __shared__ int16_t s_out[64 * 64];
  const hmgpu_pred_job jb = jobs[blockIdx.x];
  const int comp = blockIdx.y;
  predict_component<Px>(jb, comp, pl, s_tmp, s_l0, s_out);
  const int ny = jb.pu_w * jb.pu_h, nc = ny >> 2;
  int16_t* d = dst + jb.dst_offset + (comp == 0 ? 0 : (comp == 1 ? ny : ny + nc));
  const int n = comp ? nc : ny;
  for (int i = threadIdx.x; i < n; i += PR_THREADS) d[i] = s_out[i];
}

// luma prediction + distortion against the source picture: func 0 = SAD (xGetSAD*, no sub-sampling), 1 = HADS
template <typename Px>
__global__ void __launch_bounds__(PR_THREADS)
pred_error_kernel(const hmgpu_pred_job* __restrict__ jobs, PredPlanes pl, OrgView org, int func, uint32_t* __restrict__ out)
{
  __shared__ int16_t s_tmp[(64 + 7) * 64];
  __shared__ int16_t s_l0[64 * 64];
  __shared__ int16_t s_out[64 * 64];
  __shared__ uint32_t s_sum;
  const hmgpu_pred_job jb = jobs[blockIdx.x];
  if (threadIdx.x == 0) s_sum = 0;
  predict_component<Px>(jb, 0, pl, s_tmp, s_l0, s_out);      // ends with a barrier
  const int w = jb.pu_w, h = jb.pu_h;
  const Px* o = (const Px*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
  uint32_t acc = 0;
  if (func == 0)
  {
    for (int i = threadIdx.x; i < w * h; i += PR_THREADS)
    {
      const int y = i / w, x = i - y * w;
      acc += (uint32_t)hm_abs((int)o[(size_t)y * org.pitch + x] - (int)s_out[i]);
    }
  }
  else
  {
    // xGetHADs tiling (TComRdCost.cpp:1537-1604): 8x8 tiles iff both dimensions are multiples of 8, else 4x4
    const int ts = ((w & 7) == 0 && (h & 7) == 0) ? 8 : 4;
    const int tw = w / ts, nt = tw * (h / ts);
    for (int t = threadIdx.x; t < nt; t += PR_THREADS)
    {
      const int ty = (t / tw) * ts, tx = (t - (t / tw) * tw) * ts;
      int d[64];
      if (ts == 8)
      {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int k = 0; k < 8; k++) d[r * 8 + k] = (int)o[(size_t)(ty + r) * org.pitch + tx + k] - (int)s_out[(ty + r) * w + tx + k];
        acc += hm_satd8x8(d);
      }
      else
      {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int k = 0; k < 4; k++) d[r * 4 + k] = (int)o[(size_t)(ty + r) * org.pitch + tx + k] - (int)s_out[(ty + r) * w + tx + k];
        acc += hm_satd4x4(d);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum, acc);
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = s_sum >> (pl.bit_depth - 8);
}

static PredPlanes pred_planes(const hmgpu_ctx* ctx)
{
  PredPlanes pl;
  memset(&pl, 0, sizeof pl);
  for (int i = 0; i < ctx->max_refs; i++)
  {
    if (!ctx->refs[i].valid) continue;
    pl.luma[i] = (const char*)ctx->refs[i].planes + ((size_t)HMGPU_MARGIN * ctx->pitch + HMGPU_MARGIN) * ctx->px_bytes;
    if (ctx->refs[i].cb)
    {
      pl.cb[i] = (const int16_t*)ctx->refs[i].cb + (size_t)HMGPU_CMARGIN * ctx->cpitch + HMGPU_CMARGIN;
      pl.cr[i] = (const int16_t*)ctx->refs[i].cr + (size_t)HMGPU_CMARGIN * ctx->cpitch + HMGPU_CMARGIN;
    }
  }
  pl.pitch = ctx->pitch; pl.cpitch = ctx->cpitch; pl.bit_depth = ctx->bit_depth;
  return pl;
}

int hmgpu_launch_predict(hmgpu_ctx* ctx, const hmgpu_pred_job* d_jobs, int n_jobs, int with_chroma, int16_t* d_dst)
{
  const PredPlanes pl = pred_planes(ctx);
  HmgpuStage st(ctx, HMGPU_ST_MC, 1);
  const dim3 grid(n_jobs, with_chroma ? 3 : 1);
  if (ctx->px_bytes == 1) predict_kernel<uint8_t><<<grid, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, d_dst);
  else predict_kernel<uint16_t><<<grid, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, d_dst);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

int hmgpu_launch_pred_error(hmgpu_ctx* ctx, const hmgpu_pred_job* d_jobs, int n_jobs, int func, uint32_t* d_out)
{
  const PredPlanes pl = pred_planes(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  HmgpuStage st(ctx, HMGPU_ST_MC, 1);
  if (ctx->px_bytes == 1) pred_error_kernel<uint8_t><<<n_jobs, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, ov, func, d_out);
  else pred_error_kernel<uint16_t><<<n_jobs, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, ov, func, d_out);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
