// predict_impl.cuh -- device code of the PU prediction and the prediction-error costs, shared by the batch kernels of
// predict.cu and the resident mailbox server (me_single.cu), whose job lines may carry prediction-error jobs.
//
// Replaces, for one PU: TComPrediction::xPredInterBlk (TComPrediction.cpp:660-698) through the 8-tap luma / 4-tap chroma
// filters (TComInterpolationFilter.cpp:57-75, 166-251), TComYuv::addAvg (TComYuv.cpp:336-392), and the distortion of
// TEncSearch::xGetInterPredictionError (TEncSearch.cpp:2952-2972) / xGetTemplateCost (:3771-3811).
#pragma once
#include "hmgpu_internal.cuh"

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
template <typename S, int NT, int THREADS>
__device__ __forceinline__ void mc_pass(const S* src, int sstride, int16_t* dst, int dstride, int w, int h,
                                        int frac, bool vertical, bool is_first, bool is_last, int bit_depth)
{
  const int head = max(2, IF_PREC - bit_depth);
  const int max_val = (1 << bit_depth) - 1;
  if (frac == 0)
  {
    for (int i = threadIdx.x; i < w * h; i += THREADS)
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
  for (int i = threadIdx.x; i < w * h; i += THREADS)
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
template <typename S, int NT, int THREADS>
__device__ __forceinline__ void mc_block(const S* ref, int rstride, int mvx, int mvy, int w, int h, bool bi, int bit_depth,
                                         int16_t* tmp, int16_t* out)
{
  constexpr int SH = NT == 8 ? 2 : 3, HALF = NT / 2;
  const S* r = ref + (mvx >> SH) + (ptrdiff_t)(mvy >> SH) * rstride;
  const int fx = mvx & ((1 << SH) - 1), fy = mvy & ((1 << SH) - 1);
  if (fy == 0) mc_pass<S, NT, THREADS>(r, rstride, out, w, w, h, fx, false, true, !bi, bit_depth);
  else if (fx == 0) mc_pass<S, NT, THREADS>(r, rstride, out, w, w, h, fy, true, true, !bi, bit_depth);
  else
  {
    mc_pass<S, NT, THREADS>(r - (ptrdiff_t)(HALF - 1) * rstride, rstride, tmp, w, w, h + NT - 1, fx, false, true, false, bit_depth);
    __syncthreads();
    mc_pass<int16_t, NT, THREADS>(tmp + (HALF - 1) * w, w, out, w, w, h, fy, true, false, !bi, bit_depth);
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
template <typename Px, int THREADS>
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
      mc_block<Px, 8, THREADS>(ref, pl.pitch, jb.mv_x[l], jb.mv_y[l], w, h, bi, pl.bit_depth, s_tmp, out);
    }
    else
    {
      const int16_t* plane = comp == 1 ? pl.cb[jb.ref_slot[l]] : pl.cr[jb.ref_slot[l]];
      const int16_t* ref = plane + (ptrdiff_t)(jb.pu_y >> 1) * pl.cpitch + (jb.pu_x >> 1);
      mc_block<int16_t, 4, THREADS>(ref, pl.cpitch, jb.mv_x[l], jb.mv_y[l], w, h, bi, pl.bit_depth, s_tmp, out);
    }
  }
  if (bi)
  {
    // TComYuv::addAvg (TComYuv.cpp:336-392)
    const int shift = max(2, IF_PREC - pl.bit_depth) + 1;
    const int offset = (1 << (shift - 1)) + 2 * IF_OFFS;
    const int max_val = (1 << pl.bit_depth) - 1;
    for (int i = threadIdx.x; i < w * h; i += THREADS)
      s_out[i] = (int16_t)min(max_val, max(0, ((int)s_l0[i] + (int)s_out[i] + offset) >> shift));
    __syncthreads();
  }
}

// luma prediction + distortion against the source picture by the whole CTA: func 0 = SAD (xGetSAD*, no sub-sampling), 1 = HADS.
// s_tmp: (64 + 7) * 64 int16, s_l0 / s_out: 64 * 64 int16 each, s_sum: one word.  The result is returned to every thread.
template <typename Px, int THREADS>
__device__ __forceinline__ uint32_t pred_error_block(const hmgpu_pred_job& jb, const PredPlanes& pl, const OrgView& org, int func,
                                                     int16_t* s_tmp, int16_t* s_l0, int16_t* s_out, uint32_t* s_sum)
{
  if (threadIdx.x == 0) *s_sum = 0;
  predict_component<Px, THREADS>(jb, 0, pl, s_tmp, s_l0, s_out);      // ends with a barrier
  const int w = jb.pu_w, h = jb.pu_h;
  const Px* o = (const Px*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
  uint32_t acc = 0;
  if (func == 0)
  {
    for (int i = threadIdx.x; i < w * h; i += THREADS)
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
    for (int t = threadIdx.x; t < nt; t += THREADS)
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
  if ((threadIdx.x & 31) == 0) atomicAdd(s_sum, acc);
  __syncthreads();
  const uint32_t v = *s_sum >> (pl.bit_depth - 8);
  __syncthreads();                                          // s_sum / s_out may be rewritten by the caller's next job
  return v;
}

PredPlanes hmgpu_pred_planes(const hmgpu_ctx* ctx);
