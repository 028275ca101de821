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
#include "predict_impl.cuh"

#define PR_THREADS 128

template <typename Px>
__global__ void __launch_bounds__(PR_THREADS)
predict_kernel(const hmgpu_pred_job* __restrict__ jobs, PredPlanes pl, int16_t* __restrict__ dst)
{
  __shared__ int16_t s_tmp[(64 + 7) * 64];
  __shared__ int16_t s_l0[64 * 64];
  __shared__ int16_t s_out[64 * 64];
  const hmgpu_pred_job jb = jobs[blockIdx.x];
  const int comp = blockIdx.y;
  predict_component<Px, PR_THREADS>(jb, comp, pl, s_tmp, s_l0, s_out);
  const int ny = jb.pu_w * jb.pu_h, nc = ny >> 2;
  int16_t* d = dst + jb.dst_offset + (comp == 0 ? 0 : (comp == 1 ? ny : ny + nc));
  const int n = comp ? nc : ny;
  for (int i = threadIdx.x; i < n; i += PR_THREADS) d[i] = s_out[i];
}

template <typename Px>
__global__ void __launch_bounds__(PR_THREADS)
pred_error_kernel(const hmgpu_pred_job* __restrict__ jobs, PredPlanes pl, OrgView org, int func, uint32_t* __restrict__ out)
{
  __shared__ int16_t s_tmp[(64 + 7) * 64];
  __shared__ int16_t s_l0[64 * 64];
  __shared__ int16_t s_out[64 * 64];
  __shared__ uint32_t s_sum;
  const hmgpu_pred_job jb = jobs[blockIdx.x];
  const uint32_t v = pred_error_block<Px, PR_THREADS>(jb, pl, org, func, s_tmp, s_l0, s_out, &s_sum);
  if (threadIdx.x == 0) out[blockIdx.x] = v;
}

PredPlanes hmgpu_pred_planes(const hmgpu_ctx* ctx)
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
  const PredPlanes pl = hmgpu_pred_planes(ctx);
  HmgpuStage st(ctx, HMGPU_ST_MC, 1);
  const dim3 grid(n_jobs, with_chroma ? 3 : 1);
  if (ctx->px_bytes == 1) predict_kernel<uint8_t><<<grid, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, d_dst);
  else predict_kernel<uint16_t><<<grid, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, d_dst);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

int hmgpu_launch_pred_error(hmgpu_ctx* ctx, const hmgpu_pred_job* d_jobs, int n_jobs, int func, uint32_t* d_out)
{
  const PredPlanes pl = hmgpu_pred_planes(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  HmgpuStage st(ctx, HMGPU_ST_MC, 1);
  if (ctx->px_bytes == 1) pred_error_kernel<uint8_t><<<n_jobs, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, ov, func, d_out);
  else pred_error_kernel<uint16_t><<<n_jobs, PR_THREADS, 0, ctx->stream>>>(d_jobs, pl, ov, func, d_out);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
