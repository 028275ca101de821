// me_tz.cu -- batch TZ search kernel: one warp per job, TZ_WARPS jobs per CTA.
// Replaces TEncSearch::xTZSearch (TEncSearch.cpp:4027-4228); the search itself is in
// me_tz_impl.cuh.
#include "me_tz_impl.cuh"

template <typename Px, bool PACKED>
__global__ void __launch_bounds__(TZ_WARPS * 32, 8)
tz_search_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, const int16_t* __restrict__ org_blocks,
                 RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  // per-warp PU block: packed bytes (4 KB) or int16 (8 KB)
  __shared__ __align__(16) unsigned char s_org_all[TZ_WARPS][PACKED ? 4096 : 8192];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job_id = blockIdx.x * TZ_WARPS + warp;
  if (job_id >= n_jobs) return;
  const hmgpu_me_job jb = jobs[job_id];
  if (!(jb.flags & HMGPU_F_INTEGER) || (jb.flags & HMGPU_F_FULL)) return;
  hmgpu_me_result r;
  tz_search_warp<Px, PACKED>(jb, org_blocks, refs, org, s_org_all[warp], r);
  if (lane == 0) results[job_id] = r;
}

int hmgpu_launch_tz(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                    hmgpu_me_result* d_results, bool any_org_block)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int blocks = (n_jobs + TZ_WARPS - 1) / TZ_WARPS;
  HmgpuStage st(ctx, HMGPU_ST_TZ, 1);
  if (ctx->px_bytes == 1 && !any_org_block)
    tz_search_kernel<uint8_t, true><<<blocks, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  else if (ctx->px_bytes == 1)
    tz_search_kernel<uint8_t, false><<<blocks, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  else
    tz_search_kernel<uint16_t, false><<<blocks, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
