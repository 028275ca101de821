// me_full.cu -- batch full-search kernels: one CTA per job.  Replaces TEncSearch::xPatternSearch
// (TEncSearch.cpp:3932-3989); the search itself is in me_full_impl.cuh.
#include "me_full_impl.cuh"

__global__ void __launch_bounds__(FS_THREADS, 2)
full_search_packed_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, RefTable refs, OrgView org,
                          hmgpu_me_result* __restrict__ results)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned long long s_red[FS_THREADS / 32];
  const hmgpu_me_job jb = jobs[blockIdx.x];
  if (!(jb.flags & HMGPU_F_INTEGER) || !(jb.flags & HMGPU_F_FULL)) return;
  full_search_block_packed(jb, refs, org, smem, s_red, &results[blockIdx.x]);
}

template <typename Px>
__global__ void __launch_bounds__(FS_THREADS)
full_search_generic_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, const int16_t* __restrict__ org_blocks,
                           RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  __shared__ int16_t s_org[64 * 64];
  __shared__ unsigned long long s_red[FS_THREADS / 32];
  const hmgpu_me_job jb = jobs[blockIdx.x];
  if (!(jb.flags & HMGPU_F_INTEGER) || !(jb.flags & HMGPU_F_FULL)) return;
  full_search_block_generic<Px>(jb, org_blocks, refs, org, s_org, s_red, &results[blockIdx.x]);
}

int hmgpu_launch_full(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                      hmgpu_me_result* d_results, bool any_org_block, int max_win_bytes)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  HmgpuStage st(ctx, HMGPU_ST_FULL, 1);
  if (ctx->px_bytes == 1 && !any_org_block)
  {
    if (!(ctx->attr_done & HMGPU_ATTR_FULL))
    {
      HMGPU_CUDA(ctx, cudaFuncSetAttribute(full_search_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      ctx->attr_done |= HMGPU_ATTR_FULL;
    }
    if (max_win_bytes > 200 * 1024) return hmgpu_fail(ctx, HMGPU_E_INVALID, "full-search window needs %d bytes of shared memory", max_win_bytes);
    full_search_packed_kernel<<<n_jobs, FS_THREADS, max_win_bytes, ctx->stream>>>(d_jobs, n_jobs, rt, ov, d_results);
  }
  else if (ctx->px_bytes == 1)
    full_search_generic_kernel<uint8_t><<<n_jobs, FS_THREADS, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  else
    full_search_generic_kernel<uint16_t><<<n_jobs, FS_THREADS, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
