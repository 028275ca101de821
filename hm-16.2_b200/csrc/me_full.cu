// me_full.cu -- batch full-search kernels: one CTA per job.  Replaces TEncSearch::xPatternSearch
// (TEncSearch.cpp:3932-3989); the search itself is in me_full_impl.cuh.
#include <cuda.h>
#include "me_full_impl.cuh"
#include <stdlib.h>

// TMA = true: the window is staged by cp.async.bulk.tensor (maps: one tensor map per row width, passed by value -- TMA fetches
// descriptors from the parameter bank)
template <bool TMA>
__global__ void __launch_bounds__(FS_THREADS, 2)
full_search_packed_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, RefTable refs, OrgView org,
                          hmgpu_me_result* __restrict__ results, const __grid_constant__ FsMaps maps)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long s_red[FS_THREADS / 32];
  __shared__ unsigned long long s_bar;
  const hmgpu_me_job jb = jobs[blockIdx.x];
  if (!(jb.flags & HMGPU_F_INTEGER) || !(jb.flags & HMGPU_F_FULL)) return;
  full_search_block_packed(jb, refs, org, smem, s_red, &results[blockIdx.x], TMA ? &maps : NULL, &s_bar);
}

typedef CUresult (*FsEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the tensor maps of a context, built once: the planes keep their address for the life of the context (ref_alloc, api.cu)
static int fs_build_maps(hmgpu_ctx* ctx)
{
  static FsEncodeTiled s_encode = NULL;
  if (!s_encode)
  {
    void* fn = NULL;
    cudaDriverEntryPointQueryResult q;
    HMGPU_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return hmgpu_fail(ctx, HMGPU_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    s_encode = (FsEncodeTiled)fn;
  }
  void* raw = malloc(sizeof(FsMaps) + 64);
  if (!raw) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "out of host memory");
  FsMaps* maps = (FsMaps*)(((uintptr_t)raw + 63) & ~(uintptr_t)63);
  memset(maps, 0, sizeof *maps);
  for (int i = 0; i < FS_TMA_MAPS; i++)
  {
    const cuuint64_t dim[4] = { (cuuint64_t)ctx->pitch, (cuuint64_t)ctx->ph, 16, (cuuint64_t)ctx->max_refs };
    const cuuint64_t str[3] = { (cuuint64_t)ctx->pitch, (cuuint64_t)ctx->plane_elems, (cuuint64_t)ctx->slot_bytes };
    const cuuint32_t box[4] = { (cuuint32_t)((i + 2) * 16), FS_TMA_ROWS, 1, 1 };
    const cuuint32_t est[4] = { 1, 1, 1, 1 };
    const CUresult r = s_encode(&maps->m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, ctx->planes_all, dim, str, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { free(raw); return hmgpu_fail(ctx, HMGPU_E_CUDA, "cuTensorMapEncodeTiled(box %d x %d) failed: %d", (i + 2) * 16, FS_TMA_ROWS, (int)r); }
  }
  ctx->h_tmaps = raw;
  return HMGPU_OK;
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
      HMGPU_CUDA(ctx, cudaFuncSetAttribute(full_search_packed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      HMGPU_CUDA(ctx, cudaFuncSetAttribute(full_search_packed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      ctx->attr_done |= HMGPU_ATTR_FULL;
    }
    if (max_win_bytes > 200 * 1024) return hmgpu_fail(ctx, HMGPU_E_INVALID, "full-search window needs %d bytes of shared memory", max_win_bytes);
    if (ctx->tune.fs_tma && ctx->planes_all)
    {
      if (!ctx->h_tmaps) { const int rcm = fs_build_maps(ctx); if (rcm) return rcm; }
      const FsMaps* maps = (const FsMaps*)(((uintptr_t)ctx->h_tmaps + 63) & ~(uintptr_t)63);
      full_search_packed_kernel<true><<<n_jobs, FS_THREADS, max_win_bytes, ctx->stream>>>(d_jobs, n_jobs, rt, ov, d_results, *maps);
    }
    else
    {
      FsMaps none;
      memset(&none, 0, sizeof none);
      full_search_packed_kernel<false><<<n_jobs, FS_THREADS, max_win_bytes, ctx->stream>>>(d_jobs, n_jobs, rt, ov, d_results, none);
    }
  }
  else if (ctx->px_bytes == 1)
    full_search_generic_kernel<uint8_t><<<n_jobs, FS_THREADS, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  else
    full_search_generic_kernel<uint16_t><<<n_jobs, FS_THREADS, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
