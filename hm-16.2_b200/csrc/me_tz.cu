// me_tz.cu -- batch TZ search kernels.  Replaces TEncSearch::xTZSearch (TEncSearch.cpp:4027-4228);
// the search itself is in me_tz_impl.cuh.
//
//   tz_search_kernel        : one warp per job over an index list (persistent grid-stride loop; the list length is only known on
//                             the device).  In the default mapping of large 8-bit batches (hmgpu_launch_tz_thread) it searches
//                             the larger PUs and the jobs handed over by the one-thread-per-job kernels of me_tz_thread.cu --
//                             those resume from the state parked in their result slot (bit 31 of the list entry) -- with the
//                             first rounds and the refinement rounds merged into passes of up to 32 points (MERGE).  Small
//                             batches, >8-bit pictures and explicit key patterns: every TZ job, unmerged.
//   tz_selective_kernel     : xTZSearchSelective jobs (FastSearch = 2)
// Earlier mappings (four jobs per warp in lock-step, staged start neighbourhoods with one warp per job) were bit-exact but
// slower on the 1080p workload (3.6 - 5.5 ms per 1.18 M jobs against 1.4 ms) and have been removed; DESIGN.md keeps the numbers.
#include "me_tz_impl.cuh"
#include <stdlib.h>

#ifndef TZ_MIN_CTAS
#define TZ_MIN_CTAS 8
#endif
#define TZ_THREAD_MIN_JOBS 4096

int hmgpu_launch_tz_thread(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks, hmgpu_me_result* d_results);

// index list of the TZ jobs of a batch (counts[0] = its length)
__global__ void tz_classify_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, uint32_t* __restrict__ idx, uint32_t* __restrict__ counts)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool tz = false;
  if (j < n_jobs)
  {
    const hmgpu_me_job jb = jobs[j];
    tz = (jb.flags & HMGPU_F_INTEGER) && !(jb.flags & HMGPU_F_FULL) && jb.kind != HMGPU_KIND_SELECTIVE;
  }
  const int lane = threadIdx.x & 31;
  const uint32_t m = __ballot_sync(0xffffffffu, tz);
  uint32_t b = 0;
  if (lane == 0 && m) b = atomicAdd(&counts[0], (uint32_t)__popc(m));
  b = __shfl_sync(0xffffffffu, b, 0);
  if (tz) idx[b + __popc(m & ((1u << lane) - 1u))] = (uint32_t)j;
}

template <typename Px, bool PACKED, bool MERGE = false>
__global__ void __launch_bounds__(TZ_WARPS * 32, TZ_MIN_CTAS)
tz_search_kernel(const hmgpu_me_job* __restrict__ jobs, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count,
                 const int16_t* __restrict__ org_blocks, RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  // per-warp PU block: packed bytes (4 KB) or int16 (8 KB)
  __shared__ __align__(16) unsigned char s_org_all[TZ_WARPS][PACKED ? 4096 : 8192];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = *count;
  for (uint32_t k = blockIdx.x * TZ_WARPS + warp; k < n; k += gridDim.x * TZ_WARPS)
  {
    // bit 31 of a list entry: the job comes from a one-thread-per-job kernel with its state parked in the result slot
    const uint32_t e = idx[k];
    const uint32_t job_id = e & 0x7fffffffu;
    const hmgpu_me_job jb = jobs[job_id];
    hmgpu_me_result park;
    if (e >> 31) park = results[job_id];
    hmgpu_me_result r;
    tz_search_group<Px, PACKED, 32, MERGE>(jb, org_blocks, refs, org, s_org_all[warp], r, NULL, NULL, NULL, 0, (e >> 31) ? &park : NULL);
    if (lane == 0) results[job_id] = r;
    __syncwarp();
  }
}

// xTZSearchSelective jobs (FastSearch = 2): one warp per job over the whole batch, other jobs skipped
template <typename Px, bool PACKED>
__global__ void __launch_bounds__(TZ_WARPS * 32)
tz_selective_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, const int16_t* __restrict__ org_blocks,
                    RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  __shared__ __align__(16) unsigned char s_org_all[TZ_WARPS][PACKED ? 4096 : 8192];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = blockIdx.x * TZ_WARPS + warp; k < n_jobs; k += gridDim.x * TZ_WARPS)
  {
    const hmgpu_me_job jb = jobs[k];
    if (!(jb.flags & HMGPU_F_INTEGER) || (jb.flags & HMGPU_F_FULL) || jb.kind != HMGPU_KIND_SELECTIVE) continue;
    hmgpu_me_result r;
    tz_selective_warp<Px, PACKED>(jb, org_blocks, refs, org, s_org_all[warp], r);
    if (lane == 0) results[k] = r;
    __syncwarp();
  }
}

// the warp-per-job kernel (8-bit packed path) over an index list whose length is on the device, on a stream of the caller's choice
int hmgpu_launch_tz_list(hmgpu_ctx* ctx, cudaStream_t stream, const hmgpu_me_job* d_jobs, const uint32_t* idx, const uint32_t* count,
                         int n_jobs_max, const int16_t* d_org_blocks, hmgpu_me_result* d_results)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int grid = max(1, min(HMGPU_NUM_SMS * 16, (n_jobs_max + TZ_WARPS - 1) / TZ_WARPS));
  // tune.tz_merge: the rounds at distance 1, 2, 4, 8 of the first search as one pass of 28 lanes (see tz_search_group)
  const int s_merge = ctx->tune.tz_merge;
  // tune.tz_carve: shared-memory carve-out (percent) asked for this kernel.  Kernels whose L1 / shared split differs do not share
  // an SM; the one-thread-per-job kernels next to it need (almost) all of it as shared memory.
  const int s_carve = ctx->tune.tz_carve;
  if (s_carve >= 0 && !(ctx->attr_done & HMGPU_ATTR_TZ_CARVE))
  {
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(tz_search_kernel<uint8_t, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, s_carve));
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(tz_search_kernel<uint8_t, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, s_carve));
    ctx->attr_done |= HMGPU_ATTR_TZ_CARVE;
  }
  if (s_merge) tz_search_kernel<uint8_t, true, true><<<grid, TZ_WARPS * 32, 0, stream>>>(d_jobs, idx, count, d_org_blocks, rt, ov, d_results);
  else tz_search_kernel<uint8_t, true><<<grid, TZ_WARPS * 32, 0, stream>>>(d_jobs, idx, count, d_org_blocks, rt, ov, d_results);
  return HMGPU_OK;
}

int hmgpu_launch_tz(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                    hmgpu_me_result* d_results, bool any_org_block, bool any_sel)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  if (ctx->tune.tz_thread && ctx->px_bytes == 1 && !any_org_block && n_jobs >= ctx->tune.tz_thread_min)
  {
    // default for large batches: PUs up to 32x16 one THREAD per job (me_tz_thread.cu), the rest -- and what those kernels hand
    // over -- one warp per job.  Small batches keep the two-launch path below (16 launches cost more than they save there).
    HmgpuStage st(ctx, HMGPU_ST_TZ, 17);
    int rc5 = hmgpu_launch_tz_thread(ctx, d_jobs, n_jobs, d_org_blocks, d_results);
    if (rc5) return rc5;
    if (any_sel)
    {
      ctx->launches += 1; ctx->prof_launches[HMGPU_ST_TZ] += 1;
      const int grid = min(HMGPU_NUM_SMS * 16, (n_jobs + TZ_WARPS - 1) / TZ_WARPS);
      tz_selective_kernel<uint8_t, true><<<grid, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
    }
    HMGPU_CUDA(ctx, cudaGetLastError());
    return HMGPU_OK;
  }
  // one warp per job for every TZ job of the batch; the index list lives in its own scratch buffer (the fractional stage reuses d_work)
  const size_t list_bytes = ((size_t)n_jobs * sizeof(uint32_t) + 255) & ~(size_t)255;
  int rc = hmgpu_reserve_tzlist(ctx, 256 + list_bytes);
  if (rc) return rc;
  uint32_t* counts = (uint32_t*)ctx->d_tzlist;
  uint32_t* idx = (uint32_t*)((char*)ctx->d_tzlist + 256);
  HMGPU_CUDA(ctx, cudaMemsetAsync(counts, 0, 16, ctx->stream));
  const bool packed = ctx->px_bytes == 1 && !any_org_block;
  HmgpuStage st(ctx, HMGPU_ST_TZ, 2);
  tz_classify_kernel<<<(n_jobs + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, n_jobs, idx, counts);
  // persistent grid: enough CTAs to fill the machine, never more than the work could use
  const int cap = HMGPU_NUM_SMS * 16;
  const int grid_big = min(cap, (n_jobs + TZ_WARPS - 1) / TZ_WARPS);
  if (packed)
    tz_search_kernel<uint8_t, true><<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx, counts, d_org_blocks, rt, ov, d_results);
  else if (ctx->px_bytes == 1)
    tz_search_kernel<uint8_t, false><<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx, counts, d_org_blocks, rt, ov, d_results);
  else
    tz_search_kernel<uint16_t, false><<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx, counts, d_org_blocks, rt, ov, d_results);
  if (any_sel)
  {
    ctx->launches += 1; ctx->prof_launches[HMGPU_ST_TZ] += 1;
    const int grid = min(cap, (n_jobs + TZ_WARPS - 1) / TZ_WARPS);
    if (packed) tz_selective_kernel<uint8_t, true><<<grid, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
    else if (ctx->px_bytes == 1) tz_selective_kernel<uint8_t, false><<<grid, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
    else tz_selective_kernel<uint16_t, false><<<grid, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, n_jobs, d_org_blocks, rt, ov, d_results);
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
