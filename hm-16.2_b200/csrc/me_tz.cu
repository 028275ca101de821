// me_tz.cu -- batch TZ search kernels.  Replaces TEncSearch::xTZSearch (TEncSearch.cpp:4027-4228);
// the search itself is in me_tz_impl.cuh.
//
//   tz_search_kernel        : one warp per job over an index list (persistent grid-stride loop; the list length is only known on
//                             the device).  In the default mapping of large 8-bit batches (HMGPU_TZ_SPLIT=5, hmgpu_launch_tz) it
//                             searches the larger PUs and the jobs handed over by the one-thread-per-job kernels of
//                             me_tz_thread.cu -- those resume from the state parked in their result slot (bit 31 of the list
//                             entry) -- with the first rounds and the refinement rounds merged into passes of up to 32 points
//                             (MERGE).  Small batches, >8-bit pictures and explicit key patterns: every TZ job, unmerged.
//   tz_selective_kernel     : xTZSearchSelective jobs (FastSearch = 2)
//   tz_classify_kernel, tz_search_small_kernel, tz_search_win_kernel, tz_search_near_kernel (and me_tz_lock.cu): the earlier
//                             mappings HMGPU_TZ_SPLIT=1..4, bit-exact, slower, kept for study (numbers in hmgpu_launch_tz).
#include "me_tz_impl.cuh"
#include <stdlib.h>

#ifndef TZ_MIN_CTAS
#define TZ_MIN_CTAS 8
#endif
#define TZ_THREAD_MIN_JOBS 4096
#define TZ_SMALL_GS 8
#define TZ_SMALL_PIXELS 128                 // visited pixels (pu_w * rows) of the small class
#define TZ_SMALL_BYTES 256                  // staged PU bytes per small job (<= 16x16)

int hmgpu_launch_tz_thread(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks, hmgpu_me_result* d_results);
int hmgpu_launch_tz_lockstep(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, const uint32_t* d_idx, const uint32_t* d_count, uint32_t* d_cursor,
                             int n_jobs_max, hmgpu_me_result* d_results);

// class of the lock-step kernel (me_tz_lock.cu) and of tz_search_near_kernel: PUs up to 16x16
__device__ __forceinline__ bool tz_is_lockstep(const hmgpu_me_job& jb) { return jb.pu_w <= 16 && jb.pu_h <= 16; }

__device__ __forceinline__ bool tz_is_small(const hmgpu_me_job& jb)
{
  const int rows = ((jb.flags & HMGPU_F_FEN) && jb.pu_h > 8) ? jb.pu_h >> 1 : jb.pu_h;
  return jb.pu_w * rows <= TZ_SMALL_PIXELS && jb.pu_w * jb.pu_h <= TZ_SMALL_BYTES;
}

// counts[0] = number of small jobs, counts[1] = number of big jobs
__global__ void tz_classify_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, int split,
                                   uint32_t* __restrict__ idx_small, uint32_t* __restrict__ idx_big,
                                   uint32_t* __restrict__ counts)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int cls = -1;
  if (j < n_jobs)
  {
    const hmgpu_me_job jb = jobs[j];
    if ((jb.flags & HMGPU_F_INTEGER) && !(jb.flags & HMGPU_F_FULL) && jb.kind != HMGPU_KIND_SELECTIVE) cls = ((split == 1 && tz_is_small(jb)) || (split >= 2 && tz_is_lockstep(jb))) ? 0 : 1;
  }
  const int lane = threadIdx.x & 31;
  const uint32_t m0 = __ballot_sync(0xffffffffu, cls == 0), m1 = __ballot_sync(0xffffffffu, cls == 1);
  uint32_t b0 = 0, b1 = 0;
  if (lane == 0)
  {
    if (m0) b0 = atomicAdd(&counts[0], (uint32_t)__popc(m0));
    if (m1) b1 = atomicAdd(&counts[1], (uint32_t)__popc(m1));
  }
  b0 = __shfl_sync(0xffffffffu, b0, 0);
  b1 = __shfl_sync(0xffffffffu, b1, 0);
  const uint32_t below = (1u << lane) - 1u;
  if (cls == 0) idx_small[b0 + __popc(m0 & below)] = (uint32_t)j;
  if (cls == 1) idx_big[b1 + __popc(m1 & below)] = (uint32_t)j;
}

// four jobs per warp, 8 lanes each (packed 8-bit path only)
__global__ void __launch_bounds__(TZ_WARPS * 32, 8)
tz_search_small_kernel(const hmgpu_me_job* __restrict__ jobs, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count,
                       RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  constexpr int GPB = TZ_WARPS * 32 / TZ_SMALL_GS;        // job groups per block
  __shared__ __align__(16) unsigned char s_org_all[GPB][TZ_SMALL_BYTES];
  const int grp = threadIdx.x / TZ_SMALL_GS;
  const uint32_t n = *count;
  for (uint32_t base = blockIdx.x * GPB; base < n; base += gridDim.x * GPB)
  {
    const uint32_t k = base + grp;
    if (k < n)
    {
      const uint32_t job_id = idx[k];
      const hmgpu_me_job jb = jobs[job_id];
      hmgpu_me_result r;
      tz_search_group<uint8_t, true, TZ_SMALL_GS>(jb, NULL, refs, org, s_org_all[grp], r);
      if (TzGroup<TZ_SMALL_GS>::lane() == 0) results[job_id] = r;
      __syncwarp(TzGroup<TZ_SMALL_GS>::mask());
    }
  }
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

// 8-bit pictures, one warp per job, WITH the neighbourhood of the start point staged in shared memory: the ncu capture of
// the kernel above (profiles/r1d_ncu_tz_search*) shows 8.4 L1 sectors per load request -- every (point, row) pair is a
// separate cache line -- so the rounds at distance 1, 2, 4 and the two-point fill (the common case) read a window that the
// warp copied once with coalesced 16-byte loads.  ORG_BYTES / WIN_BYTES size the per-warp buffers of a PU size class.
template <int ORG_BYTES, int WIN_BYTES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 8)
tz_search_win_kernel(const hmgpu_me_job* __restrict__ jobs, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count,
                     RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  __shared__ __align__(16) unsigned char s_org_all[WARPS][ORG_BYTES];
  __shared__ __align__(16) unsigned char s_win_all[WARPS][WIN_BYTES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = *count;
  for (uint32_t k = blockIdx.x * WARPS + warp; k < n; k += gridDim.x * WARPS)
  {
    const uint32_t job_id = idx[k];
    const hmgpu_me_job jb = jobs[job_id];
    TzWindow win;
    const bool have_win = tz_window_geometry(jb, refs, win) && win.pitch * win.rows <= WIN_BYTES;
    if (have_win)
    {
      const uint8_t* src = (const uint8_t*)refs.base[jb.ref_slot] + (ptrdiff_t)(jb.pu_y + win.oy) * refs.pitch + (jb.pu_x + win.ox);
      const int c16 = win.pitch >> 4;
      for (int i = lane; i < win.rows * c16; i += 32)
      {
        const int r = i / c16, c = i - r * c16;
        *(uint4*)(s_win_all[warp] + r * win.pitch + c * 16) = __ldg((const uint4*)(src + (size_t)r * refs.pitch) + c);
      }
    }
    hmgpu_me_result r;
    tz_search_group<uint8_t, true, 32>(jb, NULL, refs, org, s_org_all[warp], r, have_win ? s_win_all[warp] : NULL, &win);   // syncs the warp after staging
    if (lane == 0) results[job_id] = r;
    __syncwarp();
  }
}

// 8-bit pictures, PUs up to 16x16 (91 % of the jobs of a picture), one warp per job.  Two changes against tz_search_kernel:
//  * the neighbourhood of the start point (+-TZN_RADIUS) is copied once into shared memory with coalesced 16-byte loads -- the
//    generic kernel spends 8.4 L1 sectors per load request because every (point, row) pair of a round is a separate line;
//  * the rounds at distance 1, 2, 4, 8 of the first search are ONE pass over that window (MERGE in tz_search_group), which
//    divides the number of dependent passes per job by about three -- the per-pass control (point generation, MV cost,
//    reductions) is what the generic kernel spends its instructions on, not the SADs.
// Points outside the window (far rings, star refinement, a zero vector far from the predictor) are read from global memory.
#ifndef TZN_MIN_CTAS
#define TZN_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(TZ_WARPS * 32, TZN_MIN_CTAS)
tz_search_near_kernel(const hmgpu_me_job* __restrict__ jobs, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count,
                      RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  __shared__ __align__(16) unsigned char s_org_all[TZ_WARPS][TZN_MAX_PU * TZN_MAX_PU];
  __shared__ __align__(16) unsigned char s_win_all[TZ_WARPS][TZN_ROWS * TZN_PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = *count;
  for (uint32_t k = blockIdx.x * TZ_WARPS + warp; k < n; k += gridDim.x * TZ_WARPS)
  {
    const uint32_t job_id = idx[k];
    const hmgpu_me_job jb = jobs[job_id];
    TzWindow win;
    int n16;
    const bool have_win = tz_near_geometry(jb, refs, win, n16);
    if (have_win)
    {
      // 8 rows per pass: lane = 4 * row + 16-byte column
      const uint8_t* src = (const uint8_t*)refs.base[jb.ref_slot] + (ptrdiff_t)(jb.pu_y + win.oy) * refs.pitch + (jb.pu_x + win.ox);
      const int c = lane & 3;
      if (c < n16)
        for (int r = lane >> 2; r < win.rows; r += 8)
          *(uint4*)(s_win_all[warp] + r * TZN_PITCH + c * 16) = __ldg((const uint4*)(src + (size_t)r * refs.pitch) + c);
    }
    hmgpu_me_result r;
    tz_search_group<uint8_t, true, 32, true>(jb, NULL, refs, org, s_org_all[warp], r, have_win ? s_win_all[warp] : NULL, &win);   // syncs the warp after staging
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
  // HMGPU_TZ_MERGE=1: the rounds at distance 1, 2, 4, 8 of the first search as one pass of 28 lanes (see tz_search_group)
  const int s_merge = getenv("HMGPU_TZ_MERGE") ? atoi(getenv("HMGPU_TZ_MERGE")) : 1;
  // HMGPU_TZ_CARVE=k: shared-memory carve-out (percent) asked for this kernel.  Kernels whose L1 / shared split differs do not share
  // an SM; the one-thread-per-job kernels next to it need (almost) all of it as shared memory.
  static const int s_carve = getenv("HMGPU_TZ_CARVE") ? atoi(getenv("HMGPU_TZ_CARVE")) : 50;
  static bool s_carve_set = false;
  if (s_carve >= 0 && !s_carve_set)
  {
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(tz_search_kernel<uint8_t, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, s_carve));
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(tz_search_kernel<uint8_t, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, s_carve));
    s_carve_set = true;
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
  // read per launch (tests switch the mapping inside one process): HMGPU_TZ_SPLIT selects it, HMGPU_TZ_THREAD_MIN the batch size
  // from which the one-thread-per-job kernels are used
  const int s_mode = getenv("HMGPU_TZ_SPLIT") ? atoi(getenv("HMGPU_TZ_SPLIT")) : 5;
  const int thread_min = getenv("HMGPU_TZ_THREAD_MIN") ? atoi(getenv("HMGPU_TZ_THREAD_MIN")) : TZ_THREAD_MIN_JOBS;
  if (s_mode == 5 && ctx->px_bytes == 1 && !any_org_block && n_jobs >= thread_min)
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
  // index lists live in their own scratch buffer (the fractional stage reuses d_work)
  const size_t list_bytes = ((size_t)n_jobs * sizeof(uint32_t) + 255) & ~(size_t)255;
  int rc = hmgpu_reserve_tzlist(ctx, 256 + 2 * list_bytes);
  if (rc) return rc;
  uint32_t* counts = (uint32_t*)ctx->d_tzlist;
  uint32_t* idx_small = (uint32_t*)((char*)ctx->d_tzlist + 256);
  uint32_t* idx_big = (uint32_t*)((char*)ctx->d_tzlist + 256 + list_bytes);
  HMGPU_CUDA(ctx, cudaMemsetAsync(counts, 0, 16, ctx->stream));
  const bool packed = ctx->px_bytes == 1 && !any_org_block;
  // HMGPU_TZ_SPLIT selects the mapping.  0 (default) = one warp per job for everything: 3.22 ms per 1.18 M jobs.  The
  // alternatives are bit-exact (the parity suite passes with each) but measured SLOWER on the 1080p workload and are kept for
  // study: 3 = one warp per job with the start neighbourhood staged in shared memory (5.46 ms: the copy costs more than the
  // ~23 near points save); 2 = PUs up to 16x16 in the lock-step kernel of me_tz_lock.cu, four jobs per warp (4.72 ms: the
  // phase-specific branches of its state machine serialise and every lane walks a whole SAD); 1 = the first
  // four-jobs-per-warp attempt, tz_search_small_kernel (3.6 ms).
  // 4 (default since round 1n) = PUs up to 16x16 in tz_search_near_kernel (window + merged first rounds), the rest one warp per job.
  const int s_split = s_mode == 5 ? 0 : s_mode;
  HmgpuStage st(ctx, HMGPU_ST_TZ, (packed && s_split) ? 3 : 2);
  tz_classify_kernel<<<(n_jobs + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, n_jobs, packed ? s_split : 0, idx_small, idx_big, counts);
  // persistent grids: enough CTAs to fill the machine, never more than the work could use
  const int cap = HMGPU_NUM_SMS * 16;
  const int grid_small = min(cap, (n_jobs + 15) / 16), grid_big = min(cap, (n_jobs + TZ_WARPS - 1) / TZ_WARPS);
  if (packed)
  {
    if (s_split == 3)
    {
      tz_search_win_kernel<256, 28 * 64, 4><<<grid_big, 4 * 32, 0, ctx->stream>>>(d_jobs, idx_small, counts + 0, rt, ov, d_results);
      tz_search_win_kernel<4096, 76 * 112, 2><<<min(cap, (n_jobs + 1) / 2), 2 * 32, 0, ctx->stream>>>(d_jobs, idx_big, counts + 1, rt, ov, d_results);
    }
    else
    {
      if (s_split == 4) tz_search_near_kernel<<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx_small, counts + 0, rt, ov, d_results);
      else if (s_split == 2) { if ((rc = hmgpu_launch_tz_lockstep(ctx, d_jobs, idx_small, counts + 0, counts + 2, n_jobs, d_results))) return rc; }
      else if (s_split == 1) tz_search_small_kernel<<<grid_small, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx_small, counts + 0, rt, ov, d_results);
      tz_search_kernel<uint8_t, true><<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx_big, counts + 1, d_org_blocks, rt, ov, d_results);
    }
  }
  else if (ctx->px_bytes == 1)
    tz_search_kernel<uint8_t, false><<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx_big, counts + 1, d_org_blocks, rt, ov, d_results);
  else
    tz_search_kernel<uint16_t, false><<<grid_big, TZ_WARPS * 32, 0, ctx->stream>>>(d_jobs, idx_big, counts + 1, d_org_blocks, rt, ov, d_results);
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
