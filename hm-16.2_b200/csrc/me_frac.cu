// me_frac.cu -- fractional (half- then quarter-pel) refinement over the precomputed phase planes.
//
// Replaces TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:4386-4422), xPatternRefinement
// (:799-852) and, through the phase planes, xExtDIFUpSamplingH/Q (:5565-5766).
//
// Batch-synchronous: the distortion of every (job, candidate, tile) triple is one independent
// work item, so lanes stay busy regardless of PU size:
//   frac_expand_kernel : one thread per job -> appends its tiles to a flat work list
//   frac_dist_kernel   : one thread per (tile, candidate): SATD (8x8 or 4x4 tiles, rule of
//                        TComRdCost.cpp:1555-1597) or SAD of the tile, accumulated per
//                        (job, candidate) with a RED.ADD
//   frac_select_kernel : one thread per job: adds the MV cost (TComRdCost::getCost, scale 1
//                        for half-pel, 0 for quarter-pel) and keeps the first strict minimum
//                        in table order (s_acMvRefineH / s_acMvRefineQ, TEncSearch.cpp:51-75)
// The half-pel phase must finish before the quarter-pel phase starts (the quarter candidates
// are centred on the best half), hence two dist/select rounds.
#include "me_frac_impl.cuh"

__global__ void frac_expand_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs,
                                   hmgpu_me_result* __restrict__ results,
                                   uint32_t* __restrict__ work, uint32_t* __restrict__ work_count,
                                   uint32_t* __restrict__ acc)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  hmgpu_me_job jb;
  if (j < n_jobs)
  {
    jb = jobs[j];
    if (!(jb.flags & HMGPU_F_INTEGER))
    {
      // integer MV supplied by the caller (xPatternSearchFracDIF called on its own)
      hmgpu_me_result r;
      r.int_x = jb.start_x; r.int_y = jb.start_y; r.int_sad = 0;
      r.half_x = r.half_y = r.qter_x = r.qter_y = 0; r.frac_cost = 0; r.n_cand = 0;
      results[j] = r;
    }
    if (jb.flags & HMGPU_F_FRAC)
    {
      const int ts = job_tile_size(jb);
      cnt = (jb.pu_w / ts) * (jb.pu_h / ts);
    }
#pragma unroll
    for (int c = 0; c < 18; c++) acc[(size_t)c * n_jobs + j] = 0;   // both phases
  }
  // warp-aggregated reservation in the flat work list
  const int lane = threadIdx.x & 31;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(work_count, (uint32_t)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t off = base + (uint32_t)(incl - cnt);
  for (int t = 0; t < cnt; t++) work[off + t] = ((uint32_t)j << WORK_TILE_BITS) | (uint32_t)t;
}

// phase 0: half-pel candidates around the integer MV; phase 1: quarter-pel around the best half
template <typename Px>
__global__ void __launch_bounds__(128)
frac_dist_kernel(const hmgpu_me_job* __restrict__ jobs, const int16_t* __restrict__ org_blocks,
                 RefTable refs, OrgView org, const hmgpu_me_result* __restrict__ results,
                 const uint32_t* __restrict__ work, const uint32_t* __restrict__ work_count,
                 uint32_t* __restrict__ acc, int phase, int n_jobs)
{
  const uint32_t n_work = *work_count;
  const uint64_t total = (uint64_t)n_work * 9u;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const int cand = (int)(i / n_work);
    const uint32_t wi = work[(uint32_t)(i - (uint64_t)cand * n_work)];
    const uint32_t j = wi >> WORK_TILE_BITS;
    const int t = (int)(wi & ((1u << WORK_TILE_BITS) - 1));
    const hmgpu_me_job jb = jobs[j];
    const hmgpu_me_result res = results[j];
    int qx, qy;
    if (phase == 0)
    {
      qx = 4 * res.int_x + 2 * c_refine_h[cand][0];
      qy = 4 * res.int_y + 2 * c_refine_h[cand][1];
    }
    else
    {
      qx = 4 * res.int_x + 2 * res.half_x + c_refine_q[cand][0];
      qy = 4 * res.int_y + 2 * res.half_y + c_refine_q[cand][1];
    }
    const int ph = (qy & 3) * 4 + (qx & 3);
    const int pitch = refs.pitch;
    const Px* ref = (const Px*)refs.base[jb.ref_slot] + (size_t)ph * refs.plane_elems
                  + (ptrdiff_t)(jb.pu_y + (qy >> 2)) * pitch + (jb.pu_x + (qx >> 2));
    const bool satd = (jb.flags & HMGPU_F_HADME) && !(jb.flags & HMGPU_F_LOSSLESS);
    uint32_t v;
    if (job_tile_size(jb) == 8)
    {
      const int tw = jb.pu_w >> 3;
      const int ty = t / tw, tx = t - ty * tw;
      v = tile_dist<Px, 8>(jb, org_blocks, org, ref, pitch, tx * 8, ty * 8, satd);
    }
    else
    {
      const int tw = jb.pu_w >> 2;
      const int ty = t / tw, tx = t - ty * tw;
      v = tile_dist<Px, 4>(jb, org_blocks, org, ref, pitch, tx * 4, ty * 4, satd);
    }
    atomicAdd(&acc[(size_t)(phase * 9 + cand) * n_jobs + j], v);
  }
}

// reuse_centre (packed path, me_frac2.cu): candidate 0 of the quarter-pel phase is the best half-pel position itself (both tables
// start with (0,0)), i.e. the same block of the same phase plane: its distortion is copied from the half-pel phase instead of being
// computed again (the dist kernel of phase 1 then skips candidate 0)
__global__ void frac_select_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs,
                                   hmgpu_me_result* __restrict__ results, uint32_t* __restrict__ acc,
                                   int bit_depth, int phase, int reuse_centre)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const hmgpu_me_job jb = jobs[j];
  if (!(jb.flags & HMGPU_F_FRAC)) return;
  hmgpu_me_result res = results[j];
  // all nine sums first (independent loads in flight together), then the costs in table order
  uint32_t d9[9];
#pragma unroll
  for (int c = 0; c < 9; c++) d9[c] = __ldcg(&acc[(size_t)(phase * 9 + c) * n_jobs + j]);
  uint32_t best = 0xffffffffu;
  int bi = 0;
#pragma unroll
  for (int c = 0; c < 9; c++)
  {
    const uint32_t dist = d9[c] >> (bit_depth - 8);
    uint32_t cost;
    if (phase == 0)
      cost = dist + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 1, 2 * res.int_x + c_refine_h[c][0], 2 * res.int_y + c_refine_h[c][1]);
    else
      cost = dist + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 0, 4 * res.int_x + 2 * res.half_x + c_refine_q[c][0],
                               4 * res.int_y + 2 * res.half_y + c_refine_q[c][1]);
    if (cost < best) { best = cost; bi = c; }
  }
  if (phase == 0)
  {
    res.half_x = c_refine_h[bi][0]; res.half_y = c_refine_h[bi][1];
    if (reuse_centre)
    {
      uint32_t dsel = d9[0];
#pragma unroll
      for (int c = 1; c < 9; c++) if (c == bi) dsel = d9[c];
      acc[(size_t)9 * n_jobs + j] = dsel;
    }
  }
  else { res.qter_x = c_refine_q[bi][0]; res.qter_y = c_refine_q[bi][1]; }
  res.frac_cost = best;
  res.n_cand += 9;
  results[j] = res;
}

int hmgpu_launch_frac(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                      hmgpu_me_result* d_results, bool any_frac)
{
  // scratch: acc[2 phases][9][n_jobs] (candidate-major: coalesced in the select kernel; cleared once by the expand kernel) | work_count[1 (+3 pad)] | work[n_jobs*64]
  const size_t acc_bytes = (size_t)n_jobs * 18 * sizeof(uint32_t);
  const size_t work_bytes = (size_t)n_jobs * 64 * sizeof(uint32_t);
  const size_t acc_al = (acc_bytes + 255) & ~(size_t)255;
  int rc = hmgpu_reserve_work(ctx, acc_al + 256 + work_bytes);
  if (rc) return rc;
  uint32_t* acc = (uint32_t*)ctx->d_work;
  uint32_t* work_count = (uint32_t*)((char*)ctx->d_work + acc_al);
  uint32_t* work = (uint32_t*)((char*)ctx->d_work + acc_al + 256);
  HMGPU_CUDA(ctx, cudaMemsetAsync(work_count, 0, 16, ctx->stream));
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int tb = 128;
  {
    HmgpuStage st(ctx, HMGPU_ST_FRAC_EXPAND, 1);
    frac_expand_kernel<<<(n_jobs + tb - 1) / tb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, d_results, work, work_count, acc);
  }
  if (any_frac)
  {
    // grid-stride over a device-side item count: enough CTAs to fill the machine
    const long long max_items = (long long)n_jobs * 64 * 9;
    long long want = (max_items + 127) / 128;
    const int grid = (int)(want < (long long)HMGPU_NUM_SMS * 16 ? (want < 1 ? 1 : want) : (long long)HMGPU_NUM_SMS * 16);
    for (int phase = 0; phase < 2; phase++)
    {
      {
      HmgpuStage st(ctx, HMGPU_ST_FRAC_DIST, 1);
      if (ctx->px_bytes == 1)
        frac_dist_kernel<uint8_t><<<grid, 128, 0, ctx->stream>>>(d_jobs, d_org_blocks, rt, ov, d_results, work, work_count, acc, phase, n_jobs);
      else
        frac_dist_kernel<uint16_t><<<grid, 128, 0, ctx->stream>>>(d_jobs, d_org_blocks, rt, ov, d_results, work, work_count, acc, phase, n_jobs);
      }
      HmgpuStage st2(ctx, HMGPU_ST_FRAC_SELECT, 1);
      frac_select_kernel<<<(n_jobs + tb - 1) / tb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, d_results, acc, ctx->bit_depth, phase, 0);
    }
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
