// me_frac2.cu -- fractional refinement, fast path for 8-bit pictures with the source picture as
// key pattern (the common case; bi-pred int16 key patterns and >8-bit video use me_frac.cu).
//
// Same contract as me_frac.cu (xPatternSearchFracDIF + xPatternRefinement over the phase planes,
// TEncSearch.cpp:4386-4422, 799-852) with a different work decomposition, chosen after the ncu
// capture of round 1 showed the per-(tile,candidate) kernel to be bound by uncoalesced loads and
// by instruction count:
//   * one thread owns one 8x8 (or 4x4) tile for ALL 9 candidates of a phase; the source tile
//     stays in registers as packed words.
//   * the horizontal Hadamard pass works on PACKED bytes: the transform is linear, so one
//     coefficient of a row of the DIFFERENCE is sum_j s_kj o_j - sum_j s_kj r_j = four chained
//     IDP.4A.U8.S8 (dot products of 4 unsigned pixels with 4 signed +-1, negated pattern for the
//     reference) -- no unpacking, no subtraction, no horizontal butterflies.  IDP issues on the
//     FMA pipe while the vertical butterflies keep the ALU pipe busy.
//   * a CTA of 128 threads (128 tiles) stages the reference rows of one candidate in shared
//     memory with loads ordered [row][tile][word], so the 3 words of a row segment and the
//     segments of horizontally adjacent tiles fall into the same 128-byte line (one L1 wavefront
//     per row segment instead of one per word).
#include "me_frac_impl.cuh"
#include <stdlib.h>

#define F2_THREADS 128

// ---- work lists: one entry per tile, 8x8-tiled and 4x4-tiled jobs in separate lists -----------------
__global__ void frac2_expand_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, hmgpu_me_result* __restrict__ results,
                                    uint32_t* __restrict__ work8, uint32_t* __restrict__ work4,
                                    uint32_t* __restrict__ counts /* [0]=n8 [1]=n4 */, uint32_t* __restrict__ acc)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt8 = 0, cnt4 = 0;
  if (j < n_jobs)
  {
    const hmgpu_me_job jb = jobs[j];
    if (!(jb.flags & HMGPU_F_INTEGER))
    {
      hmgpu_me_result r;
      r.int_x = jb.start_x; r.int_y = jb.start_y; r.int_sad = 0;
      r.half_x = r.half_y = r.qter_x = r.qter_y = 0; r.frac_cost = 0; r.n_cand = 0;
      results[j] = r;
    }
    if (jb.flags & HMGPU_F_FRAC)
    {
      if (job_tile_size(jb) == 8) cnt8 = (jb.pu_w >> 3) * (jb.pu_h >> 3);
      else cnt4 = (jb.pu_w >> 2) * (jb.pu_h >> 2);
    }
#pragma unroll
    for (int c = 0; c < 18; c++) acc[(size_t)c * n_jobs + j] = 0;   // both phases
  }
  const int lane = threadIdx.x & 31;
  // warp-aggregated reservations (packed: high half = 4x4 tiles, low half = 8x8 tiles; both < 65536 per warp)
  uint32_t incl = ((uint32_t)cnt4 << 16) | (uint32_t)cnt8;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t base8 = 0, base4 = 0;
  if (lane == 31)
  {
    if (total & 0xffffu) base8 = atomicAdd(&counts[0], total & 0xffffu);
    if (total >> 16) base4 = atomicAdd(&counts[1], total >> 16);
  }
  base8 = __shfl_sync(0xffffffffu, base8, 31);
  base4 = __shfl_sync(0xffffffffu, base4, 31);
  const uint32_t off8 = base8 + (incl & 0xffffu) - (uint32_t)cnt8;
  const uint32_t off4 = base4 + (incl >> 16) - (uint32_t)cnt4;
  for (int t = 0; t < cnt8; t++) work8[off8 + t] = ((uint32_t)j << WORK_TILE_BITS) | (uint32_t)t;
  for (int t = 0; t < cnt4; t++) work4[off4 + t] = ((uint32_t)j << WORK_TILE_BITS) | (uint32_t)t;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TS = 8: rows 8, 3 words per staged row segment; TS = 4: rows 4, 2 words
template <int TS>
__global__ void __launch_bounds__(F2_THREADS, 4)
frac2_dist_kernel(const hmgpu_me_job* __restrict__ jobs, RefTable refs, OrgView org,
                  const hmgpu_me_result* __restrict__ results, const uint32_t* __restrict__ work,
                  const uint32_t* __restrict__ work_count, uint32_t* __restrict__ acc, int phase, int n_jobs)
{
  constexpr int NW = TS / 4 + 1;                         // staged words per row segment
  constexpr int OW = TS / 4;                             // source-picture words per row (aligned)
  // every WARP stages for itself (its 32 tiles, double-buffered) and synchronises with __syncwarp only: the CTA-wide
  // barriers of the first version (two per candidate) were 25-35 % of the stall samples (profiles/r1d_ncu_frac2_dist*)
  __shared__ __align__(16) uint32_t s_ref_all[F2_THREADS / 32][2][TS][32 * NW];   // [warp][buffer][row][tile * NW + word]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t (*s_ref)[TS][32 * NW] = s_ref_all[warp];
  const uint32_t n_work = *work_count;
  const int pitch = refs.pitch;
  const int pitch_w = pitch >> 2, org_pitch_w = org.pitch >> 2;
  // loader role of this lane: elements lane + k*32 of a staged row of the warp (tile = idx / NW, word = idx % NW)
  int ld_tile[NW], ld_word[NW];
#pragma unroll
  for (int k = 0; k < NW; k++) { const int idx = lane + k * 32; ld_tile[k] = idx / NW; ld_word[k] = idx - (idx / NW) * NW; }

  for (uint32_t chunk = (blockIdx.x * (F2_THREADS / 32) + warp) * 32; chunk < n_work; chunk += gridDim.x * F2_THREADS)
  {
    const bool have = chunk + lane < n_work;
    // lanes past the end of the list shadow the last tile (loads stay in bounds, nothing is accumulated)
    const uint32_t wi = work[min(chunk + lane, n_work - 1)];
    const uint32_t j = wi >> WORK_TILE_BITS;
    const int t = (int)(wi & ((1u << WORK_TILE_BITS) - 1));
    const hmgpu_me_job jb = jobs[j];
    const hmgpu_me_result res = results[j];
    const int tw = jb.pu_w / TS;
    const int ty = (t / tw) * TS, tx = (t - (t / tw) * tw) * TS;
    const int pu_x = jb.pu_x + tx, pu_y = jb.pu_y + ty;
    const bool satd = (jb.flags & HMGPU_F_HADME) && !(jb.flags & HMGPU_F_LOSSLESS);
    const int qx0 = 4 * res.int_x + (phase ? 2 * res.half_x : 0);
    const int qy0 = 4 * res.int_y + (phase ? 2 * res.half_y : 0);
    const uint8_t* plane0 = (const uint8_t*)refs.base[jb.ref_slot];

    // source tile: aligned words straight from the source picture (pu_x, tile offsets multiples of 4)
    uint32_t ow[TS][OW];
    {
      const uint32_t* o = (const uint32_t*)((const uint8_t*)org.base + (size_t)pu_y * org.pitch + pu_x);
#pragma unroll
      for (int r = 0; r < TS; r++)
#pragma unroll
        for (int w = 0; w < OW; w++) ow[r][w] = __ldg(o + (size_t)r * org_pitch_w + w);
    }

    // candidate c+1 is fetched (cp.async into the other buffer) while candidate c is being transformed
    auto issue = [&](int cand, int buf) -> int {
      const int qx = qx0 + (phase ? c_refine_q[cand][0] : 2 * c_refine_h[cand][0]);
      const int qy = qy0 + (phase ? c_refine_q[cand][1] : 2 * c_refine_h[cand][1]);
      const uint8_t* p = plane0 + (size_t)((qy & 3) * 4 + (qx & 3)) * refs.plane_elems
                       + (ptrdiff_t)(pu_y + (qy >> 2)) * pitch + (pu_x + (qx >> 2));
      const unsigned long long base = (unsigned long long)((uintptr_t)p & ~(uintptr_t)3);   // aligned address of row 0 of MY tile
      __syncwarp();                                       // every lane is done with s_ref[buf]
#pragma unroll
      for (int k = 0; k < NW; k++)
      {
        // the lane that loads word ld_word[k] of tile ld_tile[k] takes that tile's row address from its owner
        const uint32_t* src = (const uint32_t*)(uintptr_t)__shfl_sync(0xffffffffu, base, ld_tile[k]) + ld_word[k];
#pragma unroll
        for (int r = 0; r < TS; r++) cp_async4(&s_ref[buf][r][lane + k * 32], src + (size_t)r * pitch_w);
      }
      cp_async_commit();
      return (int)((uintptr_t)p & 3) * 8;
    };
    // quarter-pel phase: candidate 0 is the best half-pel position, already costed in the half-pel phase (frac_select_kernel
    // copies that sum)
    const int c0 = phase ? 1 : 0;
    int sh_next = issue(c0, c0 & 1);
    for (int cand = c0; cand < 9; cand++)
    {
      const int sh = sh_next, buf = cand & 1;
      if (cand + 1 < 9) { sh_next = issue(cand + 1, buf ^ 1); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncwarp();
      uint32_t v;
      if (satd)
      {
        int d[TS * TS];
        int zero[TS];
#pragma unroll
        for (int k = 0; k < TS; k++) zero[k] = 0;
#pragma unroll
        for (int r = 0; r < TS; r++)
        {
          uint32_t w0, w1;
          if (TS == 4) { const uint2 ww = *(const uint2*)&s_ref[buf][r][lane * NW]; w0 = ww.x; w1 = ww.y; }   // one LDS.64: the two words of a lane are 2-way bank-conflicted as LDS.32
          else { w0 = s_ref[buf][r][lane * NW + 0]; w1 = s_ref[buf][r][lane * NW + 1]; }
          if (TS == 8)
          {
            const uint32_t w2 = s_ref[buf][r][lane * NW + NW - 1];
            int h[TS];
            had_row8<false>(ow[r][0], ow[r][OW - 1], zero, h);                       // + H(org row)
            had_row8<true>(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), h, d + r * TS);   // - H(ref row)
          }
          else
          {
            int h[TS];
            had_row4<false>(ow[r][0], zero, h);
            had_row4<true>(__funnelshift_r(w0, w1, sh), h, d + r * TS);
          }
        }
        v = TS == 8 ? had_cols8_abs(d) : had_cols4_abs(d);
      }
      else
      {
        v = 0;
#pragma unroll
        for (int r = 0; r < TS; r++)
        {
          uint32_t w0, w1;
          if (TS == 4) { const uint2 ww = *(const uint2*)&s_ref[buf][r][lane * NW]; w0 = ww.x; w1 = ww.y; }
          else { w0 = s_ref[buf][r][lane * NW + 0]; w1 = s_ref[buf][r][lane * NW + 1]; }
          v = vabsdiff4_acc(__funnelshift_r(w0, w1, sh), ow[r][0], v);
          if (TS == 8)
          {
            const uint32_t w2 = s_ref[buf][r][lane * NW + NW - 1];
            v = vabsdiff4_acc(__funnelshift_r(w1, w2, sh), ow[r][OW - 1], v);
          }
        }
      }
      if (have) atomicAdd(&acc[(size_t)(phase * 9 + cand) * n_jobs + j], v);
    }
  }
}

__global__ void frac_select_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs,
                                   hmgpu_me_result* __restrict__ results, uint32_t* __restrict__ acc,
                                   int bit_depth, int phase, int reuse_centre);

int hmgpu_launch_frac_packed(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, hmgpu_me_result* d_results, bool any_frac)
{
  // scratch: acc[2][9][n_jobs] | counts | work8[n_jobs*64] | work4[n_jobs*64 (4x4-tiled PUs have <= 48 tiles)]
  const size_t acc_al = ((size_t)n_jobs * 18 * sizeof(uint32_t) + 255) & ~(size_t)255;
  const size_t work_bytes = (size_t)n_jobs * 64 * sizeof(uint32_t);
  int rc = hmgpu_reserve_work(ctx, acc_al + 256 + 2 * work_bytes);
  if (rc) return rc;
  uint32_t* acc = (uint32_t*)ctx->d_work;
  uint32_t* counts = (uint32_t*)((char*)ctx->d_work + acc_al);
  uint32_t* work8 = (uint32_t*)((char*)ctx->d_work + acc_al + 256);
  uint32_t* work4 = (uint32_t*)((char*)ctx->d_work + acc_al + 256 + work_bytes);
  HMGPU_CUDA(ctx, cudaMemsetAsync(counts, 0, 16, ctx->stream));
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int tb = 128;
  {
    HmgpuStage st(ctx, HMGPU_ST_FRAC_EXPAND, 1);
    frac2_expand_kernel<<<(n_jobs + tb - 1) / tb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, d_results, work8, work4, counts, acc);
  }
  if (any_frac)
  {
    const long long want = ((long long)n_jobs * 16 + F2_THREADS - 1) / F2_THREADS;   // ~ tiles / 128 for typical shapes
    const int cap = HMGPU_NUM_SMS * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    for (int phase = 0; phase < 2; phase++)
    {
      {
        // the two tile sizes side by side: the 8x8 kernel is issue-bound, the 4x4 kernel waits on L1 wavefronts (ncu,
        // profiles/r1j_ncu_frac2_dist.csv), so they fill each other's gaps (HMGPU_FRAC_OVERLAP=0: one after the other)
        HmgpuStage st(ctx, HMGPU_ST_FRAC_DIST, 2);
        const int s_overlap = ctx->tune.frac_overlap;
        cudaStream_t side = ctx->stream;
        if (s_overlap)
        {
          if (!ctx->frac_stream)
          {
            HMGPU_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->frac_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; i++) HMGPU_CUDA(ctx, cudaEventCreateWithFlags(&ctx->frac_ev[i], cudaEventDisableTiming));
          }
          side = ctx->frac_stream;
          HMGPU_CUDA(ctx, cudaEventRecord(ctx->frac_ev[0], ctx->stream));
          HMGPU_CUDA(ctx, cudaStreamWaitEvent(side, ctx->frac_ev[0], 0));
        }
        frac2_dist_kernel<4><<<grid, F2_THREADS, 0, side>>>(d_jobs, rt, ov, d_results, work4, counts + 1, acc, phase, n_jobs);
        frac2_dist_kernel<8><<<grid, F2_THREADS, 0, ctx->stream>>>(d_jobs, rt, ov, d_results, work8, counts + 0, acc, phase, n_jobs);
        if (s_overlap)
        {
          HMGPU_CUDA(ctx, cudaEventRecord(ctx->frac_ev[1], side));
          HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->frac_ev[1], 0));
        }
      }
      HmgpuStage st2(ctx, HMGPU_ST_FRAC_SELECT, 1);
      frac_select_kernel<<<(n_jobs + tb - 1) / tb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, d_results, acc, ctx->bit_depth, phase, 1);
    }
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
