// me_single.cu -- fused low-latency path: ONE kernel launch per small batch, one CTA per job.
//
// HM calls xMotionEstimation synchronously from the sequential xCompressCU recursion
// (TEncCu.cpp:466, TEncSearch.cpp:3219-3227): it needs the MV of this PU before its next line.
// For such calls (a handful of jobs: one PU x its reference pictures) the batch pipeline of
// me_tz/me_full/me_frac (6 launches, 2 copies, several stream synchronisations) costs far more
// than the search.  This kernel does the whole job -- integer search (TZ by warp 0, or the
// block-wide full search) and both fractional phases -- in one launch, reads the job from and
// writes the result to MAPPED pinned host memory, and signals completion with a system-scope
// flag the host spins on, so a call is: write 48 bytes, one launch, poll.
#include "me_tz_impl.cuh"
#include "me_full_impl.cuh"
#include "me_frac_impl.cuh"
#include "predict_impl.cuh"

#define SG_THREADS FS_THREADS

template <typename Px, int TS>
__device__ __forceinline__ void sg_frac_phase(const hmgpu_me_job& jb, const int16_t* org_blocks, const RefTable& refs,
                                              const OrgView& org, const hmgpu_me_result& res, int phase, uint32_t* s_acc)
{
  const int tw = jb.pu_w / TS, nt = tw * (jb.pu_h / TS);
  const bool satd = (jb.flags & HMGPU_F_HADME) && !(jb.flags & HMGPU_F_LOSSLESS);
  const int pitch = refs.pitch;
  for (int i = threadIdx.x; i < nt * 9; i += SG_THREADS)
  {
    const int cand = i / nt, t = i - cand * nt;
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
    const Px* ref = (const Px*)refs.base[jb.ref_slot] + (size_t)ph * refs.plane_elems
                  + (ptrdiff_t)(jb.pu_y + (qy >> 2)) * pitch + (jb.pu_x + (qx >> 2));
    const int ty = t / tw, tx = t - ty * tw;
    const uint32_t v = tile_dist<Px, TS>(jb, org_blocks, org, ref, pitch, tx * TS, ty * TS, satd);
    atomicAdd(&s_acc[cand], v);
  }
}

// Small PUs: the distortion of ALL 49 quarter-pel positions (-3..3)^2 around the integer MV in one pass -- every position the
// half-pel stage or any quarter-pel stage can ask for -- so the two dependent refinement stages become two table look-ups.
template <typename Px, int TS>
__device__ __forceinline__ void sg_frac_all(const hmgpu_me_job& jb, const int16_t* org_blocks, const RefTable& refs,
                                            const OrgView& org, const hmgpu_me_result& res, uint32_t* s_acc49)
{
  const int tw = jb.pu_w / TS, nt = tw * (jb.pu_h / TS);
  const bool satd = (jb.flags & HMGPU_F_HADME) && !(jb.flags & HMGPU_F_LOSSLESS);
  const int pitch = refs.pitch;
  for (int i = threadIdx.x; i < nt * 49; i += SG_THREADS)
  {
    const int pos = i / nt, t = i - pos * nt;
    const int qx = 4 * res.int_x + (pos % 7) - 3, qy = 4 * res.int_y + (pos / 7) - 3;
    const int ph = (qy & 3) * 4 + (qx & 3);
    const Px* ref = (const Px*)refs.base[jb.ref_slot] + (size_t)ph * refs.plane_elems
                  + (ptrdiff_t)(jb.pu_y + (qy >> 2)) * pitch + (jb.pu_x + (qx >> 2));
    const int ty = t / tw, tx = t - ty * tw;
    const uint32_t v = tile_dist<Px, TS>(jb, org_blocks, org, ref, pitch, tx * TS, ty * TS, satd);
    atomicAdd(&s_acc49[pos], v);
  }
}

#define SG_WIN_BYTES (10 * 1024)   // staged TZ window: <= (64 + 12) rows x <= 112 bytes

// shared-memory scratch of one job (static part; s_dyn is the CTA's dynamic shared memory)
struct SgShared
{
  unsigned long long red[SG_THREADS / 32];
  hmgpu_me_result res;
  uint32_t acc[9];
  uint32_t acc49[49];                   // small PUs: distortion of all 7x7 quarter-pel positions around the integer MV
  TzSpec spec;
};

// optional phase trace (HMGPU_TRACE=1): globaltimer stamps of block 0, read by the host after the call
__device__ __forceinline__ void sg_stamp(unsigned long long* trace, int k)
{
  if (trace && blockIdx.x == 0 && threadIdx.x == 0)
  {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[k] = t;
  }
}

// One xMotionEstimation body by the whole CTA: integer search (TZ by warp 0 out of a staged window, or the block-wide full
// search), both fractional phases.  The result is left in sh.res (valid after the final barrier).
template <typename Px, bool PACKED>
__device__ __forceinline__ void sg_job(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks, const RefTable& refs, const OrgView& org,
                                       unsigned char* s_dyn, unsigned char* s_org, SgShared& sh, unsigned long long* trace)
{
  const int tid = threadIdx.x;
  // ---- integer search --------------------------------------------------------------------
  if (jb.flags & HMGPU_F_INTEGER)
  {
    if (jb.flags & HMGPU_F_FULL)
    {
      if (PACKED) full_search_block_packed(jb, refs, org, s_dyn, sh.red, &sh.res);
      else full_search_block_generic<Px>(jb, org_blocks, refs, org, (int16_t*)s_org, sh.red, &sh.res);
    }
    else if (jb.kind == HMGPU_KIND_SELECTIVE)
    {
      // xTZSearchSelective (FastSearch = 2): warp 0, compute-then-replay (me_tz_impl.cuh)
      if (tid < 32)
      {
        hmgpu_me_result r;
        tz_selective_warp<Px, PACKED>(jb, org_blocks, refs, org, s_org, r);
        if (tid == 0) sh.res = r;
      }
    }
    else
    {
      // TZ is a chain of dependent rounds executed by warp 0.  Every round that reads the reference from global
      // memory pays a full memory round trip, so the whole CTA first copies the neighbourhood of the start point
      // into shared memory -- one round trip -- and the rounds within TZ_WIN_RADIUS of it (the common case:
      // distances 1, 2, 4 and the two-point fill) run out of it.
      TzWindow win;
      bool have_win = false;
      if (PACKED)
      {
        have_win = tz_window_geometry(jb, refs, win) && win.pitch * win.rows <= SG_WIN_BYTES;
        if (have_win)
        {
          const uint8_t* src = (const uint8_t*)refs.base[jb.ref_slot] + (ptrdiff_t)(jb.pu_y + win.oy) * refs.pitch + (jb.pu_x + win.ox);
          const int c16 = win.pitch >> 4;
          for (int i = tid; i < win.rows * c16; i += SG_THREADS)
          {
            const int r = i / c16, c = i - r * c16;
            *(uint4*)(s_dyn + r * win.pitch + c * 16) = __ldg((const uint4*)(src + (size_t)r * refs.pitch) + c);
          }
        }
      }
      // the key pattern with all threads, then the search by warps 0..2 (speculative first rounds, see TzSpec)
      tz_stage_org<Px, PACKED>(jb, org_blocks, org, s_org, tid, SG_THREADS);
      __syncthreads();
      if (tid < 96)
      {
        hmgpu_me_result r;
        tz_search_warp<Px, PACKED>(jb, org_blocks, refs, org, s_org, r, have_win ? s_dyn : NULL, &win, &sh.spec, tid >> 5);
        if (tid == 0) sh.res = r;
      }
    }
  }
  else if (tid == 0)
  {
    hmgpu_me_result r;
    r.int_x = jb.start_x; r.int_y = jb.start_y; r.int_sad = 0;
    r.half_x = r.half_y = r.qter_x = r.qter_y = 0; r.frac_cost = 0; r.n_cand = 0;
    sh.res = r;
  }
  __syncthreads();
  sg_stamp(trace, 2);

  // ---- fractional refinement -----------------------------------------------------------------
  if (jb.flags & HMGPU_F_FRAC)
  {
    const int16_t* key = org_blocks;
    if (jb.flags & HMGPU_F_ORG_BLOCK)
    {
      // stage the int16 key pattern once (it lives in mapped host memory): tile_dist then reads
      // shared memory through a rebased pointer
      int16_t* so = (int16_t*)s_org;
      for (int i = tid; i < jb.pu_w * jb.pu_h; i += SG_THREADS) so[i] = __ldcv(org_blocks + jb.org_offset + i);
      key = so - jb.org_offset;
    }
    const int bit_depth = refs.bit_depth;
    const int ts = job_tile_size(jb);
    const int n_tiles = (jb.pu_w / ts) * (jb.pu_h / ts);
    if (n_tiles * 49 <= SG_THREADS)
    {
      // ---- all 49 positions at once, then both stages by warp 0 from the table ----
      if (tid < 49) sh.acc49[tid] = 0;
      __syncthreads();
      const hmgpu_me_result res = sh.res;
      if (ts == 8) sg_frac_all<Px, 8>(jb, key, refs, org, res, sh.acc49);
      else sg_frac_all<Px, 4>(jb, key, refs, org, res, sh.acc49);
      __syncthreads();
      if (tid < 32)
      {
        const int c = tid < 9 ? tid : 0;
        // half-pel stage (s_acMvRefineH order, TEncSearch.cpp:51-63), cost scale 1
        const int hx = c_refine_h[c][0], hy = c_refine_h[c][1];
        uint32_t cost = (sh.acc49[(2 * hy + 3) * 7 + (2 * hx + 3)] >> (bit_depth - 8))
                      + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 1, 2 * res.int_x + hx, 2 * res.int_y + hy);
        if (tid >= 9) cost = 0xffffffffu;
        uint32_t best = __reduce_min_sync(0xffffffffu, cost);
        int bi = __ffs(__ballot_sync(0xffffffffu, cost == best && tid < 9)) - 1;
        const int half_x = c_refine_h[bi][0], half_y = c_refine_h[bi][1];
        // quarter-pel stage around the best half position (s_acMvRefineQ order, :65-75), cost scale 0
        const int qx = c_refine_q[c][0], qy = c_refine_q[c][1];
        cost = (sh.acc49[(2 * half_y + qy + 3) * 7 + (2 * half_x + qx + 3)] >> (bit_depth - 8))
             + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 0, 4 * res.int_x + 2 * half_x + qx, 4 * res.int_y + 2 * half_y + qy);
        if (tid >= 9) cost = 0xffffffffu;
        best = __reduce_min_sync(0xffffffffu, cost);
        bi = __ffs(__ballot_sync(0xffffffffu, cost == best && tid < 9)) - 1;
        if (tid == 0)
        {
          sh.res.half_x = (int16_t)half_x; sh.res.half_y = (int16_t)half_y;
          sh.res.qter_x = c_refine_q[bi][0]; sh.res.qter_y = c_refine_q[bi][1];
          sh.res.frac_cost = best;
          sh.res.n_cand += 18;
        }
      }
      __syncthreads();
      sg_stamp(trace, 4);
      return;
    }
    for (int phase = 0; phase < 2; phase++)
    {
      if (tid < 9) sh.acc[tid] = 0;
      __syncthreads();
      const hmgpu_me_result res = sh.res;
      if (job_tile_size(jb) == 8) sg_frac_phase<Px, 8>(jb, key, refs, org, res, phase, sh.acc);
      else sg_frac_phase<Px, 4>(jb, key, refs, org, res, phase, sh.acc);
      __syncthreads();
      if (tid < 32)
      {
        // first strict minimum in table order (TEncSearch.cpp:816-847): lane c costs candidate c, (cost, c) minimum
        const int c = tid < 9 ? tid : 0;
        const uint32_t dist = sh.acc[c] >> (bit_depth - 8);
        uint32_t cost;
        if (phase == 0)
          cost = dist + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 1, 2 * res.int_x + c_refine_h[c][0], 2 * res.int_y + c_refine_h[c][1]);
        else
          cost = dist + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 0, 4 * res.int_x + 2 * res.half_x + c_refine_q[c][0],
                                   4 * res.int_y + 2 * res.half_y + c_refine_q[c][1]);
        if (tid >= 9) cost = 0xffffffffu;
        const uint32_t best = __reduce_min_sync(0xffffffffu, cost);
        const int bi = __ffs(__ballot_sync(0xffffffffu, cost == best && tid < 9)) - 1;
        if (tid == 0)
        {
          if (phase == 0) { sh.res.half_x = c_refine_h[bi][0]; sh.res.half_y = c_refine_h[bi][1]; }
          else { sh.res.qter_x = c_refine_q[bi][0]; sh.res.qter_y = c_refine_q[bi][1]; }
          sh.res.frac_cost = best;
          sh.res.n_cand += 9;
        }
      }
      __syncthreads();
      sg_stamp(trace, 3 + phase);
    }
  }
}

// publish: ONE warp-wide 32-byte store (6 result words, ticket, check word); the host validates it
__device__ __forceinline__ void sg_publish(const hmgpu_me_result& res, HmgpuMailSlot* slot, uint32_t ticket)
{
  const int tid = threadIdx.x;
  if (tid < 8)
  {
    const uint32_t* rw = (const uint32_t*)&res;
    const uint32_t v = tid < 6 ? rw[tid] : (tid == 6 ? ticket : hmgpu_mail_check(rw, ticket));
    ((volatile uint32_t*)slot)[tid] = v;
  }
}

template <typename Px, bool PACKED>
__global__ void __launch_bounds__(SG_THREADS)
me_single_kernel(const __grid_constant__ HmgpuJobPack pack, const int16_t* __restrict__ org_blocks, RefTable refs, OrgView org,
                 HmgpuMailSlot* slots, uint32_t ticket, unsigned long long* trace)
{
  sg_stamp(trace, 0);
  extern __shared__ __align__(16) unsigned char s_dyn[];       // packed full-search window, or the staged TZ window
  __shared__ __align__(16) unsigned char s_org[8192];          // PU block: packed bytes or int16
  __shared__ SgShared sh;
  const hmgpu_me_job jb = pack.jobs[blockIdx.x];               // kernel parameter: no PCIe read
  sg_stamp(trace, 1);
  sg_job<Px, PACKED>(jb, org_blocks, refs, org, s_dyn, s_org, sh, trace);
  sg_publish(sh.res, &slots[blockIdx.x], ticket);
  sg_stamp(trace, 5);
}

// fetch line k of the current call (ticket tk, generation gen) into s_line; false when it does not show up within idle_ns
// (the host then restarts the server and re-submits)
__device__ __forceinline__ bool srv_fetch_line(const volatile uint32_t* lines, int k, uint32_t tk, uint32_t gen, unsigned long long idle_ns,
                                               uint32_t* s_line, int* s_flag)
{
  const int tid = threadIdx.x;
  if (tid < 32)
  {
    const volatile uint32_t* line = lines + k * 16;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int ok = 0;
    for (;;)
    {
      uint32_t w = tid < 16 ? line[tid] : 0u;
      uint32_t h = 0x7F4A7C15u;
#pragma unroll
      for (int i = 0; i < 15; i++) h = (h ^ __shfl_sync(0xffffffffu, w, i)) * 0x85EBCA6Bu + (h >> 15);
      const uint32_t chk = __shfl_sync(0xffffffffu, w, 15), t = __shfl_sync(0xffffffffu, w, 12), g = __shfl_sync(0xffffffffu, w, 13);
      if (h == chk && t == tk && g == gen) { if (tid < 16) s_line[tid] = w; ok = 1; break; }
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (__shfl_sync(0xffffffffu, (int)(now - t0 > idle_ns), 0)) break;
    }
    if (tid == 0) *s_flag = ok;
  }
  __syncthreads();
  return *s_flag != 0;
}

// ---- the mailbox server: a kernel that stays resident between calls --------------------------------------------------
// CTA b polls line b of the mailbox (64 bytes of mapped pinned host memory -- or of a broker shared-memory segment -- holding
// the job, the ticket, the server generation and a check word, fetched with ONE 64-byte read).  A new valid ticket = a new
// call: the CTA runs the job of its line and then the jobs of lines b + gridDim.x, b + 2 gridDim.x, ... of the same call (a
// call may carry more lines than the server has CTAs: several encoders share the SMs of one GPU through the broker, so a
// client's server is kept small), publishing slot k for line k.  A line carries either one xMotionEstimation body or one
// prediction-error job (merge-candidate SATD / AMVP template SAD, predict_impl.cuh).  The host writes every polled line on
// every call (CTAs without a job get a no-op line), so all CTAs see all tickets and their idle clocks run together.  A CTA
// leaves when its line names another generation (the host stopped the server: before every picture upload, so the planes are
// read-only for the life of a server and __ldg stays valid) or when it has been idle for idle_ns -- it then reports
// exited[b] = generation AFTER its last poll, so the host can tell "will never answer" from "still working" and start a new
// generation for the pending call.
template <typename Px>
__global__ void __launch_bounds__(SG_THREADS)
me_server_kernel(const volatile uint32_t* lines, const int16_t* __restrict__ org_blocks, RefTable refs, OrgView org, PredPlanes pl,
                 HmgpuMailSlot* slots, volatile uint32_t* exited, uint32_t gen, uint32_t last_ticket, unsigned long long idle_ns,
                 int packed_ok)
{
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ __align__(16) unsigned char s_org[8192];
  __shared__ SgShared sh;
  __shared__ uint32_t s_line[16];
  __shared__ int s_cmd;                                        // 0 keep polling, 1 run the job in s_line, 2 leave
  __shared__ uint32_t s_sum;
  const int tid = threadIdx.x;
  const volatile uint32_t* line = lines + blockIdx.x * 16;
  unsigned long long t_last = 0;
  if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));
  // A CTA whose line has carried no job for a while polls less often: with several encoders on one GPU the 64-byte reads per
  // poll period of every server CTA add up on PCIe.  CTA 0 (every call has a job 0) always polls flat out.
  int cold = 0;
  for (;;)
  {
    if (blockIdx.x > 0 && cold > 16) __nanosleep(2000);
    if (tid < 32)
    {
      // one 64-byte read of the line by lanes 0..15, validated by its check word
      uint32_t w = tid < 16 ? line[tid] : 0u;
      uint32_t h = 0x7F4A7C15u;
#pragma unroll
      for (int i = 0; i < 15; i++) h = (h ^ __shfl_sync(0xffffffffu, w, i)) * 0x85EBCA6Bu + (h >> 15);
      const uint32_t chk = __shfl_sync(0xffffffffu, w, 15), tk = __shfl_sync(0xffffffffu, w, 12), gn = __shfl_sync(0xffffffffu, w, 13);
      int cmd = 0;
      if (h == chk)
      {
        if (gn != gen) cmd = 2;
        else if (tk != last_ticket) cmd = 1;
      }
      if (tid == 0 && cmd == 0)
      {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t_last > idle_ns) cmd = 2;
      }
      cmd = __shfl_sync(0xffffffffu, cmd, 0) == 2 ? 2 : cmd;    // lane 0 alone watches the clock
      if (cmd == 1 && tid < 16) s_line[tid] = w;
      if (tid == 0) s_cmd = cmd;
    }
    __syncthreads();
    int cmd = s_cmd;
    if (cmd == 2) break;
    if (cmd == 1)
    {
      last_ticket = s_line[12];
      const int n_lines = (int)((s_line[14] >> 8) & 0xffu);
      cold = (s_line[14] & 1u) ? 0 : cold + 1;
      for (int k = blockIdx.x; ; )
      {
        const uint32_t fl = s_line[14];
        if (fl & 2u)                                             // a prediction-error job
        {
          hmgpu_pred_job pj;
#pragma unroll
          for (int i = 0; i < 5; i++) ((uint32_t*)&pj)[i] = s_line[i];
          // scratch of the two separable passes: (h + 7) x w intermediates, the list-0 block, the prediction
          int16_t* s_tmp = (int16_t*)s_dyn;
          int16_t* s_l0 = s_tmp + (64 + 7) * 64;
          int16_t* s_out = (int16_t*)s_org;
          const uint32_t v = pred_error_block<Px, SG_THREADS>(pj, pl, org, ((fl >> 16) & 0xffu) == HMGPU_DF_HADS ? 1 : 0, s_tmp, s_l0, s_out, &s_sum);
          if (tid == 0) { sh.res.int_x = (int16_t)(v & 0xffffu); sh.res.int_y = (int16_t)(v >> 16); }   // word 0 of the slot
          __syncthreads();
          sg_publish(sh.res, &slots[k], last_ticket);
        }
        else if (fl & 1u)                                        // an xMotionEstimation body
        {
          hmgpu_me_job jb;
#pragma unroll
          for (int i = 0; i < 12; i++) ((uint32_t*)&jb)[i] = s_line[i];
          if (packed_ok && !(jb.flags & HMGPU_F_ORG_BLOCK)) sg_job<uint8_t, true>(jb, org_blocks, refs, org, s_dyn, s_org, sh, NULL);
          else sg_job<Px, false>(jb, org_blocks, refs, org, s_dyn, s_org, sh, NULL);
          sg_publish(sh.res, &slots[k], last_ticket);
        }
        k += gridDim.x;
        if (k >= n_lines) break;
        __syncthreads();                                         // s_line is rewritten
        if (!srv_fetch_line(lines, k, last_ticket, gen, idle_ns, s_line, &s_cmd)) { cmd = 2; break; }
      }
      if (cmd == 2) break;
      if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));
    }
    __syncthreads();                                           // s_cmd / s_line are rewritten by the next poll
  }
  // only after the last poll: the host may now assume this CTA will not answer
  if (tid == 0) exited[blockIdx.x] = gen;
}

int hmgpu_launch_single(hmgpu_ctx* ctx, const HmgpuJobPack& pack, int n_jobs, const int16_t* d_org_blocks,
                        HmgpuMailSlot* d_slots, uint32_t ticket, bool any_org_block, int max_win_bytes, unsigned long long* trace)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  HmgpuStage st(ctx, HMGPU_ST_SINGLE, 1);
  if (ctx->px_bytes == 1 && !any_org_block)
  {
    if (!(ctx->attr_done & HMGPU_ATTR_SINGLE))
    {
      HMGPU_CUDA(ctx, cudaFuncSetAttribute(me_single_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
      ctx->attr_done |= HMGPU_ATTR_SINGLE;
    }
    if (max_win_bytes > 180 * 1024) return hmgpu_fail(ctx, HMGPU_E_INVALID, "full-search window needs %d bytes of shared memory", max_win_bytes);
    const int dyn = max_win_bytes > SG_WIN_BYTES ? max_win_bytes : SG_WIN_BYTES;
    me_single_kernel<uint8_t, true><<<n_jobs, SG_THREADS, dyn, ctx->stream>>>(pack, d_org_blocks, rt, ov, d_slots, ticket, trace);
  }
  else if (ctx->px_bytes == 1)
    me_single_kernel<uint8_t, false><<<n_jobs, SG_THREADS, 0, ctx->stream>>>(pack, d_org_blocks, rt, ov, d_slots, ticket, trace);
  else
    me_single_kernel<uint16_t, false><<<n_jobs, SG_THREADS, 0, ctx->stream>>>(pack, d_org_blocks, rt, ov, d_slots, ticket, trace);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

// start a server generation: n_ctas CTAs, dyn_bytes of dynamic shared memory each
int hmgpu_launch_server(hmgpu_ctx* ctx, cudaStream_t stream, const uint32_t* d_lines, const int16_t* d_org_blocks, HmgpuMailSlot* d_slots,
                        uint32_t* d_exited, uint32_t gen, uint32_t last_ticket, unsigned long long idle_ns, int n_ctas, int dyn_bytes)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const PredPlanes pl = hmgpu_pred_planes(ctx);
  if (!(ctx->attr_done & HMGPU_ATTR_SERVER))
  {
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(me_server_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(me_server_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    ctx->attr_done |= HMGPU_ATTR_SERVER;
  }
  ctx->launches += 1; ctx->prof_launches[HMGPU_ST_SINGLE] += 1;
  if (ctx->px_bytes == 1)
    me_server_kernel<uint8_t><<<n_ctas, SG_THREADS, dyn_bytes, stream>>>(d_lines, d_org_blocks, rt, ov, pl, d_slots, d_exited, gen, last_ticket, idle_ns, 1);
  else
    me_server_kernel<uint16_t><<<n_ctas, SG_THREADS, dyn_bytes, stream>>>(d_lines, d_org_blocks, rt, ov, pl, d_slots, d_exited, gen, last_ticket, idle_ns, 0);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
