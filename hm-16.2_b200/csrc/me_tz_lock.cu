// me_tz_lock.cu -- batch TZ search for PUs up to 16x16 (91 % of the jobs of a picture): FOUR jobs per warp
// in lock-step, one candidate point per lane.
//
// Same contract as me_tz_impl.cuh (TEncSearch::xTZSearch with TZ_SEARCH_CONFIGURATION, TEncSearch.cpp:298-314,
// 4027-4228; xTZSearchHelp :333-424; xTZ8PointDiamondSearch :616-791; xTZ2PointSearch :429-557), different
// mapping, chosen after the ncu capture of the one-warp-per-job kernel (profiles/r1d_ncu_tz_search*):
//   * 1 930 warp instructions per job, almost all of them control overhead of the state machine that a whole
//     warp executed for ONE job.  Here a warp carries four independent searches (groups of 8 lanes); the
//     search is written as a round-level state machine whose transitions are DATA (phase, distance, centre
//     per group), so the four groups execute the same instruction stream whatever round each of them is in.
//   * 8.4 L1 sectors per load request: every (point, row) pair was a separate cache line.  Here the group
//     first copies the neighbourhood of its start point (+-6 integer positions) into shared memory with
//     coalesced 16-byte loads; the rounds at distance 1, 2, 4 and the two-point fill -- the common case --
//     read shared memory, only the far rings (distance >= 8), the raster scan and displaced start points go
//     to global memory.
//   * a group that finishes fetches its next job from a device-side queue at once (persistent grid), so a
//     warp is not held up by its slowest search.
// The reference updates its best with a strict '<' after every point in emission order; a round's result is
// therefore the minimum (cost, emission index) of the round, taken only if strictly below the running best.
// Rounds of 16 points take two passes of 8, in order.
#include "me_tz_impl.cuh"

#define TL_GS 8                         // lanes per job
#define TL_WARPS 4
#define TL_GROUPS (TL_WARPS * 32 / TL_GS)
#define TL_ORG_BYTES 256                // packed PU block, visited rows only (<= 16 rows x 16 bytes)
#define TL_WIN_ROWS (16 + 2 * TZ_WIN_RADIUS)
#define TL_WIN_PITCH 64                 // >= 16 + 2*6 + 4 + 15 (alignment) rounded to 16
#define TL_WIN_BYTES (TL_WIN_ROWS * TL_WIN_PITCH)

enum { TL_IDLE = 0, TL_START, TL_FIRST, TL_TWO1, TL_RASTER, TL_STAR, TL_TWO2, TL_FINISHED };

struct TlState
{
  int phase, d, pass;                   // current round: phase, diamond distance, pass inside the round
  int cx, cy;                           // centre of the current diamond / two-point round
  int two_nr;                           // point number the two-point round was entered with
  int rL, rT, rR, rB;                   // raster window (re-centred when the 2Nx2N MV was tested, :4083-4092)
  uint32_t best_cost; int bx, by, bdist, bround, bpnr;
  uint32_t n_cand;
};

// SAD of one candidate by ONE lane: rows visited rows of wq packed words; ref from the staged window or from the plane
__device__ __forceinline__ uint32_t tl_sad(const uint8_t* p, int pitch, const uint32_t* org_s, int wq, int rows, int row_mul, bool smem)
{
  uint32_t acc = 0;
  const uintptr_t a0 = (uintptr_t)p;
  const int sh = (int)(a0 & 3) * 8;
  const uint32_t* q0 = (const uint32_t*)(a0 & ~(uintptr_t)3);
  const int pitch_w = (pitch * row_mul) >> 2;
  if (smem)
  {
    for (int r = 0; r < rows; r++)
    {
      const uint32_t* q = q0 + r * pitch_w;
      uint32_t lo = q[0];
      for (int k = 0; k < wq; k++)
      {
        const uint32_t hi = q[k + 1];
        acc = vabsdiff4_acc(__funnelshift_r(lo, hi, sh), org_s[r * wq + k], acc);
        lo = hi;
      }
    }
  }
  else
  {
    for (int r = 0; r < rows; r++)
    {
      const uint32_t* q = q0 + (size_t)r * pitch_w;
      uint32_t lo = __ldg(q);
      for (int k = 0; k < wq; k++)
      {
        const uint32_t hi = __ldg(q + k + 1);
        acc = vabsdiff4_acc(__funnelshift_r(lo, hi, sh), org_s[r * wq + k], acc);
        lo = hi;
      }
    }
  }
  return acc;
}

__global__ void __launch_bounds__(TL_WARPS * 32, 6)
tz_lockstep_kernel(const hmgpu_me_job* __restrict__ jobs, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count,
                   uint32_t* __restrict__ cursor, RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results)
{
  __shared__ __align__(16) unsigned char s_org_all[TL_GROUPS][TL_ORG_BYTES];
  __shared__ __align__(16) unsigned char s_win_all[TL_GROUPS][TL_WIN_BYTES];
  const int lane = threadIdx.x & 31, gl = lane & (TL_GS - 1), gbase = lane & ~(TL_GS - 1);
  const int grp = threadIdx.x / TL_GS;
  const uint32_t gmask = 0xffu << gbase;
  const uint32_t n = *count;
  const uint32_t* org_s = (const uint32_t*)s_org_all[grp];
  const uint8_t* win_s = s_win_all[grp];
  const int pitch = refs.pitch;

  // per-group job data (identical in the 8 lanes of a group)
  uint32_t job_id = 0;
  hmgpu_me_job jb;
  int wq = 1, rows = 4, sub_shift = 0;
  const uint8_t* ref00 = NULL;
  bool has2n = false, have_win = false;
  TzWindow win;
  TlState st;
  st.phase = TL_IDLE; st.d = 0; st.pass = 0; st.cx = st.cy = 0; st.two_nr = 0; st.rL = st.rT = st.rR = st.rB = 0;
  st.best_cost = 0xffffffffu; st.bx = st.by = st.bdist = st.bround = st.bpnr = 0; st.n_cand = 0;

  while (__any_sync(0xffffffffu, st.phase != TL_FINISHED))
  {
    // ---- refill: an idle group takes the next job of the queue and stages its PU block and window ----------
    if (st.phase == TL_IDLE)
    {
      uint32_t k = 0;
      if (gl == 0) k = atomicAdd(cursor, 1u);
      k = __shfl_sync(gmask, k, gbase);
      if (k >= n) st.phase = TL_FINISHED;
      else
      {
        job_id = idx[k];
        jb = jobs[job_id];
        wq = jb.pu_w >> 2;
        sub_shift = ((jb.flags & HMGPU_F_FEN) && jb.pu_h > 8) ? 1 : 0;     // TEncSearch.cpp:347-353
        rows = jb.pu_h >> sub_shift;
        has2n = (jb.flags & HMGPU_F_HAS_2NX2N) != 0;
        ref00 = (const uint8_t*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;
        // PU block, visited rows only
        {
          const uint8_t* o = (const uint8_t*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
          uint32_t* so = (uint32_t*)s_org_all[grp];
          for (int i = gl; i < rows * wq; i += TL_GS)
          {
            const int r = i / wq, kk = i - r * wq;
            so[i] = __ldg((const uint32_t*)(o + (size_t)(r << sub_shift) * org.pitch) + kk);
          }
        }
        // neighbourhood of the start point
        have_win = tz_window_geometry(jb, refs, win) && win.pitch <= TL_WIN_PITCH && win.rows <= TL_WIN_ROWS;
        if (have_win)
        {
          const uint8_t* src = ref00 + (ptrdiff_t)win.oy * pitch + win.ox;
          const int c16 = win.pitch >> 4;
          for (int i = gl; i < win.rows * c16; i += TL_GS)
          {
            const int r = i / c16, c = i - r * c16;
            *(uint4*)(s_win_all[grp] + r * TL_WIN_PITCH + c * 16) = __ldg((const uint4*)(src + (size_t)r * pitch) + c);
          }
        }
        st.phase = TL_START; st.d = 0; st.pass = 0;
        st.best_cost = 0xffffffffu; st.bx = st.by = st.bdist = st.bround = st.bpnr = 0; st.n_cand = 0;
        st.rL = jb.win_l; st.rT = jb.win_t; st.rR = jb.win_r; st.rB = jb.win_b;
      }
      __syncwarp(gmask);
    }

    // ---- this lane's point of the current round ----------------------------------------------------------
    int x = 0, y = 0, pnr = 0, dist = 0;
    bool valid = false;
    if (st.phase == TL_START)
    {
      // start points (TEncSearch.cpp:4045-4093): clipped MVP >> 2, zero, clipped 2Nx2N integer MV; not window-checked
      if (gl == 0)
      {
        x = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)jb.start_x)) >> 2;
        y = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)jb.start_y)) >> 2;
      }
      else if (gl == 2)
      {
        x = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(jb.i2n_x << 2))) >> 2;
        y = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(jb.i2n_y << 2))) >> 2;
      }
      valid = gl < (has2n ? 3 : 2);
    }
    else if (st.phase == TL_FIRST || st.phase == TL_STAR)
    {
      const int npts = st.d == 1 ? 4 : (st.d <= 8 ? 8 : 16);
      const int i = st.pass * TL_GS + gl;
      tz_diamond_point(st.cx, st.cy, st.d, i, x, y, pnr, dist);
      TzJob J; J.L = jb.win_l; J.T = jb.win_t; J.R = jb.win_r; J.B = jb.win_b;
      valid = i < npts && tz_in_window(J, st.cx, st.cy, x, y);
    }
    else if (st.phase == TL_TWO1 || st.phase == TL_TWO2)
    {
      x = st.cx + c_two_point[st.two_nr][gl == 0 ? 0 : 2];
      y = st.cy + c_two_point[st.two_nr][gl == 0 ? 1 : 3];
      dist = 2;
      TzJob J; J.L = jb.win_l; J.T = jb.win_t; J.R = jb.win_r; J.B = jb.win_b;
      valid = gl < 2 && tz_in_window(J, st.cx, st.cy, x, y);
    }
    else if (st.phase == TL_RASTER)
    {
      // raster scan, step 5 (:4144-4154); st.d holds the number of points per row, st.pass the pass
      const int nx = st.d, total = nx * ((st.rB - st.rT) / 5 + 1);
      const int i = st.pass * TL_GS + gl;
      const int gy = i / nx, gx = i - gy * nx;
      x = st.rL + gx * 5; y = st.rT + gy * 5; dist = 5;
      valid = i < total;
    }

    // ---- evaluate -------------------------------------------------------------------------------------------
    uint32_t cost = 0xffffffffu;
    if (valid)
    {
      const bool in_win = have_win && x >= win.x0 && x <= win.x1 && y >= win.y0 && y <= win.y1;
      uint32_t sad;
      if (in_win) sad = tl_sad(win_s + (y - win.oy) * TL_WIN_PITCH + (x - win.ox), TL_WIN_PITCH, org_s, wq, rows, 1 << sub_shift, true);
      else sad = tl_sad(ref00 + (ptrdiff_t)y * pitch + x, pitch, org_s, wq, rows, 1 << sub_shift, false);
      cost = hm_sad_norm(sad, sub_shift, 8) + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, x, y);
    }
    // group minimum of (cost, emission index); lanes are in emission order inside a pass
    uint32_t mc = cost; int ml = gl;
#pragma unroll
    for (int o = TL_GS / 2; o > 0; o >>= 1)
    {
      const uint32_t oc = __shfl_xor_sync(0xffffffffu, mc, o);
      const int ol = __shfl_xor_sync(0xffffffffu, ml, o);
      if (oc < mc || (oc == mc && ol < ml)) { mc = oc; ml = ol; }
    }
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    {
      // every lane takes part in the shuffles (idle groups too); only searching groups use the result
      const int src = gbase + ml;
      const int wx = __shfl_sync(0xffffffffu, x, src), wy = __shfl_sync(0xffffffffu, y, src);
      const int wd = __shfl_sync(0xffffffffu, dist, src), wp = __shfl_sync(0xffffffffu, pnr, src);
      if (st.phase != TL_IDLE && st.phase != TL_FINISHED)
      {
        st.n_cand += __popc(vmask & gmask);
        if (mc < st.best_cost)                               // strict '<' (TEncSearch.cpp:414)
        {
          st.best_cost = mc; st.bx = wx; st.by = wy; st.bdist = wd; st.bpnr = wp; st.bround = 0;
        }
      }
    }

    // ---- advance the state machine (uniform inside a group) ---------------------------------------------------
    // `next` = what follows the round that just ended: 0 nothing yet, 1 after the first search, 2 after the first
    // two-point fill, 3 the star-refinement loop test
    int next = 0;
    if (st.phase == TL_START)
    {
      if (has2n)
      {
        const int px = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(st.bx << 2)));
        const int py = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(st.by << 2)));
        const int sr4 = jb.search_range << 2;
        st.rL = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(px - sr4))) >> 2;
        st.rT = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(py - sr4))) >> 2;
        st.rR = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(px + sr4))) >> 2;
        st.rB = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(py + sr4))) >> 2;
      }
      // first search: diamonds at distance 1, 2, 4, ... around the best start point (:4095-4116)
      st.cx = st.bx; st.cy = st.by; st.phase = TL_FIRST; st.d = 1; st.pass = 0; st.bround += 1;
    }
    else if (st.phase == TL_FIRST || st.phase == TL_STAR)
    {
      const int npts = st.d == 1 ? 4 : (st.d <= 8 ? 8 : 16);
      if ((st.pass + 1) * TL_GS < npts) st.pass += 1;        // second pass of a 16-point round
      else
      {
        st.pass = 0;
        const bool first = st.phase == TL_FIRST;
        const bool stop = first ? (st.bround >= 3 || (st.d << 1) > jb.search_range) : ((st.d << 1) >= jb.search_range + 1);
        if (!stop) { st.d <<= 1; st.bround += 1; }
        else if (first) next = 1;
        else
        {
          // end of a star refinement pass (:4205-4222)
          if (st.bdist == 1)
          {
            st.bdist = 0;
            if (st.bpnr >= 1 && st.bpnr <= 8) { st.phase = TL_TWO2; st.two_nr = st.bpnr; st.cx = st.bx; st.cy = st.by; }
            else next = 3;
          }
          else next = 3;
        }
      }
    }
    else if (st.phase == TL_TWO1) next = 2;
    else if (st.phase == TL_TWO2) next = 3;
    else if (st.phase == TL_RASTER)
    {
      const int total = st.d * ((st.rB - st.rT) / 5 + 1);
      if ((st.pass + 1) * TL_GS < total) st.pass += 1;
      else next = 3;
    }
    if (next == 1)
    {
      // two-point fill when the best is a distance-1 neighbour (:4137-4141)
      if (st.bdist == 1)
      {
        st.bdist = 0;
        if (st.bpnr >= 1 && st.bpnr <= 8) { st.phase = TL_TWO1; st.two_nr = st.bpnr; st.cx = st.bx; st.cy = st.by; st.pass = 0; next = 0; }
        else next = 2;
      }
      else next = 2;
    }
    if (next == 2)
    {
      // raster search when the best is far from the start (:4144-4154)
      next = 3;
      if (st.bdist > 5)
      {
        st.bdist = 5;
        const int nx = (st.rR - st.rL) / 5 + 1, ny = (st.rB - st.rT) / 5 + 1;
        if (st.rR >= st.rL && st.rB >= st.rT && nx * ny > 0) { st.phase = TL_RASTER; st.d = nx; st.pass = 0; next = 0; }
      }
    }
    if (next == 3)
    {
      // star refinement loop (:4189-4223): restart the diamonds around the best point while it keeps moving
      if (st.bdist > 0)
      {
        st.cx = st.bx; st.cy = st.by; st.bdist = 0; st.bpnr = 0;
        st.phase = TL_STAR; st.d = 1; st.pass = 0; st.bround += 1;
      }
      else
      {
        if (gl == 0)
        {
          hmgpu_me_result r;
          r.int_x = (int16_t)st.bx; r.int_y = (int16_t)st.by;
          r.int_sad = st.best_cost - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, st.bx, st.by);
          r.half_x = r.half_y = r.qter_x = r.qter_y = 0; r.frac_cost = 0; r.n_cand = st.n_cand;
          results[job_id] = r;
        }
        st.phase = TL_IDLE;
        __syncwarp(gmask);                                   // every lane is done with the group's shared memory
      }
    }
  }
}

int hmgpu_launch_tz_lockstep(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, const uint32_t* d_idx, const uint32_t* d_count, uint32_t* d_cursor,
                             int n_jobs_max, hmgpu_me_result* d_results)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int want = (n_jobs_max + TL_GROUPS - 1) / TL_GROUPS;
  const int cap = HMGPU_NUM_SMS * 6;
  const int grid = want < 1 ? 1 : (want < cap ? want : cap);
  tz_lockstep_kernel<<<grid, TL_WARPS * 32, 0, ctx->stream>>>(d_jobs, d_idx, d_count, d_cursor, rt, ov, d_results);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
