// me_tz_thread.cu -- TZ search of the small and medium PUs (up to 32x16 / 16x32, 8-bit pictures), ONE THREAD per job.
// Replaces TEncSearch::xTZSearch (TEncSearch.cpp:4027-4228) with xTZSearchHelp (:333-424),
// xTZ8PointDiamondSearch (:616-791) and xTZ2PointSearch (:429-557) for those jobs.
//
// Why: the warp-per-job kernel (me_tz.cu) spends ~1 500 warp instructions on a job whose useful work is ~60 (ncu,
// profiles/r1k_ncu_tz_search*): the per-round control of the search -- point generation, MV cost, reductions, the update of
// the best -- is executed by a whole warp for a handful of points.  Here a warp instruction serves 32 jobs: every thread runs
// the reference's sequential loop for its own job, exactly as written (strict '<' after every point, in emission order), so
// there is nothing to reduce and nothing to replay (~230 warp instructions per job).  What makes that affordable:
//   * jobs are binned by PU shape (tzt_classify_kernel), one launch per shape with (words per row, visited rows, row step) as
//     template parameters: the threads of a warp run the same trip counts and the SAD of a point is straight-line code;
//   * the neighbourhood of each job's likely start points (+-TZT_R samples around the predictor, or around the point between
//     the predictor and the 2Nx2N integer MV) is copied by the WHOLE warp, coalesced, with cp.async, into a private window in
//     shared memory (row-major, odd word stride between the jobs of a warp); the key pattern lives in registers;
//   * a thread has no parallelism of its own, so the points of the near rings are costed four at a time as independent
//     instruction streams (updates applied afterwards in emission order), and a far ring reads the first row of all its points
//     in one round trip and finishes only the few that have not already lost;
//   * what would make 31 threads wait for one is handed over: star refinement, raster scan, a best start point that is not
//     near the window centre, windows that touch the edge of the padded plane.  Those jobs (about one in ten on the 1080p
//     workload) are appended to the list of the warp-per-job kernel, which searches them from scratch -- exact either way.
//     (A second one-thread-per-job pass for them exists, P2 / HMGPU_TZ_P2=1; measured slower: its kernels are a few very long
//     warps each.)
//   * the kernels of the 14 shapes and the warp-per-job kernel of the larger PUs run on side streams, so that the tail of one
//     overlaps the body of the next.  (Tried and dropped: the shapes of similar window size in ONE launch, their tasks dealt out
//     round-robin -- 2.15 ms instead of 1.40 ms for the stage: five instantiations in one kernel issue 11 % more instructions and
//     run slower than the five launches.)
// Measured (1080p, 4 references, 1.18 M jobs): TZ stage 2.65 ms (one warp per job for everything) -> 1.40 ms; results are
// byte-identical (profiles/tz_ab.py prints an MD5 of the result array for either mapping).
#include "me_tz_impl.cuh"
#include <stdlib.h>
#include <type_traits>

#define TZT_R 6                      // window radius: the rings up to distance TZT_NEAR around a best start <= 1 away from its centre
#define TZT_NEAR 4                   // rings up to this distance are costed from the window, the others from global memory
#define TZT_PITCH(WQ) (((2 * TZT_R + 3) >> 2) + (WQ) + 1)   // words per window row: 3 (alignment) + 2R + 4 WQ + 3 (funnel look-ahead) bytes
#define TZT_MAX_REFINE 3              // star-refinement rounds a thread runs before it hands the job over
#define TZT_CLASSES 14               // PU shapes with their own launch; list TZT_CLASSES = everything else

struct TztShape { int w, h, wq, vr, rm; };
// (width, height) -> words per row, visited rows, row step (2 = FEN sub-sampling, TEncSearch.cpp:347-353)
static const TztShape k_tzt_shapes[TZT_CLASSES] = {
  { 4, 8, 1, 8, 1 }, { 4, 16, 1, 8, 2 }, { 8, 4, 2, 4, 1 }, { 8, 8, 2, 8, 1 }, { 8, 16, 2, 8, 2 },
  { 12, 16, 3, 8, 2 }, { 16, 4, 4, 4, 1 }, { 16, 8, 4, 8, 1 }, { 16, 12, 4, 6, 2 }, { 16, 16, 4, 8, 2 },
  { 32, 8, 8, 8, 1 }, { 8, 32, 2, 16, 2 }, { 16, 32, 4, 16, 2 }, { 32, 16, 8, 8, 2 } };

__device__ __forceinline__ int tzt_class(const hmgpu_me_job& jb)
{
  const int w = jb.pu_w, h = jb.pu_h;
  int c = -1;
  if (w == 4) c = h == 8 ? 0 : (h == 16 ? 1 : -1);
  else if (w == 8) c = h == 4 ? 2 : (h == 8 ? 3 : (h == 16 ? 4 : (h == 32 ? 11 : -1)));
  else if (w == 12) c = h == 16 ? 5 : -1;
  else if (w == 16) c = h == 4 ? 6 : (h == 8 ? 7 : (h == 12 ? 8 : (h == 16 ? 9 : (h == 32 ? 12 : -1))));
  else if (w == 32) c = h == 8 ? 10 : (h == 16 ? 13 : -1);
  if (c >= 0 && h > 8 && !(jb.flags & HMGPU_F_FEN)) c = -1;      // all rows visited: the key pattern would not fit the registers
  return c;
}

// lists[c * cap ..] = indices of the TZ jobs of class c (c = TZT_CLASSES: the warp-per-job kernel), counts[c] their number
__global__ void tzt_classify_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, uint32_t* __restrict__ lists, uint32_t cap,
                                    uint32_t* __restrict__ counts)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int cls = -1;
  if (j < n_jobs)
  {
    const hmgpu_me_job jb = jobs[j];
    if ((jb.flags & HMGPU_F_INTEGER) && !(jb.flags & HMGPU_F_FULL) && jb.kind != HMGPU_KIND_SELECTIVE)
    {
      cls = tzt_class(jb);
      if (cls < 0) cls = TZT_CLASSES;
    }
  }
  const uint32_t active = __ballot_sync(0xffffffffu, cls >= 0);
  if (cls >= 0)
  {
    const uint32_t peers = __match_any_sync(active, cls);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&counts[cls], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    lists[(size_t)cls * cap + base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)j;
  }
}

// two untested neighbours of best-point number 1..8 (xTZ2PointSearch, TEncSearch.cpp:429-557), as in c_two_point of
// me_tz_impl.cuh but packed: byte nr-1 holds (dx0+1) | (dy0+1) << 2 | (dx1+1) << 4 | (dy1+1) << 6
__device__ __forceinline__ void tzt_two_point_offsets(int nr, int i, int& dx, int& dy)
{
  //            nr:      1            2            3            4            5            6            7            8
  // (dx0,dy0,dx1,dy1): -1,0,0,-1   -1,-1,1,-1   0,-1,1,0    -1,1,-1,-1   1,-1,1,1     -1,0,0,1     -1,1,1,1     1,0,0,1
  const unsigned long long T =
      (unsigned long long)(0 | 1 << 2 | 1 << 4 | 0 << 6)        | (unsigned long long)(0 | 0 << 2 | 2 << 4 | 0 << 6) << 8 |
      (unsigned long long)(1 | 0 << 2 | 2 << 4 | 1 << 6) << 16  | (unsigned long long)(0 | 2 << 2 | 0 << 4 | 0 << 6) << 24 |
      (unsigned long long)(2 | 0 << 2 | 2 << 4 | 2 << 6) << 32  | (unsigned long long)(0 | 1 << 2 | 1 << 4 | 2 << 6) << 40 |
      (unsigned long long)(0 | 2 << 2 | 2 << 4 | 2 << 6) << 48  | (unsigned long long)(2 | 1 << 2 | 1 << 4 | 2 << 6) << 56;
  const uint32_t e = (uint32_t)(T >> (8 * (nr - 1))) >> (4 * i);
  dx = (int)(e & 3u) - 1;
  dy = (int)((e >> 2) & 3u) - 1;
}

// P2 = false: first pass over the jobs of one shape.  Jobs that need the star refinement, or whose best start point is too far
//   from the predictor for the window, are appended to the shape's second list (their best start point parked in the result
//   slot); raster scans and windows touching the plane edge go to the warp-per-job kernel (rest list).
// P2 = true: second pass over that second list: every thread of a warp now has such a job, so the refinement loop runs
//   convergently.  The search restarts from scratch with the window centred on the parked start point.
// counts[0..TZT_CLASSES) = first-pass list lengths, counts[TZT_CLASSES] = larger PUs (classifier), counts[TZT_CLASSES + 1] = rest
// list (jobs handed over by these kernels), counts[16 + c] = second-pass list of shape c,
// which lives at lists2 + sum of the first-pass lengths of the shapes before c.
// resident 1-warp CTAs per SM that shared memory allows (capped): the register budget of the kernel follows from it
#define TZT_ITEMS(WQ, VR, RM) (((VR) * (RM) + 2 * TZT_R) * TZT_PITCH(WQ))
#define TZT_PER_SM(WQ, VR, RM) ((227 * 1024) / ((TZT_ITEMS(WQ, VR, RM) | 1) * 128 + 1024) > TZT_MAX_PER_SM ? TZT_MAX_PER_SM : (227 * 1024) / ((TZT_ITEMS(WQ, VR, RM) | 1) * 128 + 1024))
#ifndef TZT_MAX_PER_SM
#define TZT_MAX_PER_SM 16
#endif
template <int WQ, int VR, int RM, bool P2>
__global__ void __launch_bounds__(32, TZT_PER_SM(WQ, VR, RM))
tzt_search_kernel(const hmgpu_me_job* __restrict__ jobs, const uint32_t* __restrict__ idx, uint32_t* __restrict__ counts, int cls,
                  RefTable refs, OrgView org, hmgpu_me_result* __restrict__ results,
                  uint32_t* __restrict__ lists2, uint32_t* __restrict__ rest_idx, int use_p2)
{
  constexpr int H = VR * RM;
  constexpr int R = TZT_R;
  constexpr int PITCH = TZT_PITCH(WQ);
  constexpr int WROWS = H + 2 * R;
  constexpr int ITEMS = WROWS * PITCH;
  constexpr int S = ITEMS | 1;                        // odd word stride between the windows of a warp
  constexpr int SUB = RM == 2 ? 1 : 0;
  extern __shared__ uint32_t s_win[];
  const int lane = threadIdx.x;
  const uint32_t* my_win = s_win + lane * S;
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_win);
  const int gpitch = refs.pitch;                      // bytes (8-bit planes)
  uint32_t off2 = 0;
  for (int c = 0; c < cls; c++) off2 += counts[c];
  uint32_t* const my_list2 = lists2 + off2;
  uint32_t* const count2 = counts + 16 + cls;
  uint32_t* const rest_count = counts + TZT_CLASSES + 1;
  if (P2) idx = my_list2;
  const uint32_t n = P2 ? *count2 : counts[cls];
  // staging items of this lane: word i = it * 32 + lane of a window -> global byte offset / validity
  int soff[(ITEMS + 31) / 32];
#pragma unroll
  for (int it = 0; it < (ITEMS + 31) / 32; it++)
  {
    const int i = it * 32 + lane;
    const int r = i / PITCH, w = i - r * PITCH;
    soff[it] = i < ITEMS ? r * gpitch + w * 4 : -1;
  }

  for (uint32_t t0 = blockIdx.x * 32u; t0 < n; t0 += gridDim.x * 32u)
  {
    const bool have = t0 + lane < n;
    uint32_t job_id = 0;
    hmgpu_me_job jb;
    memset(&jb, 0, sizeof(jb));
    if (have) { job_id = idx[t0 + lane]; jb = jobs[job_id]; }

    // ---- window geometry: +-R around the clipped predictor (the first start point, TEncSearch.cpp:4045-4046) ----
    const int mvp_x = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)jb.start_x)) >> 2;
    const int mvp_y = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)jb.start_y)) >> 2;
    // window centre: the clipped predictor, or -- the two likely best start points being the predictor and the 2Nx2N integer MV --
    // the point between them when they are at most 2 (TZT_R - TZT_NEAR) apart (both are then within TZT_R - TZT_NEAR of it)
    const bool has2n = (jb.flags & HMGPU_F_HAS_2NX2N) != 0;
    const int i2n_x = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(jb.i2n_x << 2))) >> 2;
    const int i2n_y = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(jb.i2n_y << 2))) >> 2;
    int sx = mvp_x, sy = mvp_y;
    if (has2n && abs(i2n_x - mvp_x) <= 2 * (TZT_R - TZT_NEAR) && abs(i2n_y - mvp_y) <= 2 * (TZT_R - TZT_NEAR))
    {
      sx = (mvp_x + i2n_x) >> 1; sy = (mvp_y + i2n_y) >> 1;
    }
    if (P2 && have) { const hmgpu_me_result park = results[job_id]; sx = park.int_x; sy = park.int_y; }
    const uint8_t* plane_pu = (const uint8_t*)refs.base[have ? jb.ref_slot : 0] + (ptrdiff_t)jb.pu_y * gpitch + jb.pu_x;
    const int gx0 = jb.pu_x + sx - R, gy0 = jb.pu_y + sy - R;
    const int mis = (gx0 + HMGPU_MARGIN) & 3;         // rows start 64-byte aligned at x = -80
    const uint8_t* a4 = plane_pu + (ptrdiff_t)(sy - R) * gpitch + (sx - R) - mis;
    // the whole window must lie inside the padded plane (always true for clipMv-conformant jobs of PUs up to 16x16
    // unless the CU is larger than the PU grid assumed here; checked, not assumed)
    const bool ok = have && gx0 >= -HMGPU_MARGIN && gy0 >= -HMGPU_MARGIN && gy0 + WROWS <= refs.pic_h + HMGPU_MARGIN &&
                    gx0 - mis + 4 * PITCH <= gpitch - HMGPU_MARGIN;

    // ---- the warp copies the 32 windows, 32 consecutive words per instruction, with cp.async (no register staging, one wait) ----
    const uint32_t okm = __ballot_sync(0xffffffffu, ok);
    for (int j = 0; j < 32; j++)
    {
      if (!((okm >> j) & 1u)) continue;
      const unsigned long long aj = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)a4, j);
      const uint32_t dst = s_base + (uint32_t)(j * S + lane) * 4u;
#pragma unroll
      for (int it = 0; it < (ITEMS + 31) / 32; it++)
        if (soff[it] >= 0)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + it * 128u), "l"(aj + (unsigned long long)soff[it]) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- key pattern in registers: the visited rows only ----
    uint32_t o[VR * WQ];
    {
      const uint8_t* op = (const uint8_t*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;    // pu_x % 4 == 0, pitch % 4 == 0
#pragma unroll
      for (int r = 0; r < VR; r++)
#pragma unroll
        for (int k = 0; k < WQ; k++) o[r * WQ + k] = ok ? __ldg((const uint32_t*)(op + (size_t)(r * RM) * org.pitch) + k) : 0u;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    uint32_t best_cost = 0xffffffffu, n_cand = 0;
    int bx = 0, by = 0, bdist = 0, bround = 0, bpnr = 0;

    // xTZSearchHelp (TEncSearch.cpp:333-424, the non-selective branch).  A thread has no other parallelism than what sits between
    // two updates of its best, so the SADs of the points of a batch are computed as independent instruction streams and the
    // strict-'<' updates are then applied in emission order.
    auto update = [&](uint32_t cost, int x, int y, int pnr, int dist)
    {
      if (cost < best_cost)                                      // strict '<' (TEncSearch.cpp:414)
      {
        best_cost = cost; bx = x; by = y; bdist = dist; bround = 0; bpnr = pnr;
      }
    };
    // one point read from global memory, all rows from r0 on in flight together (kept as ONE copy of the code: a rolled loop
    // over the few points that need it -- the unrolled version made the kernel three times larger and starved the
    // instruction cache, 22 % of the warp cycles in ncu)
    auto sad_global = [&](int x, int y, uint32_t a, int r0) -> uint32_t
    {
      const uintptr_t ga = (uintptr_t)(plane_pu + (ptrdiff_t)y * gpitch + x);
      const int sh = (int)(ga & 3) * 8;
      const uint32_t* q0 = (const uint32_t*)(ga & ~(uintptr_t)3);
      uint32_t w[VR][WQ + 1];
#pragma unroll
      for (int r = 0; r < VR; r++)
        if (r >= r0)
#pragma unroll
          for (int k = 0; k <= WQ; k++) w[r][k] = __ldg(q0 + (size_t)(r * RM) * (gpitch >> 2) + k);
#pragma unroll
      for (int r = 0; r < VR; r++)
        if (r >= r0)
#pragma unroll
          for (int k = 0; k < WQ; k++) a = vabsdiff4_acc(__funnelshift_r(w[r][k], w[r][k + 1], sh), o[r * WQ + k], a);
      return a;
    };
    // up to NP points near the window: window points unconditionally (at clamped coordinates when the point is not valid or
    // outside), the others afterwards, one by one, from global memory -- first row, then the rest unless it has already lost
    auto evaln = [&](auto n_tag, const int* px, const int* py, const bool* pv, const int* ppnr, const int* pdist)
    {
      constexpr int NP = decltype(n_tag)::value;
      uint32_t cost[NP];
      uint32_t gmask = 0;
#pragma unroll
      for (int p = 0; p < NP; p++)
      {
        const uint32_t mvc = hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, px[p], py[p]);
        const int dx = px[p] - sx + R, dy = py[p] - sy + R;      // window coordinates of the candidate
        const bool inw = (unsigned)dx <= 2u * R && (unsigned)dy <= 2u * R;
        const int bo = min(max(dx, 0), 2 * R) + mis;
        const uint32_t* q = my_win + min(max(dy, 0), 2 * R) * PITCH + (bo >> 2);
        const int sh = (bo & 3) * 8;
        uint32_t a0 = 0, a1 = 0;
#pragma unroll
        for (int r = 0; r < VR; r++)
        {
          uint32_t w[WQ + 1];
#pragma unroll
          for (int k = 0; k <= WQ; k++) w[k] = q[r * RM * PITCH + k];
#pragma unroll
          for (int k = 0; k < WQ; k++)
            if ((r * WQ + k) & 1) a1 = vabsdiff4_acc(__funnelshift_r(w[k], w[k + 1], sh), o[r * WQ + k], a1);
            else a0 = vabsdiff4_acc(__funnelshift_r(w[k], w[k + 1], sh), o[r * WQ + k], a0);
        }
        cost[p] = (pv[p] && inw) ? ((a0 + a1) << SUB) + mvc : 0xffffffffu;    // 8-bit pictures: no distortion shift
        if (pv[p] && !inw) gmask |= 1u << p;
      }
      if (gmask == 0)
      {
#pragma unroll
        for (int p = 0; p < NP; p++)
          if (pv[p]) { n_cand++; update(cost[p], px[p], py[p], ppnr[p], pdist[p]); }
      }
      else
      {
        // rare: in emission order, the global points costed on the way against the best of the moment
        for (int p = 0; p < NP; p++)
        {
          int x = 0, y = 0, pnr = 0, dist = 0; uint32_t c = 0; bool v = false;
#pragma unroll
          for (int k = 0; k < NP; k++) if (k == p) { x = px[k]; y = py[k]; pnr = ppnr[k]; dist = pdist[k]; c = cost[k]; v = pv[k]; }
          if (!v) continue;
          n_cand++;
          if ((gmask >> p) & 1u)
          {
            const uint32_t mvc = hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, x, y);
            if (mvc >= best_cost) continue;
            // smallest raw sum whose normalised value reaches best_cost - mvc: the point has lost once a partial sum gets there
            const uint32_t need = best_cost - mvc;
            const uint32_t raw_bound = (need >> SUB) + ((need & ((1u << SUB) - 1u)) ? 1u : 0u);
            uint32_t a = sad_global(x, y, 0u, VR - 1);           // row VR-1 first ...
            if (VR > 1 && a < raw_bound) a = sad_global(x, y, 0u, 0);   // ... then, unless it has lost, everything
            c = (a << SUB) + mvc;
          }
          update(c, x, y, pnr, dist);
        }
      }
    };
    auto eval4 = [&](const int (&px)[4], const int (&py)[4], const bool (&pv)[4], const int (&ppnr)[4], const int (&pdist)[4])
    {
      evaln(std::integral_constant<int, 4>(), px, py, pv, ppnr, pdist);
    };

    // The NF points of a ring at distance > TZT_NEAR around (rcx, rcy): (almost) all outside the window and almost all losers --
    // far from the predictor, so expensive in MV bits, and rarely a match.  The first visited row of all of them is read
    // together (one round trip for the ring); the few whose first row has not already lost against the best known when the
    // ring started are then costed completely, in emission order.
    auto eval_far = [&](auto n_tag, int rcx, int rcy, int d, const TzJob& J)
    {
      constexpr int NF = decltype(n_tag)::value;
      uint32_t surv = 0, valid = 0;
#pragma unroll
      for (int p = 0; p < NF; p++)
      {
        int x, y, pnr, dist;
        tz_diamond_point(rcx, rcy, d, p, x, y, pnr, dist);
        const bool v = tz_in_window(J, rcx, rcy, x, y);
        const uint32_t mvc = hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, x, y);
        const bool need = v && mvc < best_cost;
        const uintptr_t ga = (uintptr_t)(plane_pu + (ptrdiff_t)(need ? y : mvp_y) * gpitch + (need ? x : mvp_x));
        const int sh = (int)(ga & 3) * 8;
        const uint32_t* q = (const uint32_t*)(ga & ~(uintptr_t)3);
        uint32_t w[WQ + 1];
#pragma unroll
        for (int k = 0; k <= WQ; k++) w[k] = __ldg(q + k);
        uint32_t a = 0;
#pragma unroll
        for (int k = 0; k < WQ; k++) a = vabsdiff4_acc(__funnelshift_r(w[k], w[k + 1], sh), o[k], a);
        const uint32_t nd = best_cost - mvc;
        const uint32_t raw_bound = (nd >> SUB) + ((nd & ((1u << SUB) - 1u)) ? 1u : 0u);
        if (v) valid |= 1u << p;
        if (need && a < raw_bound) surv |= 1u << p;
      }
      n_cand += __popc(valid);
      while (surv)
      {
        const int p = __ffs(surv) - 1;
        surv &= surv - 1u;
        int x, y, pnr, dist;
        tz_diamond_point(rcx, rcy, d, p, x, y, pnr, dist);
        const uint32_t mvc = hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, x, y);
        if (mvc >= best_cost) continue;
        const uint32_t a = sad_global(x, y, 0u, 0);
        update((a << SUB) + mvc, x, y, pnr, dist);
      }
    };

    // where the job goes if it is not finished here: 1 = second pass of this shape, 2 = warp-per-job kernel
    int ovf = (have && !ok) ? 2 : 0;
    int park_x = 0, park_y = 0;
    bool resumable = false;                                      // handed over with the first search done: state parked for the warp
    bool past_raster = false;                                    // ... in the middle of the star refinement
    if (ok)
    {
      // ---- start points (:4045-4093): clipped predictor, zero, clipped 2Nx2N integer MV ----
      {
        const int x2 = i2n_x, y2 = i2n_y;
        // the predictor first and alone: the other two, usually outside the window, are then read against its cost
        const int zero4[4] = { 0, 0, 0, 0 };
        const int px1[1] = { mvp_x }, py1[1] = { mvp_y };
        const bool pv1[1] = { true };
        evaln(std::integral_constant<int, 1>(), px1, py1, pv1, zero4, zero4);
        const int px[2] = { 0, has2n ? x2 : 0 }, py[2] = { 0, has2n ? y2 : 0 };
        const bool pv[2] = { true, has2n };
        evaln(std::integral_constant<int, 2>(), px, py, pv, zero4, zero4);
      }
      const int cx = bx, cy = by;
      park_x = cx; park_y = cy;
      if (!P2 && (abs(cx - sx) > R - TZT_NEAR || abs(cy - sy) > R - TZT_NEAR)) ovf = use_p2 ? 1 : 2;    // the near rings would leave the window
      else
      {
        TzJob J;                                                 // only the window fields are read by tz_in_window
        J.L = jb.win_l; J.T = jb.win_t; J.R = jb.win_r; J.B = jb.win_b;
        auto ring = [&](int rcx, int rcy, int d)                 // xTZ8PointDiamondSearch (:616-791)
        {
          bround += 1;
          if (d > TZT_NEAR)
          {
            if (d <= 8) eval_far(std::integral_constant<int, 8>(), rcx, rcy, d, J);
            else eval_far(std::integral_constant<int, 16>(), rcx, rcy, d, J);
            return;
          }
          const int np = d == 1 ? 4 : 8;
          for (int i0 = 0; i0 < np; i0 += 4)
          {
            int px[4], py[4], ppnr[4], pdist[4];
            bool pv[4];
#pragma unroll
            for (int p = 0; p < 4; p++)
            {
              tz_diamond_point(rcx, rcy, d, i0 + p, px[p], py[p], ppnr[p], pdist[p]);
              pv[p] = tz_in_window(J, rcx, rcy, px[p], py[p]);
            }
            eval4(px, py, pv, ppnr, pdist);
          }
        };
        auto two_point = [&]()                                   // xTZ2PointSearch (:429-557)
        {
          const int nr = bpnr, tcx = bx, tcy = by;
          if (nr >= 1 && nr <= 8)
          {
            int px[2] = { 0, 0 }, py[2] = { 0, 0 };
            bool pv[2] = { false, false };
            const int zero4[4] = { 0, 0, 0, 0 }, two4[4] = { 2, 2, 2, 2 };
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
              int dx, dy;
              tzt_two_point_offsets(nr, i, dx, dy);
              px[i] = tcx + dx; py[i] = tcy + dy;
              pv[i] = tz_in_window(J, tcx, tcy, px[i], py[i]);
            }
            evaln(std::integral_constant<int, 2>(), px, py, pv, zero4, two4);
          }
        };
        // ---- first search (:4100-4115): diamonds at distance 1, 2, 4, .. around the best start point, stop 3 rounds after the
        // last improvement; two-point fill (:4137-4141); raster scan when the best is far (:4144-4154, handed over); then the
        // star refinement (:4189-4223): all rings around the best, two-point fill, until a round brings nothing.  One loop
        // for both, so that the ring and two-point code exists once.
        int rcx = cx, rcy = cy;
        bool first = true;
        for (int iter = 0; ; iter++)
        {
          for (int d = 1; d <= jb.search_range; d <<= 1)
          {
            ring(rcx, rcy, d);
            if (first && bround >= 3) break;
          }
          if (bdist == 1) { bdist = 0; two_point(); }            // (the refinement skips it for point number 0; so does two_point)
          if (first && bdist > 5) { ovf = 2; resumable = true; break; }   // raster scan: warp-per-job kernel
          first = false;
          if (bdist == 0) break;
          if (!P2) { ovf = use_p2 ? 1 : 2; resumable = !use_p2; break; }   // needs the refinement: second pass (or the warp-per-job kernel)
          // a long walk (the optimum far from every start point) would keep 31 threads waiting for one: after a few
          // rounds the job goes to the warp-per-job kernel, which costs the points of a round in parallel
          if (iter == TZT_MAX_REFINE) { ovf = 2; resumable = true; past_raster = true; break; }
          rcx = bx; rcy = by;
          bdist = 0; bpnr = 0;
        }
      }
    }

    if (ok && !ovf)
    {
      hmgpu_me_result res;
      res.int_x = (int16_t)bx; res.int_y = (int16_t)by;
      res.int_sad = best_cost - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, bx, by);
      res.half_x = res.half_y = res.qter_x = res.qter_y = 0;
      res.frac_cost = 0;
      res.n_cand = n_cand;
      results[job_id] = res;
    }
    // jobs not finished here: appended to the second list of the shape (start point parked in the result slot) or to the rest list
    if (!P2)
    {
      const uint32_t m1 = __ballot_sync(0xffffffffu, ovf == 1);
      if (m1)
      {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(count2, (uint32_t)__popc(m1));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ovf == 1)
        {
          my_list2[base + __popc(m1 & ((1u << lane) - 1u))] = job_id;
          hmgpu_me_result park;
          memset(&park, 0, sizeof(park));
          park.int_x = (int16_t)park_x; park.int_y = (int16_t)park_y;
          results[job_id] = park;
        }
      }
    }
    const uint32_t m2 = __ballot_sync(0xffffffffu, ovf == 2);
    if (m2)
    {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(rest_count, (uint32_t)__popc(m2));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (ovf == 2)
      {
        rest_idx[base + __popc(m2 & ((1u << lane) - 1u))] = job_id | (resumable ? 0x80000000u : 0u);
        if (resumable)
        {
          // what tz_search_group needs to go on from here (see its `park` argument)
          hmgpu_me_result park;
          park.int_x = (int16_t)bx; park.int_y = (int16_t)by; park.int_sad = best_cost;
          park.half_x = (int16_t)bdist; park.half_y = (int16_t)bpnr;
          park.qter_x = (int16_t)park_x; park.qter_y = (int16_t)park_y;
          park.frac_cost = past_raster ? 1u : 0u; park.n_cand = n_cand;
          results[job_id] = park;
        }
      }
    }
    __syncwarp();                                                // the windows are rewritten by the next task
  }
}

template <int WQ, int VR, int RM, bool P2>
static int tzt_launch_one(hmgpu_ctx* ctx, cudaStream_t stream, const hmgpu_me_job* d_jobs, const uint32_t* idx, uint32_t* counts, int cls,
                          int n_jobs, const RefTable& rt, const OrgView& ov, hmgpu_me_result* d_results, uint32_t* lists2, uint32_t* rest_idx,
                          int use_p2)
{
  constexpr int ITEMS = (VR * RM + 2 * TZT_R) * TZT_PITCH(WQ);
  constexpr int smem = (ITEMS | 1) * 32 * 4;
  const uint32_t attr_bit = 1u << (2 * cls + (P2 ? 1 : 0));
  if (!(ctx->attr_tzt & attr_bit))
  {
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(tzt_search_kernel<WQ, VR, RM, P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->attr_tzt |= attr_bit;
  }
  const int per_sm = TZT_PER_SM(WQ, VR, RM);
  const int grid = max(1, min(HMGPU_NUM_SMS * per_sm, (n_jobs + 31) / 32));
  tzt_search_kernel<WQ, VR, RM, P2><<<grid, 32, smem, stream>>>(d_jobs, idx, counts, cls, rt, ov, d_results, lists2, rest_idx, use_p2);
  return HMGPU_OK;
}

int hmgpu_launch_tz_list(hmgpu_ctx* ctx, cudaStream_t stream, const hmgpu_me_job* d_jobs, const uint32_t* idx, const uint32_t* count,
                         int n_jobs_max, const int16_t* d_org_blocks, hmgpu_me_result* d_results);

// The TZ stage of a batch of 8-bit jobs without explicit key blocks: classify, then -- on side streams, so that the kernels of
// different shapes and the warp-per-job kernel of the larger PUs overlap (each of them alone leaves most of the GPU idle at its
// tail) -- both passes of every small shape and the larger PUs, and finally what those kernels handed over.
int hmgpu_launch_tz_thread(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks, hmgpu_me_result* d_results)
{
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const size_t cap = ((size_t)n_jobs + 63) & ~(size_t)63;
  // first-pass lists of the shapes, the larger PUs, the rest list, one region for all second-pass lists
  int rc = hmgpu_reserve_tzlist(ctx, 256 + (TZT_CLASSES + 3) * cap * sizeof(uint32_t));
  if (rc) return rc;
  if (!ctx->tz_ev[0])
  {
    for (int i = 0; i < HMGPU_TZ_STREAMS; i++) HMGPU_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->tz_streams[i], cudaStreamNonBlocking));
    for (int i = 0; i <= HMGPU_TZ_STREAMS; i++) HMGPU_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tz_ev[i], cudaEventDisableTiming));
  }
  uint32_t* counts = (uint32_t*)ctx->d_tzlist;
  uint32_t* lists = (uint32_t*)((char*)ctx->d_tzlist + 256);
  HMGPU_CUDA(ctx, cudaMemsetAsync(counts, 0, 128, ctx->stream));
  tzt_classify_kernel<<<(n_jobs + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, n_jobs, lists, (uint32_t)cap, counts);
  uint32_t* big_idx = lists + (size_t)TZT_CLASSES * cap;
  uint32_t* rest_idx = lists + (size_t)(TZT_CLASSES + 1) * cap;
  uint32_t* lists2 = lists + (size_t)(TZT_CLASSES + 2) * cap;
  HMGPU_CUDA(ctx, cudaEventRecord(ctx->tz_ev[HMGPU_TZ_STREAMS], ctx->stream));
  for (int i = 0; i < HMGPU_TZ_STREAMS; i++) HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->tz_streams[i], ctx->tz_ev[HMGPU_TZ_STREAMS], 0));
  // the larger PUs (one warp per job) on side stream 0, the shapes round-robin on the others, each second pass behind its first
  // HMGPU_TZ_P2=0: no second pass, the jobs that need the refinement go to the warp-per-job kernel with the rest
  const int s_use_p2 = ctx->tune.tz_p2;
  static const int order[TZT_CLASSES] = { 0, 2, 3, 7, 4, 5, 8, 1, 6, 9, 10, 11, 12, 13 };      // most jobs first
  for (int k = 0; k < TZT_CLASSES; k++)
  {
    const int c = order[k];
    cudaStream_t stream = ctx->tz_streams[1 + k % (HMGPU_TZ_STREAMS - 1)];
    const uint32_t* idx = lists + (size_t)c * cap;
    const TztShape& s = k_tzt_shapes[c];
    for (int pass = 0; pass < (s_use_p2 ? 2 : 1); pass++)
    {
#define TZT_CASE(WQ_, VR_, RM_) \
      if (s.wq == WQ_ && s.vr == VR_ && s.rm == RM_) \
        rc = pass == 0 ? tzt_launch_one<WQ_, VR_, RM_, false>(ctx, stream, d_jobs, idx, counts, c, n_jobs, rt, ov, d_results, lists2, rest_idx, s_use_p2) \
                       : tzt_launch_one<WQ_, VR_, RM_, true>(ctx, stream, d_jobs, idx, counts, c, n_jobs, rt, ov, d_results, lists2, rest_idx, s_use_p2)
      TZT_CASE(1, 8, 1); TZT_CASE(1, 8, 2); TZT_CASE(2, 4, 1); TZT_CASE(2, 8, 1); TZT_CASE(2, 8, 2);
      TZT_CASE(3, 8, 2); TZT_CASE(4, 4, 1); TZT_CASE(4, 8, 1); TZT_CASE(4, 6, 2); TZT_CASE(4, 8, 2);
      TZT_CASE(8, 8, 1); TZT_CASE(2, 16, 2); TZT_CASE(4, 16, 2); TZT_CASE(8, 8, 2);
#undef TZT_CASE
      if (rc) return rc;
    }
  }
  // launched last: its persistent CTAs would otherwise take every register of the machine before the first shape kernel starts
  if ((rc = hmgpu_launch_tz_list(ctx, ctx->tz_streams[0], d_jobs, big_idx, counts + TZT_CLASSES, n_jobs, d_org_blocks, d_results))) return rc;
  // the hand-over list is complete when the shape kernels are done; it does not wait for the larger PUs (side stream 0),
  // whose kernel it overlaps
  for (int i = 1; i < HMGPU_TZ_STREAMS; i++)
  {
    HMGPU_CUDA(ctx, cudaEventRecord(ctx->tz_ev[i], ctx->tz_streams[i]));
    HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->tz_ev[i], 0));
  }
  if ((rc = hmgpu_launch_tz_list(ctx, ctx->stream, d_jobs, rest_idx, counts + TZT_CLASSES + 1, n_jobs, d_org_blocks, d_results))) return rc;
  HMGPU_CUDA(ctx, cudaEventRecord(ctx->tz_ev[0], ctx->tz_streams[0]));
  HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->tz_ev[0], 0));
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
