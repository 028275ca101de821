// me_tz_impl.cuh -- integer TZ search executed by a group of GS lanes per job (device code shared
// by the batch kernels in me_tz.cu and the fused low-latency kernel in me_single.cu).
//
// Replaces TEncSearch::xTZSearch with TZ_SEARCH_CONFIGURATION (TEncSearch.cpp:298-314,
// 4027-4228), xTZSearchHelp (:333-424), xTZ8PointDiamondSearch (:616-791), xTZ2PointSearch
// (:429-557).
//
// The state machine is uniform inside a lane group; the points of one diamond round are
// evaluated concurrently by sub-groups of lanes.  Because the reference updates its best with a
// strict '<' after every point in emission order, the result of a round is the minimum cost of
// the round, earliest emission index among ties, taken only if it is strictly below the running
// best -- which is what a (cost, lane) min over lanes ordered by emission index yields.  Rounds
// with more points than the group can hold are evaluated in consecutive passes, in order.
//
// GS = 32: one warp per job (large PUs: the lanes of a point split its rows).  The code is written for any power-of-two group
// size (all warp primitives take the group's member mask); only GS = 32 is instantiated -- small PUs in large batches go to the
// one-thread-per-job kernels of me_tz_thread.cu instead.
//
// SAD arithmetic: 8-bit pictures use packed bytes and VABSDIFF4.U8.ACC against the PU block
// staged in shared memory; >8-bit pictures and explicit int16 key patterns (bi-pred
// 2*org-pred, values outside the pixel range) use 16-bit elements and scalar |a-b|.
#pragma once
#include "hmgpu_internal.cuh"

#define TZ_WARPS 4

struct TzJob
{
  int pu_w, pu_h, sub_shift, rows;     // rows = visited rows (pu_h >> sub_shift)
  int pred_x, pred_y;
  uint32_t ui_cost;
  int L, T, R, B;
  int bit_depth;
  // optional copy of the reference around the start point in shared memory (fused low-latency kernel only):
  // candidates with wx0 <= x <= wx1, wy0 <= y <= wy1 read every sample from it
  const uint8_t* win_s; int win_pitch, wx0, wy0, wx1, wy1, win_ox, win_oy;
};

// geometry of that window, computed the same way by the threads that fill it and by the search
struct TzWindow
{
  int ox, oy;          // picture-relative position (bytes / rows from the PU origin) of window byte (0,0): ox is 16-aligned in memory
  int pitch, rows;     // bytes per staged row (multiple of 16), staged rows
  int x0, y0, x1, y1;  // candidate range fully served by the window
};
#define TZ_WIN_RADIUS 6
__device__ __forceinline__ int tz_clip_q(int v, int lo, int hi) { return min(hi, max(lo, v)); }
// returns false when no useful window exists (start point too close to the padded border)
__device__ __forceinline__ bool tz_window_geometry(const hmgpu_me_job& jb, const RefTable& refs, TzWindow& w)
{
  const int sx = tz_clip_q(jb.start_x, jb.clip_hmin, jb.clip_hmax) >> 2, sy = tz_clip_q(jb.start_y, jb.clip_vmin, jb.clip_vmax) >> 2;
  // absolute picture coordinates of the samples needed by candidates sx-R..sx+R, sy-R..sy+R (+4 bytes funnel look-ahead)
  int ax0 = jb.pu_x + sx - TZ_WIN_RADIUS, ax1 = jb.pu_x + sx + TZ_WIN_RADIUS + jb.pu_w + 4;
  int ay0 = jb.pu_y + sy - TZ_WIN_RADIUS, ay1 = jb.pu_y + sy + TZ_WIN_RADIUS + jb.pu_h;
  ax0 = max(ax0, -HMGPU_MARGIN); ay0 = max(ay0, -HMGPU_MARGIN);
  ax1 = min(ax1, refs.pic_w + HMGPU_MARGIN); ay1 = min(ay1, refs.pic_h + HMGPU_MARGIN);
  const int al0 = (ax0 + HMGPU_MARGIN) & ~15;               // rows start 64-byte aligned at x = -80
  const int al1 = (ax1 + HMGPU_MARGIN + 15) & ~15;          // the row pitch has >= 32 bytes of slack
  w.ox = al0 - HMGPU_MARGIN - jb.pu_x; w.oy = ay0 - jb.pu_y;
  w.pitch = al1 - al0; w.rows = ay1 - ay0;
  // candidate x reads bytes [x & ~3 (in memory), x + pu_w + 3]: inside [al0, al1) when
  w.x0 = w.ox; w.x1 = w.ox + w.pitch - jb.pu_w - 4;
  w.y0 = w.oy; w.y1 = w.oy + w.rows - jb.pu_h;
  return w.x1 >= w.x0 && w.y1 >= w.y0 && w.rows > 0 && w.pitch > 0;
}

struct TzBest
{
  uint32_t cost;
  int x, y;
  int dist, round, pnr;
  uint32_t n_cand;
};

template <int GS> struct TzGroup
{
  // lane index inside the group, member mask of the group, first lane of the group
  __device__ __forceinline__ static int lane() { return threadIdx.x & (GS - 1); }
  __device__ __forceinline__ static int base() { return (threadIdx.x & 31) & ~(GS - 1); }
  __device__ __forceinline__ static uint32_t mask() { return GS == 32 ? 0xffffffffu : (((1u << GS) - 1u) << base()); }
};

// ---- per-lane partial SAD over rows r0, r0+rstep, ... (indices of visited rows) -------------

// packed 8-bit path. org_s: pu_h rows of pu_w bytes (row pitch wq words); ref = plane pointer at
// the PU origin displaced by the candidate.
// raw_bound: the smallest raw sum at which this point can no longer beat the running best (the
// reference's compare is a strict '<', TEncSearch.cpp:414): a lane stops as soon as its own partial
// sum reaches it -- the point then loses whatever the remaining rows add, so the truncated value is
// never selected and the result stays exact.  The far rings of the star refinement end after a
// row or two this way.
// WQ > 0: words per row known at compile time (the row unrolls, every operand is base + immediate); WQ = 0: run-time wq
template <bool SMEM, int WQ>
__device__ __forceinline__ uint32_t sad_rows_packed_w(const uint32_t* q0, int sh, int pitch_w, const uint32_t* org_s,
                                                      int wq_rt, int rows, int row_mul, int r0, int rstep, uint32_t raw_bound)
{
  const int wq = WQ ? WQ : wq_rt;
  uint32_t acc = 0;
  for (int r = r0; r < rows; r += rstep)
  {
    const uint32_t* q = q0 + (r * row_mul) * pitch_w;     // < 2^31 words inside a plane
    const uint32_t* o = org_s + (r * row_mul) * wq;
    if (WQ)
    {
      uint32_t w[WQ + 1];
#pragma unroll
      for (int k = 0; k <= WQ; k++) w[k] = SMEM ? q[k] : __ldg(q + k);
#pragma unroll
      for (int k = 0; k < WQ; k++) acc = vabsdiff4_acc(__funnelshift_r(w[k], w[k + 1], sh), o[k], acc);
    }
    else
    {
      uint32_t lo = SMEM ? q[0] : __ldg(q);
      for (int k = 0; k < wq; k++)
      {
        const uint32_t hi = SMEM ? q[k + 1] : __ldg(q + k + 1);
        acc = vabsdiff4_acc(__funnelshift_r(lo, hi, sh), o[k], acc);
        lo = hi;
      }
    }
    if (acc >= raw_bound) break;
  }
  return acc;
}

template <bool SMEM = false>
__device__ __forceinline__ uint32_t sad_rows_packed(const uint8_t* ref, int pitch, const uint32_t* org_s,
                                                    int wq, int rows, int row_mul, int r0, int rstep, uint32_t raw_bound)
{
  // SMEM: `ref` may have travelled through a struct member that can be NULL, which hides its address space from the compiler
  // (it then emits generic LD instead of LDS): rebuild the pointer from the 32-bit shared-window address
  const uintptr_t a0 = SMEM ? (uintptr_t)__cvta_generic_to_shared(ref) : (uintptr_t)ref;
  const int sh = (int)(a0 & 3) * 8;
  const uint32_t* q0 = SMEM ? (const uint32_t*)__cvta_shared_to_generic((size_t)(a0 & ~(uintptr_t)3)) : (const uint32_t*)(a0 & ~(uintptr_t)3);
  const int pitch_w = pitch >> 2;                      // pitch is a multiple of 4 bytes
  // the common PU widths (8, 16, 4, 32 samples) get unrolled rows; wq is uniform across the lanes of a job
  if (wq == 2) return sad_rows_packed_w<SMEM, 2>(q0, sh, pitch_w, org_s, wq, rows, row_mul, r0, rstep, raw_bound);
  if (wq == 4) return sad_rows_packed_w<SMEM, 4>(q0, sh, pitch_w, org_s, wq, rows, row_mul, r0, rstep, raw_bound);
  if (wq == 1) return sad_rows_packed_w<SMEM, 1>(q0, sh, pitch_w, org_s, wq, rows, row_mul, r0, rstep, raw_bound);
  if (wq == 8) return sad_rows_packed_w<SMEM, 8>(q0, sh, pitch_w, org_s, wq, rows, row_mul, r0, rstep, raw_bound);
  return sad_rows_packed_w<SMEM, 0>(q0, sh, pitch_w, org_s, wq, rows, row_mul, r0, rstep, raw_bound);
}

// generic path: org int16 in shared memory (row pitch pu_w), ref elements of type Px
template <typename Px>
__device__ __forceinline__ uint32_t sad_rows_generic(const Px* ref, int pitch, const int16_t* org_s,
                                                     int w, int rows, int row_mul, int r0, int rstep, uint32_t raw_bound)
{
  uint32_t acc = 0;
  for (int r = r0; r < rows; r += rstep)
  {
    const Px* p = ref + (size_t)(r * row_mul) * pitch;
    const int16_t* o = org_s + (r * row_mul) * w;
    for (int k = 0; k < w; k++) acc += (uint32_t)hm_abs((int)o[k] - (int)__ldg(p + k));
    if (acc >= raw_bound) break;
  }
  return acc;
}

// Cost (SAD + MV cost) of THIS lane's point, summed over the `lanes_per_point` lanes that share it; 0xffffffff when the point is
// not valid or cannot beat `best_cost` (MV cost alone too high, or the row sums reached the bound -- see sad_rows_packed_w).
template <typename Px, bool PACKED, int GS>
__device__ __forceinline__ uint32_t tz_point_cost(const TzJob& J, const Px* ref00, int pitch, const void* org_s,
                                                  int lanes_per_point, int x, int y, bool valid, uint32_t best_cost)
{
  const uint32_t gm = TzGroup<GS>::mask();
  const int sub = TzGroup<GS>::lane() & (lanes_per_point - 1);
  uint32_t part = 0;
  const uint32_t mvc = hm_mv_cost(J.ui_cost, J.pred_x, J.pred_y, 2, x, y);
  if (valid && mvc < best_cost)
  {
    // smallest raw sum whose normalised value (sum << sub_shift) >> (bitDepth-8) reaches best.cost - mvc
    const uint32_t need = best_cost - mvc;
    uint32_t raw_bound;
    if (PACKED) raw_bound = (need >> J.sub_shift) + ((need & ((1u << J.sub_shift) - 1u)) ? 1u : 0u);   // 8-bit pictures: no 64-bit math
    else
    {
      const int shift = J.bit_depth - 8;
      const uint64_t nb = (((uint64_t)need << shift) + ((1u << J.sub_shift) - 1u)) >> J.sub_shift;
      raw_bound = nb > 0xffffffffull ? 0xffffffffu : (uint32_t)nb;
    }
    const Px* ref = ref00 + (ptrdiff_t)y * pitch + x;
    if (PACKED && J.win_s && x >= J.wx0 && x <= J.wx1 && y >= J.wy0 && y <= J.wy1)
      part = sad_rows_packed<true>(J.win_s + (y - J.win_oy) * J.win_pitch + (x - J.win_ox), J.win_pitch, (const uint32_t*)org_s, J.pu_w >> 2, J.rows,
                                   1 << J.sub_shift, sub, lanes_per_point, raw_bound);
    else if (PACKED)
      part = sad_rows_packed((const uint8_t*)ref, pitch, (const uint32_t*)org_s, J.pu_w >> 2, J.rows, 1 << J.sub_shift, sub, lanes_per_point, raw_bound);
    else
      part = sad_rows_generic<Px>(ref, pitch, (const int16_t*)org_s, J.pu_w, J.rows, 1 << J.sub_shift, sub, lanes_per_point, raw_bound);
  }
  for (int o = lanes_per_point >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(gm, part, o);
  uint32_t cost = 0xffffffffu;
  if (valid) cost = mvc < best_cost ? hm_sad_norm(part, J.sub_shift, J.bit_depth) + mvc : 0xffffffffu;
  return cost;
}

// Evaluate up to GS / lanes_per_point points at once.  Sub-group lane / lanes_per_point owns one
// point; (x, y, valid, pnr, dist) are the data of THIS lane's point.  Updates `best` uniformly
// across the group.
template <typename Px, bool PACKED, int GS>
__device__ __forceinline__ void tz_eval(const TzJob& J, const Px* ref00, int pitch, const void* org_s,
                                        int lanes_per_point, int x, int y, bool valid, int pnr, int dist,
                                        TzBest& best)
{
  const uint32_t gm = TzGroup<GS>::mask();
  const int sub = TzGroup<GS>::lane() & (lanes_per_point - 1);
  const uint32_t cost = tz_point_cost<Px, PACKED, GS>(J, ref00, pitch, org_s, lanes_per_point, x, y, valid, best.cost);
  best.n_cand += __popc(__ballot_sync(gm, valid && sub == 0));
  const uint32_t m = __reduce_min_sync(gm, cost);
  if (m < best.cost)                                   // strict '<' (TEncSearch.cpp:414)
  {
    const int src = __ffs(__ballot_sync(gm, valid && cost == m)) - 1;   // absolute lane of the earliest minimum
    best.cost = m;
    best.x = __shfl_sync(gm, x, src);
    best.y = __shfl_sync(gm, y, src);
    best.dist = __shfl_sync(gm, dist, src);
    best.pnr = __shfl_sync(gm, pnr, src);
    best.round = 0;
  }
}

// side-specific window test of the reference: a coordinate is tested only against the bound it
// moved towards (TEncSearch.cpp:634-790, 443-549)
__device__ __forceinline__ bool tz_in_window(const TzJob& J, int cx, int cy, int x, int y)
{
  if (x < cx && x < J.L) return false;
  if (x > cx && x > J.R) return false;
  if (y < cy && y < J.T) return false;
  if (y > cy && y > J.B) return false;
  return true;
}

// Point i (emission order) of the diamond at distance d around (cx, cy): xTZ8PointDiamondSearch
// (TEncSearch.cpp:616-791).  Every offset of a round is a small multiple of one unit -- d for d = 1, d/2 for 2 <= d <= 8
// (axis points at 2 units, diagonal points at 1), d/4 for d > 8 (16 points on the diamond of radius 4 units) -- so the
// points come branch-free out of nibble-packed tables (one table per class: signed 4-bit multipliers, point numbers,
// distance multipliers), indexed by the lane's emission index.
__device__ __forceinline__ int tz_nib(unsigned long long t, int i) { return (int)((t >> (4 * i)) & 15ull); }
__device__ __forceinline__ int tz_snib(unsigned long long t, int i) { return (tz_nib(t, i) ^ 8) - 8; }   // sign-extend 4 bits
__device__ __forceinline__ void tz_diamond_point(int cx, int cy, int d, int i, int& x, int& y, int& pnr, int& dist)
{
  // nibble i of each table = value for emission index i (two's-complement nibbles for the multipliers)
  //   d == 1 : (0,-1) (-1,0) (1,0) (0,1)                                   pnr 2 4 5 7
  //   d <= 8 : (0,-2) (-1,-1) (1,-1) (-2,0) (2,0) (-1,1) (1,1) (0,2)        pnr 2 1 3 4 5 6 8 7, dist 2 1 1 2 2 1 1 2 units
  //   d  > 8 : (0,-4) (-4,0) (4,0) (0,4) then idx = 1..3: (-q,-(4-q)) (q,-(4-q)) (-q,4-q) (q,4-q) with q = idx
  unsigned long long mx, my, pn, dm;
  int unit;
  if (d == 1)      { mx = 0x01F0ull; my = 0x100Full; pn = 0x7542ull; dm = 0x1111ull; unit = 1; }
  else if (d <= 8) { mx = 0x01F2E1F0ull; my = 0x21100FFEull; pn = 0x78654312ull; dm = 0x21122112ull; unit = d >> 1; }
  else             { mx = 0x3D3D2E2E1F1F04C0ull; my = 0x11FF22EE33DD400Cull; pn = 0ull; dm = 0x4444444444444444ull; unit = d >> 2; }
  x = cx + tz_snib(mx, i) * unit;
  y = cy + tz_snib(my, i) * unit;
  pnr = tz_nib(pn, i);
  dist = tz_nib(dm, i) * unit;
}

// largest power of two <= v (v >= 1)
__device__ __forceinline__ int pow2_floor(int v) { return 1 << (31 - __clz(v)); }

// xTZ8PointDiamondSearch (TEncSearch.cpp:616-791): all points of the round, in emission order
template <typename Px, bool PACKED, int GS>
__device__ __forceinline__ void tz_diamond(const TzJob& J, const Px* ref00, int pitch, const void* org_s,
                                           int cx, int cy, int d, TzBest& best)
{
  const int gl = TzGroup<GS>::lane();
  best.round += 1;
  const int n = d == 1 ? 4 : (d <= 8 ? 8 : 16);
  // lanes per point: as many as the group allows for n points, never more than the visited rows
  const int per_pass = min(n, GS);                       // points evaluated concurrently
  const int lpp = pow2_floor(min(GS / per_pass, J.rows));
  const int slot = gl >> (31 - __clz(lpp));              // lpp is a power of two
  for (int i0 = 0; i0 < n; i0 += per_pass)
  {
    const int i = i0 + slot;
    int x, y, pnr, dist;
    tz_diamond_point(cx, cy, d, i, x, y, pnr, dist);
    const bool valid = (slot < per_pass) && (i < n) && tz_in_window(J, cx, cy, x, y);
    tz_eval<Px, PACKED, GS>(J, ref00, pitch, org_s, lpp, x, y, valid, pnr, dist, best);
  }
}

// xTZ2PointSearch (TEncSearch.cpp:429-557): offsets (dx0,dy0,dx1,dy1) of the two untested
// neighbours for best-point numbers 1..8, in the reference's evaluation order
static __constant__ int8_t c_two_point[9][4] = {
  { 0, 0, 0, 0 },
  { -1, 0, 0, -1 }, { -1, -1, 1, -1 }, { 0, -1, 1, 0 },
  { -1, 1, -1, -1 }, { 1, -1, 1, 1 },
  { -1, 0, 0, 1 }, { -1, 1, 1, 1 }, { 1, 0, 0, 1 } };

template <typename Px, bool PACKED, int GS>
__device__ __forceinline__ void tz_two_point(const TzJob& J, const Px* ref00, int pitch, const void* org_s, TzBest& best)
{
  const int nr = best.pnr;
  if (nr < 1 || nr > 8) return;
  const int lpp = pow2_floor(min(GS / 2, J.rows));
  const int i = TzGroup<GS>::lane() >> (31 - __clz(lpp));
  const int cx = best.x, cy = best.y;
  const int x = cx + c_two_point[nr][i == 0 ? 0 : 2];
  const int y = cy + c_two_point[nr][i == 0 ? 1 : 3];
  const bool valid = (i < 2) && tz_in_window(J, cx, cy, x, y);
  tz_eval<Px, PACKED, GS>(J, ref00, pitch, org_s, lpp, x, y, valid, 0, 2, best);
}

// stage the key pattern of a job into shared memory: packed bytes (PACKED) or int16; executed by `nthr` threads, this one is `t`
template <typename Px, bool PACKED>
__device__ __forceinline__ void tz_stage_org(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks, const OrgView& org,
                                             unsigned char* s_org, int t, int nthr)
{
  if (PACKED)
  {
    const int wq = jb.pu_w >> 2;
    const uint8_t* o = (const uint8_t*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
    uint32_t* so = (uint32_t*)s_org;
    for (int i = t; i < jb.pu_h * wq; i += nthr)
    {
      const int r = i / wq, k = i - r * wq;
      so[i] = __ldg((const uint32_t*)(o + (size_t)r * org.pitch) + k);   // pu_x % 4 == 0, pitch % 4 == 0
    }
  }
  else
  {
    int16_t* so = (int16_t*)s_org;
    if (jb.flags & HMGPU_F_ORG_BLOCK)
    {
      // key patterns may live in mapped host memory that a resident server kernel sees rewritten call after call:
      // ld.cv (never served from a stale cache line)
      const int16_t* o = org_blocks + jb.org_offset;
      for (int i = t; i < jb.pu_h * jb.pu_w; i += nthr) so[i] = __ldcv(o + i);
    }
    else
    {
      const Px* o = (const Px*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
      for (int i = t; i < jb.pu_h * jb.pu_w; i += nthr)
      {
        const int r = i / jb.pu_w, k = i - r * jb.pu_w;
        so[i] = (int16_t)o[(size_t)r * org.pitch + k];
      }
    }
  }
}

// Speculative first rounds (fused low-latency kernel only): the diamonds at distance 1, 2, 4 of the first search all sit
// around the SAME centre (the best start point), so three warps evaluate them at once and warp 0 replays the reference's
// sequential bookkeeping on their results.  A round reports its minimum only if it beats the best start point; whether it
// also beats what the earlier rounds found is decided in the replay -- exactly the strict-'<' scan of the reference.
struct TzSpecRound { uint32_t cost; int x, y, dist, pnr; uint32_t n_valid; };
struct TzSpec { TzSpecRound r[3]; };

// The whole TZ search of one job, executed by ONE GROUP of GS lanes (all of them call it together).
// s_org: per-group shared memory for the PU block: pu_w*pu_h bytes (PACKED) or 2*pu_w*pu_h (int16).
// The first lane of the group returns the result in `out` (integer MV, SAD without MV cost,
// candidate count).
template <typename Px, bool PACKED, int GS, bool MERGE = false>
__device__ __forceinline__ void tz_search_group(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks,
                                                const RefTable& refs, const OrgView& org, unsigned char* s_org,
                                                hmgpu_me_result& out, const uint8_t* win_s = NULL, const TzWindow* win = NULL,
                                                TzSpec* spec = NULL, int role = 0, const hmgpu_me_result* park = NULL)
{
  const int gl = TzGroup<GS>::lane();
  const uint32_t gm = TzGroup<GS>::mask();

  TzJob J;
  J.pu_w = jb.pu_w; J.pu_h = jb.pu_h;
  J.sub_shift = ((jb.flags & HMGPU_F_FEN) && jb.pu_h > 8) ? 1 : 0;   // TEncSearch.cpp:347-353
  J.rows = J.pu_h >> J.sub_shift;
  J.pred_x = jb.pred_x; J.pred_y = jb.pred_y; J.ui_cost = jb.ui_cost;
  J.L = jb.win_l; J.T = jb.win_t; J.R = jb.win_r; J.B = jb.win_b;
  J.bit_depth = refs.bit_depth;
  J.win_s = NULL; J.win_pitch = 0; J.wx0 = J.wy0 = 0; J.wx1 = J.wy1 = -1; J.win_ox = J.win_oy = 0;
  if (win_s && win)
  {
    J.win_s = win_s; J.win_pitch = win->pitch; J.wx0 = win->x0; J.wy0 = win->y0; J.wx1 = win->x1; J.wy1 = win->y1;
    J.win_ox = win->ox; J.win_oy = win->oy;
  }

  const int pitch = refs.pitch;
  const Px* ref00 = (const Px*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;

  // ---- stage the key pattern (unless the caller did it with more threads) ---------------------
  if (!spec) tz_stage_org<Px, PACKED>(jb, org_blocks, org, s_org, gl, GS);
  __syncwarp(gm);

  TzBest best; best.cost = 0xffffffffu; best.x = 0; best.y = 0; best.dist = 0; best.round = 0; best.pnr = 0; best.n_cand = 0;

  // ---- start points (TEncSearch.cpp:4045-4093): clipped MVP>>2, zero, clipped 2Nx2N integer MV.
  // Evaluated sequentially in the reference; concurrently here with emission order = sub-group index.
  const bool has2n = (jb.flags & HMGPU_F_HAS_2NX2N) != 0;
  // park != NULL: the one-thread-per-job kernel (me_tz_thread.cu) has run the start points, the first search and its two-point
  // fill of this job and parked the state in the result slot (int_x/int_y/int_sad = best point and its cost, half_x/half_y =
  // its distance and point number, qter_x/qter_y = the best START point, n_cand): only the raster scan and the star refinement
  // remain -- the parts where a warp per job pays.
  const bool resumed = park != NULL;
  int bsx = 0, bsy = 0;                                  // best start point
  if (resumed)
  {
    best.cost = park->int_sad; best.x = park->int_x; best.y = park->int_y;
    best.dist = park->half_x; best.pnr = park->half_y; best.n_cand = park->n_cand;
    bsx = park->qter_x; bsy = park->qter_y;
  }
  else
  {
    const int lpp = pow2_floor(min(GS / 4, J.rows));
    const int i = gl >> (31 - __clz(lpp));
    int x = 0, y = 0;
    if (i == 0)
    {
      x = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)jb.start_x)) >> 2;
      y = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)jb.start_y)) >> 2;
    }
    else if (i == 2)
    {
      x = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(jb.i2n_x << 2))) >> 2;
      y = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(jb.i2n_y << 2))) >> 2;
    }
    const bool valid = i < (has2n ? 3 : 2);
    tz_eval<Px, PACKED, GS>(J, ref00, pitch, s_org, lpp, x, y, valid, 0, 0, best);
    bsx = best.x; bsy = best.y;
  }
  // raster window: re-centred on the best start point when the 2Nx2N MV was tested (:4083-4092)
  int rL = J.L, rT = J.T, rR = J.R, rB = J.B;
  if (has2n)
  {
    const int px = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(bsx << 2)));
    const int py = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(bsy << 2)));
    const int sr4 = jb.search_range << 2;
    rL = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(px - sr4))) >> 2;
    rT = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(py - sr4))) >> 2;
    rR = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(px + sr4))) >> 2;
    rB = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(py + sr4))) >> 2;
  }

  // ---- first search: diamonds at distance 1,2,4,.. ; stop 3 rounds after the last improvement
  int cx = best.x, cy = best.y;
  int d0 = resumed ? jb.search_range + 1 : 1;
  if (spec)
  {
    // warps `role` = 0, 1, 2 of the CTA run this function together (same job, same start evaluation)
    if (jb.search_range < 4) { if (role != 0) return; }
    else
    {
      TzBest b2 = best;
      tz_diamond<Px, PACKED, GS>(J, ref00, pitch, s_org, cx, cy, 1 << role, b2);
      if (gl == 0)
      {
        TzSpecRound rr;
        rr.cost = b2.cost < best.cost ? b2.cost : 0xffffffffu;
        rr.x = b2.x; rr.y = b2.y; rr.dist = b2.dist; rr.pnr = b2.pnr; rr.n_valid = b2.n_cand - best.n_cand;
        spec->r[role] = rr;
      }
      asm volatile("bar.sync 1, 96;" ::: "memory");         // the three role warps
      if (role != 0) return;
      for (int w = 0; w < 3; w++)                           // replay: rounds d = 1, 2, 4 in order
      {
        const TzSpecRound rr = spec->r[w];
        best.round += 1;
        best.n_cand += rr.n_valid;
        if (rr.cost < best.cost) { best.cost = rr.cost; best.x = rr.x; best.y = rr.y; best.dist = rr.dist; best.pnr = rr.pnr; best.round = 0; }
      }
      d0 = best.round >= 3 ? jb.search_range + 1 : 8;       // the reference stops 3 rounds after the last improvement
    }
  }
  if (MERGE && GS == 32 && !resumed)
  {
    // Merged first rounds (batch kernel of the small PUs): the diamonds at distance 1, 2, 4, 8 all sit around the SAME centre, so
    // their 4 + 8 + 8 + 8 = 28 points are costed in ONE pass, one lane per point in emission order, against the best START point
    // (a point that cannot beat it cannot beat anything later either -- the best only decreases), and the reference's sequential
    // bookkeeping -- strict '<' per round, the round counter, the stop three rounds after the last improvement, the candidate
    // count of the rounds it actually runs -- is replayed on the 28 costs.
    const int rr = gl < 4 ? 0 : (gl < 12 ? 1 : (gl < 20 ? 2 : 3));         // lanes 0-3: d = 1, 4-11: d = 2, 12-19: d = 4, 20-27: d = 8
    int x, y, pnr, dist;
    tz_diamond_point(cx, cy, 1 << rr, rr == 0 ? gl : ((gl + 4) & 7), x, y, pnr, dist);
    const bool valid = gl < 28 && (1 << rr) <= jb.search_range && tz_in_window(J, cx, cy, x, y);
    const uint32_t cost = tz_point_cost<Px, PACKED, GS>(J, ref00, pitch, s_org, 1, x, y, valid, best.cost);
    const uint32_t vm = __ballot_sync(gm, valid);
    d0 = 16;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
      if ((1 << r) > jb.search_range) { d0 = jb.search_range + 1; break; }
      const uint32_t rmask = r == 0 ? 0xFu : (0xFFu << (8 * r - 4));
      best.round += 1;
      best.n_cand += __popc(vm & rmask);
      const uint32_t c = ((rmask >> gl) & 1u) ? cost : 0xffffffffu;
      const uint32_t m = __reduce_min_sync(gm, c);
      if (m < best.cost)
      {
        const int src = __ffs(__ballot_sync(gm, c == m)) - 1;
        best.cost = m;
        best.x = __shfl_sync(gm, x, src);
        best.y = __shfl_sync(gm, y, src);
        best.dist = __shfl_sync(gm, dist, src);
        best.pnr = __shfl_sync(gm, pnr, src);
        best.round = 0;
      }
      if (best.round >= 3) { d0 = jb.search_range + 1; break; }
    }
  }
  for (int d = d0; d <= jb.search_range; d <<= 1)
  {
    tz_diamond<Px, PACKED, GS>(J, ref00, pitch, s_org, cx, cy, d, best);
    if (best.round >= 3) break;
  }
  if (best.dist == 1 && !resumed)                       // :4137-4141
  {
    best.dist = 0;
    tz_two_point<Px, PACKED, GS>(J, ref00, pitch, s_org, best);
  }
  if (best.dist > 5 && !(resumed && park->frac_cost))   // raster, step 5 (:4144-4154); parked in the middle of the refinement: already behind it
  {
    best.dist = 5;
    const int nx = (rR - rL) / 5 + 1, ny = (rB - rT) / 5 + 1;
    const int total = (rR >= rL && rB >= rT) ? nx * ny : 0;
    for (int base = 0; base < total; base += GS)
    {
      const int i = base + gl;
      const int gy = i / nx, gx = i - gy * nx;
      tz_eval<Px, PACKED, GS>(J, ref00, pitch, s_org, 1, rL + gx * 5, rT + gy * 5, i < total, 0, 5, best);
    }
  }
  while (best.dist > 0)                                 // star refinement (:4189-4223)
  {
    cx = best.x; cy = best.y;
    best.dist = 0; best.pnr = 0;
    if (MERGE && GS == 32)
    {
      // All rings of a refinement round sit around the same centre and there is no early stop, so their points are costed
      // together, one lane per point in emission order: distances 1, 2, 4, 8 (28 points) in one pass, then two 16-point rings
      // per pass.  The strict-'<' scan over the whole round keeps the first minimum, which is what the passes, taken in order,
      // keep (tz_eval).  Three passes instead of seven for SearchRange 64.
      {
        const int rr = gl < 4 ? 0 : (gl < 12 ? 1 : (gl < 20 ? 2 : 3));
        int x, y, pnr, dist;
        tz_diamond_point(cx, cy, 1 << rr, rr == 0 ? gl : ((gl + 4) & 7), x, y, pnr, dist);
        const bool valid = gl < 28 && (1 << rr) <= jb.search_range && tz_in_window(J, cx, cy, x, y);
        tz_eval<Px, PACKED, GS>(J, ref00, pitch, s_org, 1, x, y, valid, pnr, dist, best);
      }
      for (int d = 16; d <= jb.search_range; d <<= 2)
      {
        const int dd = gl < 16 ? d : d << 1;
        int x, y, pnr, dist;
        tz_diamond_point(cx, cy, dd, gl & 15, x, y, pnr, dist);
        const bool valid = dd <= jb.search_range && tz_in_window(J, cx, cy, x, y);
        tz_eval<Px, PACKED, GS>(J, ref00, pitch, s_org, 1, x, y, valid, pnr, dist, best);
      }
    }
    else
      for (int d = 1; d < jb.search_range + 1; d <<= 1)
        tz_diamond<Px, PACKED, GS>(J, ref00, pitch, s_org, cx, cy, d, best);
    if (best.dist == 1)
    {
      best.dist = 0;
      if (best.pnr != 0) tz_two_point<Px, PACKED, GS>(J, ref00, pitch, s_org, best);
    }
  }

  if (gl == 0)
  {
    out.int_x = (int16_t)best.x; out.int_y = (int16_t)best.y;
    out.int_sad = best.cost - hm_mv_cost(J.ui_cost, J.pred_x, J.pred_y, 2, best.x, best.y);
    out.half_x = out.half_y = out.qter_x = out.qter_y = 0;
    out.frac_cost = 0;
    out.n_cand = best.n_cand;
  }
}

// one warp per job (kept for the fused low-latency kernel)
template <typename Px, bool PACKED>
__device__ __forceinline__ void tz_search_warp(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks,
                                               const RefTable& refs, const OrgView& org, unsigned char* s_org,
                                               hmgpu_me_result& out, const uint8_t* win_s = NULL, const TzWindow* win = NULL,
                                               TzSpec* spec = NULL, int role = 0)
{
  tz_search_group<Px, PACKED, 32>(jb, org_blocks, refs, org, s_org, out, win_s, win, spec, role);
}

// =====================================================================================================================
// xTZSearchSelective (FastSearch = 2; TEncSearch.cpp:4231-4383 with SEL_SEARCH_CONFIGURATION :317-330), one warp per job.
//
// Its xTZSearchHelp branch (:360-406) costs a point progressively -- a coarse row-sampled SAD, then the rows in between level
// by level -- and abandons the point when an ESTIMATE (not a bound) fails to beat the running best, so whether a point is
// accepted depends on the best cost at the moment the reference reaches it.  The device therefore separates arithmetic from
// decisions: every lane computes, for ITS point, the raw row sums of all levels (no early exit), and the warp then replays the
// reference's decisions point by point in emission order on those sums (sel_replay).  This keeps the mode bit-exact; it is not
// tuned (no BASELINE configuration uses FastSearch = 2).
// =====================================================================================================================

struct SelPoint { int x, y, pnr, dist; bool valid; };

// raw row sums of one point by one lane: level 0 = rows 0, 2^S, ..; level l = rows at offset 2^(S-l), step 2^(S-l+1)
template <typename Px, bool PACKED>
__device__ __forceinline__ void sel_point_sums(const TzJob& J, const Px* ref00, int pitch, const void* org_s, int S,
                                               const SelPoint& p, uint32_t (&raw)[5])
{
#pragma unroll
  for (int l = 0; l < 5; l++) raw[l] = 0;
  if (!p.valid) return;
  const Px* ref = ref00 + (ptrdiff_t)p.y * pitch + p.x;
  for (int r = 0; r < J.pu_h; r++)
  {
    uint32_t rs = 0;
    if (PACKED)
    {
      const uint8_t* q8 = (const uint8_t*)ref + (ptrdiff_t)r * pitch;
      const uintptr_t a0 = (uintptr_t)q8;
      const int sh = (int)(a0 & 3) * 8;
      const uint32_t* q = (const uint32_t*)(a0 & ~(uintptr_t)3);
      const uint32_t* o = (const uint32_t*)org_s + r * (J.pu_w >> 2);
      uint32_t lo = __ldg(q);
      for (int k = 0; k < (J.pu_w >> 2); k++)
      {
        const uint32_t hi = __ldg(q + k + 1);
        rs = vabsdiff4_acc(__funnelshift_r(lo, hi, sh), o[k], rs);
        lo = hi;
      }
    }
    else
    {
      const Px* pr = ref + (ptrdiff_t)r * pitch;
      const int16_t* o = (const int16_t*)org_s + r * J.pu_w;
      for (int k = 0; k < J.pu_w; k++) rs += (uint32_t)hm_abs((int)o[k] - (int)__ldg(pr + k));
    }
    // level of row r: multiples of 2^S are level 0, otherwise S - ctz(r)
    const int tzc = __ffs(r | (1 << S)) - 1;               // min(ctz(r), S)
    const int lvl = tzc >= S ? 0 : S - tzc;
#pragma unroll
    for (int l = 0; l < 5; l++) if (lvl == l) raw[l] += rs;
  }
}

// replay of xTZSearchHelp's selective branch over the points held by lanes 0 .. n-1, in lane (= emission) order
__device__ __forceinline__ void sel_replay(const TzJob& J, int S, int n, const SelPoint& p, const uint32_t (&raw)[5], TzBest& best)
{
  const int bds = J.bit_depth - 8;
  const uint32_t bit = hm_mv_cost(J.ui_cost, J.pred_x, J.pred_y, 2, p.x, p.y);
  for (int i = 0; i < n; i++)
  {
    const bool v = __shfl_sync(0xffffffffu, (int)p.valid, i) != 0;
    if (!v) continue;                                      // uniform
    const int x = __shfl_sync(0xffffffffu, p.x, i), y = __shfl_sync(0xffffffffu, p.y, i);
    const int pnr = __shfl_sync(0xffffffffu, p.pnr, i), dist = __shfl_sync(0xffffffffu, p.dist, i);
    const uint32_t bc = __shfl_sync(0xffffffffu, bit, i);
    uint32_t r[5];
#pragma unroll
    for (int l = 0; l < 5; l++) r[l] = __shfl_sync(0xffffffffu, raw[l], i);
    best.n_cand += 1;
    int sh = S;
    uint32_t tmp = (r[0] << sh) >> bds;
    if (tmp + bc < best.cost)
    {
      uint32_t sad = tmp >> sh;
      int l = 1;
      while (sh > 0)
      {
        const int is = sh - 1;
        const uint32_t rl = l == 1 ? r[1] : (l == 2 ? r[2] : (l == 3 ? r[3] : r[4]));
        tmp = (rl << sh) >> bds;
        sad += tmp >> sh;
        if (((sad << is) + bc) > best.cost) break;
        sh--; l++;
      }
      if (sh == 0)
      {
        sad += bc;
        if (sad < best.cost) { best.cost = sad; best.x = x; best.y = y; best.dist = dist; best.pnr = pnr; best.round = 0; }
      }
    }
  }
}

template <typename Px, bool PACKED>
__device__ __forceinline__ void sel_batch(const TzJob& J, const Px* ref00, int pitch, const void* org_s, int S, int n,
                                          const SelPoint& p, TzBest& best)
{
  uint32_t raw[5];
  sel_point_sums<Px, PACKED>(J, ref00, pitch, org_s, S, p, raw);
  sel_replay(J, S, n, p, raw, best);
}

// side data of a selective job: m_acMvPredictors[MD_LEFT, MD_ABOVE, MD_ABOVE_RIGHT] as 6 int16 at org_blocks[org_offset]
template <typename Px, bool PACKED>
__device__ __forceinline__ void tz_selective_warp(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks,
                                                  const RefTable& refs, const OrgView& org, unsigned char* s_org,
                                                  hmgpu_me_result& out)
{
  const int lane = threadIdx.x & 31;
  TzJob J;
  J.pu_w = jb.pu_w; J.pu_h = jb.pu_h; J.sub_shift = 0; J.rows = jb.pu_h;
  J.pred_x = jb.pred_x; J.pred_y = jb.pred_y; J.ui_cost = jb.ui_cost;
  J.L = jb.win_l; J.T = jb.win_t; J.R = jb.win_r; J.B = jb.win_b;
  J.bit_depth = refs.bit_depth;
  J.win_s = NULL; J.win_pitch = 0; J.wx0 = J.wy0 = 0; J.wx1 = J.wy1 = -1; J.win_ox = J.win_oy = 0;
  const int pitch = refs.pitch;
  const Px* ref00 = (const Px*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;
  tz_stage_org<Px, PACKED>(jb, org_blocks, org, s_org, lane, 32);
  __syncwarp();
  const int S = jb.pu_h > 32 ? 4 : (jb.pu_h > 16 ? 3 : (jb.pu_h > 8 ? 2 : 1));      // :367-374
  TzBest best; best.cost = 0xffffffffu; best.x = 0; best.y = 0; best.dist = 0; best.round = 0; best.pnr = 0; best.n_cand = 0;
  const bool has2n = (jb.flags & HMGPU_F_HAS_2NX2N) != 0;
  auto clipq = [&](int vx, int vy, int& ox, int& oy) {
    ox = min((int)jb.clip_hmax, max((int)jb.clip_hmin, vx)) >> 2;
    oy = min((int)jb.clip_vmax, max((int)jb.clip_vmin, vy)) >> 2;
  };

  // ---- start points (:4267-4295): the MVP, the three spatial predictors, zero, the 2Nx2N integer MV ----
  {
    SelPoint p; p.x = 0; p.y = 0; p.pnr = 0; p.dist = 0; p.valid = lane < (has2n ? 6 : 5);
    if (lane == 0) clipq(jb.start_x, jb.start_y, p.x, p.y);
    else if (lane >= 1 && lane <= 3)
      clipq((int)__ldcv(org_blocks + jb.org_offset + 2 * (lane - 1)), (int)__ldcv(org_blocks + jb.org_offset + 2 * (lane - 1) + 1), p.x, p.y);
    else if (lane == 5) clipq((int)(int16_t)(jb.i2n_x << 2), (int)(int16_t)(jb.i2n_y << 2), p.x, p.y);
    sel_batch<Px, PACKED>(J, ref00, pitch, s_org, S, has2n ? 6 : 5, p, best);
  }
  int rL = J.L, rT = J.T, rR = J.R, rB = J.B;              // iSrchRng*: re-centred when the 2Nx2N MV was tested (:4297-4306)
  if (has2n)
  {
    const int px = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(best.x << 2)));
    const int py = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(best.y << 2)));
    const int sr4 = jb.search_range << 2;
    rL = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(px - sr4))) >> 2;
    rT = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(py - sr4))) >> 2;
    rR = min((int)jb.clip_hmax, max((int)jb.clip_hmin, (int)(int16_t)(px + sr4))) >> 2;
    rB = min((int)jb.clip_vmax, max((int)jb.clip_vmin, (int)(int16_t)(py + sr4))) >> 2;
  }

  // ---- initial search (:4308-4324): step-4 grid of centres, each with the diamonds at distance 1 and 2 (13 points) ----
  const int bx0 = best.x, by0 = best.y;
  {
    const int sri = jb.search_range >> 2;
    const int fl = max(bx0 - sri, rL), ft = max(by0 - sri, rT), fr = min(bx0 + sri, rR), fb = min(by0 + sri, rB);
    const int gnx = fr >= fl ? (fr - fl) / 4 + 1 : 0, gny = fb >= ft ? (fb - ft) / 4 + 1 : 0;
    const int n_centres = gnx * gny;
    for (int c0 = 0; c0 < n_centres; c0 += 2)              // two centres (26 points) per pass, emission order = lane order
    {
      const int ci = c0 + (lane >= 13 ? 1 : 0), k = lane >= 13 ? lane - 13 : lane;
      SelPoint p; p.x = 0; p.y = 0; p.pnr = 0; p.dist = 0; p.valid = false;
      if (lane < 26 && ci < n_centres)
      {
        const int cx = fl + (ci % gnx) * 4, cy = ft + (ci / gnx) * 4;
        if (k == 0) { p.x = cx; p.y = cy; p.valid = true; }
        else
        {
          const int d = k <= 4 ? 1 : 2, i = k <= 4 ? k - 1 : k - 5;
          tz_diamond_point(cx, cy, d, i, p.x, p.y, p.pnr, p.dist);
          p.valid = tz_in_window(J, cx, cy, p.x, p.y);
        }
      }
      sel_batch<Px, PACKED>(J, ref00, pitch, s_org, S, 26, p, best);
    }
  }

  const bool far_from_pred = abs(best.x - bx0) > 8 || abs(best.y - by0) > 8;          // iMVDistThresh (:4326)
  if (far_from_pred)
  {
    // every position of the window, raster order, distance 1 (:4329-4338)
    const int nx = rR - rL + 1, ny = rB - rT + 1;
    const int total = (nx > 0 && ny > 0) ? nx * ny : 0;
    for (int b0 = 0; b0 < total; b0 += 32)
    {
      const int i = b0 + lane;
      SelPoint p; p.pnr = 0; p.dist = 1; p.valid = i < total;
      p.x = rL + (p.valid ? i % nx : 0); p.y = rT + (p.valid ? i / nx : 0);
      sel_batch<Px, PACKED>(J, ref00, pitch, s_org, S, 32, p, best);
    }
  }
  else
  {
    while (best.dist > 0)                                  // star refinement (:4340-4375)
    {
      const int cx = best.x, cy = best.y;
      best.dist = 0; best.pnr = 0;
      for (int d = 1; d < jb.search_range + 1; d <<= 1)
      {
        const int npts = d == 1 ? 4 : (d <= 8 ? 8 : 16);
        SelPoint p; p.x = 0; p.y = 0; p.pnr = 0; p.dist = 0; p.valid = false;
        if (lane < npts)
        {
          tz_diamond_point(cx, cy, d, lane, p.x, p.y, p.pnr, p.dist);
          p.valid = tz_in_window(J, cx, cy, p.x, p.y);
        }
        sel_batch<Px, PACKED>(J, ref00, pitch, s_org, S, npts, p, best);
      }
      if (best.dist == 1)
      {
        best.dist = 0;
        if (best.pnr >= 1 && best.pnr <= 8)
        {
          const int nr = best.pnr, tcx = best.x, tcy = best.y;
          SelPoint p; p.pnr = 0; p.dist = 2; p.valid = false; p.x = 0; p.y = 0;
          if (lane < 2)
          {
            p.x = tcx + c_two_point[nr][lane == 0 ? 0 : 2];
            p.y = tcy + c_two_point[nr][lane == 0 ? 1 : 3];
            p.valid = tz_in_window(J, tcx, tcy, p.x, p.y);
          }
          sel_batch<Px, PACKED>(J, ref00, pitch, s_org, S, 2, p, best);
        }
      }
    }
  }

  if (lane == 0)
  {
    out.int_x = (int16_t)best.x; out.int_y = (int16_t)best.y;
    out.int_sad = best.cost - hm_mv_cost(J.ui_cost, J.pred_x, J.pred_y, 2, best.x, best.y);
    out.half_x = out.half_y = out.qter_x = out.qter_y = 0;
    out.frac_cost = 0;
    out.n_cand = best.n_cand;
  }
}
