// me_fracw.cu -- fractional refinement of 8-bit pictures, batch path: ONE kernel, one CTA per GROUP of jobs = all the jobs of
// one batch that lie in the same CTU (64x64 luma samples) and search the same reference slot.
//
// Replaces TEncSearch::xPatternSearchFracDIF + xPatternRefinement (TEncSearch.cpp:4386-4422, 799-852) over the 16 quarter-pel
// phase planes (planes.cu), like me_frac2.cu, whose ncu captures (profiles/r1n_ncu_frac2_dist.csv) showed both of its kernels
// bound by L1: every tile fetched the rows of every candidate on its own (4-byte cp.async with per-thread address arithmetic,
// 213 L1 wavefronts per warp and candidate against 168 issue cycles), the planes were read 1.9 x from DRAM, and the per-candidate
// sums and the half-pel winner travelled through global memory between six launches.  Here:
//
//  * GROUPS.  A counting sort (bin / scan / scatter kernels, one thread per job) orders the batch by (reference slot, CTU).  The
//    jobs of a group look at the same neighbourhood of the same planes: the PUs of the CTU displaced by vectors that differ by a
//    few samples.  Groups are handed out in (slot, CTU raster) order through an atomic counter, so the CTAs in flight walk along
//    a CTU row of one reference: the planes come from DRAM once.
//  * WINDOWS BY TMA.  The CTA stages ONE window per phase plane for the whole group: a FW_WW x FW_WH box (144 x 96 bytes)
//    around the CTU displaced by the mean integer vector of the group, copied by cp.async.bulk.tensor from the 4-D tensor map
//    over the planes of all reference slots (x, y, phase plane, slot) against an mbarrier -- 4 boxes (the even-even planes)
//    for the nine half-pel candidates, 12 boxes for the eight quarter-pel candidates.  No thread computes a global address on
//    the fast path.  A job whose footprint leaves the window (a vector more than ~15 samples from the mean), or that wants SAD
//    instead of SATD (HADME off / lossless), takes the SLOW path of the same kernel: per-tile loads from global memory.
//  * ONE LAUNCH.  Per-candidate sums are shared-memory accumulators of the group, the half-pel winner is chosen between two
//    __syncthreads and never leaves the SM, the result is written once.
//  * BANKS.  The window pitch is 36 words, so row r starts at bank 4r: vertically adjacent 8x8 tiles (8 rows apart) would hit the
//    same banks.  The Hadamard transform only changes the SIGNS of its coefficients when the input rows are permuted by r -> r^c,
//    and SATD sums absolute values, so every lane walks its rows in the order r^c with c chosen from the tile row (source rows
//    permuted alike): the tiles of a PU read 16 different banks at every step.
//
// Arithmetic per tile and candidate as in me_frac2.cu: horizontal Hadamard pass on packed bytes with IDP.4A (linearity:
// H(org) - H(ref)), vertical pass and |.| sum in registers, source tile in registers for the 2-3 candidates of a work item.
#include <cuda.h>
#include "me_frac_impl.cuh"
#include <stdlib.h>

#define FW_THREADS 512
#define FW_WW 144                      // window row bytes (TMA box inner dimension, multiple of 16)
#define FW_WH 96                       // window rows
#define FW_PLANE_BYTES (FW_WW * FW_WH) // 13 824 = 108 x 128 (TMA destinations are 128-byte aligned)
#define FW_NBUF 12
#define FW_GMAX 640                    // jobs per pass over a group
#define FW_T8MAX 2048                  // 8x8 tiles per pass
#define FW_T4MAX 3072                  // 4x4 tiles per pass
#define FW_XOFF 32                     // the window starts this far left of (CTU + mean vector), before the 16-byte alignment
#define FW_YOFF 16
#define FW_ORG_PITCH 17                // words per row of the staged source CTU (odd: rows start on different banks)

enum { FWF_FAST = 1, FWF_SATD = 2, FWF_T8 = 4 };

struct __align__(16) FwJob
{
  int16_t xw, yw;       // PU origin displaced by the integer vector, window coordinates
  int16_t mvx, mvy;     // integer vector
  uint16_t inv_tw;      // ceil(32768 / tiles per row): tile row = (tile * inv_tw) >> 15, exact for tile < 64
  uint8_t rx, ry;       // PU origin inside the CTU
  uint8_t w, h;
  uint8_t flags;
  uint8_t tiles;
};                      // 16 bytes: one LDS.128

struct FwMisc
{
  unsigned long long bar;
  int group, sumx, sumy, m, n_slow, next;
  uint32_t warp_tot[FW_THREADS / 32];
};

struct FwSmem
{
  unsigned char planes[FW_NBUF][FW_PLANE_BYTES];
  uint32_t org[64 * FW_ORG_PITCH];
  uint32_t acc[9][FW_GMAX];
  FwJob job[FW_GMAX];
  uint32_t jidx[FW_GMAX];
  uint32_t off[FW_GMAX];               // exclusive tile offsets of the pass (low half: 8x8 list, high half: 4x4 list)
  uint16_t t8[FW_T8MAX];
  uint16_t t4[FW_T4MAX];
  uint16_t slow[FW_GMAX];
  uint8_t hsel[FW_GMAX];               // half-pel winner of a job: (hx + 1) | (hy + 1) << 2
  FwMisc misc;
};

// buffer of phase plane p = fy * 4 + fx in the quarter-pel pass, one nibble per plane (f: even-even plane, not a quarter-pel
// candidate): 1 -> 0, 3 -> 1, 4..7 -> 2..5, 9 -> 6, 11 -> 7, 12..15 -> 8..11
#define FW_SLOT_LUT 0xBA987F6F54321F0Full
__host__ __device__ __forceinline__ int fw_slot(int p) { return (int)((FW_SLOT_LUT >> (4 * p)) & 15ull); }

// ---- counting sort by (slot, CTU) -------------------------------------------------------------------------------------------
__device__ __forceinline__ int fw_bin(const hmgpu_me_job& jb, int ctus_x, int n_ctus) { return jb.ref_slot * n_ctus + (jb.pu_y >> 6) * ctus_x + (jb.pu_x >> 6); }

__global__ void fracw_bin_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, hmgpu_me_result* __restrict__ results,
                                 uint32_t* __restrict__ bin_count, int ctus_x, int n_ctus)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int bin = -1 - (int)(threadIdx.x & 31);                 // lanes without a job to count: distinct keys
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
    if (jb.flags & HMGPU_F_FRAC) bin = fw_bin(jb, ctus_x, n_ctus);
  }
  // neighbouring jobs are the partitions of one CU x the reference pictures: a warp counts into a handful of bins
  const uint32_t same = __match_any_sync(0xffffffffu, bin);
  if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&bin_count[bin], (uint32_t)__popc(same));
}

// one CTA: exclusive scan of the job counts of the bins and the list of the non-empty bins; totals[0] = jobs, totals[1] = groups
__global__ void __launch_bounds__(1024)
fracw_scan_kernel(const uint32_t* __restrict__ bin_count, int n_bins, uint32_t* __restrict__ bin_start,
                  uint32_t* __restrict__ group_list, uint32_t* __restrict__ totals)
{
  __shared__ uint32_t s_j[1024], s_g[1024];
  const int t = threadIdx.x;
  const int per = (n_bins + 1023) / 1024;
  const int b0 = min(n_bins, t * per), b1 = min(n_bins, b0 + per);
  uint32_t sj = 0, sg = 0;
  for (int b = b0; b < b1; b++) { const uint32_t c = bin_count[b]; sj += c; sg += c ? 1u : 0u; }
  s_j[t] = sj; s_g[t] = sg;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1)
  {
    const uint32_t aj = t >= o ? s_j[t - o] : 0, ag = t >= o ? s_g[t - o] : 0;
    __syncthreads();
    s_j[t] += aj; s_g[t] += ag;
    __syncthreads();
  }
  uint32_t oj = s_j[t] - sj, og = s_g[t] - sg;
  for (int b = b0; b < b1; b++)
  {
    const uint32_t c = bin_count[b];
    bin_start[b] = oj;
    if (c) group_list[og++] = (uint32_t)b;
    oj += c;
  }
  if (t == 1023) { totals[0] = s_j[t]; totals[1] = s_g[t]; }
}

__global__ void fracw_scatter_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, const uint32_t* __restrict__ bin_start,
                                     uint32_t* __restrict__ bin_fill, uint32_t* __restrict__ sorted, int ctus_x, int n_ctus)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int b = -1 - lane;
  if (j < n_jobs)
  {
    const hmgpu_me_job jb = jobs[j];
    if (jb.flags & HMGPU_F_FRAC) b = fw_bin(jb, ctus_x, n_ctus);
  }
  const uint32_t same = __match_any_sync(0xffffffffu, b);
  const int leader = __ffs(same) - 1;
  uint32_t pos = 0;
  if (b >= 0 && lane == leader) pos = bin_start[b] + atomicAdd(&bin_fill[b], (uint32_t)__popc(same));
  pos = __shfl_sync(0xffffffffu, pos, leader);
  if (b >= 0) sorted[pos + __popc(same & ((1u << lane) - 1u))] = (uint32_t)j;
}

// ---- TMA / mbarrier ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fw_tma_box(uint32_t dst, const void* tmap, int x, int y, int z, int slot, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(slot), "r"(bar) : "memory");
}
__device__ __forceinline__ void fw_mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "FW_WAIT:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra FW_DONE;\n\t"
    "bra FW_WAIT;\n\t"
    "FW_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// ---- one tile, one candidate, rows from a staged window -----------------------------------------------------------------------
// win: the plane's window; (xw, yw): top-left sample of the candidate's tile in window coordinates; ow: source tile, rows already
// permuted by r -> r ^ c
template <int TS>
__device__ __forceinline__ uint32_t fw_tile_satd(const unsigned char* win, int xw, int yw, const uint32_t (&ow)[TS][TS / 4], const int (&ro)[TS])
{
  const int sh = (xw & 3) * 8;
  const uint32_t* q = (const uint32_t*)(win + yw * FW_WW + (xw & ~3));
  int d[TS * TS];
  int zero[TS];
#pragma unroll
  for (int k = 0; k < TS; k++) zero[k] = 0;
#pragma unroll
  for (int r = 0; r < TS; r++)
  {
    const uint32_t* row = q + ro[r];
    const uint32_t w0 = row[0], w1 = row[1];
    int h[TS];
    if (TS == 8)
    {
      const uint32_t w2 = row[2];
      had_row8<false>(ow[r][0], ow[r][1], zero, h);
      had_row8<true>(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), h, d + r * TS);
    }
    else
    {
      had_row4<false>(ow[r][0], zero, h);
      had_row4<true>(__funnelshift_r(w0, w1, sh), h, d + r * TS);
    }
  }
  return TS == 8 ? had_cols8_abs(d) : had_cols4_abs(d);
}

// ---- one tile, TWO candidates at once --------------------------------------------------------------------------------------
// The horizontal pass leaves (H(org) - H(ref B)) * 65536 + (H(org) - H(ref A)) in one register per coefficient: the IDP.4A chain of
// candidate A simply continues on the shifted result of candidate B, and the row transform of the source tile is computed once
// for both.  A register then holds the pair as the INTEGER 65536 * hi + lo with lo signed ("carry-save" halves: the raw upper
// half is hi - (lo < 0)); plain 32-bit additions and subtractions act on both halves at once without interference as long as
// |lo| < 32768 -- the coefficients of an 8x8 Hadamard transform of 9-bit differences stay below 16 384.  The last butterfly
// stage adds 0x80008000, which turns both halves into offset-binary values, so that VIMNMX.U16x2 compares them as signed
// numbers; |a + b| + |a - b| = 2 max(|a|, |b|) = 2 max(max(a, b), -min(a, b)) folds the final stage into two packed maxima, one
// packed minimum and one subtraction (-v in offset binary is 0x10000 - v per half, i.e. 0x00010000 - x for the pair).
#define FW_BIAS 0x80008000u
__device__ __forceinline__ uint32_t fw_pair_absmax(uint32_t ba, uint32_t bb)
{
  const uint32_t hi = __vmaxu2(ba, bb), lo = __vminu2(ba, bb);
  return __vmaxu2(hi, 0x00010000u - lo);        // offset-binary max(|a|, |b|) of both halves
}

template <int TS>
__device__ __forceinline__ void fw_tile_satd_pair(const unsigned char* win_a, int xa, int ya, const unsigned char* win_b, int xb, int yb,
                                                  const uint32_t (&ow)[TS][TS / 4], const int (&ro)[TS], uint32_t& out_a, uint32_t& out_b)
{
  const int sha = (xa & 3) * 8, shb = (xb & 3) * 8;
  const uint32_t* qa = (const uint32_t*)(win_a + ya * FW_WW + (xa & ~3));
  const uint32_t* qb = (const uint32_t*)(win_b + yb * FW_WW + (xb & ~3));
  int x[TS * TS];
  int zero[TS];
#pragma unroll
  for (int k = 0; k < TS; k++) zero[k] = 0;
#pragma unroll
  for (int r = 0; r < TS; r++)
  {
    const uint32_t* ra = qa + ro[r];
    const uint32_t* rb = qb + ro[r];
    int h[TS], t[TS];
    if (TS == 8)
    {
      const uint32_t a0 = ra[0], a1 = ra[1], a2 = ra[2], b0 = rb[0], b1 = rb[1], b2 = rb[2];
      had_row8<false>(ow[r][0], ow[r][1], zero, h);
      had_row8<true>(__funnelshift_r(b0, b1, shb), __funnelshift_r(b1, b2, shb), h, t);
#pragma unroll
      for (int k = 0; k < TS; k++) t[k] = t[k] * 65536 + h[k];
      had_row8<true>(__funnelshift_r(a0, a1, sha), __funnelshift_r(a1, a2, sha), t, x + r * TS);
    }
    else
    {
      const uint32_t a0 = ra[0], a1 = ra[1], b0 = rb[0], b1 = rb[1];
      had_row4<false>(ow[r][0], zero, h);
      had_row4<true>(__funnelshift_r(b0, b1, shb), h, t);
#pragma unroll
      for (int k = 0; k < TS; k++) t[k] = t[k] * 65536 + h[k];
      had_row4<true>(__funnelshift_r(a0, a1, sha), t, x + r * TS);
    }
  }
  uint32_t sum = 0;                              // packed: both halves stay below 65 536
  if (TS == 8)
  {
    uint32_t part[4] = { 0, 0, 0, 0 };           // each collects 8 maxima of at most 8 160
#pragma unroll
    for (int col = 0; col < 8; col++)
    {
      int* v = x + col;
#pragma unroll
      for (int i = 0; i < 4; i++) { const int a = v[i * 8], b = v[(i + 4) * 8]; v[i * 8] = a + b; v[(i + 4) * 8] = a - b; }
      uint32_t m[4];
#pragma unroll
      for (int i = 0; i < 8; i += 4)
      {
        // second stage with the bias folded in, third stage inside the maxima
        const uint32_t p0 = (uint32_t)(v[i * 8] + v[(i + 2) * 8]) + FW_BIAS, p2 = (uint32_t)(v[i * 8] - v[(i + 2) * 8]) + FW_BIAS;
        const uint32_t p1 = (uint32_t)(v[(i + 1) * 8] + v[(i + 3) * 8]) + FW_BIAS, p3 = (uint32_t)(v[(i + 1) * 8] - v[(i + 3) * 8]) + FW_BIAS;
        m[i / 2] = fw_pair_absmax(p0, p1);
        m[i / 2 + 1] = fw_pair_absmax(p2, p3);
      }
      // two offset-binary halves + two more - 0x00010000: the biases leave through the carries (see the file header)
      part[col & 3] += (m[0] + m[1] + 0xffff0000u) + (m[2] + m[3] + 0xffff0000u);
    }
    const uint32_t lo = (part[0] & 0xffffu) + (part[1] & 0xffffu) + (part[2] & 0xffffu) + (part[3] & 0xffffu);
    const uint32_t hi = (part[0] >> 16) + (part[1] >> 16) + (part[2] >> 16) + (part[3] >> 16);
    out_a = (2u * lo + 2) >> 2; out_b = (2u * hi + 2) >> 2;
  }
  else
  {
#pragma unroll
    for (int col = 0; col < 4; col++)
    {
      const int* v = x + col;
      const uint32_t a0 = (uint32_t)(v[0] + v[8]) + FW_BIAS, a1 = (uint32_t)(v[4] + v[12]) + FW_BIAS;
      const uint32_t a2 = (uint32_t)(v[0] - v[8]) + FW_BIAS, a3 = (uint32_t)(v[4] - v[12]) + FW_BIAS;
      sum += fw_pair_absmax(a0, a1) + fw_pair_absmax(a2, a3) + 0xffff0000u;     // 8 maxima of at most 2 040 in all
    }
    out_a = (2u * (sum & 0xffffu) + 1) >> 1; out_b = (2u * (sum >> 16) + 1) >> 1;
  }
}

// one warp-wide batch of work items of one list (8x8 or 4x4 tiles) in one phase: item = (tile, group of CPI candidates), idx = the
// lane's item (items of one candidate group are consecutive: the lanes of a warp hold consecutive tiles)
template <int TS, int PHASE, int NP>
__device__ __forceinline__ void fw_items(FwSmem& S, const uint16_t* tab, int n_tiles, int idx, int lane)
{
  // An item = one tile with NP pairs of candidates (8x8 tiles: one pair; 4x4 tiles: two, so that the per-item work -- table
  // look-up, tile geometry, source rows, the reduction below -- is paid once per four candidates).  Half-pel: candidate 0 alone
  // (group 0), then the pairs (1,2) (3,4) (5,6) (7,8); quarter-pel: the four pairs (candidate 0 is the half-pel winner, whose
  // sum is reused).
  constexpr int NGRP = PHASE ? 4 / NP : 1 + 4 / NP;
  const bool valid = idx < n_tiles * NGRP;
  uint32_t ji = 0xffffu;
  uint32_t v[2 * NP];
#pragma unroll
  for (int k = 0; k < 2 * NP; k++) v[k] = 0;
  int grp = 0;
  if (valid)
  {
#pragma unroll
    for (int g = 1; g < NGRP; g++) grp += idx >= g * n_tiles;
    const uint32_t e = tab[idx - grp * n_tiles];
    ji = e >> 6;
    const int t = (int)(e & 63u);
    const FwJob jb = S.job[ji];
    const int hx = PHASE ? (int)(S.hsel[ji] & 3) - 1 : 0, hy = PHASE ? (int)(S.hsel[ji] >> 2) - 1 : 0;
    const int tw = jb.w / TS;
    const int ty = (t * jb.inv_tw) >> 15, tx = t - ty * tw;
    const int c = TS == 8 ? (2 * ty) & 7 : ty & 3;
    uint32_t ow[TS][TS / 4];
    int ro[TS];                                   // window row offsets (words) in the lane's row order r ^ c
    {
      const uint32_t* o = S.org + (jb.ry + ty * TS) * FW_ORG_PITCH + ((jb.rx + tx * TS) >> 2);
#pragma unroll
      for (int r = 0; r < TS; r++)
      {
        ro[r] = (r ^ c) * (FW_WW / 4);
#pragma unroll
        for (int w = 0; w < TS / 4; w++) ow[r][w] = o[(r ^ c) * FW_ORG_PITCH + w];
      }
    }
    const int x0 = jb.xw + tx * TS, y0 = jb.yw + ty * TS;
    if (PHASE == 0 && grp == 0) v[0] = fw_tile_satd<TS>(S.planes[0], x0, y0, ow, ro);
    else
    {
#pragma unroll
      for (int p = 0; p < NP; p++)
      {
        const int ca = PHASE ? 1 + 2 * (grp * NP + p) : 2 * ((grp - 1) * NP + p) + 1;
        int qxa, qya, qxb, qyb, sa, sb;
        if (PHASE == 0)
        {
          qxa = 2 * c_refine_h[ca][0]; qya = 2 * c_refine_h[ca][1]; qxb = 2 * c_refine_h[ca + 1][0]; qyb = 2 * c_refine_h[ca + 1][1];
          sa = ((qya & 3) >> 1) * 2 + ((qxa & 3) >> 1); sb = ((qyb & 3) >> 1) * 2 + ((qxb & 3) >> 1);
        }
        else
        {
          qxa = 2 * hx + c_refine_q[ca][0]; qya = 2 * hy + c_refine_q[ca][1]; qxb = 2 * hx + c_refine_q[ca + 1][0]; qyb = 2 * hy + c_refine_q[ca + 1][1];
          sa = fw_slot((qya & 3) * 4 + (qxa & 3)); sb = fw_slot((qyb & 3) * 4 + (qxb & 3));
        }
        fw_tile_satd_pair<TS>(S.planes[sa], x0 + (qxa >> 2), y0 + (qya >> 2), S.planes[sb], x0 + (qxb >> 2), y0 + (qyb >> 2), ow, ro,
                              v[2 * p], v[2 * p + 1]);
      }
    }
  }
  // The lanes of a (job, candidate group) add up their tiles with a segmented shuffle reduction -- they are consecutive lanes --
  // and the first lane of the run issues one shared-memory atomic per candidate.  (__reduce_add_sync on the match mask was 29 %
  // of the kernel's stall samples: REDUX with a partial mask is a loop.)
  const uint32_t key = ji | ((uint32_t)grp << 16);
  const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = lane == 0 || prev != key;
  const uint32_t heads = __ballot_sync(0xffffffffu, head);
  if (heads != 0xffffffffu)                                 // (every lane its own job: nothing to add up)
  {
    const uint32_t above = heads & ~((2u << lane) - 1u);    // heads at higher lanes (lane 31: 2u << 31 == 0, mask of all lanes)
    const int run_end = above ? __ffs(above) - 2 : 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
#pragma unroll
      for (int k = 0; k < 2 * NP; k++)
      {
        const uint32_t t = __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane + o <= run_end) v[k] += t;
      }
    }
  }
  if (head && valid)
  {
    if (PHASE == 0 && grp == 0) atomicAdd(&S.acc[0][ji], v[0]);
    else
    {
#pragma unroll
      for (int p = 0; p < NP; p++)
      {
        const int ca = PHASE ? 1 + 2 * (grp * NP + p) : 2 * ((grp - 1) * NP + p) + 1;
        atomicAdd(&S.acc[ca][ji], v[2 * p]);
        atomicAdd(&S.acc[ca + 1][ji], v[2 * p + 1]);
      }
    }
  }
}

// all the fast work of one phase: the warps draw batches of 32 items from a shared counter, 8x8 tiles first (the heavier items)
template <int PHASE>
__device__ __forceinline__ void fw_run_phase(FwSmem& S, int n8, int n4)
{
  constexpr int NGRP8 = PHASE ? 2 : 3, NGRP4 = PHASE ? 2 : 3;
  const int lane = threadIdx.x & 31;
  const int end8 = (n8 * NGRP8 + 31) & ~31, end = end8 + n4 * NGRP4;
  // the next batch is drawn before the current one is worked on: the round trip of the atomic hides under the arithmetic
  int next = 0;
  if (lane == 0) next = atomicAdd(&S.misc.next, 32);
  for (;;)
  {
    const int base = __shfl_sync(0xffffffffu, next, 0);
    if (base >= end) break;
    if (lane == 0) next = atomicAdd(&S.misc.next, 32);
    if (base < end8) fw_items<8, PHASE, 2>(S, S.t8, n8, base + lane, lane);
    else fw_items<4, PHASE, 2>(S, S.t4, n4, base - end8 + lane, lane);
  }
}

// jobs the windows do not serve: per (tile, candidate) from global memory (tile_dist, me_frac_impl.cuh)
template <int PHASE>
__device__ __forceinline__ void fw_run_slow(FwSmem& S, const hmgpu_me_job* __restrict__ jobs, const RefTable& refs, const OrgView& org)
{
  const int ncand = PHASE ? 8 : 9;
  for (int s = 0; s < S.misc.n_slow; s++)
  {
    const uint32_t ji = S.slow[s];
    const FwJob fj = S.job[ji];
    const int hx = (int)(S.hsel[ji] & 3) - 1, hy = (int)(S.hsel[ji] >> 2) - 1;
    const hmgpu_me_job jb = jobs[S.jidx[ji]];
    const bool satd = (fj.flags & FWF_SATD) != 0;
    const int ts = (fj.flags & FWF_T8) ? 8 : 4;
    const int tw = jb.pu_w / ts;
    const int n_it = (int)fj.tiles * ncand;
    for (int it = threadIdx.x; it < n_it; it += FW_THREADS)
    {
      const int ci = it / fj.tiles, t = it - ci * fj.tiles;
      const int cand = PHASE ? ci + 1 : ci;
      const int ty = t / tw, tx = t - ty * tw;
      int qx, qy;
      if (PHASE == 0) { qx = 4 * fj.mvx + 2 * c_refine_h[cand][0]; qy = 4 * fj.mvy + 2 * c_refine_h[cand][1]; }
      else { qx = 4 * fj.mvx + 2 * hx + c_refine_q[cand][0]; qy = 4 * fj.mvy + 2 * hy + c_refine_q[cand][1]; }
      const uint8_t* ref = (const uint8_t*)refs.base[jb.ref_slot] + (size_t)((qy & 3) * 4 + (qx & 3)) * refs.plane_elems
                         + (ptrdiff_t)(jb.pu_y + (qy >> 2)) * refs.pitch + (jb.pu_x + (qx >> 2));
      uint32_t v;
      if (ts == 8) v = tile_dist<uint8_t, 8>(jb, NULL, org, ref, refs.pitch, tx * 8, ty * 8, satd);
      else v = tile_dist<uint8_t, 4>(jb, NULL, org, ref, refs.pitch, tx * 4, ty * 4, satd);
      atomicAdd(&S.acc[cand][ji], v);
    }
  }
}

__global__ void __launch_bounds__(FW_THREADS, 1)
fracw_group_kernel(const hmgpu_me_job* __restrict__ jobs, hmgpu_me_result* __restrict__ results, RefTable refs, OrgView org,
                   const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ bin_start, const uint32_t* __restrict__ bin_count,
                   const uint32_t* __restrict__ group_list, const uint32_t* __restrict__ totals, uint32_t* __restrict__ counter,
                   const __grid_constant__ CUtensorMap tmap, int ctus_x, int n_ctus, int bit_depth)
{
  // no integer round trip on this pointer: the compiler must see a shared-memory address (LDS / ATOMS with 32-bit addresses; through
  // a cast the whole structure was accessed with generic LD.E and 64-bit address arithmetic, 14 % of the executed instructions)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FwSmem& S = *reinterpret_cast<FwSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = fw_smem_u32(&S.misc.bar);
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t parity = 0;
  const uint32_t n_groups = totals[1];
  const int pic_w = refs.pic_w, pic_h = refs.pic_h;

  for (;;)
  {
    __syncthreads();                                     // everybody is done with the previous group (misc, tables)
    if (tid == 0) S.misc.group = (int)atomicAdd(counter, 1u);
    __syncthreads();
    const uint32_t g = (uint32_t)S.misc.group;
    if (g >= n_groups) break;
    const int bin = (int)group_list[g];
    const int n = (int)bin_count[bin];
    const uint32_t first = bin_start[bin];
    const int slot = bin / n_ctus, ctu = bin - slot * n_ctus;
    const int ctu_y = (ctu / ctus_x) * 64, ctu_x = (ctu - (ctu / ctus_x) * ctus_x) * 64;

    // source CTU (rows / words past the picture are never read by a PU)
    for (int i = tid; i < 64 * 16; i += FW_THREADS)
    {
      const int y = ctu_y + (i >> 4), x = ctu_x + (i & 15) * 4;
      S.org[(i >> 4) * FW_ORG_PITCH + (i & 15)] = (y < pic_h && x + 4 <= org.pitch) ? __ldg((const uint32_t*)((const uint8_t*)org.base + (size_t)y * org.pitch + x)) : 0u;
    }

    for (int c0 = 0; c0 < n;)
    {
      const int cnt = min(FW_GMAX, n - c0);
      if (tid == 0) { S.misc.sumx = 0; S.misc.sumy = 0; S.misc.m = cnt; S.misc.n_slow = 0; S.misc.next = 0; }
      __syncthreads();
      // ---- A: the jobs of the pass, their integer vectors
      {
        int sx = 0, sy = 0;
        for (int i = tid; i < cnt; i += FW_THREADS)
        {
          const uint32_t j = sorted[first + c0 + i];
          const hmgpu_me_job jb = jobs[j];
          const hmgpu_me_result res = results[j];
          FwJob fj;
          fj.mvx = res.int_x; fj.mvy = res.int_y;
          fj.xw = 0; fj.yw = 0;
          fj.rx = (uint8_t)(jb.pu_x - ctu_x); fj.ry = (uint8_t)(jb.pu_y - ctu_y);
          fj.w = jb.pu_w; fj.h = jb.pu_h;
          const bool t8 = job_tile_size(jb) == 8;
          const int tw = t8 ? jb.pu_w >> 3 : jb.pu_w >> 2;
          fj.inv_tw = (uint16_t)((32768 + tw - 1) / tw);
          fj.tiles = (uint8_t)(tw * (t8 ? jb.pu_h >> 3 : jb.pu_h >> 2));
          fj.flags = (uint8_t)((t8 ? FWF_T8 : 0) | (((jb.flags & HMGPU_F_HADME) && !(jb.flags & HMGPU_F_LOSSLESS)) ? FWF_SATD : 0));
          S.job[i] = fj;
          S.jidx[i] = j;
          sx += res.int_x; sy += res.int_y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
        if (lane == 0 && (sx | sy)) { atomicAdd(&S.misc.sumx, sx); atomicAdd(&S.misc.sumy, sy); }
      }
      __syncthreads();
      // ---- B: window around the mean vector, copies of the four even-even planes in flight; which jobs it serves; tile offsets
      const int cx = (int)floorf((float)S.misc.sumx / (float)cnt + 0.5f), cy = (int)floorf((float)S.misc.sumy / (float)cnt + 0.5f);
      const int wx0 = (ctu_x + cx - FW_XOFF) & ~15, wy0 = ctu_y + cy - FW_YOFF;
      if (tid == 0)
      {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(4 * FW_PLANE_BYTES)) : "memory");
        for (int b = 0; b < 4; b++)
          fw_tma_box(fw_smem_u32(S.planes[b]), &tmap, wx0 + HMGPU_MARGIN, wy0 + HMGPU_MARGIN, (b >> 1) * 8 + (b & 1) * 2, slot, bar);
      }
      {
        // two consecutive jobs per thread, block-wide exclusive scan of (8x8 tiles | 4x4 tiles << 16)
        uint32_t v2[2] = { 0, 0 };
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
          const int i = 2 * tid + k;
          if (i < cnt)
          {
            FwJob fj = S.job[i];
            const int px = ctu_x + fj.rx + fj.mvx, py = ctu_y + fj.ry + fj.mvy;   // PU displaced by the integer vector, picture coordinates
            fj.xw = (int16_t)(px - wx0); fj.yw = (int16_t)(py - wy0);
            // candidates reach one sample left / above the integer position and none right / below (TEncSearch.cpp:4386-4422)
            const bool inside = fj.xw >= 1 && fj.yw >= 1 && fj.xw + fj.w <= FW_WW && fj.yw + fj.h <= FW_WH
                             && fj.rx + fj.w <= 64 && fj.ry + fj.h <= 64;
            if (inside && (fj.flags & FWF_SATD)) fj.flags |= FWF_FAST;
            S.job[i] = fj;
            if (fj.flags & FWF_FAST) v2[k] = (fj.flags & FWF_T8) ? (uint32_t)fj.tiles : (uint32_t)fj.tiles << 16;
          }
        }
        uint32_t incl = v2[0] + v2[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) S.misc.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; w++) before += S.misc.warp_tot[w];
        const uint32_t ex0 = before + incl - v2[0] - v2[1];
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
          const int i = 2 * tid + k;
          if (i < cnt)
          {
            const uint32_t ex = k ? ex0 + v2[0] : ex0, in = ex + v2[k];
            S.off[i] = ex;
            if ((in & 0xffffu) > FW_T8MAX || (in >> 16) > FW_T4MAX) atomicMin(&S.misc.m, i);
          }
        }
      }
      __syncthreads();
      const int m = S.misc.m;                              // jobs of this pass (>= 1: one job has at most 64 tiles)
      // ---- C: tile tables, slow list, cleared accumulators
      for (int i = tid; i < m; i += FW_THREADS)
      {
        const FwJob fj = S.job[i];
        if (fj.flags & FWF_FAST)
        {
          const uint32_t ex = S.off[i];
          uint16_t* tab = (fj.flags & FWF_T8) ? S.t8 + (ex & 0xffffu) : S.t4 + (ex >> 16);
          for (int t = 0; t < fj.tiles; t++) tab[t] = (uint16_t)((i << 6) | t);
        }
        else S.slow[atomicAdd(&S.misc.n_slow, 1)] = (uint16_t)i;
#pragma unroll
        for (int c = 0; c < 9; c++) S.acc[c][i] = 0;
      }
      __syncthreads();
      int n8, n4;
      {
        // totals of the pass: exclusive offset + own count of its last fast job = offsets a job at index m would get
        uint32_t end = 0;
        const FwJob fl = S.job[m - 1];
        end = S.off[m - 1] + ((fl.flags & FWF_FAST) ? ((fl.flags & FWF_T8) ? (uint32_t)fl.tiles : (uint32_t)fl.tiles << 16) : 0u);
        n8 = (int)(end & 0xffffu); n4 = (int)(end >> 16);
      }
      // ---- D: half-pel candidates
      fw_mbar_wait(bar, parity); parity ^= 1;
      fw_run_phase<0>(S, n8, n4);
      fw_run_slow<0>(S, jobs, refs, org);
      __syncthreads();
      // the windows of the twelve other planes replace them (every thread has left the half-pel pass)
      if (tid == 0)
      {
        S.misc.next = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(12 * FW_PLANE_BYTES)) : "memory");
        for (int p = 0; p < 16; p++)
          if ((p & 5) != 0) fw_tma_box(fw_smem_u32(S.planes[fw_slot(p)]), &tmap, wx0 + HMGPU_MARGIN, wy0 + HMGPU_MARGIN, p, slot, bar);
      }
      // ---- E: half-pel winner per job (xPatternRefinement: first strict minimum in table order)
      for (int i = tid; i < m; i += FW_THREADS)
      {
        const hmgpu_me_job jb = jobs[S.jidx[i]];
        const FwJob fj = S.job[i];
        uint32_t best = 0xffffffffu; int bi = 0;
#pragma unroll
        for (int c = 0; c < 9; c++)
        {
          const uint32_t cost = (S.acc[c][i] >> (bit_depth - 8))
                              + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 1, 2 * fj.mvx + c_refine_h[c][0], 2 * fj.mvy + c_refine_h[c][1]);
          if (cost < best) { best = cost; bi = c; }
        }
        S.hsel[i] = (uint8_t)((c_refine_h[bi][0] + 1) | ((c_refine_h[bi][1] + 1) << 2));
        // candidate 0 of the quarter-pel table is the half-pel winner itself: same block of the same plane
        const uint32_t dsel = S.acc[bi][i];
#pragma unroll
        for (int c = 1; c < 9; c++) S.acc[c][i] = 0;
        S.acc[0][i] = dsel;
      }
      __syncthreads();
      // ---- F: quarter-pel candidates
      fw_mbar_wait(bar, parity); parity ^= 1;
      fw_run_phase<1>(S, n8, n4);
      fw_run_slow<1>(S, jobs, refs, org);
      __syncthreads();
      // ---- G: quarter-pel winner, result
      for (int i = tid; i < m; i += FW_THREADS)
      {
        const uint32_t j = S.jidx[i];
        const hmgpu_me_job jb = jobs[j];
        const FwJob fj = S.job[i];
        const int hx = (int)(S.hsel[i] & 3) - 1, hy = (int)(S.hsel[i] >> 2) - 1;
        uint32_t best = 0xffffffffu; int bi = 0;
#pragma unroll
        for (int c = 0; c < 9; c++)
        {
          const uint32_t cost = (S.acc[c][i] >> (bit_depth - 8))
                              + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 0, 4 * fj.mvx + 2 * hx + c_refine_q[c][0], 4 * fj.mvy + 2 * hy + c_refine_q[c][1]);
          if (cost < best) { best = cost; bi = c; }
        }
        hmgpu_me_result res = results[j];
        res.half_x = (int16_t)hx; res.half_y = (int16_t)hy;
        res.qter_x = c_refine_q[bi][0]; res.qter_y = c_refine_q[bi][1];
        res.frac_cost = best;
        res.n_cand += 18;
        results[j] = res;
      }
      c0 += m;
      __syncthreads();                                     // the tables and the windows are free for the next pass
    }
  }
}

// ---- host -------------------------------------------------------------------------------------------------------------------
typedef CUresult (*FwEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the tensor map of a context, built once: the planes keep their address for the life of the context (ref_alloc, api.cu)
static int fw_build_map(hmgpu_ctx* ctx)
{
  static FwEncodeTiled s_encode = NULL;
  if (!s_encode)
  {
    void* fn = NULL;
    cudaDriverEntryPointQueryResult q;
    HMGPU_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return hmgpu_fail(ctx, HMGPU_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    s_encode = (FwEncodeTiled)fn;
  }
  void* raw = malloc(sizeof(CUtensorMap) + 64);
  if (!raw) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "out of host memory");
  CUtensorMap* map = (CUtensorMap*)(((uintptr_t)raw + 63) & ~(uintptr_t)63);
  const cuuint64_t dim[4] = { (cuuint64_t)ctx->pitch, (cuuint64_t)ctx->ph, 16, (cuuint64_t)ctx->max_refs };
  const cuuint64_t str[3] = { (cuuint64_t)ctx->pitch, (cuuint64_t)ctx->plane_elems, (cuuint64_t)ctx->slot_bytes };
  const cuuint32_t box[4] = { FW_WW, FW_WH, 1, 1 };
  const cuuint32_t est[4] = { 1, 1, 1, 1 };
  const CUresult r = s_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, ctx->planes_all, dim, str, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { free(raw); return hmgpu_fail(ctx, HMGPU_E_CUDA, "cuTensorMapEncodeTiled(box %d x %d) failed: %d", FW_WW, FW_WH, (int)r); }
  ctx->h_fw_tmap = raw;
  return HMGPU_OK;
}

int hmgpu_launch_frac_window(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, hmgpu_me_result* d_results, bool any_frac)
{
  const int ctus_x = (ctx->pic_w + 63) >> 6, n_ctus = ctus_x * ((ctx->pic_h + 63) >> 6);
  const int n_bins = n_ctus * ctx->max_refs;
  // scratch: bin_count[n_bins] | bin_fill[n_bins] | totals[2] counter[1] (cleared together) | bin_start[n_bins] | group_list[n_bins] | sorted[n_jobs]
  const size_t clr_bytes = ((size_t)(2 * n_bins + 4) * sizeof(uint32_t) + 255) & ~(size_t)255;
  const size_t bins_al = ((size_t)n_bins * sizeof(uint32_t) + 255) & ~(size_t)255;
  int rc = hmgpu_reserve_work(ctx, clr_bytes + 2 * bins_al + (size_t)n_jobs * sizeof(uint32_t));
  if (rc) return rc;
  uint32_t* bin_count = (uint32_t*)ctx->d_work;
  uint32_t* bin_fill = bin_count + n_bins;
  uint32_t* totals = bin_fill + n_bins;
  uint32_t* counter = totals + 2;
  uint32_t* bin_start = (uint32_t*)((char*)ctx->d_work + clr_bytes);
  uint32_t* group_list = (uint32_t*)((char*)ctx->d_work + clr_bytes + bins_al);
  uint32_t* sorted = (uint32_t*)((char*)ctx->d_work + clr_bytes + 2 * bins_al);
  if (!ctx->h_fw_tmap) { rc = fw_build_map(ctx); if (rc) return rc; }
  const CUtensorMap* map = (const CUtensorMap*)(((uintptr_t)ctx->h_fw_tmap + 63) & ~(uintptr_t)63);
  const size_t smem = sizeof(FwSmem);
  if (!(ctx->attr_done & HMGPU_ATTR_FRACW))
  {
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(fracw_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attr_done |= HMGPU_ATTR_FRACW;
  }
  HMGPU_CUDA(ctx, cudaMemsetAsync(bin_count, 0, clr_bytes, ctx->stream));
  const RefTable rt = hmgpu_ref_table(ctx);
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int tb = 256, nb = (n_jobs + tb - 1) / tb;
  {
    HmgpuStage st(ctx, HMGPU_ST_FRAC_EXPAND, any_frac ? 3 : 1);
    fracw_bin_kernel<<<nb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, d_results, bin_count, ctus_x, n_ctus);
    if (any_frac)
    {
      fracw_scan_kernel<<<1, 1024, 0, ctx->stream>>>(bin_count, n_bins, bin_start, group_list, totals);
      fracw_scatter_kernel<<<nb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, bin_start, bin_fill, sorted, ctus_x, n_ctus);
    }
  }
  if (any_frac)
  {
    HmgpuStage st(ctx, HMGPU_ST_FRAC_DIST, 1);
    const int grid = n_jobs < HMGPU_NUM_SMS ? n_jobs : HMGPU_NUM_SMS;
    fracw_group_kernel<<<grid, FW_THREADS, smem, ctx->stream>>>(d_jobs, d_results, rt, ov, sorted, bin_start, bin_count, group_list, totals,
                                                                counter, *map, ctus_x, n_ctus, ctx->bit_depth);
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
