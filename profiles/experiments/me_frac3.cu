// me_frac3.cu -- fractional refinement of 8-bit pictures with the source picture as key pattern: ONE fused kernel per batch,
// candidate blocks staged in shared memory by TMA.
//
// Replaces TEncSearch::xPatternSearchFracDIF + xPatternRefinement (TEncSearch.cpp:4386-4422, 799-852) over the 16 quarter-pel
// phase planes (planes.cu), like me_frac2.cu, whose ncu capture (profiles/r1n_ncu_frac2_dist.csv) showed where its time goes:
// rows staged four bytes per cp.async with per-thread address arithmetic (10.8 long-scoreboard stalls per issued instruction),
// six launches with the per-candidate sums and the half-pel winner travelling through global memory, and 1.9 x the algorithmic
// DRAM traffic because every PU shape sweeps the picture on its own.  Here:
//
//  * JOB ORDER.  A counting sort puts the jobs in (64-row band of the picture, PU shape) order.  Warps take GROUPS of
//    consecutive jobs of one shape from that list through an atomic counter, so the groups in flight at any moment touch a few
//    bands of the phase planes: the planes are read from DRAM about once per batch instead of once per shape.
//  * MAPPING.  A warp owns a group: lane = (job slot, SATD tile), as many jobs per group as fill the 32 lanes (16 jobs of
//    8x4, 8 of 16x16, one 64x32 ...; a 64x64 job takes two passes).  A lane keeps its source tile in registers for all 17
//    candidates; the horizontal Hadamard pass works on packed bytes with IDP.4A (linearity: H(org) - H(ref)), as in me_frac2.cu.
//  * TMA.  The reference stores what one search needs as (W+1) x (H+1) blocks of m_filteredBlock[4][4]; so does this kernel:
//    the nine half-pel candidates are four boxes (planes (0,0) (0,2) (2,0) (2,2), all at integer offset (-1,-1) of the best
//    integer position, each serving 1 / 2 / 2 / 4 candidates), the eight quarter-pel candidates around the half-pel winner are
//    one box each.  The warp issues one cp.async.bulk.tensor per job and box (a 4-D tensor map per box shape: x, y,
//    phase plane, reference slot) into a ring of shared-memory buffers guarded by mbarriers, one or two boxes ahead of the
//    arithmetic.  (A box must start at a 16-byte aligned byte of the row, so it is up to 15 bytes wider than the block.)  Shared memory is written with the 128-byte TMA swizzle: the boxes of the job slots of a warp
//    sit at 128-byte multiples, and without it the lanes of different slots would read the same banks (16-way conflicts).
//  * FUSION.  Per-candidate sums are shared-memory atomics of the warp, the half-pel winner is chosen by the slot's first lane
//    and read by the others, the result goes to global memory once.  Launches per batch: count, scan, scatter, search.
#include "me_frac_impl.cuh"
#include <cuda.h>

#define F3_WARPS 4
#define F3_BANDS_MAX 128                         // 64-row bands of a picture of up to 8184 rows
#define F3_SHAPES 256                            // (w/4 - 1) * 16 + (h/4 - 1)
#define F3_MAX_SLOTS 32

// ---- geometry of a PU shape (host and device agree on it) ---------------------------------------------------------------
struct F3Shape
{
  int w, h, ts, tw, tiles;      // SATD tile size (xGetHADs, TComRdCost.cpp:1555-1597), tiles per row, tiles per job
  int ib, br;                   // TMA box: bytes per row (multiple of 16, >= w + 16), rows (h + 1)
  int stride;                   // bytes between the boxes of consecutive job slots (multiple of 128: TMA destination alignment)
  int lanes_per_job, jobs;      // lanes a job occupies (min(tiles, 32)), jobs per group
  int map;                      // index of the box shape's tensor map: inner class * 16 + (h/4 - 1)
};

__host__ __device__ static inline F3Shape f3_shape(int w, int h, int buf_bytes)
{
  F3Shape s;
  s.w = w; s.h = h;
  s.ts = ((w & 7) == 0 && (h & 7) == 0) ? 8 : 4;
  s.tw = w / s.ts; s.tiles = s.tw * (h / s.ts);
  // A TMA box starts at a 16-byte aligned byte of the row (tile mode: coordinate 0 x element size must be a multiple of 16;
  // anything else raises "illegal instruction", profiles/probes/tma_probe.cu), so a box carries up to 15 bytes in front of
  // the W + 1 it is fetched for.  Box widths with a tensor map: 32, 48, 64, 80.
  s.ib = (w + 16 + 15) & ~15; s.br = h + 1;
  s.stride = (s.ib * s.br + 127) & ~127;
  s.lanes_per_job = s.tiles < 32 ? s.tiles : 32;
  int j = 32 / s.lanes_per_job, cap = buf_bytes / s.stride;
  if (cap < 1) cap = 1;
  s.jobs = j < cap ? j : cap;
  s.map = (s.ib / 16 - 2) * 16 + (h / 4 - 1);
  return s;
}

// ---- work list: counting sort by (band, shape) ---------------------------------------------------------------------------
__device__ __forceinline__ int f3_bin(const hmgpu_me_job& jb) { return (jb.pu_y >> 6) * F3_SHAPES + ((jb.pu_w >> 2) - 1) * 16 + ((jb.pu_h >> 2) - 1); }

__global__ void frac3_count_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, hmgpu_me_result* __restrict__ results, uint32_t* __restrict__ bin_count)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const hmgpu_me_job jb = jobs[j];
  if (!(jb.flags & HMGPU_F_INTEGER))
  {
    hmgpu_me_result r;
    r.int_x = jb.start_x; r.int_y = jb.start_y; r.int_sad = 0;
    r.half_x = r.half_y = r.qter_x = r.qter_y = 0; r.frac_cost = 0; r.n_cand = 0;
    results[j] = r;
  }
  if (jb.flags & HMGPU_F_FRAC) atomicAdd(&bin_count[f3_bin(jb)], 1u);
}

// one CTA: exclusive scans of the job counts and of the group counts of the bins; totals[0] = jobs, totals[1] = groups
__global__ void __launch_bounds__(1024)
frac3_scan_kernel(const uint32_t* __restrict__ bin_count, int n_bins, int buf_bytes, uint32_t* __restrict__ bin_start,
                  uint32_t* __restrict__ grp_start, uint32_t* __restrict__ totals)
{
  __shared__ uint32_t s_j[1024], s_g[1024];
  const int t = threadIdx.x;
  const int per = (n_bins + 1023) / 1024;
  const int b0 = t * per, b1 = min(n_bins, b0 + per);
  uint32_t sj = 0, sg = 0;
  for (int b = b0; b < b1; b++)
  {
    const uint32_t c = bin_count[b];
    if (c)
    {
      const int sh = b % F3_SHAPES;
      const F3Shape s = f3_shape(((sh >> 4) + 1) * 4, ((sh & 15) + 1) * 4, buf_bytes);
      sj += c; sg += (c + s.jobs - 1) / s.jobs;
    }
  }
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
    bin_start[b] = oj; grp_start[b] = og;
    const uint32_t c = bin_count[b];
    if (c)
    {
      const int sh = b % F3_SHAPES;
      const F3Shape s = f3_shape(((sh >> 4) + 1) * 4, ((sh & 15) + 1) * 4, buf_bytes);
      oj += c; og += (c + s.jobs - 1) / s.jobs;
    }
  }
  if (t == 1023) { bin_start[n_bins] = s_j[t]; grp_start[n_bins] = s_g[t]; totals[0] = s_j[t]; totals[1] = s_g[t]; totals[2] = 0; }
}

__global__ void frac3_scatter_kernel(const hmgpu_me_job* __restrict__ jobs, int n_jobs, const uint32_t* __restrict__ bin_start,
                                     uint32_t* __restrict__ bin_cursor, uint32_t* __restrict__ sorted)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const hmgpu_me_job jb = jobs[j];
  if (!(jb.flags & HMGPU_F_FRAC)) return;
  const int b = f3_bin(jb);
  sorted[bin_start[b] + atomicAdd(&bin_cursor[b], 1u)] = (uint32_t)j;
}

// ---- PTX: mbarrier + TMA ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f3_mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void f3_mbar_expect(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f3_mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "F3_WAIT:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra F3_DONE;\n\t"
    "bra F3_WAIT;\n\t"
    "F3_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// one box of the 4-D tensor (x bytes, y rows, z phase plane, reference slot) into shared memory, completion counted on the mbarrier
__device__ __forceinline__ void f3_tma_box(uint32_t dst, const void* tmap, int x, int y, int z, int slot, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(slot), "r"(bar) : "memory");
}

// the tensor maps of a context: 4 box widths (32, 48, 64, 80 bytes) x 16 box heights (5, 9, ... 65 rows) over the phase planes of
// all reference slots.  Passed to the kernel BY VALUE as a __grid_constant__ parameter (8 KB): TMA fetches descriptors from the
// parameter / constant bank; a descriptor array in global memory raised "illegal instruction" on this driver.
struct F3Maps { CUtensorMap m[64]; };
// CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk index (address bits 4..6) is XORed with address bits 7..9
template <bool SWZ> __device__ __forceinline__ uint32_t f3_swz(uint32_t off) { return SWZ ? off ^ ((off >> 3) & 0x70u) : off; }

// ---- Hadamard helpers (packed bytes, IDP.4A): as in me_frac2.cu -----------------------------------------------------------
__host__ __device__ constexpr uint32_t f3_pat4(int k, bool neg)
{
  uint32_t w = 0;
  for (int j = 0; j < 4; j++)
  {
    int s = 0;
    for (int b = 0; b < 3; b++) s ^= ((k >> b) & (j >> b) & 1);
    const bool minus = (s != 0) != neg;
    w |= (minus ? 0xffu : 0x01u) << (8 * j);
  }
  return w;
}
template <int K, bool NEG> struct F3Pat
{
  static constexpr uint32_t lo = f3_pat4(K, NEG);
  static constexpr uint32_t hi = f3_pat4(K, NEG != ((K & 4) != 0));
};
template <bool NEG>
__device__ __forceinline__ void f3_row8(uint32_t w0, uint32_t w1, const int* init, int* out)
{
  out[0] = hm_dp4a_us(w1, F3Pat<0, NEG>::hi, hm_dp4a_us(w0, F3Pat<0, NEG>::lo, init[0]));
  out[1] = hm_dp4a_us(w1, F3Pat<1, NEG>::hi, hm_dp4a_us(w0, F3Pat<1, NEG>::lo, init[1]));
  out[2] = hm_dp4a_us(w1, F3Pat<2, NEG>::hi, hm_dp4a_us(w0, F3Pat<2, NEG>::lo, init[2]));
  out[3] = hm_dp4a_us(w1, F3Pat<3, NEG>::hi, hm_dp4a_us(w0, F3Pat<3, NEG>::lo, init[3]));
  out[4] = hm_dp4a_us(w1, F3Pat<4, NEG>::hi, hm_dp4a_us(w0, F3Pat<4, NEG>::lo, init[4]));
  out[5] = hm_dp4a_us(w1, F3Pat<5, NEG>::hi, hm_dp4a_us(w0, F3Pat<5, NEG>::lo, init[5]));
  out[6] = hm_dp4a_us(w1, F3Pat<6, NEG>::hi, hm_dp4a_us(w0, F3Pat<6, NEG>::lo, init[6]));
  out[7] = hm_dp4a_us(w1, F3Pat<7, NEG>::hi, hm_dp4a_us(w0, F3Pat<7, NEG>::lo, init[7]));
}
template <bool NEG>
__device__ __forceinline__ void f3_row4(uint32_t w0, const int* init, int* out)
{
  out[0] = hm_dp4a_us(w0, F3Pat<0, NEG>::lo, init[0]);
  out[1] = hm_dp4a_us(w0, F3Pat<1, NEG>::lo, init[1]);
  out[2] = hm_dp4a_us(w0, F3Pat<2, NEG>::lo, init[2]);
  out[3] = hm_dp4a_us(w0, F3Pat<3, NEG>::lo, init[3]);
}
__device__ __forceinline__ uint32_t f3_cols8_abs(int* d)
{
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++)
  {
    int* v = d + c;
#pragma unroll
    for (int i = 0; i < 4; i++) { const int a = v[i * 8], b = v[(i + 4) * 8]; v[i * 8] = a + b; v[(i + 4) * 8] = a - b; }
#pragma unroll
    for (int i = 0; i < 8; i += 4)
#pragma unroll
      for (int j = i; j < i + 2; j++) { const int a = v[j * 8], b = v[(j + 2) * 8]; v[j * 8] = a + b; v[(j + 2) * 8] = a - b; }
#pragma unroll
    for (int i = 0; i < 8; i += 2) s += 2u * (uint32_t)max(hm_abs(v[i * 8]), hm_abs(v[(i + 1) * 8]));
  }
  return (s + 2) >> 2;
}
__device__ __forceinline__ uint32_t f3_cols4_abs(int* d)
{
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
  {
    int* v = d + c;
    const int a0 = v[0] + v[8], a1 = v[4] + v[12], a2 = v[0] - v[8], a3 = v[4] - v[12];
    s += 2u * (uint32_t)max(hm_abs(a0), hm_abs(a1)) + 2u * (uint32_t)max(hm_abs(a2), hm_abs(a3));
  }
  return (s + 1) >> 1;
}

// distortion of one TS x TS tile against the block at (row0, col0 bytes) of a staged box (row pitch ib) whose first byte is at
// offset `box` of the warp's buffer `wb` (1024-byte aligned): SATD (xCalcHADs8x8 / 4x4) or SAD
template <int TS, bool SWZ>
__device__ __forceinline__ uint32_t f3_tile(const unsigned char* wb, uint32_t box, int ib, int row0, int col0, const uint32_t (&ow)[TS][TS / 4], bool satd)
{
  const int sh = (col0 & 3) * 8;
  const uint32_t a0 = box + (uint32_t)(row0 * ib + (col0 & ~3));
  if (satd)
  {
    int d[TS * TS];
    int zero[TS];
#pragma unroll
    for (int k = 0; k < TS; k++) zero[k] = 0;
#pragma unroll
    for (int r = 0; r < TS; r++)
    {
      const uint32_t a = a0 + (uint32_t)(r * ib);
      const uint32_t w0 = *(const uint32_t*)(wb + f3_swz<SWZ>(a)), w1 = *(const uint32_t*)(wb + f3_swz<SWZ>(a + 4));
      int h[TS];
      if (TS == 8)
      {
        const uint32_t w2 = *(const uint32_t*)(wb + f3_swz<SWZ>(a + 8));
        f3_row8<false>(ow[r][0], ow[r][TS / 4 - 1], zero, h);
        f3_row8<true>(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), h, d + r * TS);
      }
      else
      {
        f3_row4<false>(ow[r][0], zero, h);
        f3_row4<true>(__funnelshift_r(w0, w1, sh), h, d + r * TS);
      }
    }
    return TS == 8 ? f3_cols8_abs(d) : f3_cols4_abs(d);
  }
  uint32_t v = 0;
#pragma unroll
  for (int r = 0; r < TS; r++)
  {
    const uint32_t a = a0 + (uint32_t)(r * ib);
    const uint32_t w0 = *(const uint32_t*)(wb + f3_swz<SWZ>(a)), w1 = *(const uint32_t*)(wb + f3_swz<SWZ>(a + 4));
    v = vabsdiff4_acc(__funnelshift_r(w0, w1, sh), ow[r][0], v);
    if (TS == 8)
    {
      const uint32_t w2 = *(const uint32_t*)(wb + f3_swz<SWZ>(a + 8));
      v = vabsdiff4_acc(__funnelshift_r(w1, w2, sh), ow[r][TS / 4 - 1], v);
    }
  }
  return v;
}

// candidates of the four half-pel boxes: table index c of s_acMvRefineH and the in-box offset (bit 0: x, bit 1: y) of its block
//   box 0 = plane (0,0): c0;  box 1 = plane (x 2, y 0): c3, c4;  box 2 = plane (x 0, y 2): c1, c2;  box 3 = plane (2,2): c5..c8
__constant__ uint8_t c_f3_half_n[4] = { 1, 2, 2, 4 };
__constant__ uint8_t c_f3_half_c[4][4] = { { 0, 0, 0, 0 }, { 3, 4, 0, 0 }, { 1, 2, 0, 0 }, { 5, 6, 7, 8 } };
__constant__ uint8_t c_f3_half_o[4][4] = { { 3, 0, 0, 0 }, { 2, 3, 0, 0 }, { 1, 3, 0, 0 }, { 0, 1, 2, 3 } };

struct F3Warp                                    // per-warp bookkeeping in shared memory
{
  uint32_t acc[F3_MAX_SLOTS][9];
  uint32_t best[F3_MAX_SLOTS];                   // half-pel winner of the slot: table index
  unsigned long long bar[4];
};

template <int NBUF, int BUF_BYTES, bool SWZ>
__global__ void __launch_bounds__(F3_WARPS * 32)
frac3_kernel(const hmgpu_me_job* __restrict__ jobs, hmgpu_me_result* __restrict__ results, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ bin_start, const uint32_t* __restrict__ grp_start, int n_bins, uint32_t* __restrict__ totals,
             const __grid_constant__ F3Maps maps, OrgView org, int bit_depth)
{
  extern __shared__ unsigned char f3_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = (unsigned char*)(((uintptr_t)f3_dyn + 1023) & ~(uintptr_t)1023);
  unsigned char* wb = base + (size_t)warp * NBUF * BUF_BYTES;          // this warp's ring of box buffers
  F3Warp* ws = (F3Warp*)(base + (size_t)F3_WARPS * NBUF * BUF_BYTES) + warp;
  const uint32_t wb_s = f3_smem(wb);
  if (lane == 0)
  {
    for (int b = 0; b < NBUF; b++) f3_mbar_init(f3_smem(&ws->bar[b]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;                                                  // bit b: phase parity buffer b is waited on next
  const uint32_t n_groups = totals[1];
  const int org_pitch_w = org.pitch >> 2;

  for (;;)
  {
    uint32_t gid = 0;
    if (lane == 0) gid = atomicAdd(&totals[2], 1u);
    gid = __shfl_sync(0xffffffffu, gid, 0);
    if (gid >= n_groups) break;
    // the bin of this group: last bin whose first group is <= gid
    int lo = 0, hi = n_bins;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (grp_start[mid] <= gid) lo = mid; else hi = mid; }
    const int bin = lo, shp = bin % F3_SHAPES;
    const F3Shape S = f3_shape(((shp >> 4) + 1) * 4, ((shp & 15) + 1) * 4, BUF_BYTES);
    const uint32_t first = bin_start[bin] + (gid - grp_start[bin]) * S.jobs;
    const int n_in = min((int)(bin_start[bin + 1] - first), S.jobs);
    const int slot = lane / S.lanes_per_job, tile0 = lane - slot * S.lanes_per_job;
    const bool active = slot < n_in;
    const bool leader = active && tile0 == 0;
    const uint32_t ji = sorted[first + (active ? slot : 0)];
    const hmgpu_me_job jb = jobs[ji];
    const hmgpu_me_result res = results[ji];
    const bool satd = (jb.flags & HMGPU_F_HADME) && !(jb.flags & HMGPU_F_LOSSLESS);
    const uint32_t box_bytes = (uint32_t)(S.ib * S.br);
    const uint32_t slot_off = (uint32_t)((active ? slot : 0) * S.stride);   // idle lanes shadow slot 0 (reads stay inside the buffer)
    const int passes = (S.tiles + 31) >> 5;
    // box s of this job into buffer s % NBUF: s = 0..3 half-pel boxes, 4..11 quarter-pel boxes around half-pel winner (hx, hy)
    int hx = 0, hy = 0;
    // A TMA copy is ONE warp-level instruction with uniform operands (UTMALDG reads uniform registers), so the boxes of the
    // job slots are issued one after the other: every lane packs the box of its own job, the loop fetches slot k's word from
    // the slot's first lane and lane 0 issues the copy.
    auto issue = [&](int s) {
      const int b = s % NBUF;
      const uint32_t bar = f3_smem(&ws->bar[b]);
      if (lane == 0) f3_mbar_expect(bar, box_bytes * (uint32_t)n_in);
      int x, y, z;
      if (s < 4)
      {
        x = jb.pu_x + res.int_x - 1; y = jb.pu_y + res.int_y - 1;
        z = (s >> 1) * 8 + (s & 1) * 2;
      }
      else
      {
        const int qx = 4 * res.int_x + 2 * hx + c_refine_q[s - 3][0], qy = 4 * res.int_y + 2 * hy + c_refine_q[s - 3][1];
        x = jb.pu_x + (qx >> 2); y = jb.pu_y + (qy >> 2);
        z = (qy & 3) * 4 + (qx & 3);
      }
      // x, y < 8184 + 160: 14 bits each; plane 4 bits (the reference slot travels in a second word); x rounded down to 16 bytes
      const uint32_t packed = (uint32_t)((x + HMGPU_MARGIN) & ~15) | ((uint32_t)(y + HMGPU_MARGIN) << 14) | ((uint32_t)z << 28);
      const uint32_t dst0 = wb_s + (uint32_t)(b * BUF_BYTES);
      __syncwarp();
      for (int k = 0; k < n_in; k++)
      {
        const uint32_t pk = __shfl_sync(0xffffffffu, packed, k * S.lanes_per_job);
        const int rs = __shfl_sync(0xffffffffu, (int)jb.ref_slot, k * S.lanes_per_job);
        if (lane == 0)
          f3_tma_box(dst0 + (uint32_t)(k * S.stride), &maps.m[S.map], (int)(pk & 0x3fffu), (int)((pk >> 14) & 0x3fffu), (int)(pk >> 28), rs, bar);
      }
    };
    auto wait = [&](int s) {
      const int b = s % NBUF;
      f3_mbar_wait(f3_smem(&ws->bar[b]), (parity >> b) & 1u);
      parity ^= 1u << b;
    };
    for (int s = 0; s < NBUF && s < 4; s++) issue(s);
    for (int i = lane; i < F3_MAX_SLOTS * 9; i += 32) (&ws->acc[0][0])[i] = 0;
    __syncwarp();

    for (int pass = 0; pass < passes; pass++)
    {
      // (a 64x64 job runs the whole box sequence once per 32 tiles; its boxes are re-fetched: 1 job in 600)
      const int t = tile0 + pass * 32;
      const int ty = (t / S.tw) * S.ts, tx = (t - (t / S.tw) * S.tw) * S.ts;
      if (pass > 0)
      {
        __syncwarp();
        for (int s = 0; s < NBUF && s < 4; s++) issue(s);
      }
      // source tile: aligned words straight from the source picture (pu_x and the tile offsets are multiples of 4)
      uint32_t ow8[8][2], ow4[4][1];
      {
        const uint32_t* o = (const uint32_t*)((const uint8_t*)org.base + (size_t)(jb.pu_y + ty) * org.pitch + jb.pu_x + tx);
        if (S.ts == 8)
        {
#pragma unroll
          for (int r = 0; r < 8; r++) { ow8[r][0] = __ldg(o + (size_t)r * org_pitch_w); ow8[r][1] = __ldg(o + (size_t)r * org_pitch_w + 1); }
        }
        else
        {
#pragma unroll
          for (int r = 0; r < 4; r++) ow4[r][0] = __ldg(o + (size_t)r * org_pitch_w);
        }
      }
      // ---- half-pel phase: four boxes ----
      const int xo_h = (jb.pu_x + res.int_x - 1 + HMGPU_MARGIN) & 15;      // bytes of the box in front of the block's column -1
      for (int s = 0; s < 4; s++)
      {
        wait(s);
        const uint32_t box = (uint32_t)((s % NBUF) * BUF_BYTES) + slot_off;
        const int nc = c_f3_half_n[s];
        for (int k = 0; k < nc; k++)
        {
          const int c = c_f3_half_c[s][k], o = c_f3_half_o[s][k];
          const uint32_t v = S.ts == 8 ? f3_tile<8, SWZ>(wb, box, S.ib, ty + (o >> 1), tx + (o & 1) + xo_h, ow8, satd)
                                       : f3_tile<4, SWZ>(wb, box, S.ib, ty + (o >> 1), tx + (o & 1) + xo_h, ow4, satd);
          if (active) atomicAdd(&ws->acc[slot][c], v);
        }
        __syncwarp();
        if (s + NBUF < 4) issue(s + NBUF);
      }
      if (pass + 1 < passes) continue;                          // a job larger than the warp: next 32 tiles first
    }
    // ---- half-pel selection by the slot's first lane: first strict minimum in table order (TEncSearch.cpp:816-847) ----
    uint32_t best_cost = 0xffffffffu, centre_dist = 0;
    if (leader)
    {
      int bi = 0;
#pragma unroll
      for (int c = 0; c < 9; c++)
      {
        const uint32_t cost = (ws->acc[slot][c] >> (bit_depth - 8))
                            + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 1, 2 * res.int_x + c_refine_h[c][0], 2 * res.int_y + c_refine_h[c][1]);
        if (cost < best_cost) { best_cost = cost; bi = c; }
      }
      ws->best[slot] = (uint32_t)bi;
      centre_dist = ws->acc[slot][bi];
    }
    __syncwarp();
    {
      const int bi = (int)ws->best[active ? slot : 0];
      hx = c_refine_h[bi][0]; hy = c_refine_h[bi][1];
    }
    __syncwarp();
    if (leader) { for (int c = 0; c < 9; c++) ws->acc[slot][c] = 0; }
    // ---- quarter-pel phase: eight boxes, one candidate each (candidate 0 is the half-pel winner itself) ----
    for (int pass = 0; pass < passes; pass++)
    {
      const int t = tile0 + pass * 32;
      const int ty = (t / S.tw) * S.ts, tx = (t - (t / S.tw) * S.tw) * S.ts;
      __syncwarp();
      for (int s = 4; s < 4 + NBUF; s++) issue(s);
      uint32_t ow8[8][2], ow4[4][1];
      {
        const uint32_t* o = (const uint32_t*)((const uint8_t*)org.base + (size_t)(jb.pu_y + ty) * org.pitch + jb.pu_x + tx);
        if (S.ts == 8)
        {
#pragma unroll
          for (int r = 0; r < 8; r++) { ow8[r][0] = __ldg(o + (size_t)r * org_pitch_w); ow8[r][1] = __ldg(o + (size_t)r * org_pitch_w + 1); }
        }
        else
        {
#pragma unroll
          for (int r = 0; r < 4; r++) ow4[r][0] = __ldg(o + (size_t)r * org_pitch_w);
        }
      }
      for (int s = 4; s < 12; s++)
      {
        wait(s);
        const uint32_t box = (uint32_t)((s % NBUF) * BUF_BYTES) + slot_off;
        const int xo = (jb.pu_x + ((4 * res.int_x + 2 * hx + c_refine_q[s - 3][0]) >> 2) + HMGPU_MARGIN) & 15;
        const uint32_t v = S.ts == 8 ? f3_tile<8, SWZ>(wb, box, S.ib, ty, tx + xo, ow8, satd) : f3_tile<4, SWZ>(wb, box, S.ib, ty, tx + xo, ow4, satd);
        if (active) atomicAdd(&ws->acc[slot][s - 3], v);
        __syncwarp();
        if (s + NBUF < 12) issue(s + NBUF);
      }
    }
    if (leader)
    {
      // quarter-pel selection (s_acMvRefineQ order, TEncSearch.cpp:65-75); candidate 0 keeps the half-pel winner's distortion
      uint32_t best = 0xffffffffu;
      int bi = 0;
#pragma unroll
      for (int c = 0; c < 9; c++)
      {
        const uint32_t dist = (c == 0 ? centre_dist : ws->acc[slot][c]) >> (bit_depth - 8);
        const uint32_t cost = dist + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 0, 4 * res.int_x + 2 * hx + c_refine_q[c][0],
                                                4 * res.int_y + 2 * hy + c_refine_q[c][1]);
        if (cost < best) { best = cost; bi = c; }
      }
      hmgpu_me_result r = res;
      r.half_x = (int16_t)hx; r.half_y = (int16_t)hy;
      r.qter_x = c_refine_q[bi][0]; r.qter_y = c_refine_q[bi][1];
      r.frac_cost = best;
      r.n_cand = res.n_cand + 18;
      results[ji] = r;
    }
    __syncwarp();
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*F3EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the 64 tensor maps of a context (4 box widths x 16 box heights) over the phase planes of all its reference slots: x = byte in the
// padded row, y = padded row, z = phase plane, w = reference slot.  Built once: the planes keep their address for the life of the
// context (ref_alloc, api.cu).
static int f3_build_maps(hmgpu_ctx* ctx)
{
  static F3EncodeTiled s_encode = NULL;
  if (!s_encode)
  {
    void* fn = NULL;
    cudaDriverEntryPointQueryResult q;
    HMGPU_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return hmgpu_fail(ctx, HMGPU_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    s_encode = (F3EncodeTiled)fn;
  }
  if (!ctx->h_tmaps) ctx->h_tmaps = malloc(sizeof(F3Maps) + 64);
  if (!ctx->h_tmaps) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "out of host memory");
  F3Maps* maps = (F3Maps*)(((uintptr_t)ctx->h_tmaps + 63) & ~(uintptr_t)63);
  static const int k_ib[4] = { 32, 48, 64, 80 };
  const int swz = ctx->tune.frac3_swizzle;
  for (int ic = 0; ic < 4; ic++)
    for (int hc = 0; hc < 16; hc++)
    {
      const cuuint64_t dim[4] = { (cuuint64_t)ctx->pitch, (cuuint64_t)ctx->ph, 16, (cuuint64_t)ctx->max_refs };
      const cuuint64_t str[3] = { (cuuint64_t)ctx->pitch, (cuuint64_t)ctx->plane_elems, (cuuint64_t)ctx->slot_bytes };
      const cuuint32_t box[4] = { (cuuint32_t)k_ib[ic], (cuuint32_t)((hc + 1) * 4 + 1), 1, 1 };
      const cuuint32_t est[4] = { 1, 1, 1, 1 };
      const CUresult r = s_encode(&maps->m[ic * 16 + hc], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, ctx->planes_all, dim, str, box, est,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return hmgpu_fail(ctx, HMGPU_E_CUDA, "cuTensorMapEncodeTiled(box %d x %d) failed: %d", k_ib[ic], (hc + 1) * 4 + 1, (int)r);
    }
  ctx->tmaps_swz = swz + 1;
  return HMGPU_OK;
}

template <int NBUF, int BUF_BYTES, bool SWZ>
static int f3_launch(hmgpu_ctx* ctx, cudaStream_t stream, int grid, const hmgpu_me_job* d_jobs, hmgpu_me_result* d_results, const uint32_t* sorted,
                     const uint32_t* bin_start, const uint32_t* grp_start, int n_bins, uint32_t* totals, const OrgView& ov)
{
  const int smem = 1024 + F3_WARPS * NBUF * BUF_BYTES + F3_WARPS * (int)sizeof(F3Warp);
  HMGPU_CUDA(ctx, cudaFuncSetAttribute(frac3_kernel<NBUF, BUF_BYTES, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const F3Maps* maps = (const F3Maps*)(((uintptr_t)ctx->h_tmaps + 63) & ~(uintptr_t)63);
  frac3_kernel<NBUF, BUF_BYTES, SWZ><<<grid, F3_WARPS * 32, smem, stream>>>(d_jobs, d_results, sorted, bin_start, grp_start, n_bins, totals,
                                                                           *maps, ov, ctx->bit_depth);
  return HMGPU_OK;
}

int hmgpu_launch_frac_tma(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, hmgpu_me_result* d_results, bool any_frac)
{
  const int variant = ctx->tune.frac3_variant;               // 0: 3 x 6 KB, 1: 2 x 6 KB, 2: 3 x 8 KB, 3: 2 x 8 KB (default), 4: 3 x 4 KB, 5: 2 x 4 KB
  const int buf_bytes = variant < 2 ? 6144 : (variant < 4 ? 8192 : 4096);
  const int swz = ctx->tune.frac3_swizzle;
  if (!ctx->planes_all) return hmgpu_fail(ctx, HMGPU_E_STATE, "no reference picture uploaded");
  if (ctx->tmaps_swz != swz + 1)
  {
    const int rcm = f3_build_maps(ctx);
    if (rcm) return rcm;
  }
  const int n_bands = (ctx->pic_h + 63) >> 6;
  const int n_bins = n_bands * F3_SHAPES;
  // scratch: bin_count | bin_cursor | bin_start[+1] | grp_start[+1] | totals | sorted[n_jobs]
  const size_t bins_al = (((size_t)n_bins + 1) * sizeof(uint32_t) + 255) & ~(size_t)255;
  int rc = hmgpu_reserve_work(ctx, 4 * bins_al + 256 + (size_t)n_jobs * sizeof(uint32_t));
  if (rc) return rc;
  char* w = (char*)ctx->d_work;
  uint32_t* bin_count = (uint32_t*)w;
  uint32_t* bin_cursor = (uint32_t*)(w + bins_al);
  uint32_t* bin_start = (uint32_t*)(w + 2 * bins_al);
  uint32_t* grp_start = (uint32_t*)(w + 3 * bins_al);
  uint32_t* totals = (uint32_t*)(w + 4 * bins_al);
  uint32_t* sorted = (uint32_t*)(w + 4 * bins_al + 256);
  HMGPU_CUDA(ctx, cudaMemsetAsync(bin_count, 0, 2 * bins_al, ctx->stream));
  OrgView ov; ov.base = ctx->d_org; ov.pitch = ctx->org_pitch;
  const int tb = 256;
  {
    HmgpuStage st(ctx, HMGPU_ST_FRAC_EXPAND, any_frac ? 3 : 1);
    frac3_count_kernel<<<(n_jobs + tb - 1) / tb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, d_results, bin_count);
    if (any_frac)
    {
      frac3_scan_kernel<<<1, 1024, 0, ctx->stream>>>(bin_count, n_bins, buf_bytes, bin_start, grp_start, totals);
      frac3_scatter_kernel<<<(n_jobs + tb - 1) / tb, tb, 0, ctx->stream>>>(d_jobs, n_jobs, bin_start, bin_cursor, sorted);
    }
  }
  if (any_frac)
  {
    HmgpuStage st(ctx, HMGPU_ST_FRAC_DIST, 1);
    // persistent grid: the CTAs one SM can hold (shared memory bound), never more than the batch could use
    const int per_sm = (variant == 0 || variant == 2) ? 2 : 3;      // shared memory (80 / 104 KB per CTA) or registers (165 x 128) bound
    long long want = ((long long)n_jobs + F3_WARPS - 1) / F3_WARPS;
    const int grid = (int)(want < (long long)HMGPU_NUM_SMS * per_sm ? (want < 1 ? 1 : want) : (long long)HMGPU_NUM_SMS * per_sm);
#define F3_GO(NB, BB) (swz ? f3_launch<NB, BB, true>(ctx, ctx->stream, grid, d_jobs, d_results, sorted, bin_start, grp_start, n_bins, totals, ov) \
                           : f3_launch<NB, BB, false>(ctx, ctx->stream, grid, d_jobs, d_results, sorted, bin_start, grp_start, n_bins, totals, ov))
    switch (variant)
    {
    case 1: rc = F3_GO(2, 6144); break;
    case 2: rc = F3_GO(3, 8192); break;
    case 3: rc = F3_GO(2, 8192); break;
    case 4: rc = F3_GO(3, 4096); break;
    case 5: rc = F3_GO(2, 4096); break;
    default: rc = F3_GO(3, 6144); break;
    }
#undef F3_GO
    if (rc) return rc;
  }
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}
