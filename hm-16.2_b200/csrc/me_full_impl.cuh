// me_full_impl.cuh -- exhaustive integer search by one CTA per job (device code shared by the
// batch kernels in me_full.cu and the fused low-latency kernel in me_single.cu).
//
// Replaces TEncSearch::xPatternSearch (TEncSearch.cpp:3932-3989): every candidate of the
// window [L..R] x [T..B] is costed with SAD (sub-sampled rows under FEN, :3950-3956) plus
// TComRdCost::getCost(x, y) at scale 2, and the FIRST strict minimum in raster order
// (y outer, x inner, :3959-3983) wins.  The reduction therefore carries the 64-bit key
// (cost << 32 | raster index) and takes its minimum.
//
// The reference window (W + R - L) x (H' rows, all of them needed) is staged once in shared
// memory with aligned 128-bit loads; the PU block is staged next to it.  8-bit pictures use
// packed bytes: a thread owns the four candidates that share one aligned 32-bit column of the
// window, builds their byte-shifted operands with one funnel shift each and accumulates with
// VABSDIFF4.U8.ACC.  >8-bit pictures and int16 key patterns (bi-pred) take a scalar path.
#pragma once
#include <cuda.h>
#include "hmgpu_internal.cuh"
#include <type_traits>

#define FS_THREADS 256
#define FS_YT 4            // candidate rows per work item of the packed full search

__device__ __forceinline__ unsigned long long fs_block_min(unsigned long long key, unsigned long long* s_red)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other < key ? other : key;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = key;
  __syncthreads();
  if (warp == 0)
  {
    key = lane < (FS_THREADS / 32) ? s_red[lane] : ~0ull;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
    {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other < key ? other : key;
    }
  }
  return key; // valid in warp 0
}

// ---- SAD of one work item of the packed full search (see full_search_block_packed) --------------------------
struct FsItem
{
  const unsigned char* s_win; const uint32_t* s_org;
  int spitch, wq, rows, rmul, cy0, g, win_h;
};

// window row u of the item (u = 0 .. rows + FS_YT - 2), word chunk [kc, kc + KC): candidate rows JLO..JHI are
// active, candidate row j sees PU row u - j
template <int KC, int JLO, int JHI>
__device__ __forceinline__ void fs_step(const FsItem& it, int u, int kc, uint32_t (&acc)[FS_YT][4])
{
  // (clamped: rows past the window belong to candidate rows >= ny, which the caller discards)
  const int wr = min(it.cy0 + u * it.rmul, it.win_h - 1);
  const uint32_t* wrow = (const uint32_t*)(it.s_win + (size_t)wr * it.spitch) + it.g + kc;
  uint32_t w[KC + 1];
#pragma unroll
  for (int k = 0; k <= KC; k++) w[k] = wrow[k];
  uint32_t o[JHI - JLO + 1][KC];
#pragma unroll
  for (int j = JLO; j <= JHI; j++)
  {
    const uint32_t* orow = it.s_org + (u - j) * it.wq + kc;
    if (KC == 4) { const uint4 q = *(const uint4*)orow; o[j - JLO][0] = q.x; o[j - JLO][1] = q.y; o[j - JLO][2] = q.z; o[j - JLO][3] = q.w; }
    else
    {
#pragma unroll
      for (int k = 0; k < KC; k++) o[j - JLO][k] = orow[k];
    }
  }
#pragma unroll
  for (int k = 0; k < KC; k++)
  {
    const uint32_t v1 = __funnelshift_r(w[k], w[k + 1], 8), v2 = __funnelshift_r(w[k], w[k + 1], 16), v3 = __funnelshift_r(w[k], w[k + 1], 24);
#pragma unroll
    for (int j = JLO; j <= JHI; j++)
    {
      acc[j][0] = vabsdiff4_acc(w[k], o[j - JLO][k], acc[j][0]);
      acc[j][1] = vabsdiff4_acc(v1, o[j - JLO][k], acc[j][1]);
      acc[j][2] = vabsdiff4_acc(v2, o[j - JLO][k], acc[j][2]);
      acc[j][3] = vabsdiff4_acc(v3, o[j - JLO][k], acc[j][3]);
    }
  }
}

template <int KC>
__device__ __forceinline__ void fs_item_sad(const FsItem& it, uint32_t (&acc)[FS_YT][4])
{
  static_assert(FS_YT == 4, "the ramps below are written for 4 candidate rows");
  for (int kc = 0; kc < it.wq; kc += KC)
  {
    // ramp up (rows >= 4 always), branch-free steady state with all FS_YT candidate rows active, ramp down
    fs_step<KC, 0, 0>(it, 0, kc, acc); fs_step<KC, 0, 1>(it, 1, kc, acc); fs_step<KC, 0, 2>(it, 2, kc, acc);
    for (int u = FS_YT - 1; u < it.rows; u++) fs_step<KC, 0, 3>(it, u, kc, acc);
    fs_step<KC, 1, 3>(it, it.rows, kc, acc); fs_step<KC, 2, 3>(it, it.rows + 1, kc, acc); fs_step<KC, 3, 3>(it, it.rows + 2, kc, acc);
  }
}

// ---- TMA staging of the search window ---------------------------------------------------------------------------------------
// The window of xPatternSearch is a 2-D box out of a pitched plane -- what cp.async.bulk.tensor copies.  One tensor map per row
// width (a multiple of 16 bytes, 32 .. 256) over the phase planes of all reference slots (x = byte of the padded row, y = padded
// row, z = phase plane, w = slot; api.cu keeps the slots in one allocation), box = that width x FS_TMA_ROWS rows.  Thread 0 issues
// ceil(rows / FS_TMA_ROWS) copies against one mbarrier and the CTA stages the PU block meanwhile.  Rules learnt the hard way
// (profiles/probes/tma_probe.cu): the first coordinate x element size must be a multiple of 16 bytes -- anything else is an
// "illegal instruction" -- and the copy is ONE warp-level instruction with uniform operands (UTMALDG).
#define FS_TMA_ROWS 32
#define FS_TMA_MAPS 15
struct FsMaps { CUtensorMap m[FS_TMA_MAPS + 1]; };

__device__ __forceinline__ uint32_t fs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fs_tma_box(uint32_t dst, const void* tmap, int x, int y, int z, int slot, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(slot), "r"(bar) : "memory");
}
__device__ __forceinline__ void fs_mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "FS_WAIT:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra FS_DONE;\n\t"
    "bra FS_WAIT;\n\t"
    "FS_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// shared memory of one job of the packed full search: PU block (visited rows), then the window rows at 128-byte alignment
// (TMA destination), padded to whole copies of FS_TMA_ROWS rows
__host__ __device__ static inline int fs_packed_smem_bytes(int pu_w, int pu_h, int nx, int ny, bool fen)
{
  const int rows = (fen && pu_h > 8) ? pu_h >> 1 : pu_h;
  const int win_w = 15 + nx - 1 + pu_w + 4;
  const int spitch = ((win_w + 15) >> 4) * 16 + 16;
  const int win_h = ny - 1 + pu_h;
  return ((rows * pu_w + 127) & ~127) + spitch * ((win_h + FS_TMA_ROWS - 1) / FS_TMA_ROWS * FS_TMA_ROWS) + 128;
}

// ---- packed 8-bit path -------------------------------------------------------------------------
// Executed by one CTA of FS_THREADS threads; smem: dynamic shared memory sized by
// fs_packed_smem_bytes(); thread 0 writes *out.
__device__ __forceinline__ void full_search_block_packed(const hmgpu_me_job& jb, const RefTable& refs, const OrgView& org,
                                                         unsigned char* smem_raw, unsigned long long* s_red, hmgpu_me_result* out,
                                                         const FsMaps* maps = NULL, unsigned long long* s_bar = NULL)
{
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const int W = jb.pu_w, H = jb.pu_h;
  const int sub = ((jb.flags & HMGPU_F_FEN) && H > 8) ? 1 : 0;
  const int rows = H >> sub, rmul = 1 << sub;
  const int L = jb.win_l, T = jb.win_t, R = jb.win_r, B = jb.win_b;
  const int nx = R - L + 1, ny = B - T + 1;
  if (nx <= 0 || ny <= 0)
  {
    if (threadIdx.x == 0)
    {
      hmgpu_me_result r; memset(&r, 0, sizeof r);
      r.int_sad = 0xffffffffu - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, 0, 0);
      *out = r;
    }
    return;
  }
  const int pitch = refs.pitch;
  const uint8_t* ref00 = (const uint8_t*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;
  // window origin, aligned down to 16 bytes in memory
  const uint8_t* wl = ref00 + (ptrdiff_t)T * pitch + L;
  const int mis = (int)((uintptr_t)wl & 15);
  const uint8_t* wbase = wl - mis;
  const int win_w = mis + nx - 1 + W + 4;                 // bytes needed per row (+4 for the funnel look-ahead)
  const int rw16 = (win_w + 15) >> 4;                     // 16-byte chunks per row
  const int spitch = rw16 * 16 + 16;                      // odd multiple of 16 B keeps rows on different banks
  const int win_h = ny - 1 + H;
  uint32_t* s_org = (uint32_t*)smem;                      // rows * W bytes (visited rows only)
  unsigned char* s_win = smem + ((rows * W + 127) & ~127);
  const bool tma = maps != NULL && spitch <= 256;           // box widths up to 256 bytes have a tensor map
  if (tma)
  {
    if (threadIdx.x == 0)
    {
      const uint32_t bar = fs_smem_u32(s_bar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const int n_ops = (win_h + FS_TMA_ROWS - 1) / FS_TMA_ROWS;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(n_ops * FS_TMA_ROWS * spitch)) : "memory");
      const int x = HMGPU_MARGIN + jb.pu_x + L - mis, y = HMGPU_MARGIN + jb.pu_y + T;   // wbase in padded-plane coordinates
      for (int op = 0; op < n_ops; op++)
        fs_tma_box(fs_smem_u32(s_win) + (uint32_t)(op * FS_TMA_ROWS * spitch), &maps->m[(spitch >> 4) - 2], x, y + op * FS_TMA_ROWS, 0, jb.ref_slot, bar);
    }
    __syncthreads();                                       // the barrier is initialised for everybody who waits on it
  }

  // stage PU block (visited rows only) and window
  {
    const int wq = W >> 2;
    const uint8_t* o = (const uint8_t*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
    for (int i = threadIdx.x; i < rows * wq; i += FS_THREADS)
    {
      const int r = i / wq, k = i - r * wq;
      s_org[i] = __ldg((const uint32_t*)(o + (size_t)(r * rmul) * org.pitch) + k);
    }
    // window: 16 lanes per row (a row is 10..15 chunks of 16 bytes), 16 rows per pass of the CTA, no index division
    if (tma) fs_mbar_wait(fs_smem_u32(s_bar), 0);
    else
    {
      const int rr = threadIdx.x >> 4, k = threadIdx.x & 15;
      if (k < rw16)
        for (int r = rr; r < win_h; r += FS_THREADS / 16)
          *(uint4*)(s_win + (size_t)r * spitch + k * 16) = __ldg((const uint4*)(wbase + (size_t)r * pitch) + k);
      for (int k2 = 16 + k; k2 < rw16; k2 += 16)             // rows wider than 256 bytes (not reached by 8..64-wide PUs at SR 64)
        for (int r = rr; r < win_h; r += FS_THREADS / 16)
          *(uint4*)(s_win + (size_t)r * spitch + k2 * 16) = __ldg((const uint4*)(wbase + (size_t)r * pitch) + k2);
    }
  }
  __syncthreads();

  // Work item = FS_YT candidate rows x the 4 candidates of one aligned 32-bit column group g (window byte offsets
  // 4g .. 4g+3, i.e. x = L + 4g - mis + s).  The FS_YT candidate rows cy0 + j*rmul of an item visit the SAME window
  // rows (candidate row c pairs PU row r with window row c + r*rmul), so the three funnel shifts that build the
  // byte-shifted operands of a window word are shared by FS_YT * 4 candidates: per window word 1 load + 3 SHF feed
  // 4 * FS_YT VABSDIFF4.U8.ACC -- the ALU pipe spends 16 of 19 slots on SADs (the one-row version spent 4 of 7).
  // The PU row is walked in chunks of KC words with KC a template parameter, so the inner loops unroll completely
  // and every shared-memory operand is a base register + immediate.
  const int wq = W >> 2;
  const int ng = (mis + nx + 3) >> 2;
  const int ngy = ((ny + FS_YT * rmul - 1) / (FS_YT * rmul)) * rmul;   // row groups: block b, parity p -> gy = b*rmul + p
  unsigned long long best = ~0ull;
  for (int it = threadIdx.x; it < ngy * ng; it += FS_THREADS)
  {
    const int gy = it / ng, g = it - gy * ng;
    const int cy0 = (gy / rmul) * (FS_YT * rmul) + (gy % rmul);
    uint32_t acc[FS_YT][4];
#pragma unroll
    for (int j = 0; j < FS_YT; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0; }
    const FsItem item = { s_win, s_org, spitch, wq, rows, rmul, cy0, g, win_h };
    if ((wq & 3) == 0) fs_item_sad<4>(item, acc);
    else if (wq == 2) fs_item_sad<2>(item, acc);
    else if (wq == 1) fs_item_sad<1>(item, acc);
    else fs_item_sad<3>(item, acc);                          // 12- and 24-wide AMP parts
    // MV cost: the exp-Golomb length of a component depends on x or on y alone (TComRdCost.h:171-188)
    uint32_t bx[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) bx[s4] = hm_component_bits(((L + 4 * g - mis + s4) << 2) - jb.pred_x);
    // the 16 candidates of the item, ordered (cost, raster position): inside an item the raster order is (j, s4), so
    // (cost << 4 | j*4 + s4) is a 32-bit key (8-bit video: cost < 2^21) and the item minimum takes 15 VIMNMX;
    // one 64-bit (cost, raster index) key per item then joins the thread's running minimum
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < FS_YT; j++)
    {
      const int cy = cy0 + j * rmul;
      const uint32_t by = hm_component_bits(((T + cy) << 2) - jb.pred_y);
#pragma unroll
      for (int s4 = 0; s4 < 4; s4++)
      {
        const int xi = 4 * g - mis + s4;                     // candidate index along x
        const bool ok = cy < ny && xi >= 0 && xi < nx;
        const uint32_t cost = hm_sad_norm(acc[j][s4], sub, 8) + ((jb.ui_cost * (bx[s4] + by)) >> 16);
        m = min(m, ok ? ((cost << 4) | (uint32_t)(j * 4 + s4)) : 0xffffffffu);
      }
    }
    if (m != 0xffffffffu)
    {
      const int j = (m >> 2) & 3, s4 = m & 3;
      const unsigned long long key = ((unsigned long long)(m >> 4) << 32) | (uint32_t)((cy0 + j * rmul) * nx + (4 * g - mis + s4));
      best = key < best ? key : best;
    }
  }
  best = fs_block_min(best, s_red);
  if (threadIdx.x == 0)
  {
    const uint32_t idx = (uint32_t)best, cost = (uint32_t)(best >> 32);
    const int by = T + (int)(idx / nx), bx = L + (int)(idx % nx);
    hmgpu_me_result r; memset(&r, 0, sizeof r);
    r.int_x = (int16_t)bx; r.int_y = (int16_t)by;
    r.int_sad = cost - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, bx, by);
    r.n_cand = (uint32_t)(nx * ny);
    *out = r;
  }
}

// ---- generic path: Px elements, int16 key pattern ------------------------------------------------
// Executed by one CTA of FS_THREADS threads; s_org: 64*64 int16 of shared memory; thread 0 writes *out.
template <typename Px>
__device__ __forceinline__ void full_search_block_generic(const hmgpu_me_job& jb, const int16_t* __restrict__ org_blocks,
                                                          const RefTable& refs, const OrgView& org, int16_t* s_org,
                                                          unsigned long long* s_red, hmgpu_me_result* out)
{
  const int W = jb.pu_w, H = jb.pu_h;
  const int sub = ((jb.flags & HMGPU_F_FEN) && H > 8) ? 1 : 0;
  const int rows = H >> sub, rmul = 1 << sub;
  const int L = jb.win_l, T = jb.win_t, R = jb.win_r, B = jb.win_b;
  const int nx = R - L + 1, ny = B - T + 1;
  if (nx <= 0 || ny <= 0)
  {
    if (threadIdx.x == 0)
    {
      hmgpu_me_result r; memset(&r, 0, sizeof r);
      r.int_sad = 0xffffffffu - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, 0, 0);
      *out = r;
    }
    return;
  }
  const int pitch = refs.pitch;
  const Px* ref00 = (const Px*)refs.base[jb.ref_slot] + (ptrdiff_t)jb.pu_y * pitch + jb.pu_x;
  if (jb.flags & HMGPU_F_ORG_BLOCK)
  {
    const int16_t* o = org_blocks + jb.org_offset;
    for (int i = threadIdx.x; i < W * H; i += FS_THREADS) s_org[i] = __ldcv(o + i);   // see me_tz_impl.cuh: may be rewritten host memory
  }
  else
  {
    const Px* o = (const Px*)org.base + (size_t)jb.pu_y * org.pitch + jb.pu_x;
    for (int i = threadIdx.x; i < W * H; i += FS_THREADS)
    {
      const int r = i / W, k = i - r * W;
      s_org[i] = (int16_t)o[(size_t)r * org.pitch + k];
    }
  }
  __syncthreads();
  unsigned long long best = ~0ull;
  for (int it = threadIdx.x; it < nx * ny; it += FS_THREADS)
  {
    const int cy = it / nx, cx = it - cy * nx;
    const Px* p = ref00 + (ptrdiff_t)(T + cy) * pitch + (L + cx);
    uint32_t acc = 0;
    for (int r = 0; r < rows; r++)
    {
      const Px* pr = p + (size_t)(r * rmul) * pitch;
      const int16_t* o = s_org + (r * rmul) * W;
      for (int k = 0; k < W; k++) acc += (uint32_t)hm_abs((int)o[k] - (int)__ldg(pr + k));
    }
    const uint32_t cost = hm_sad_norm(acc, sub, refs.bit_depth) + hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, L + cx, T + cy);
    const unsigned long long key = ((unsigned long long)cost << 32) | (uint32_t)it;
    best = key < best ? key : best;
  }
  best = fs_block_min(best, s_red);
  if (threadIdx.x == 0)
  {
    const uint32_t idx = (uint32_t)best, cost = (uint32_t)(best >> 32);
    const int by = T + (int)(idx / nx), bx = L + (int)(idx % nx);
    hmgpu_me_result r; memset(&r, 0, sizeof r);
    r.int_x = (int16_t)bx; r.int_y = (int16_t)by;
    r.int_sad = cost - hm_mv_cost(jb.ui_cost, jb.pred_x, jb.pred_y, 2, bx, by);
    r.n_cand = (uint32_t)(nx * ny);
    *out = r;
  }
}

