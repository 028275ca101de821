// rdoq.cu -- rate-distortion optimised quantisation of a batch of TUs.
//
// Replaces TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1974-2520) as called from TComTrQuant::xQuant (:1074-1118) when
// RDOQ = 1, for the TUs of one batch: every TU brings its own quantiser parameters, lambda and the index of the CABAC bit
// estimates (estBitsSbacStruct, TComTrQuant.h:59-73) that apply to it, so TUs of different CUs / QPs / components share a launch.
//
// Mapping: one launch per TU size; a TU belongs to a group of LANES lanes of a warp (4x4: 1 lane, 8x8: 2, 16x16: 8, 32x32: the
// whole warp), its workspace (44 bytes per coefficient) lives in shared memory.  The lanes of the group share everything that
// does not depend on the level coder's running state (rdoq_impl.cuh, phases A, C, D and the stores); the group's leader runs the
// sequential decisions (phase B) -- double-precision sums in the reference's order, so the chosen levels are bit-identical.
// The parallelism that fills the machine is across TUs: 148 SMs x 4 warps x 32 / LANES TUs in flight.
#include "hmgpu_internal.cuh"
#include "rdoq_impl.cuh"

#define RDOQ_WARPS 4            // warps per CTA: 4 x 45 KB of workspace for 32x32 TUs

template <int LOG2, int LANES>
__global__ void __launch_bounds__(RDOQ_WARPS * 32)
rdoq_kernel(const hmgpu_rdoq_job* __restrict__ jobs, const int* __restrict__ list, int n_list, const hmgpu_rdoq_bits* __restrict__ bits,
            const uint16_t* __restrict__ scan_tab, const int32_t* __restrict__ coef_all, int32_t* __restrict__ level_all, int32_t* __restrict__ abs_sum)
{
  constexpr int N_COEF = 1 << (2 * LOG2), PER_WARP = 32 / LANES;
  constexpr int WORK = N_COEF * RQ_WORK_BYTES_PER_COEF + (LANES == 1 ? 8 : 0);      // (one more word pair per TU: fewer bank conflicts)
  extern __shared__ __align__(16) unsigned char s_work[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / LANES, gl = lane % LANES;
  const unsigned grp_mask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (grp * LANES));
  RqWork w = rq_carve(s_work + (size_t)(warp * PER_WARP + grp) * WORK, N_COEF);

  for (int first = (blockIdx.x * RDOQ_WARPS + warp) * PER_WARP; first < n_list; first += gridDim.x * RDOQ_WARPS * PER_WARP)
  {
    const int item = first + grp;
    const bool active = item < n_list;
    const int ji = active ? list[item] : 0;
    hmgpu_rdoq_job j;
    if (active) j = jobs[ji];
    else { j.scan = 0; j.coef_offset = 0; j.bits_index = 0; j.flags = 0; }
    const uint16_t* scan = scan_tab + rq_scan_base(j.scan, LOG2 - 2);
    const uint16_t* scan_cg = scan_tab + rq_cg_base(j.scan, LOG2 - 2);
    const int32_t* coef = coef_all + j.coef_offset;
    int32_t* level = level_all + j.coef_offset;

    // A: per-coefficient terms, last position
    int last_pos = active ? rq_prepass(j, scan, coef, w, gl, LANES) : -1;
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) last_pos = max(last_pos, __shfl_xor_sync(0xffffffffu, last_pos, d));
    __syncwarp();

    // B: the leader decides
    int best_end = 0;
    if (active && gl == 0 && last_pos >= 0) best_end = rq_decide(j, bits + j.bits_index, scan, scan_cg, w, last_pos);
    __syncwarp();
    best_end = __shfl_sync(0xffffffffu, best_end, grp * LANES);

    // C: signs and the absolute sum
    int sum = (active && last_pos >= 0) ? rq_finish(j, scan, coef, w, best_end, last_pos, gl, LANES) : 0;
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    __syncwarp();

    // D: sign-bit hiding
    if (active && (j.flags & HMGPU_RDOQ_SIGN_HIDE) && sum >= 2) rq_hide_signs(j, scan, coef, w, best_end, gl, LANES);
    __syncwarp();

    if (active)
    {
      for (int i = gl; i < N_COEF; i += LANES) level[i] = w.lv[i];
      if (gl == 0) abs_sum[ji] = sum;
    }
    __syncwarp();
    (void)grp_mask;
  }
}

template <int LOG2, int LANES>
static int launch_class(hmgpu_ctx* ctx, uint32_t attr_bit, const hmgpu_rdoq_job* d_jobs, const int* d_list, int n_list, const hmgpu_rdoq_bits* d_bits,
                        const uint16_t* d_scan, const int32_t* d_coef, int32_t* d_level, int32_t* d_abs_sum)
{
  if (n_list == 0) return HMGPU_OK;
  constexpr int N_COEF = 1 << (2 * LOG2), PER_WARP = 32 / LANES;
  constexpr int WORK = N_COEF * RQ_WORK_BYTES_PER_COEF + (LANES == 1 ? 8 : 0);
  const int smem = RDOQ_WARPS * PER_WARP * WORK;
  if (!(ctx->attr_done & attr_bit))
  {
    HMGPU_CUDA(ctx, cudaFuncSetAttribute(rdoq_kernel<LOG2, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->attr_done |= attr_bit;
  }
  const int per_cta = RDOQ_WARPS * PER_WARP;
  int grid = (n_list + per_cta - 1) / per_cta;
  if (grid > 148 * 8) grid = 148 * 8;
  rdoq_kernel<LOG2, LANES><<<grid, RDOQ_WARPS * 32, smem, ctx->stream>>>(d_jobs, d_list, n_list, d_bits, d_scan, d_coef, d_level, d_abs_sum);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

// ---- one thread per TU (rdoq_impl.cuh, second half) ---------------------------------------------------------------------------
// Warp i of the launch takes 32 TUs of one size class -- the classes follow one another in the warp numbering, largest TUs first,
// so the long chains start first and the short ones fill the machine around them -- and one launch covers all four sizes.  The
// workspace of a warp (24 bytes per coefficient and lane, [scan position][lane]) lives in global memory: resident warps x 768 KB
// for 32x32 TUs, written and read once in full lines.
#define RDOQ_TU_WARPS 2         // warps per CTA
struct RdoqTuPlan
{
  int first_warp[5];            // warps [first_warp[c], first_warp[c + 1]) work on size class 3 - c (32x32 first)
  int first_item[4];            // where the items of that class start in the list
  int n_item[4];
  unsigned long long slot_bytes;
};

__global__ void __launch_bounds__(RDOQ_TU_WARPS * 32)
rdoq_tu_kernel(const hmgpu_rdoq_job* __restrict__ jobs, const int* __restrict__ list, RdoqTuPlan plan, const hmgpu_rdoq_bits* __restrict__ bits,
               const uint16_t* __restrict__ scan_tab, const int32_t* __restrict__ coef_all, int32_t* __restrict__ level_all, int32_t* __restrict__ abs_sum,
               unsigned char* __restrict__ work)
{
  __shared__ int32_t s_bits[RDOQ_TU_WARPS][RQ2_BITS_WORDS * 32];    // 24 KB per warp
  const int lane = threadIdx.x & 31, slot = blockIdx.x * RDOQ_TU_WARPS + (threadIdx.x >> 5);
  for (int wi = slot; wi < plan.first_warp[4]; wi += gridDim.x * RDOQ_TU_WARPS)
  {
    const int c = wi < plan.first_warp[1] ? 0 : wi < plan.first_warp[2] ? 1 : wi < plan.first_warp[3] ? 2 : 3;
    const int log2 = 5 - c, at = (wi - plan.first_warp[c]) * 32 + lane;
    const bool has_tu = at < plan.n_item[c];
    const int ji = has_tu ? list[plan.first_item[c] + at] : 0;
    hmgpu_rdoq_job j;
    if (has_tu) j = jobs[ji];
    else { memset(&j, 0, sizeof j); j.qbits = 14; j.err_scale = 1.0; j.lambda = 1.0; }
    const uint16_t* scan = scan_tab + rq_scan_base(j.scan, log2 - 2);
    const uint16_t* scan_cg = scan_tab + rq_cg_base(j.scan, log2 - 2);
    Rq2Work w = rq2_carve(work + (size_t)slot * plan.slot_bytes, 1 << (2 * log2), lane);
    // the lane's set of bit estimates into shared memory, [word][lane]
    int32_t* my_bits = s_bits[threadIdx.x >> 5] + lane;
    {
      const int4* src = (const int4*)(bits + j.bits_index);
#pragma unroll 8
      for (int i = 0; i < RQ2_BITS_WORDS / 4; i++)
      {
        const int4 v = __ldg(src + i);
        my_bits[(4 * i) * 32] = v.x; my_bits[(4 * i + 1) * 32] = v.y; my_bits[(4 * i + 2) * 32] = v.z; my_bits[(4 * i + 3) * 32] = v.w;
      }
    }
    __syncwarp();
    Rq2Bits eb; eb.p = my_bits;
    const int sum = rq2_tu(j, has_tu, log2, eb, scan, scan_cg, coef_all + j.coef_offset, level_all + j.coef_offset, w);
    if (has_tu) abs_sum[ji] = sum;
    __syncwarp();
  }
}

static int launch_tu(hmgpu_ctx* ctx, const hmgpu_rdoq_job* d_jobs, const int* d_list, const int n_class[4], const hmgpu_rdoq_bits* d_bits,
                     const uint16_t* d_scan, const int32_t* d_coef, int32_t* d_level, int32_t* d_abs_sum)
{
  RdoqTuPlan plan;
  int item = n_class[0] + n_class[1] + n_class[2] + n_class[3], warp = 0, largest = -1;
  for (int c = 0; c < 4; c++)                                      // class 3 - c
  {
    const int n = n_class[3 - c];
    item -= n;
    plan.first_warp[c] = warp; plan.first_item[c] = item; plan.n_item[c] = n;
    warp += (n + 31) / 32;
    if (n > 0 && largest < 0) largest = 3 - c;
  }
  plan.first_warp[4] = warp;
  if (warp == 0) return HMGPU_OK;
  plan.slot_bytes = (unsigned long long)RQ2_BYTES_PER_COEF * 32 * (16u << (2 * largest));
  int grid = (warp + RDOQ_TU_WARPS - 1) / RDOQ_TU_WARPS;
  if (grid > 148 * 4) grid = 148 * 4;                              // at most 1184 warps' worth of workspace (0.9 GB for 32x32 TUs)
  int rc;
  if ((rc = hmgpu_reserve_work(ctx, (size_t)grid * RDOQ_TU_WARPS * plan.slot_bytes))) return rc;
  HmgpuStage st(ctx, HMGPU_ST_QUANT, 1);
  rdoq_tu_kernel<<<grid, RDOQ_TU_WARPS * 32, 0, ctx->stream>>>(d_jobs, d_list, plan, d_bits, d_scan, d_coef, d_level, d_abs_sum, (unsigned char*)ctx->d_work);
  HMGPU_CUDA(ctx, cudaGetLastError());
  return HMGPU_OK;
}

// d_list: the job indices sorted by TU size, n_class[s] of size class s = log2 - 2, one class after the other.
// d_level must be zero where no level is written (rdoq_tu_kernel stores the non-zero levels only).
int hmgpu_launch_rdoq(hmgpu_ctx* ctx, const hmgpu_rdoq_job* d_jobs, const int* d_list, const int n_class[4], const hmgpu_rdoq_bits* d_bits,
                      const uint16_t* d_scan, const int32_t* d_coef, int32_t* d_level, int32_t* d_abs_sum)
{
  if (ctx->tune.rdoq_tu) return launch_tu(ctx, d_jobs, d_list, n_class, d_bits, d_scan, d_coef, d_level, d_abs_sum);
  const int launches = (n_class[0] > 0) + (n_class[1] > 0) + (n_class[2] > 0) + (n_class[3] > 0);
  HmgpuStage st(ctx, HMGPU_ST_QUANT, launches);
  int rc;
  // the largest TUs first: their launch has the fewest warps and the longest sequential chains
  const int* l3 = d_list + n_class[0] + n_class[1] + n_class[2];
  const int* l2 = d_list + n_class[0] + n_class[1];
  const int* l1 = d_list + n_class[0];
  if ((rc = launch_class<5, 32>(ctx, HMGPU_ATTR_RDOQ5, d_jobs, l3, n_class[3], d_bits, d_scan, d_coef, d_level, d_abs_sum))) return rc;
  if ((rc = launch_class<4, 8>(ctx, HMGPU_ATTR_RDOQ4, d_jobs, l2, n_class[2], d_bits, d_scan, d_coef, d_level, d_abs_sum))) return rc;
  if ((rc = launch_class<3, 2>(ctx, HMGPU_ATTR_RDOQ3, d_jobs, l1, n_class[1], d_bits, d_scan, d_coef, d_level, d_abs_sum))) return rc;
  if ((rc = launch_class<2, 1>(ctx, HMGPU_ATTR_RDOQ2, d_jobs, d_list, n_class[0], d_bits, d_scan, d_coef, d_level, d_abs_sum))) return rc;
  return HMGPU_OK;
}

void hmgpu_rdoq_scan_table(uint16_t* tab) { rq_build_scan_table(tab); }
