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

// d_list: the job indices sorted by TU size, n_class[s] of size class s = log2 - 2, one class after the other
int hmgpu_launch_rdoq(hmgpu_ctx* ctx, const hmgpu_rdoq_job* d_jobs, const int* d_list, const int n_class[4], const hmgpu_rdoq_bits* d_bits,
                      const uint16_t* d_scan, const int32_t* d_coef, int32_t* d_level, int32_t* d_abs_sum)
{
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
