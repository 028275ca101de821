// prof.cu -- per-stage device timing and the integer-pipe microbenchmarks the roofline needs.
//
// The reference has no profiler hooks (clock() only).  bench.py needs the
// duration of each kernel family measured with CUDA events ON THE STREAM THE KERNELS RUN ON;
// that stream belongs to the context, so the events are recorded here.
//
// hmgpu_microbench measures what MEASURED_PEAKS.json does not contain: the sustained rate of
// the two instructions the search kernels are made of -- a dependent-free stream of 32-bit
// integer adds (IADD3 / IMAD.IADD, both integer pipes) and of VABSDIFF4.U8.ACC (packed
// 4-pixel SAD).  These are the denominators of the INT32-pipe roofline.
#include "hmgpu_internal.cuh"

static const char* const k_stage_names[HMGPU_ST_COUNT] = {
  "planes", "org", "tz", "full", "frac_expand", "frac_dist", "frac_select", "dist", "transform", "quant", "mc", "single" };

static void prof_drain(hmgpu_ctx* ctx)
{
  if (ctx->prof_n == 0) return;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->lane_store[1 - ctx->cur_lane].stream) cudaStreamSynchronize(ctx->lane_store[1 - ctx->cur_lane].stream);
  for (int i = 0; i < ctx->prof_n; i++)
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]) == cudaSuccess)
      ctx->prof_ms[ctx->prof_stage[i]] += ms;
  }
  ctx->prof_n = 0;
}

void hmgpu_prof_begin(hmgpu_ctx* ctx, int stage)
{
  if (ctx->prof_n == 1024) prof_drain(ctx);
  const int i = ctx->prof_n;
  if (!ctx->prof_ev[2 * i])
  {
    cudaEventCreate(&ctx->prof_ev[2 * i]);
    cudaEventCreate(&ctx->prof_ev[2 * i + 1]);
  }
  ctx->prof_stage[i] = stage;
  cudaEventRecord(ctx->prof_ev[2 * i], ctx->stream);
}

void hmgpu_prof_end(hmgpu_ctx* ctx)
{
  cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + 1], ctx->stream);
  ctx->prof_n++;
}

// ---- microbenchmarks -----------------------------------------------------------------------------

// 8 independent accumulators per thread, ITER * 8 adds: ptxas keeps them as IADD3 / IMAD.IADD
__global__ void __launch_bounds__(256) mb_iadd_kernel(uint32_t* out, uint32_t seed, int iters)
{
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
  const uint32_t b = seed | 1;
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int u = 0; u < 8; u++)
    {
      a0 += b ^ a1; a1 += b ^ a2; a2 += b ^ a3; a3 += b ^ a4; a4 += b ^ a5; a5 += b ^ a6; a6 += b ^ a7; a7 += b ^ a0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) mb_vabsdiff4_kernel(uint32_t* out, uint32_t seed, int iters)
{
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  uint32_t x = seed + threadIdx.x * 0x01010101u, y = seed * 7 + blockIdx.x;
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int u = 0; u < 8; u++)
    {
      a0 = vabsdiff4_acc(x, y, a0); a1 = vabsdiff4_acc(y, a0, a1); a2 = vabsdiff4_acc(x, a1, a2); a3 = vabsdiff4_acc(y, a2, a3);
      a4 = vabsdiff4_acc(x, a3, a4); a5 = vabsdiff4_acc(y, a4, a5); a6 = vabsdiff4_acc(x, a5, a6); a7 = vabsdiff4_acc(y, a6, a7);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// IDP.4A.U8.S8 stream (the horizontal Hadamard pass of the SATD kernel is built from it)
__global__ void __launch_bounds__(256) mb_dp4a_kernel(uint32_t* out, uint32_t seed, int iters)
{
  int a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  const uint32_t x = seed + threadIdx.x * 0x01010101u;
  const int y = (int)(seed * 7 + blockIdx.x);
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int u = 0; u < 8; u++)
    {
      a0 = hm_dp4a_us(x, (uint32_t)y, a0); a1 = hm_dp4a_us(x, (uint32_t)a0, a1); a2 = hm_dp4a_us(x, (uint32_t)a1, a2); a3 = hm_dp4a_us(x, (uint32_t)a2, a3);
      a4 = hm_dp4a_us(x, (uint32_t)a3, a4); a5 = hm_dp4a_us(x, (uint32_t)a4, a5); a6 = hm_dp4a_us(x, (uint32_t)a5, a6); a7 = hm_dp4a_us(x, (uint32_t)a6, a7);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7);
}

// dp4a + independent add stream: do the FMA-side IDP and the ALU-side adds overlap?
__global__ void __launch_bounds__(256) mb_dp4a_add_kernel(uint32_t* out, uint32_t seed, int iters)
{
  int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  uint32_t b0 = seed, b1 = seed * 3, b2 = seed * 5, b3 = seed * 7;
  const uint32_t x = seed + threadIdx.x * 0x01010101u;
  const int y = (int)(seed * 7 + blockIdx.x);
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int u = 0; u < 16; u++)
    {
      a0 = hm_dp4a_us(x, (uint32_t)y, a0); a1 = hm_dp4a_us(x, (uint32_t)a0, a1); a2 = hm_dp4a_us(x, (uint32_t)a1, a2); a3 = hm_dp4a_us(x, (uint32_t)a2, a3);
      b0 = max(b0, b1) ^ b2; b1 = max(b1, b2) ^ b3; b2 = max(b2, b3) ^ b0; b3 = max(b3, b0) ^ b1;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(a0 + a1 + a2 + a3) + b0 + b1 + b2 + b3;
}

extern "C" {

int hmgpu_profile_enable(hmgpu_ctx* ctx, int on)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (ctx->remote) return hmgpu_fail(ctx, HMGPU_E_STATE, "profiling is not available through the broker daemon");
  prof_drain(ctx);
  ctx->prof_on = on != 0;
  return HMGPU_OK;
}

int hmgpu_profile_stage_count(void) { return HMGPU_ST_COUNT; }

const char* hmgpu_profile_stage_name(int stage)
{
  return (stage >= 0 && stage < HMGPU_ST_COUNT) ? k_stage_names[stage] : "";
}

// accumulated device milliseconds and launch counts per stage since the last reset
int hmgpu_profile_read(hmgpu_ctx* ctx, double* ms, uint64_t* launches, int reset)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (ctx->remote) return hmgpu_fail(ctx, HMGPU_E_STATE, "profiling is not available through the broker daemon");
  prof_drain(ctx);
  for (int i = 0; i < HMGPU_ST_COUNT; i++)
  {
    if (ms) ms[i] = ctx->prof_ms[i];
    if (launches) launches[i] = ctx->prof_launches[i];
    if (reset) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
  }
  return HMGPU_OK;
}

// which: 0 = 32-bit integer add stream (IADD3 + logic op per element pair), 1 = VABSDIFF4.U8.ACC.
// Returns giga warp-lane operations per second (lane-ops: one per thread per instruction).
int hmgpu_microbench(hmgpu_ctx* ctx, int which, double* gops)
{
  if (!ctx || !gops) return HMGPU_E_INVALID;
  if (ctx->remote) return hmgpu_fail(ctx, HMGPU_E_STATE, "the microbenchmark is not available through the broker daemon");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const int blocks = HMGPU_NUM_SMS * 8, threads = 256, iters = 4096;
  int rc = hmgpu_reserve_work(ctx, sizeof(uint32_t) * (size_t)blocks * threads);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  HMGPU_CUDA(ctx, cudaEventCreate(&e0));
  HMGPU_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++)
  {
    HMGPU_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (which == 0) mb_iadd_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->d_work, 12345u + rep, iters);
    else if (which == 1) mb_vabsdiff4_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->d_work, 12345u + rep, iters);
    else if (which == 2) mb_dp4a_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->d_work, 12345u + rep, iters);
    else mb_dp4a_add_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->d_work, 12345u + rep, iters);
    HMGPU_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    HMGPU_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    HMGPU_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  ctx->launches += 5;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  // instructions per thread: iadd kernel = 2 per statement (xor + add) * 64 statements per iteration
  // dp4a kernel: 64 per iteration; dp4a+add kernel: 16 * (4 dp4a + 8 alu) = 192 per iteration
  const double per_thread = (which == 0 ? 128.0 : which == 3 ? 192.0 : 64.0) * iters;
  *gops = per_thread * blocks * threads / (best * 1e-3) / 1e9;
  return HMGPU_OK;
}

} // extern "C"
