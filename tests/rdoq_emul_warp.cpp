// tests/rdoq_emul_warp.cpp -- TEST INFRASTRUCTURE ONLY (never part of libhmgpu.so).
// A warp of rdoq_tu_kernel on the CPU: 32 host threads are the 32 lanes, each runs the __host__ __device__ body rq2_tu
// (hm-16.2_b200/csrc/rdoq_impl.cuh) on its own TU; the warp-wide maximum / vote the body asks for (RQ_WARP_MAX / RQ_WARP_ANY) are
// real collectives across the threads (a barrier each), and the workspace and the bit estimates are laid out as on the device:
// [scan position][lane] and [word][lane].  So the lockstep structure of the kernel -- loop bounds that are the warp's, predicates
// that are the lane's, breaks and continues decided by votes -- runs with 32 DIFFERENT TUs side by side, as on the GPU, where there
// is no GPU (tests/test_rdoq_emul.py).  Built by the test with g++ -O2 -ffp-contract=off -pthread -shared.
#include <pthread.h>
#include <cstring>
#include <thread>
#include <vector>
#include <stdint.h>

static pthread_barrier_t g_bar;
static int g_val[32];
static thread_local int t_lane = 0;
static int emul_warp_max(int v)
{
  g_val[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  int m = g_val[0];
  for (int i = 1; i < 32; i++) m = g_val[i] > m ? g_val[i] : m;
  pthread_barrier_wait(&g_bar);                                    // (nobody overwrites a slot before everybody has read)
  return m;
}
#define RQ_WARP_MAX(v) emul_warp_max((int)(v))
#define RQ_WARP_ANY(p) (emul_warp_max((p) ? 1 : 0) != 0)
#define RQ2_STRIDE 32
#define RQ2_EB_STRIDE 32
#include "../hm-16.2_b200/csrc/rdoq_impl.cuh"

// one warp: lane l < n_live carries jobs[l] (all of one size class, log2); coef / level are the batch arrays the jobs index
extern "C" void rdoq_emul_warp(const hmgpu_rdoq_job* jobs, int n_live, int log2, const hmgpu_rdoq_bits* bits, const int32_t* coef, int32_t* level, int32_t* abs_sum)
{
  static uint16_t tab[RQ_SCAN_WORDS];
  static bool built = false;
  if (!built) { rq_build_scan_table(tab); built = true; }
  const int n_coef = 1 << (2 * log2);
  std::vector<double> slot((size_t)n_coef * 32 * RQ2_BYTES_PER_COEF / 8);
  memset(slot.data(), 0xA5, slot.size() * 8);
  std::vector<int32_t> sbits((size_t)RQ2_BITS_WORDS * 32);
  pthread_barrier_init(&g_bar, NULL, 32);
  std::vector<std::thread> lanes;
  for (int l = 0; l < 32; l++)
    lanes.emplace_back([&, l]() {
      t_lane = l;
      const bool has_tu = l < n_live;
      hmgpu_rdoq_job j;
      if (has_tu) j = jobs[l];
      else { memset(&j, 0, sizeof j); j.qbits = 14; j.err_scale = 1.0; j.lambda = 1.0; }      // (as rdoq_tu_kernel fills an empty lane)
      const int32_t* src = (const int32_t*)(bits + j.bits_index);
      for (int i = 0; i < RQ2_BITS_WORDS; i++) sbits[(size_t)i * 32 + l] = src[i];
      Rq2Bits eb; eb.p = sbits.data() + l;
      Rq2Work w = rq2_carve(slot.data(), n_coef, l);
      const int sum = rq2_tu(j, has_tu, log2, eb, tab + rq_scan_base(j.scan, log2 - 2), tab + rq_cg_base(j.scan, log2 - 2),
                             coef + j.coef_offset, level + j.coef_offset, w);
      if (has_tu) abs_sum[l] = sum;
    });
  for (auto& t : lanes) t.join();
  pthread_barrier_destroy(&g_bar);
}
