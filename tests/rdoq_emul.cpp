// tests/rdoq_emul.cpp -- TEST INFRASTRUCTURE ONLY (never part of libhmgpu.so).
// Runs the __host__ __device__ body of the RDOQ kernel (hm-16.2_b200/csrc/rdoq_impl.cuh) lane by lane on the CPU, in the phase
// order rdoq.cu gives it, so that the kernel's logic -- the split of the reference's loop into lane-parallel phases and a
// sequential one, the scan-position workspace, the order of the floating-point operations -- can be checked against the
// reference's dumped calls where there is no GPU (tests/test_rdoq_emul.py).  Built by the test with
// g++ -O2 -ffp-contract=off -shared; the GPU tests check the kernel itself.
#include <vector>
#include <cstring>
#include "../hm-16.2_b200/csrc/rdoq_impl.cuh"

extern "C" int rdoq_emul(const hmgpu_rdoq_job* job, const hmgpu_rdoq_bits* bits, const int32_t* coef, int32_t* level, int lanes)
{
  static uint16_t tab[RQ_SCAN_WORDS];
  static bool built = false;
  if (!built) { rq_build_scan_table(tab); built = true; }
  const hmgpu_rdoq_job& j = *job;
  const int n_coef = 1 << (2 * j.log2_size);
  std::vector<double> store((size_t)n_coef * RQ_WORK_BYTES_PER_COEF / 8 + 1);
  // garbage in the workspace: nothing may depend on what an earlier TU left behind
  memset(store.data(), 0xA5, store.size() * 8);
  RqWork w = rq_carve(store.data(), n_coef);
  const uint16_t* scan = tab + rq_scan_base(j.scan, j.log2_size - 2);
  const uint16_t* scan_cg = tab + rq_cg_base(j.scan, j.log2_size - 2);
  int last_pos = -1;
  for (int l = 0; l < lanes; l++) { const int v = rq_prepass(j, scan, coef, w, l, lanes); if (v > last_pos) last_pos = v; }
  int best_end = 0, sum = 0;
  if (last_pos >= 0)
  {
    best_end = rq_decide(j, bits, scan, scan_cg, w, last_pos);
    for (int l = 0; l < lanes; l++) sum += rq_finish(j, scan, coef, w, best_end, last_pos, l, lanes);
    if ((j.flags & HMGPU_RDOQ_SIGN_HIDE) && sum >= 2)
      for (int l = 0; l < lanes; l++) rq_hide_signs(j, scan, coef, w, best_end, l, lanes);
  }
  for (int i = 0; i < n_coef; i++) level[i] = w.lv[i];
  return sum;
}

extern "C" void rdoq_emul_scan_table(uint16_t* tab) { rq_build_scan_table(tab); }

// the one-thread-per-TU body (rq2_tu); ghost != 0: as if a neighbouring lane of the warp carried a TU that keeps every lockstep
// loop running to its end, so the lane's predicates decide alone
extern "C" int rdoq_emul_tu(const hmgpu_rdoq_job* job, const hmgpu_rdoq_bits* bits, const int32_t* coef, int32_t* level, int ghost)
{
  static uint16_t tab[RQ_SCAN_WORDS];
  static bool built = false;
  if (!built) { rq_build_scan_table(tab); built = true; }
  const hmgpu_rdoq_job& j = *job;
  const int n_coef = 1 << (2 * j.log2_size);
  std::vector<double> store((size_t)n_coef * RQ2_BYTES_PER_COEF / 8);
  memset(store.data(), 0xA5, store.size() * 8);
  Rq2Work w = rq2_carve(store.data(), n_coef, 0);
  memset(level, 0, sizeof(int32_t) * n_coef);                      // (the kernel's level buffer starts out as zeros)
  rq_ghost_top = ghost ? n_coef - 1 : -1;
  Rq2Bits eb; eb.p = (const int32_t*)bits;
  const int sum = rq2_tu(j, true, j.log2_size, eb, tab + rq_scan_base(j.scan, j.log2_size - 2), tab + rq_cg_base(j.scan, j.log2_size - 2), coef, level, w);
  rq_ghost_top = -1;
  return sum;
}
