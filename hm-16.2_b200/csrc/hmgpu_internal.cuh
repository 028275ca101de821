// hmgpu_internal.cuh -- shared declarations of libhmgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/hmgpu.h"
#include "broker_proto.h"

#define HMGPU_MARGIN 80          // luma padding, g_uiMaxCUWidth + 16 (TComPicYuv.cpp:87-88)
#define HMGPU_CMARGIN 40         // chroma padding (4:2:0)
#define HMGPU_MAX_REFS 16
#define HMGPU_NUM_SMS 148
#define HMGPU_MAIL_JOBS 32       // jobs per call of the low-latency path (me_single.cu)
#define HMGPU_TZ_STREAMS 15      // side streams of the TZ stage (me_tz_thread.cu)
#define HMGPU_SERVER_CTAS 16     // most CTAs a mailbox server runs with (hmgpu_ctx::srv_ctas of them)
#define HMGPU_MAIL_LINES 32      // job lines of the mailbox = jobs a call through the server can carry

// Device view of the reference planes of one context, passed to kernels by value.
// plane(slot, phase) sample (x, y), x in [-80, W+80): base[slot] + phase*plane_elems + y*pitch + x
struct RefTable
{
  const void* base[HMGPU_MAX_REFS]; // pointer to sample (0,0) of phase plane 0 (integer plane)
  int64_t plane_elems;              // elements between consecutive phase planes
  int32_t pitch;                    // elements per padded row
  int32_t pic_w, pic_h;
  int32_t bit_depth;
};

struct OrgView
{
  const void* base;    // sample (0,0) of the source picture (Px elements)
  int32_t pitch;
};

struct RefSlot
{
  bool  valid;
  bool  has_chroma;    // cb / cr hold this picture (hmgpu_ref_upload with chroma planes)
  void* planes;        // 16 padded planes, Px elements
  void* cb;            // padded chroma planes (int16), may be null
  void* cr;
};

// A pipeline lane: its own stream and its own staging / scratch buffers.  Lane 0 is the context's
// ordinary stream; lane 1 exists so that large hmgpu_me_search batches can copy chunk k+1 in and
// chunk k-1 out while chunk k is being searched.  The launchers always use the CURRENT view
// (ctx->stream, ctx->d_work, ...); hmgpu_use_lane() swaps a lane into that view.
struct HmgpuLane
{
  cudaStream_t stream;
  void* h_pin; size_t h_pin_bytes;
  void* d_stage; size_t d_stage_bytes;
  void* d_work; size_t d_work_bytes;
  void* d_tzlist; size_t d_tzlist_bytes;
};

// Tuning knobs, read from the environment ONCE (hmgpu_create) and changed afterwards only through hmgpu_set_option:
// nothing on a launch path calls getenv.
struct HmgpuTuning
{
  int tz_thread;        // HMGPU_TZ_SPLIT != 0: one-thread-per-job TZ kernels for large 8-bit batches (me_tz_thread.cu)
  int tz_thread_min;    // HMGPU_TZ_THREAD_MIN: batch size from which they are used
  int tz_merge;         // HMGPU_TZ_MERGE: merged passes in the warp-per-job kernel
  int tz_carve;         // HMGPU_TZ_CARVE: shared-memory carve-out (percent) of the warp-per-job kernel, < 0: driver default
  int tz_p2;            // HMGPU_TZ_P2: second one-thread-per-job pass for hand-over jobs
  int frac_v1;          // HMGPU_FRAC_V1: generic fractional kernels for 8-bit pictures too
  int fs_tma;           // HMGPU_FS_TMA: the full-search window is staged by TMA (cp.async.bulk.tensor) instead of per-thread loads
  int frac_overlap;     // HMGPU_FRAC_OVERLAP: 8x8 / 4x4 tile kernels side by side
  int frac_win;         // HMGPU_FRAC_WIN: fractional stage of large 8-bit batches by CTU groups with TMA-staged windows (me_fracw.cu)
  int frac_win_min;     // HMGPU_FRAC_WIN_MIN: batch size from which it is used
  int pipe_chunk;       // HMGPU_PIPE_CHUNK: jobs per chunk of the pipelined batch path (0: short first / last chunk around two long ones)
  int pipe_edge;        // HMGPU_PIPE_EDGE: the first and the last chunk are 1 / pipe_edge of the batch (default 8)
  int pipeline;         // !HMGPU_NO_PIPELINE
  int fastpath;         // !HMGPU_NO_FASTPATH
  int server;           // HMGPU_SERVER: resident mailbox server for calls of <= HMGPU_SERVER_CTAS jobs
  int server_idle_us;   // HMGPU_SERVER_IDLE_US
  int trace;            // HMGPU_TRACE
  int server_stats;     // HMGPU_SERVER_STATS
  int rdoq_tu;          // HMGPU_RDOQ_TU: RDOQ with one thread per TU (rdoq_tu_kernel) instead of one lane group per TU
};

struct hmgpu_ctx
{
  int device, pic_w, pic_h, bit_depth, max_refs;
  HmgpuTuning tune;
  uint32_t attr_tzt;            // same, one bit per (shape, pass) instantiation of tzt_search_kernel
  uint32_t attr_done;           // per-context record of the cudaFuncSetAttribute calls made (attributes are per device)
  int px_bytes;                 // 1 (8-bit) or 2
  int pw, ph, pitch;            // padded luma geometry (elements)
  size_t plane_elems;
  int cpw, cph, cpitch;         // padded chroma geometry
  cudaStream_t stream;
  RefSlot refs[HMGPU_MAX_REFS];
  void* d_org; int org_pitch;   // source picture, Px
  void* planes_all; size_t slot_bytes;   // the phase planes of all reference slots (one allocation), bytes per slot
  void* h_tmaps;                         // CUtensorMap[16] over planes_all for the full-search window (me_full.cu), built on first use
  void* h_fw_tmap;                       // CUtensorMap over planes_all for the windows of the fractional stage (me_fracw.cu)
  // staging (grow on demand)
  void* h_pin; size_t h_pin_bytes;     // pinned host
  void* d_stage; size_t d_stage_bytes; // device
  void* d_work; size_t d_work_bytes;   // device scratch for the search kernels
  void* d_tzlist; size_t d_tzlist_bytes; // device index lists of the TZ size classes (me_tz.cu)
  cudaStream_t tz_streams[HMGPU_TZ_STREAMS]; cudaEvent_t tz_ev[HMGPU_TZ_STREAMS + 1]; // side streams of the TZ stage (me_tz_thread.cu): kernels of different PU shapes overlap
  cudaStream_t frac_stream; cudaEvent_t frac_ev[2];  // side stream of the fractional stage (me_frac2.cu)
  HmgpuLane lane_store[2]; int cur_lane; // parked lanes (the current one lives in the fields above)
  cudaEvent_t lane_done[2], scan_done[2], fork_ev;
  cudaStream_t copy_stream;              // H2D of the next chunk + its validation scan
  void* d_scan; void* h_scan;            // per-lane verdict of the scan (device / pinned host copy)
  void* d_orgblk; size_t d_orgblk_bytes; // bi-pred key patterns of a pipelined batch
  void* h_mail; uint32_t mail_ticket;  // mapped pinned mailbox of the low-latency path (me_single.cu), struct Mailbox
  void* d_mail;                        // the same memory as the device addresses it
  bool  mail_external;                 // h_mail lives in a broker shared-memory segment (not freed with the context)
  struct HmgpuRemote* remote;          // != NULL: client of a broker daemon (remote.cu); this process makes no CUDA call
  int   srv_ctas;                      // CTAs of this context's server kernel (1..HMGPU_SERVER_CTAS)
  // mailbox server (me_server_kernel): a kernel that stays resident between calls
  hmgpu_me_job pend_jobs[HMGPU_MAIL_JOBS]; int pend_n;   // jobs of a submit that has not been waited for (server path)
  hmgpu_pred_job pend_pred[HMGPU_MAIL_LINES]; uint8_t pend_pfunc[HMGPU_MAIL_LINES]; int pend_np;   // its prediction-error jobs
  int defer_n, defer_np, defer_org_n; int16_t* defer_org; size_t defer_org_cap;   // submits that run inside the wait
  cudaStream_t srv_stream; bool srv_alive; uint32_t srv_gen; int srv_dyn; uint32_t srv_calls, srv_starts;
  uint64_t launches;
  // optional per-stage device timing (hmgpu_profile_enable): CUDA events on ctx->stream
  bool prof_on;
  int prof_n;                          // event pairs in flight
  cudaEvent_t prof_ev[2 * 1024];
  int prof_stage[1024];
  double prof_ms[16];
  uint64_t prof_launches[16];
  char err[512];
};

// ---- low-latency path (me_single.cu): mapped pinned mailbox --------------------------------------
// the jobs of one call travel as a kernel parameter (no PCIe read on the device side)
struct HmgpuJobPack { hmgpu_me_job jobs[HMGPU_MAIL_JOBS]; };
// One result slot = 32 bytes written by ONE warp-wide store of 8 consecutive words, so it crosses PCIe as a
// single write.  No system-scope fence is issued (it cost 4 us per call): instead the slot validates itself --
// the host accepts it only when the ticket matches and the check word matches the six result words.
struct HmgpuMailSlot { hmgpu_me_result r; uint32_t ticket; uint32_t check; };
// The mailbox: mapped pinned host memory (or a broker shared-memory segment registered with CUDA) polled by the resident server
// kernel.  A job line is 64 bytes: words 0..11 the job (hmgpu_me_job, or a hmgpu_pred_job in words 0..4), 12 the ticket of the
// call, 13 the server generation, 14 flags (bit 0: the line carries a job; bit 1: it is a prediction-error job, func in bits
// 16..23; bits 8..15: number of lines of the call), 15 the check word over words 0..14.
struct Mailbox
{
  HmgpuMailSlot   slots[HMGPU_MAIL_LINES];
  uint32_t        lines[HMGPU_MAIL_LINES][16];    // host -> server
  uint32_t        exited[HMGPU_SERVER_CTAS];      // server -> host: generation of the server CTA that stopped polling
  unsigned long long trace[8];          // HMGPU_TRACE: globaltimer stamps of the kernel phases
  int16_t         org_blocks[HMGPU_MAIL_JOBS * 64 * 64];
};

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline uint32_t hmgpu_mail_check(const uint32_t* w, uint32_t ticket)
{
  uint32_t h = ticket * 0x9E3779B9u + 0x7F4A7C15u;
  for (int i = 0; i < 6; i++) h = (h ^ w[i]) * 0x85EBCA6Bu + (h >> 15);
  return h;
}

enum
{
  HMGPU_ST_PLANES = 0, HMGPU_ST_ORG, HMGPU_ST_TZ, HMGPU_ST_FULL, HMGPU_ST_FRAC_EXPAND, HMGPU_ST_FRAC_DIST,
  HMGPU_ST_FRAC_SELECT, HMGPU_ST_DIST, HMGPU_ST_TRANSFORM, HMGPU_ST_QUANT, HMGPU_ST_MC, HMGPU_ST_SINGLE, HMGPU_ST_COUNT
};

void hmgpu_prof_begin(hmgpu_ctx* ctx, int stage);
void hmgpu_prof_end(hmgpu_ctx* ctx);

// RAII: counts the launches of one stage and, when profiling is on, brackets them with events
struct HmgpuStage
{
  hmgpu_ctx* ctx;
  HmgpuStage(hmgpu_ctx* c, int stage, int n_launches) : ctx(c)
  {
    c->launches += n_launches;
    c->prof_launches[stage] += n_launches;
    if (c->prof_on) hmgpu_prof_begin(c, stage);
  }
  ~HmgpuStage() { if (ctx->prof_on) hmgpu_prof_end(ctx); }
};

int hmgpu_fail(hmgpu_ctx* ctx, int code, const char* fmt, ...);
// device / pinned-host allocations of the library.  In pool mode (the broker daemon) a freed block is kept for the next request
// of a similar size instead of going back to CUDA: cudaFree / cudaFreeHost synchronise the whole device and would wait for the
// resident server kernels of every other client.
cudaError_t hmgpu_dmalloc(void** p, size_t bytes);
void hmgpu_dfree(void* p);
cudaError_t hmgpu_hmalloc(void** p, size_t bytes);
void hmgpu_hfree(void* p);
int hmgpu_reserve_pinned(hmgpu_ctx* ctx, size_t bytes);
int hmgpu_reserve_stage(hmgpu_ctx* ctx, size_t bytes);
int hmgpu_reserve_work(hmgpu_ctx* ctx, size_t bytes);
int hmgpu_reserve_tzlist(hmgpu_ctx* ctx, size_t bytes);
RefTable hmgpu_ref_table(const hmgpu_ctx* ctx);
void hmgpu_use_lane(hmgpu_ctx* ctx, int lane);

// ---- client of a broker daemon (remote.cu; no CUDA call is made by a process in this mode) ----------------------------
// connect to the daemon's socket, create the remote context, map its shared segment; the returned hmgpu_ctx carries the picture
// geometry, the mailbox (h_mail = the segment's Mailbox), srv_ctas and the tuning values the daemon dictates
int  hmgpu_remote_create(const char* socket_path, int pic_w, int pic_h, int bit_depth, int max_refs, hmgpu_ctx** out);
void hmgpu_remote_destroy(hmgpu_ctx* ctx);
// one round trip on the control socket; a / text / reply may be NULL.  Returns the daemon's rc (its error text lands in ctx->err)
int  hmgpu_remote_call(hmgpu_ctx* ctx, int op, const int32_t a[6], const char* text, BrokerReply* reply);
// which = 0: upload area, 1: batch area of the shared segment
void* hmgpu_remote_area(hmgpu_ctx* ctx, int which, size_t* bytes);
// ---- used by the daemon (hmgpud.cu) on its in-process contexts ---------------------------------------------------------
extern "C" {
void hmgpu_internal_pool_mode(int on);
// like hmgpu_create, with the mailbox placed at `mail` (host address inside a cudaHostRegister'ed segment, sizeof(Mailbox) bytes)
int  hmgpu_internal_create_shared(int device, int pic_w, int pic_h, int bit_depth, int max_refs, void* mail, int srv_ctas, hmgpu_ctx** out);
int  hmgpu_internal_server_launch(hmgpu_ctx* ctx, uint32_t gen, uint32_t last_ticket, int dyn_bytes);
int  hmgpu_internal_server_sync(hmgpu_ctx* ctx);      // cudaStreamSynchronize of the server stream
int  hmgpu_internal_server_query(hmgpu_ctx* ctx);
int  hmgpu_internal_server_kill(hmgpu_ctx* ctx);      // make the server kernel leave whatever the client's state, and wait
size_t hmgpu_internal_mailbox_bytes(void);
}

// bits of hmgpu_ctx::attr_done
enum { HMGPU_ATTR_SINGLE = 1, HMGPU_ATTR_SERVER = 2, HMGPU_ATTR_FULL = 4, HMGPU_ATTR_TZT = 8, HMGPU_ATTR_TZ_CARVE = 16, HMGPU_ATTR_FRAC3 = 32, HMGPU_ATTR_FRACW = 64,
       HMGPU_ATTR_RDOQ2 = 128, HMGPU_ATTR_RDOQ3 = 256, HMGPU_ATTR_RDOQ4 = 512, HMGPU_ATTR_RDOQ5 = 1024 };

#define HMGPU_CUDA(ctx, call)                                                              \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return hmgpu_fail((ctx), HMGPU_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,    \
                        cudaGetErrorString(e__));                                          \
  } while (0)

// launchers implemented in the .cu files (all asynchronous on ctx->stream)
int hmgpu_launch_planes(hmgpu_ctx* ctx, int slot, const int16_t* d_src, int src_stride);
int hmgpu_launch_org(hmgpu_ctx* ctx, const int16_t* d_src, int src_stride);
int hmgpu_launch_me(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                    hmgpu_me_result* d_results, bool any_org_block, bool any_full, bool any_tz, bool any_frac,
                    int max_win_bytes, bool any_sel);
int hmgpu_launch_dist(hmgpu_ctx* ctx, const int16_t* d_org, const int16_t* d_cur,
                      const hmgpu_dist_item* d_items, int n_items, uint32_t* d_out);
int hmgpu_launch_intra_costs(hmgpu_ctx* ctx, const hmgpu_intra_job* d_jobs, int n_jobs, const int16_t* d_org, const int16_t* d_lines, uint32_t* d_dist);
int hmgpu_launch_sao_stats(hmgpu_ctx* ctx, const int16_t* d_rec, int rec_stride, const int16_t* d_org, int org_stride, int width, int height,
                           int ctu_w, int ctu_h, const uint8_t* d_flags, const int32_t* skip_r, const int32_t* skip_b, long long* d_stats);
int hmgpu_launch_sao_apply(hmgpu_ctx* ctx, const int16_t* d_src, int stride, int width, int height, int ctu_w, int ctu_h, const uint8_t* d_flags,
                           const int8_t* d_types, const int32_t* d_offsets, int16_t* d_dst);
int hmgpu_launch_deblock(hmgpu_ctx* ctx, int16_t* d_y, int16_t* d_cb, int16_t* d_cr, int w, int h, int bd_luma, int bd_chroma,
                         const uint8_t* d_bs_ver, const uint8_t* d_bs_hor, const int8_t* d_qp, const uint8_t* d_nf,
                         int beta_off2, int tc_off2, int cb_off, int cr_off);
int hmgpu_launch_mc_luma(hmgpu_ctx* ctx, const hmgpu_mc_job* d_jobs, int n_jobs, int16_t* d_dst);
int hmgpu_launch_predict(hmgpu_ctx* ctx, const hmgpu_pred_job* d_jobs, int n_jobs, int with_chroma, int16_t* d_dst);
int hmgpu_launch_pred_error(hmgpu_ctx* ctx, const hmgpu_pred_job* d_jobs, int n_jobs, int func, uint32_t* d_out);
int hmgpu_launch_inv_transform(hmgpu_ctx* ctx, const int32_t* d_coeff, int n_tus, int n, int use_dst, int16_t* d_resi);
int hmgpu_launch_fwd_transform(hmgpu_ctx* ctx, const int16_t* d_resi, int n_tus, int n, int use_dst, int32_t* d_coeff);
int hmgpu_launch_quant(hmgpu_ctx* ctx, const int32_t* d_coeff, int n_tus, int n, int qp_per, int qp_rem,
                       int is_intra, int32_t* d_level, int32_t* d_delta, uint32_t* d_abs_sum);
int hmgpu_launch_rdoq(hmgpu_ctx* ctx, const hmgpu_rdoq_job* d_jobs, const int* d_list, const int n_class[4], const hmgpu_rdoq_bits* d_bits,
                      const uint16_t* d_scan, const int32_t* d_coef, int32_t* d_level, int32_t* d_abs_sum);
int hmgpu_launch_dequant(hmgpu_ctx* ctx, const int32_t* d_level, int n_tus, int n, const hmgpu_rdoq_job* d_jobs, int qp_per, int qp_rem, int32_t* d_coef);
void hmgpu_rdoq_scan_table(uint16_t* tab);      // 4335 words, layout in rdoq_impl.cuh

// ---------------------------------------------------------------------------------------
// device helpers shared by the search kernels
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Exp-Golomb length, TComRdCost::xGetComponentBits (TComRdCost.cpp:278-292):
// 2*floor(log2(v<=0 ? -2v+1 : 2v)) + 1
__device__ __forceinline__ uint32_t hm_component_bits(int v)
{
  const uint32_t t = (v <= 0) ? (uint32_t)((-v << 1) + 1) : (uint32_t)(v << 1);
  return 2u * (31u - (uint32_t)__clz((int)t)) + 1u;
}

// TComRdCost::getCost(x,y) (TComRdCost.h:171-188): uint32 wrap-around (m_uiCost*bits)>>16
__device__ __forceinline__ uint32_t hm_mv_cost(uint32_t ui_cost, int pred_x, int pred_y, int scale, int x, int y)
{
  const uint32_t bits = hm_component_bits((x << scale) - pred_x) + hm_component_bits((y << scale) - pred_y);
  return (ui_cost * bits) >> 16;
}

// sum of |a_i - b_i| over 4 packed unsigned bytes, accumulated: one VABSDIFF4.U8.ACC
__device__ __forceinline__ uint32_t vabsdiff4_acc(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t r;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// 4-way dot product of UNSIGNED bytes a with SIGNED bytes b, accumulated: one IDP.4A.U8.S8
__device__ __forceinline__ int hm_dp4a_us(uint32_t a, uint32_t b, int c)
{
  int r;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

__device__ __forceinline__ int hm_abs(int v) { return v < 0 ? -v : v; }

// integer-pel SAD normalisation: (sum << iSubShift) >> (bitDepth-8)  (TComRdCost.cpp:520-521)
__device__ __forceinline__ uint32_t hm_sad_norm(uint32_t sum, int sub_shift, int bit_depth)
{
  return (sum << sub_shift) >> (bit_depth - 8);
}

// 8-point Hadamard on 8 ints with stride S inside a register array (fully unrolled)
template <int S>
__device__ __forceinline__ void hm_hadamard8(int* v)
{
#pragma unroll
  for (int i = 0; i < 4; i++) { const int a = v[i * S], b = v[(i + 4) * S]; v[i * S] = a + b; v[(i + 4) * S] = a - b; }
#pragma unroll
  for (int i = 0; i < 8; i += 4)
#pragma unroll
    for (int j = i; j < i + 2; j++) { const int a = v[j * S], b = v[(j + 2) * S]; v[j * S] = a + b; v[(j + 2) * S] = a - b; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) { const int a = v[i * S], b = v[(i + 1) * S]; v[i * S] = a + b; v[(i + 1) * S] = a - b; }
}

// SATD of one 8x8 tile of differences d[64] (row-major), rounding (s+2)>>2
// (xCalcHADs8x8, TComRdCost.cpp:1439-1534).  The last vertical stage is folded into the
// absolute-value sum with |a+b| + |a-b| = 2*max(|a|,|b|).
__device__ __forceinline__ uint32_t hm_satd8x8(int* d)
{
#pragma unroll
  for (int r = 0; r < 8; r++) hm_hadamard8<1>(d + r * 8);
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

// SATD of one 4x4 tile, rounding (s+1)>>1 (xCalcHADs4x4, TComRdCost.cpp:1343-1437)
__device__ __forceinline__ uint32_t hm_satd4x4(int* d)
{
#pragma unroll
  for (int r = 0; r < 4; r++)
  {
    int* v = d + r * 4;
    const int a0 = v[0] + v[2], a1 = v[1] + v[3], a2 = v[0] - v[2], a3 = v[1] - v[3];
    v[0] = a0 + a1; v[1] = a0 - a1; v[2] = a2 + a3; v[3] = a2 - a3;
  }
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

#endif // __CUDACC__
