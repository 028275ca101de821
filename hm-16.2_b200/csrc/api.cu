// api.cu -- extern "C" entry points of libhmgpu.so (see include/hmgpu.h for the reference
// interfaces each one replaces).  Host buffers are staged through one pinned buffer and one
// device buffer per context; everything runs on the context's stream.
#include "hmgpu_internal.cuh"
#include <cuda.h>
#include "me_full_impl.cuh"     // fs_packed_smem_bytes: shared memory of one full-search job
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>
#include <new>
#include <mutex>
#include <unordered_map>
#include <vector>

int hmgpu_launch_chroma(hmgpu_ctx* ctx, int16_t* d_dst, const int16_t* d_src, int src_stride);
int hmgpu_launch_tz(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                    hmgpu_me_result* d_results, bool any_org_block, bool any_sel);
int hmgpu_launch_full(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                      hmgpu_me_result* d_results, bool any_org_block, int max_win_bytes);
int hmgpu_launch_frac(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                      hmgpu_me_result* d_results, bool any_frac);
int hmgpu_launch_frac_packed(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, hmgpu_me_result* d_results, bool any_frac);
int hmgpu_launch_frac_window(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, hmgpu_me_result* d_results, bool any_frac);
int hmgpu_launch_single(hmgpu_ctx* ctx, const HmgpuJobPack& pack, int n_jobs, const int16_t* d_org_blocks,
                        HmgpuMailSlot* d_slots, uint32_t ticket, bool any_org_block, int max_win_bytes,
                        unsigned long long* trace);

// mailbox of the low-latency path: up to MAIL_JOBS jobs per call
#define MAIL_JOBS HMGPU_MAIL_JOBS
int hmgpu_launch_server(hmgpu_ctx* ctx, cudaStream_t stream, const uint32_t* d_lines, const int16_t* d_org_blocks, HmgpuMailSlot* d_slots,
                        uint32_t* d_exited, uint32_t gen, uint32_t last_ticket, unsigned long long idle_ns, int n_ctas, int dyn_bytes);

// HMGPU_TRACE=1: where the time of a low-latency call goes (host clock around the launch / the poll, device
// globaltimer inside the kernel); printed by hmgpu_destroy
struct TraceAcc { double host_prep, host_launch, host_wait, host_copy, dev[5]; unsigned long long n; };
static TraceAcc g_trace;
static double now_us() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }

extern "C" { static int hmgpu_server_stop(hmgpu_ctx* ctx); }
extern "C" { static void server_write_lines(const hmgpu_ctx* ctx, Mailbox* mb, const hmgpu_me_job* jobs, int n_jobs, const hmgpu_pred_job* pjobs,
                               const uint8_t* pfuncs, int n_pred, uint32_t ticket, uint32_t gen); }
extern "C" { static int hmgpu_server_launch(hmgpu_ctx* ctx, uint32_t gen, uint32_t last_ticket, int dyn_bytes); }

static char g_create_err[512] = "";

int hmgpu_fail(hmgpu_ctx* ctx, int code, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
  va_end(ap);
  return code;
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- allocation pool (broker daemon) ------------------------------------------------------------------------------------
// cudaFree / cudaFreeHost synchronise the device: with the resident server kernels of other clients spinning they would not
// return before every one of them has gone idle.  In pool mode freed blocks are therefore parked and handed to the next request
// they fit (a daemon serves encoders of a few picture formats over and over, so the sizes repeat).
struct PoolState
{
  std::mutex mu;
  bool on = false;
  std::unordered_map<void*, size_t> size[2];               // live + parked blocks -> bytes; [0] device, [1] pinned host
  std::vector<std::pair<void*, size_t> > parked[2];
};
static PoolState& pool() { static PoolState* p = new PoolState; return *p; }

static cudaError_t pool_alloc(int kind, void** out, size_t bytes)
{
  PoolState& P = pool();
  {
    std::lock_guard<std::mutex> g(P.mu);
    if (P.on)
    {
      int best = -1;
      for (size_t i = 0; i < P.parked[kind].size(); i++)
      {
        const size_t b = P.parked[kind][i].second;
        if (b >= bytes && b <= bytes + bytes / 2 + 4096 && (best < 0 || b < P.parked[kind][best].second)) best = (int)i;
      }
      if (best >= 0)
      {
        *out = P.parked[kind][best].first;
        P.parked[kind].erase(P.parked[kind].begin() + best);
        return cudaSuccess;
      }
    }
  }
  const cudaError_t e = kind == 0 ? cudaMalloc(out, bytes) : cudaMallocHost(out, bytes);
  if (e == cudaSuccess) { std::lock_guard<std::mutex> g(P.mu); if (P.on) P.size[kind][*out] = bytes; }
  return e;
}

static void pool_free(int kind, void* p)
{
  if (!p) return;
  PoolState& P = pool();
  {
    std::lock_guard<std::mutex> g(P.mu);
    auto it = P.size[kind].find(p);
    if (P.on && it != P.size[kind].end()) { P.parked[kind].push_back(std::make_pair(p, it->second)); return; }
    if (it != P.size[kind].end()) P.size[kind].erase(it);
  }
  if (kind == 0) cudaFree(p); else cudaFreeHost(p);
}

cudaError_t hmgpu_dmalloc(void** p, size_t bytes) { return pool_alloc(0, p, bytes); }
void hmgpu_dfree(void* p) { pool_free(0, p); }
cudaError_t hmgpu_hmalloc(void** p, size_t bytes) { return pool_alloc(1, p, bytes); }
void hmgpu_hfree(void* p) { pool_free(1, p); }

static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }

// the environment is read here, once per context, and nowhere else
static void tuning_from_env(HmgpuTuning* t)
{
  t->tz_thread      = env_int("HMGPU_TZ_SPLIT", 5) != 0;
  t->tz_thread_min  = env_int("HMGPU_TZ_THREAD_MIN", 4096);
  t->tz_merge       = env_int("HMGPU_TZ_MERGE", 1);
  t->tz_carve       = env_int("HMGPU_TZ_CARVE", 50);
  t->tz_p2          = env_int("HMGPU_TZ_P2", 0);
  t->frac_v1        = env_int("HMGPU_FRAC_V1", 0);
  t->frac_overlap   = env_int("HMGPU_FRAC_OVERLAP", 1);
  t->frac_win       = env_int("HMGPU_FRAC_WIN", 1);
  t->frac_win_min   = env_int("HMGPU_FRAC_WIN_MIN", 4096);
  t->fs_tma         = env_int("HMGPU_FS_TMA", 1);
  t->rdoq_tu        = env_int("HMGPU_RDOQ_TU", 1);
  t->pipe_chunk     = env_int("HMGPU_PIPE_CHUNK", 0);
  t->pipe_edge      = env_int("HMGPU_PIPE_EDGE", 8);
  t->pipeline       = !env_int("HMGPU_NO_PIPELINE", 0);
  t->fastpath       = !env_int("HMGPU_NO_FASTPATH", 0);
  t->server         = env_int("HMGPU_SERVER", 1);
  t->server_idle_us = env_int("HMGPU_SERVER_IDLE_US", 200);
  t->trace          = getenv("HMGPU_TRACE") != NULL;
  t->server_stats   = getenv("HMGPU_SERVER_STATS") != NULL;
}

struct TuneName { const char* name; int HmgpuTuning::* field; };
static const TuneName k_tune_names[] = {
  { "tz_thread", &HmgpuTuning::tz_thread }, { "tz_thread_min", &HmgpuTuning::tz_thread_min }, { "tz_merge", &HmgpuTuning::tz_merge },
  { "tz_carve", &HmgpuTuning::tz_carve }, { "tz_p2", &HmgpuTuning::tz_p2 }, { "frac_v1", &HmgpuTuning::frac_v1 },
  { "frac_overlap", &HmgpuTuning::frac_overlap }, { "frac_win", &HmgpuTuning::frac_win }, { "frac_win_min", &HmgpuTuning::frac_win_min }, { "fs_tma", &HmgpuTuning::fs_tma }, { "pipe_chunk", &HmgpuTuning::pipe_chunk }, { "pipe_edge", &HmgpuTuning::pipe_edge }, { "pipeline", &HmgpuTuning::pipeline },
  { "fastpath", &HmgpuTuning::fastpath }, { "server", &HmgpuTuning::server }, { "server_idle_us", &HmgpuTuning::server_idle_us },
  { "trace", &HmgpuTuning::trace }, { "server_stats", &HmgpuTuning::server_stats }, { "rdoq_tu", &HmgpuTuning::rdoq_tu } };

// true when p is page-locked host memory known to CUDA (hmgpu_host_alloc or the caller's own cudaHostAlloc)
static bool is_pinned(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int hmgpu_reserve_pinned(hmgpu_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->h_pin_bytes) return HMGPU_OK;
  if (ctx->h_pin) hmgpu_hfree(ctx->h_pin);
  ctx->h_pin = NULL; ctx->h_pin_bytes = 0;
  bytes = round_up(bytes + bytes / 4, 1 << 20);
  HMGPU_CUDA(ctx, hmgpu_hmalloc((void**)&ctx->h_pin, bytes));
  ctx->h_pin_bytes = bytes;
  return HMGPU_OK;
}

int hmgpu_reserve_stage(hmgpu_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->d_stage_bytes) return HMGPU_OK;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_stage) hmgpu_dfree(ctx->d_stage);
  ctx->d_stage = NULL; ctx->d_stage_bytes = 0;
  bytes = round_up(bytes + bytes / 4, 1 << 20);
  HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&ctx->d_stage, bytes));
  ctx->d_stage_bytes = bytes;
  return HMGPU_OK;
}

int hmgpu_reserve_work(hmgpu_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->d_work_bytes) return HMGPU_OK;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_work) hmgpu_dfree(ctx->d_work);
  ctx->d_work = NULL; ctx->d_work_bytes = 0;
  bytes = round_up(bytes + bytes / 4, 1 << 20);
  HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&ctx->d_work, bytes));
  ctx->d_work_bytes = bytes;
  return HMGPU_OK;
}

int hmgpu_reserve_tzlist(hmgpu_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->d_tzlist_bytes) return HMGPU_OK;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_tzlist) hmgpu_dfree(ctx->d_tzlist);
  ctx->d_tzlist = NULL; ctx->d_tzlist_bytes = 0;
  bytes = round_up(bytes + bytes / 4, 1 << 20);
  HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&ctx->d_tzlist, bytes));
  ctx->d_tzlist_bytes = bytes;
  return HMGPU_OK;
}

void hmgpu_use_lane(hmgpu_ctx* ctx, int lane)
{
  if (lane == ctx->cur_lane) return;
  HmgpuLane& cur = ctx->lane_store[ctx->cur_lane];
  cur.stream = ctx->stream;
  cur.h_pin = ctx->h_pin; cur.h_pin_bytes = ctx->h_pin_bytes;
  cur.d_stage = ctx->d_stage; cur.d_stage_bytes = ctx->d_stage_bytes;
  cur.d_work = ctx->d_work; cur.d_work_bytes = ctx->d_work_bytes;
  cur.d_tzlist = ctx->d_tzlist; cur.d_tzlist_bytes = ctx->d_tzlist_bytes;
  const HmgpuLane& nx = ctx->lane_store[lane];
  ctx->stream = nx.stream;
  ctx->h_pin = nx.h_pin; ctx->h_pin_bytes = nx.h_pin_bytes;
  ctx->d_stage = nx.d_stage; ctx->d_stage_bytes = nx.d_stage_bytes;
  ctx->d_work = nx.d_work; ctx->d_work_bytes = nx.d_work_bytes;
  ctx->d_tzlist = nx.d_tzlist; ctx->d_tzlist_bytes = nx.d_tzlist_bytes;
  ctx->cur_lane = lane;
}

RefTable hmgpu_ref_table(const hmgpu_ctx* ctx)
{
  RefTable t;
  memset(&t, 0, sizeof t);
  for (int i = 0; i < ctx->max_refs; i++)
    if (ctx->refs[i].valid)
      t.base[i] = (const char*)ctx->refs[i].planes + ((size_t)HMGPU_MARGIN * ctx->pitch + HMGPU_MARGIN) * ctx->px_bytes;
  t.plane_elems = (int64_t)ctx->plane_elems;
  t.pitch = ctx->pitch;
  t.pic_w = ctx->pic_w; t.pic_h = ctx->pic_h;
  t.bit_depth = ctx->bit_depth;
  return t;
}

int hmgpu_launch_me(hmgpu_ctx* ctx, const hmgpu_me_job* d_jobs, int n_jobs, const int16_t* d_org_blocks,
                    hmgpu_me_result* d_results, bool any_org_block, bool any_full, bool any_tz, bool any_frac,
                    int max_win_bytes, bool any_sel)
{
  int rc;
  if (any_tz && (rc = hmgpu_launch_tz(ctx, d_jobs, n_jobs, d_org_blocks, d_results, any_org_block, any_sel))) return rc;
  if (any_full && (rc = hmgpu_launch_full(ctx, d_jobs, n_jobs, d_org_blocks, d_results, any_org_block, max_win_bytes))) return rc;
  if (ctx->px_bytes == 1 && !any_org_block && !ctx->tune.frac_v1)
  {
    if (ctx->tune.frac_win && n_jobs >= ctx->tune.frac_win_min && ctx->planes_all)
    {
      if ((rc = hmgpu_launch_frac_window(ctx, d_jobs, n_jobs, d_results, any_frac))) return rc;
    }
    else if ((rc = hmgpu_launch_frac_packed(ctx, d_jobs, n_jobs, d_results, any_frac))) return rc;
  }
  else if ((rc = hmgpu_launch_frac(ctx, d_jobs, n_jobs, d_org_blocks, d_results, any_frac))) return rc;
  return HMGPU_OK;
}

// entry points a broker client does not have (test / measurement surface, device pointers): refused, never emulated
#define HMGPU_NOT_REMOTE(ctx, what) \
  do { if ((ctx)->remote) return hmgpu_fail((ctx), HMGPU_E_STATE, "%s is not available through the broker daemon (HMGPU_BROKER is set)", what); } while (0)

extern "C" {

int hmgpu_abi_version(void) { return HMGPU_ABI_VERSION; }

static_assert(sizeof(hmgpu_me_job) == 48 && sizeof(hmgpu_me_result) == 24 && sizeof(hmgpu_dist_item) == 20 &&
              sizeof(hmgpu_mc_job) == 16, "ABI struct layout changed");
static_assert(sizeof(hmgpu_pred_job) == 20, "ABI struct layout changed");
void hmgpu_struct_sizes(int out[5])
{
  out[0] = (int)sizeof(hmgpu_me_job); out[1] = (int)sizeof(hmgpu_me_result);
  out[2] = (int)sizeof(hmgpu_dist_item); out[3] = (int)sizeof(hmgpu_mc_job);
  out[4] = (int)sizeof(hmgpu_pred_job);
}

const char* hmgpu_last_error(const hmgpu_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

uint64_t hmgpu_launch_count(const hmgpu_ctx* ctx)
{
  if (!ctx) return 0;
  if (ctx->remote)
  {
    BrokerReply r;
    return hmgpu_remote_call((hmgpu_ctx*)ctx, HMB_OP_LAUNCH_COUNT, NULL, NULL, &r) == HMGPU_OK ? r.v64 : 0;
  }
  return ctx->launches;
}

void* hmgpu_stream(const hmgpu_ctx* ctx) { return ctx && !ctx->remote ? (void*)ctx->stream : NULL; }

int hmgpu_host_alloc(hmgpu_ctx* ctx, size_t bytes, void** out)
{
  if (!ctx || !out) return HMGPU_E_INVALID;
  HMGPU_NOT_REMOTE(ctx, "hmgpu_host_alloc");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  HMGPU_CUDA(ctx, cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return HMGPU_OK;
}

int hmgpu_host_free(hmgpu_ctx* ctx, void* p)
{
  if (!ctx) return HMGPU_E_INVALID;
  HMGPU_NOT_REMOTE(ctx, "hmgpu_host_free");
  HMGPU_CUDA(ctx, cudaFreeHost(p));
  return HMGPU_OK;
}

int hmgpu_set_option(hmgpu_ctx* ctx, const char* name, int value)
{
  if (!ctx || !name) return HMGPU_E_INVALID;
  for (size_t i = 0; i < sizeof k_tune_names / sizeof k_tune_names[0]; i++)
    if (!strcmp(name, k_tune_names[i].name))
    {
      if (&(ctx->tune.*k_tune_names[i].field) == &ctx->tune.server && !value) { const int rc = hmgpu_server_stop(ctx); if (rc) return rc; }
      ctx->tune.*k_tune_names[i].field = value;
      // the kernel-mapping switches of a broker client live in the daemon's context
      if (ctx->remote) { const int32_t a[6] = { value, 0, 0, 0, 0, 0 }; return hmgpu_remote_call(ctx, HMB_OP_SET_OPTION, a, name, NULL); }
      return HMGPU_OK;
    }
  return hmgpu_fail(ctx, HMGPU_E_INVALID, "unknown option '%s'", name);
}

int hmgpu_synchronize(hmgpu_ctx* ctx)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (ctx->remote) return HMGPU_OK;                        // every remote operation is complete when its call returns
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HMGPU_OK;
}

static int create_local(int device, int pic_w, int pic_h, int bit_depth, int max_refs, void* ext_mail, int srv_ctas, hmgpu_ctx** out);

int hmgpu_create(int device, int pic_w, int pic_h, int bit_depth, int max_refs, hmgpu_ctx** out)
{
  // HMGPU_BROKER=<socket path>: this process becomes a client of the per-GPU daemon (hmgpud) and never creates a CUDA context;
  // `device` is then the daemon's business
  const char* broker = getenv("HMGPU_BROKER");
  if (broker && *broker)
  {
    if (!out) return hmgpu_fail(NULL, HMGPU_E_INVALID, "out is NULL");
    *out = NULL;
    return hmgpu_remote_create(broker, pic_w, pic_h, bit_depth, max_refs, out);
  }
  return create_local(device, pic_w, pic_h, bit_depth, max_refs, NULL, 0, out);
}

int hmgpu_internal_create_shared(int device, int pic_w, int pic_h, int bit_depth, int max_refs, void* mail, int srv_ctas, hmgpu_ctx** out)
{
  return create_local(device, pic_w, pic_h, bit_depth, max_refs, mail, srv_ctas, out);
}

void hmgpu_internal_pool_mode(int on) { std::lock_guard<std::mutex> g(pool().mu); pool().on = on != 0; }
int  hmgpu_internal_server_launch(hmgpu_ctx* ctx, uint32_t gen, uint32_t last_ticket, int dyn_bytes)
{
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  return hmgpu_server_launch(ctx, gen, last_ticket, dyn_bytes);
}
int  hmgpu_internal_server_sync(hmgpu_ctx* ctx)
{
  if (ctx->srv_stream) HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->srv_stream));
  return HMGPU_OK;
}
int  hmgpu_internal_server_query(hmgpu_ctx* ctx)
{
  if (!ctx->srv_stream) return HMGPU_OK;
  const cudaError_t e = cudaStreamQuery(ctx->srv_stream);
  if (e != cudaErrorNotReady && e != cudaSuccess) return hmgpu_fail(ctx, HMGPU_E_CUDA, "mailbox server failed: %s", cudaGetErrorString(e));
  return HMGPU_OK;
}
size_t hmgpu_internal_mailbox_bytes(void) { return sizeof(Mailbox); }
// the daemon's last word to a client's server kernel, whatever the client's state was: a generation no kernel runs as
int  hmgpu_internal_server_kill(hmgpu_ctx* ctx)
{
  if (!ctx->h_mail || !ctx->srv_stream) return HMGPU_OK;
  const int keep = ctx->srv_ctas;
  ctx->srv_ctas = HMGPU_SERVER_CTAS;
  server_write_lines(ctx, (Mailbox*)ctx->h_mail, NULL, 0, NULL, NULL, 0, 0u, 0xffffffffu);
  ctx->srv_ctas = keep;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->srv_stream));
  return HMGPU_OK;
}

static int create_local(int device, int pic_w, int pic_h, int bit_depth, int max_refs, void* ext_mail, int srv_ctas, hmgpu_ctx** out)
{
  if (!out) return hmgpu_fail(NULL, HMGPU_E_INVALID, "out is NULL");
  *out = NULL;
  if (pic_w < 8 || pic_h < 8 || (pic_w & 3) || (pic_h & 3) || pic_w > 8184 || pic_h > 8184)
    return hmgpu_fail(NULL, HMGPU_E_INVALID, "picture size %dx%d unsupported (multiples of 4, 8..8184: quarter-pel clipMv bounds are int16)", pic_w, pic_h);
  if (bit_depth < 8 || bit_depth > 12) return hmgpu_fail(NULL, HMGPU_E_INVALID, "bit depth %d unsupported (8..12)", bit_depth);
  if (max_refs < 1 || max_refs > HMGPU_MAX_REFS) return hmgpu_fail(NULL, HMGPU_E_INVALID, "max_refs %d not in 1..%d", max_refs, HMGPU_MAX_REFS);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return hmgpu_fail(NULL, HMGPU_E_CUDA, "no CUDA device (%s); libhmgpu has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return hmgpu_fail(NULL, HMGPU_E_INVALID, "device %d not in 0..%d", device, n_dev - 1);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return hmgpu_fail(NULL, HMGPU_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

  hmgpu_ctx* ctx = new (std::nothrow) hmgpu_ctx;
  if (!ctx) return hmgpu_fail(NULL, HMGPU_E_NOMEM, "out of host memory");
  memset(ctx, 0, sizeof *ctx);
  ctx->device = device; ctx->pic_w = pic_w; ctx->pic_h = pic_h; ctx->bit_depth = bit_depth; ctx->max_refs = max_refs;
  tuning_from_env(&ctx->tune);
  ctx->srv_ctas = srv_ctas > 0 ? (srv_ctas > HMGPU_SERVER_CTAS ? HMGPU_SERVER_CTAS : srv_ctas) : env_int("HMGPU_SERVER_CTAS", HMGPU_SERVER_CTAS);
  if (ctx->srv_ctas < 1 || ctx->srv_ctas > HMGPU_SERVER_CTAS) ctx->srv_ctas = HMGPU_SERVER_CTAS;
  ctx->px_bytes = bit_depth == 8 ? 1 : 2;
  ctx->pw = pic_w + 2 * HMGPU_MARGIN; ctx->ph = pic_h + 2 * HMGPU_MARGIN;
  ctx->pitch = (int)round_up(ctx->pw + 32, 64);          // slack for aligned look-ahead loads
  ctx->plane_elems = (size_t)ctx->pitch * ctx->ph;
  ctx->cpw = pic_w / 2 + 2 * HMGPU_CMARGIN; ctx->cph = pic_h / 2 + 2 * HMGPU_CMARGIN;
  ctx->cpitch = (int)round_up(ctx->cpw, 32);
  ctx->org_pitch = (int)round_up(pic_w + 32, 64);
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete ctx; return hmgpu_fail(NULL, HMGPU_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  e = hmgpu_dmalloc((void**)&ctx->d_org, (size_t)ctx->org_pitch * (pic_h + 1) * ctx->px_bytes + 256);
  if (e != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return hmgpu_fail(NULL, HMGPU_E_NOMEM, "cudaMalloc org: %s", cudaGetErrorString(e)); }
  cudaMemsetAsync(ctx->d_org, 0, (size_t)ctx->org_pitch * (pic_h + 1) * ctx->px_bytes + 256, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  if (ext_mail)
  {
    // the daemon's view of a client: the mailbox is part of the shared segment (already registered with CUDA)
    ctx->h_mail = ext_mail; ctx->mail_external = true;
    e = cudaHostGetDevicePointer(&ctx->d_mail, ext_mail, 0);
    if (e != cudaSuccess) { hmgpu_dfree(ctx->d_org); cudaStreamDestroy(ctx->stream); delete ctx; return hmgpu_fail(NULL, HMGPU_E_CUDA, "cudaHostGetDevicePointer(mailbox): %s", cudaGetErrorString(e)); }
    ctx->tune.fastpath = 0;                               // batches of any size go to the batch kernels: the mailbox belongs to the client
  }
  *out = ctx;
  return HMGPU_OK;
}

void hmgpu_destroy(hmgpu_ctx* ctx)
{
  if (!ctx) return;
  if (ctx->remote)
  {
    hmgpu_server_stop(ctx);
    if (ctx->tune.trace || ctx->tune.server_stats)
      fprintf(stderr, "[hmgpu server] %u calls served by %u server generations (through the broker)\n", ctx->srv_calls, ctx->srv_starts);
    free(ctx->defer_org);
    hmgpu_remote_destroy(ctx);
    return;
  }
  if (g_trace.n)
  {
    const double n = (double)g_trace.n;
    fprintf(stderr, "[hmgpu trace] %llu low-latency calls, us/call: host prep %.2f, launch %.2f, wait %.2f, copy-out %.2f | device (block 0): "
                    "job fetch %.2f, integer %.2f, half %.2f, quarter %.2f, publish %.2f\n", g_trace.n, g_trace.host_prep / n, g_trace.host_launch / n,
            g_trace.host_wait / n, g_trace.host_copy / n, g_trace.dev[0] / n, g_trace.dev[1] / n, g_trace.dev[2] / n, g_trace.dev[3] / n, g_trace.dev[4] / n);
    memset(&g_trace, 0, sizeof g_trace);
  }
  cudaSetDevice(ctx->device);
  hmgpu_server_stop(ctx);
  if (ctx->srv_stream) { cudaStreamSynchronize(ctx->srv_stream); cudaStreamDestroy(ctx->srv_stream); }
  if (ctx->tune.trace || ctx->tune.server_stats)
    fprintf(stderr, "[hmgpu server] %u calls served by %u server generations\n", ctx->srv_calls, ctx->srv_starts);
  hmgpu_use_lane(ctx, 0);
  cudaStreamSynchronize(ctx->stream);
  {
    HmgpuLane& l1 = ctx->lane_store[1];
    if (l1.stream) cudaStreamSynchronize(l1.stream);
    if (l1.d_stage) hmgpu_dfree(l1.d_stage);
    if (l1.d_work) hmgpu_dfree(l1.d_work);
    if (l1.d_tzlist) hmgpu_dfree(l1.d_tzlist);
    if (l1.h_pin) hmgpu_hfree(l1.h_pin);
    if (l1.stream) cudaStreamDestroy(l1.stream);
    for (int i = 0; i < 2; i++) if (ctx->lane_done[i]) cudaEventDestroy(ctx->lane_done[i]);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    for (int i = 0; i < 2; i++) if (ctx->scan_done[i]) cudaEventDestroy(ctx->scan_done[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->d_scan) hmgpu_dfree(ctx->d_scan);
    if (ctx->h_scan) hmgpu_hfree(ctx->h_scan);
    if (ctx->d_orgblk) hmgpu_dfree(ctx->d_orgblk);
  }
  for (int i = 0; i < HMGPU_MAX_REFS; i++)
  {
    if (ctx->refs[i].cb) hmgpu_dfree(ctx->refs[i].cb);
    if (ctx->refs[i].cr) hmgpu_dfree(ctx->refs[i].cr);
  }
  if (ctx->frac_stream) { cudaStreamSynchronize(ctx->frac_stream); cudaStreamDestroy(ctx->frac_stream); }
  for (int i = 0; i < 2; i++) if (ctx->frac_ev[i]) cudaEventDestroy(ctx->frac_ev[i]);
  for (int i = 0; i < HMGPU_TZ_STREAMS; i++) if (ctx->tz_streams[i]) { cudaStreamSynchronize(ctx->tz_streams[i]); cudaStreamDestroy(ctx->tz_streams[i]); }
  for (int i = 0; i <= HMGPU_TZ_STREAMS; i++) if (ctx->tz_ev[i]) cudaEventDestroy(ctx->tz_ev[i]);
  if (ctx->d_org) hmgpu_dfree(ctx->d_org);
  if (ctx->planes_all) hmgpu_dfree(ctx->planes_all);
  free(ctx->h_tmaps);
  free(ctx->h_fw_tmap);
  if (ctx->d_stage) hmgpu_dfree(ctx->d_stage);
  if (ctx->d_work) hmgpu_dfree(ctx->d_work);
  if (ctx->d_tzlist) hmgpu_dfree(ctx->d_tzlist);
  if (ctx->h_pin) hmgpu_hfree(ctx->h_pin);
  if (ctx->h_mail && !ctx->mail_external) cudaFreeHost(ctx->h_mail);
  free(ctx->defer_org);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int ref_alloc(hmgpu_ctx* ctx, int slot, bool chroma)
{
  RefSlot& s = ctx->refs[slot];
  if (!ctx->planes_all)
  {
    // the phase planes of ALL reference slots are one allocation: a slot is then a coordinate of the TMA tensor maps
    // (me_frac3.cu: x, y, phase plane, slot), which are built once and travel as a kernel parameter
    ctx->slot_bytes = (ctx->plane_elems * 16 * ctx->px_bytes + 512 + 255) & ~(size_t)255;
    HMGPU_CUDA(ctx, hmgpu_dmalloc(&ctx->planes_all, ctx->slot_bytes * ctx->max_refs));
    HMGPU_CUDA(ctx, cudaMemsetAsync(ctx->planes_all, 0, ctx->slot_bytes * ctx->max_refs, ctx->stream));
  }
  if (!s.planes) s.planes = (char*)ctx->planes_all + (size_t)slot * ctx->slot_bytes;
  if (chroma && !s.cb)
  {
    const size_t bytes = (size_t)ctx->cpitch * ctx->cph * sizeof(int16_t);
    HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&s.cb, bytes));
    HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&s.cr, bytes));
  }
  return HMGPU_OK;
}

// copy a strided host plane into pinned memory (tight), then to the device staging buffer
static int stage_plane(hmgpu_ctx* ctx, const int16_t* src, int stride, int w, int h, size_t pin_off, size_t dev_off)
{
  if (stride == w && is_pinned(src))
  {
    HMGPU_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_stage + dev_off, src, sizeof(int16_t) * (size_t)w * h, cudaMemcpyHostToDevice, ctx->stream));
    return HMGPU_OK;
  }
  int16_t* pin = (int16_t*)((char*)ctx->h_pin + pin_off);
  for (int y = 0; y < h; y++) memcpy(pin + (size_t)y * w, src + (size_t)y * stride, sizeof(int16_t) * w);
  HMGPU_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_stage + dev_off, pin, sizeof(int16_t) * (size_t)w * h, cudaMemcpyHostToDevice, ctx->stream));
  return HMGPU_OK;
}

// broker client: the picture travels through the upload area of the shared segment (tight rows), one round trip
static int remote_upload(hmgpu_ctx* ctx, int op, int slot, const int16_t* luma, int luma_stride, const int16_t* cb, const int16_t* cr, int chroma_stride)
{
  int rc = hmgpu_server_stop(ctx);                         // the server reads the planes through the read-only path
  if (rc) return rc;
  size_t cap = 0;
  int16_t* area = (int16_t*)hmgpu_remote_area(ctx, 0, &cap);
  const size_t yel = (size_t)ctx->pic_w * ctx->pic_h, cel = (size_t)(ctx->pic_w / 2) * (ctx->pic_h / 2);
  const bool chroma = cb && cr;
  if ((yel + (chroma ? 2 * cel : 0)) * sizeof(int16_t) > cap) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "picture does not fit the broker's upload area");
  for (int y = 0; y < ctx->pic_h; y++) memcpy(area + (size_t)y * ctx->pic_w, luma + (size_t)y * luma_stride, sizeof(int16_t) * ctx->pic_w);
  if (chroma)
    for (int y = 0; y < ctx->pic_h / 2; y++)
    {
      memcpy(area + yel + (size_t)y * (ctx->pic_w / 2), cb + (size_t)y * chroma_stride, sizeof(int16_t) * (ctx->pic_w / 2));
      memcpy(area + yel + cel + (size_t)y * (ctx->pic_w / 2), cr + (size_t)y * chroma_stride, sizeof(int16_t) * (ctx->pic_w / 2));
    }
  const int32_t a[6] = { slot, chroma ? 1 : 0, 0, 0, 0, 0 };
  if ((rc = hmgpu_remote_call(ctx, op, a, NULL, NULL))) return rc;
  if (op == HMB_OP_REF_UPLOAD) { ctx->refs[slot].valid = true; ctx->refs[slot].has_chroma = chroma; }
  return HMGPU_OK;
}

int hmgpu_ref_upload(hmgpu_ctx* ctx, int slot, const int16_t* luma, int luma_stride,
                     const int16_t* cb, const int16_t* cr, int chroma_stride)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (slot < 0 || slot >= ctx->max_refs || !luma) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad slot %d or NULL luma", slot);
  if (ctx->remote) return remote_upload(ctx, HMB_OP_REF_UPLOAD, slot, luma, luma_stride, cb, cr, chroma_stride);
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  { const int rcs = hmgpu_server_stop(ctx); if (rcs) return rcs; }   // the server reads the planes through the read-only path
  const bool chroma = cb && cr;
  int rc = ref_alloc(ctx, slot, chroma);
  if (rc) return rc;
  const size_t ybytes = sizeof(int16_t) * (size_t)ctx->pic_w * ctx->pic_h;
  const size_t cbytes = sizeof(int16_t) * (size_t)(ctx->pic_w / 2) * (ctx->pic_h / 2);
  const size_t total = ybytes + (chroma ? 2 * cbytes : 0);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the pinned buffer is reused
  if ((rc = hmgpu_reserve_pinned(ctx, total))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, total))) return rc;
  if ((rc = stage_plane(ctx, luma, luma_stride, ctx->pic_w, ctx->pic_h, 0, 0))) return rc;
  if ((rc = hmgpu_launch_planes(ctx, slot, (const int16_t*)ctx->d_stage, ctx->pic_w))) return rc;
  if (chroma)
  {
    if ((rc = stage_plane(ctx, cb, chroma_stride, ctx->pic_w / 2, ctx->pic_h / 2, ybytes, ybytes))) return rc;
    if ((rc = stage_plane(ctx, cr, chroma_stride, ctx->pic_w / 2, ctx->pic_h / 2, ybytes + cbytes, ybytes + cbytes))) return rc;
    if ((rc = hmgpu_launch_chroma(ctx, (int16_t*)ctx->refs[slot].cb, (const int16_t*)((char*)ctx->d_stage + ybytes), ctx->pic_w / 2))) return rc;
    if ((rc = hmgpu_launch_chroma(ctx, (int16_t*)ctx->refs[slot].cr, (const int16_t*)((char*)ctx->d_stage + ybytes + cbytes), ctx->pic_w / 2))) return rc;
  }
  ctx->refs[slot].valid = true;
  ctx->refs[slot].has_chroma = chroma;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HMGPU_OK;
}

int hmgpu_ref_upload_device(hmgpu_ctx* ctx, int slot, const void* d_luma, int luma_stride)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (slot < 0 || slot >= ctx->max_refs || !d_luma) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad slot %d or NULL luma", slot);
  HMGPU_NOT_REMOTE(ctx, "hmgpu_ref_upload_device");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  { const int rcs = hmgpu_server_stop(ctx); if (rcs) return rcs; }   // the server reads the planes through the read-only path
  int rc = ref_alloc(ctx, slot, false);
  if (rc) return rc;
  if ((rc = hmgpu_launch_planes(ctx, slot, (const int16_t*)d_luma, luma_stride))) return rc;
  ctx->refs[slot].valid = true;
  ctx->refs[slot].has_chroma = false;
  return HMGPU_OK;
}

int hmgpu_ref_release(hmgpu_ctx* ctx, int slot)
{
  if (!ctx || slot < 0 || slot >= HMGPU_MAX_REFS) return HMGPU_E_INVALID;
  if (ctx->remote) { const int32_t a[6] = { slot, 0, 0, 0, 0, 0 }; const int rc = hmgpu_remote_call(ctx, HMB_OP_REF_RELEASE, a, NULL, NULL); if (rc) return rc; }
  ctx->refs[slot].valid = false;   // memory is kept for the next picture that takes the slot
  return HMGPU_OK;
}

int hmgpu_ref_download_plane(hmgpu_ctx* ctx, int slot, int frac_x, int frac_y, int16_t* dst)
{
  if (!ctx || !dst) return HMGPU_E_INVALID;
  HMGPU_NOT_REMOTE(ctx, "hmgpu_ref_download_plane");
  if (slot < 0 || slot >= ctx->max_refs || !ctx->refs[slot].valid) return hmgpu_fail(ctx, HMGPU_E_STATE, "slot %d holds no reference", slot);
  if (frac_x < 0 || frac_x > 3 || frac_y < 0 || frac_y > 3) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad phase");
  const size_t bytes = ctx->plane_elems * ctx->px_bytes;
  int rc = hmgpu_reserve_pinned(ctx, bytes);
  if (rc) return rc;
  const char* src = (const char*)ctx->refs[slot].planes + (size_t)(frac_y * 4 + frac_x) * bytes;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int y = 0; y < ctx->ph; y++)
    for (int x = 0; x < ctx->pw; x++)
      dst[(size_t)y * ctx->pw + x] = ctx->px_bytes == 1 ? (int16_t)((const uint8_t*)ctx->h_pin)[(size_t)y * ctx->pitch + x]
                                                        : (int16_t)((const uint16_t*)ctx->h_pin)[(size_t)y * ctx->pitch + x];
  return HMGPU_OK;
}

int hmgpu_org_upload(hmgpu_ctx* ctx, const int16_t* luma, int luma_stride)
{
  if (!ctx || !luma) return HMGPU_E_INVALID;
  if (ctx->remote) return remote_upload(ctx, HMB_OP_ORG_UPLOAD, 0, luma, luma_stride, NULL, NULL, 0);
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  { const int rcs = hmgpu_server_stop(ctx); if (rcs) return rcs; }   // the server reads the planes through the read-only path
  const size_t ybytes = sizeof(int16_t) * (size_t)ctx->pic_w * ctx->pic_h;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, ybytes))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, ybytes))) return rc;
  if ((rc = stage_plane(ctx, luma, luma_stride, ctx->pic_w, ctx->pic_h, 0, 0))) return rc;
  if ((rc = hmgpu_launch_org(ctx, (const int16_t*)ctx->d_stage, ctx->pic_w))) return rc;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HMGPU_OK;
}

int hmgpu_org_upload_device(hmgpu_ctx* ctx, const void* d_luma, int luma_stride)
{
  if (!ctx || !d_luma) return HMGPU_E_INVALID;
  HMGPU_NOT_REMOTE(ctx, "hmgpu_org_upload_device");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  { const int rcs = hmgpu_server_stop(ctx); if (rcs) return rcs; }   // the server reads the planes through the read-only path
  return hmgpu_launch_org(ctx, (const int16_t*)d_luma, luma_stride);
}

// ---- motion search --------------------------------------------------------------------------

// PU shapes the search kernels take: sides multiples of 4 in 4..64 and at most 64 SATD tiles (8x8 tiles iff both sides are
// multiples of 8, else 4x4: TComRdCost.cpp:1555-1597) -- the fractional stage packs (job, tile) with 6 tile bits.  Every HEVC PU
// shape qualifies (4x4-tiled HEVC shapes have at most 12x16/16 = 12 tiles); 36x32 or 60x64 do not.
__host__ __device__ static inline bool hmgpu_pu_shape_ok(int w, int h)
{
  if (w < 4 || w > 64 || h < 4 || h > 64 || (w & 3) || (h & 3)) return false;
  const int ts = ((w & 7) == 0 && (h & 7) == 0) ? 8 : 4;
  return (w / ts) * (h / ts) <= 64;
}

static int validate_jobs(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n, int n_org_elems,
                         bool* any_org, bool* any_full, bool* any_tz, bool* any_frac, int* max_win_bytes, int first = 0,
                         bool* any_sel = NULL)
{
  *any_org = *any_full = *any_tz = *any_frac = false;
  if (any_sel) *any_sel = false;
  *max_win_bytes = 0;
  jobs -= first;                                           // messages carry the index in the caller's array
  for (int i = first; i < first + n; i++)
  {
    const hmgpu_me_job& j = jobs[i];
    if (!hmgpu_pu_shape_ok(j.pu_w, j.pu_h))
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: PU %dx%d unsupported", i, j.pu_w, j.pu_h);
    if (j.pu_x < 0 || j.pu_y < 0 || j.pu_x + j.pu_w > ctx->pic_w || j.pu_y + j.pu_h > ctx->pic_h || (j.pu_x & 3))
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: PU at (%d,%d) outside the picture or x not a multiple of 4", i, j.pu_x, j.pu_y);
    if (j.ref_slot >= ctx->max_refs || !ctx->refs[j.ref_slot].valid)
      return hmgpu_fail(ctx, HMGPU_E_STATE, "job %d: reference slot %d not uploaded", i, j.ref_slot);
    // every sample the search can touch must lie inside the 80-sample padding
    const int lo_x = j.pu_x + (j.clip_hmin >> 2) - 4, hi_x = j.pu_x + j.pu_w + (j.clip_hmax >> 2) + 4;
    const int lo_y = j.pu_y + (j.clip_vmin >> 2) - 4, hi_y = j.pu_y + j.pu_h + (j.clip_vmax >> 2) + 4;
    if (lo_x < -HMGPU_MARGIN || lo_y < -HMGPU_MARGIN || hi_x > ctx->pic_w + HMGPU_MARGIN || hi_y > ctx->pic_h + HMGPU_MARGIN)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: clip bounds reach outside the padded reference", i);
    if ((j.win_l << 2) < j.clip_hmin - 3 || (j.win_r << 2) > j.clip_hmax || (j.win_t << 2) < j.clip_vmin - 3 || (j.win_b << 2) > j.clip_vmax)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: search window outside the clip bounds", i);
    if (!(j.flags & HMGPU_F_INTEGER))
    {
      if ((j.start_x << 2) < j.clip_hmin - 3 || (j.start_x << 2) > j.clip_hmax || (j.start_y << 2) < j.clip_vmin - 3 || (j.start_y << 2) > j.clip_vmax)
        return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: integer MV outside the clip bounds", i);
    }
    if (j.flags & HMGPU_F_ORG_BLOCK)
    {
      *any_org = true;
      if ((size_t)j.org_offset + (size_t)j.pu_w * j.pu_h > (size_t)n_org_elems)
        return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: org block outside org_blocks", i);
    }
    if (j.kind != HMGPU_KIND_DEFAULT)
    {
      if (j.kind != HMGPU_KIND_SELECTIVE || !(j.flags & HMGPU_F_INTEGER) || (j.flags & (HMGPU_F_FULL | HMGPU_F_ORG_BLOCK)))
        return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: kind %d needs an integer TZ search without a key-pattern block", i, j.kind);
      if (!any_sel) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: selective search is not available on this entry point", i);
      if ((size_t)j.org_offset + 6 > (size_t)n_org_elems)
        return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: the three MV predictors of the selective search lie outside org_blocks", i);
      *any_sel = true;
    }
    if (j.flags & HMGPU_F_FRAC) *any_frac = true;
    if (j.flags & HMGPU_F_INTEGER)
    {
      if (j.flags & HMGPU_F_FULL)
      {
        *any_full = true;
        const int nx = j.win_r - j.win_l + 1, ny = j.win_b - j.win_t + 1;
        if (nx > 0 && ny > 0)
        {
          const int bytes = fs_packed_smem_bytes(j.pu_w, j.pu_h, nx, ny, (j.flags & HMGPU_F_FEN) != 0);
          if (bytes > *max_win_bytes) *max_win_bytes = bytes;
        }
      }
      else
      {
        *any_tz = true;
        if (j.search_range < 1 || j.search_range > 512) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: search range %d", i, j.search_range);
      }
    }
  }
  return HMGPU_OK;
}

// ---- large batches -----------------------------------------------------------------------------------
// Chunks alternate between the two lanes: the H2D copy of chunk k+1 and the D2H copy of chunk k-1
// overlap the search kernels of chunk k.  The per-job range checks of validate_jobs() cost more host
// time per job than the device needs for the whole search (measured: ~10 ns/job on the host, 5 ns/job
// on the B200), so for these batches the same checks run as a kernel on the freshly copied chunk
// (job_scan_kernel, on the copy stream) and the host only reads its 16-byte verdict.
#define PIPE_MIN_JOBS 65536

struct ScanOut { uint32_t first_bad; uint32_t flags_any; uint32_t max_win; uint32_t int_kinds; /* bit0 tz, bit1 full, bit2 selective */ };

enum { SCAN_OK = 0, SCAN_PU_SIZE, SCAN_PU_POS, SCAN_SLOT, SCAN_CLIP, SCAN_WINDOW, SCAN_START, SCAN_ORG, SCAN_RANGE, SCAN_KIND };
static const char* const k_scan_msg[] = {
  "ok", "PU size unsupported", "PU outside the picture or x not a multiple of 4", "reference slot not uploaded",
  "clip bounds reach outside the padded reference", "search window outside the clip bounds", "integer MV outside the clip bounds",
  "org block outside org_blocks", "search range not in 1..512",
  "job kind needs an integer TZ search without a key-pattern block and its MV predictors inside org_blocks" };

__host__ __device__ static inline int job_check(const hmgpu_me_job& j, int pic_w, int pic_h, uint32_t valid_slots, unsigned long long n_org_elems)
{
  if (!hmgpu_pu_shape_ok(j.pu_w, j.pu_h)) return SCAN_PU_SIZE;
  if (j.pu_x < 0 || j.pu_y < 0 || j.pu_x + j.pu_w > pic_w || j.pu_y + j.pu_h > pic_h || (j.pu_x & 3)) return SCAN_PU_POS;
  if (j.ref_slot >= 32 || !((valid_slots >> j.ref_slot) & 1u)) return SCAN_SLOT;
  const int lo_x = j.pu_x + (j.clip_hmin >> 2) - 4, hi_x = j.pu_x + j.pu_w + (j.clip_hmax >> 2) + 4;
  const int lo_y = j.pu_y + (j.clip_vmin >> 2) - 4, hi_y = j.pu_y + j.pu_h + (j.clip_vmax >> 2) + 4;
  if (lo_x < -HMGPU_MARGIN || lo_y < -HMGPU_MARGIN || hi_x > pic_w + HMGPU_MARGIN || hi_y > pic_h + HMGPU_MARGIN) return SCAN_CLIP;
  if ((j.win_l << 2) < j.clip_hmin - 3 || (j.win_r << 2) > j.clip_hmax || (j.win_t << 2) < j.clip_vmin - 3 || (j.win_b << 2) > j.clip_vmax) return SCAN_WINDOW;
  if (!(j.flags & HMGPU_F_INTEGER) &&
      ((j.start_x << 2) < j.clip_hmin - 3 || (j.start_x << 2) > j.clip_hmax || (j.start_y << 2) < j.clip_vmin - 3 || (j.start_y << 2) > j.clip_vmax)) return SCAN_START;
  if ((j.flags & HMGPU_F_ORG_BLOCK) && (unsigned long long)j.org_offset + (unsigned long long)j.pu_w * j.pu_h > n_org_elems) return SCAN_ORG;
  if ((j.flags & HMGPU_F_INTEGER) && !(j.flags & HMGPU_F_FULL) && (j.search_range < 1 || j.search_range > 512)) return SCAN_RANGE;
  if (j.kind != HMGPU_KIND_DEFAULT &&
      (j.kind != HMGPU_KIND_SELECTIVE || !(j.flags & HMGPU_F_INTEGER) || (j.flags & (HMGPU_F_FULL | HMGPU_F_ORG_BLOCK)) ||
       (unsigned long long)j.org_offset + 6ull > n_org_elems)) return SCAN_KIND;
  return SCAN_OK;
}

__host__ __device__ static inline int job_full_window_bytes(const hmgpu_me_job& j)
{
  const int nx = j.win_r - j.win_l + 1, ny = j.win_b - j.win_t + 1;
  if (nx <= 0 || ny <= 0) return 0;
  return fs_packed_smem_bytes(j.pu_w, j.pu_h, nx, ny, (j.flags & HMGPU_F_FEN) != 0);
}

__global__ void job_scan_kernel(const hmgpu_me_job* __restrict__ jobs, int n, int pic_w, int pic_h, uint32_t valid_slots,
                                unsigned long long n_org_elems, ScanOut* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t bad = 0xffffffffu, fl = 0, win = 0, kinds = 0;
  if (i < n)
  {
    const hmgpu_me_job j = jobs[i];
    const int code = job_check(j, pic_w, pic_h, valid_slots, n_org_elems);
    if (code) bad = ((uint32_t)i << 4) | (uint32_t)code;       // i < 2^26: (index, reason) ordered by index
    fl = j.flags;
    if (j.flags & HMGPU_F_INTEGER)
    {
      if (j.flags & HMGPU_F_FULL) { kinds = 2; if (!code) win = (uint32_t)job_full_window_bytes(j); }
      else kinds = j.kind == HMGPU_KIND_SELECTIVE ? 5 : 1;
    }
  }
  bad = __reduce_min_sync(0xffffffffu, bad);
  fl = __reduce_or_sync(0xffffffffu, fl);
  win = __reduce_max_sync(0xffffffffu, win);
  kinds = __reduce_or_sync(0xffffffffu, kinds);
  if ((threadIdx.x & 31) == 0)
  {
    if (bad != 0xffffffffu) atomicMin(&out->first_bad, bad);
    atomicOr(&out->flags_any, fl);
    if (win) atomicMax(&out->max_win, win);
    atomicOr(&out->int_kinds, kinds);
  }
}

static int me_search_pipelined(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs,
                               const int16_t* org_blocks, int n_org_elems, hmgpu_me_result* results)
{
  const int s_chunk = ctx->tune.pipe_chunk;
  // Chunk schedule.  What the pipeline cannot hide is the copy-in of the FIRST chunk and the copy-out of the LAST one, while
  // the kernels want large chunks (the TZ stage of a chunk is ~17 launches on side streams whose tails overlap best when the
  // chunk is large, the fractional stage stages one window set per CTU group and chunk).  Default: a short first and last chunk
  // around two long ones -- 1/8, 3/8, 3/8, 1/8 of the batch (measured end to end on the 1.18 M-job bench batch: equal quarters
  // 3.86 ms, equal halves 3.77 ms, profiles/r2n_e2e_chunk_sweep.txt).  HMGPU_PIPE_CHUNK=<jobs>: equal chunks of that size.
  enum { MAX_CHUNKS = 4096 };
  int n_chunks, chunk;                       // chunk = the longest chunk (sizes the staging buffers)
  std::vector<int> c_first, c_cnt;
  if (s_chunk > 0 || n_jobs < 4 * 32768)
  {
    chunk = s_chunk > 0 ? s_chunk : 32768;
    chunk = (chunk + 255) & ~255;
    if ((n_jobs + chunk - 1) / chunk > MAX_CHUNKS) chunk = ((n_jobs + MAX_CHUNKS - 1) / MAX_CHUNKS + 255) & ~255;
    for (int f = 0; f < n_jobs; f += chunk) { c_first.push_back(f); c_cnt.push_back(n_jobs - f < chunk ? n_jobs - f : chunk); }
  }
  else
  {
    const int den = ctx->tune.pipe_edge >= 4 ? ctx->tune.pipe_edge : 8;    // HMGPU_PIPE_EDGE: first / last chunk = 1 / den of the batch
    const int eighth = ((n_jobs + den - 1) / den + 255) & ~255;
    int longc = (n_jobs - 2 * eighth + 1) / 2;
    if (longc > 786432) longc = 786432;      // very large batches: more long chunks instead of longer ones
    int f = 0;
    c_first.push_back(f); c_cnt.push_back(eighth); f += eighth;
    while (n_jobs - f > eighth + longc / 2)
    {
      const int c = n_jobs - f - eighth < longc ? n_jobs - f - eighth : longc;
      c_first.push_back(f); c_cnt.push_back(c); f += c;
    }
    c_first.push_back(f); c_cnt.push_back(n_jobs - f);
    chunk = 0;
    for (size_t i = 0; i < c_cnt.size(); i++) chunk = c_cnt[i] > chunk ? c_cnt[i] : chunk;
  }
  n_chunks = (int)c_first.size();
  if (!ctx->lane_store[1].stream)
  {
    HMGPU_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lane_store[1].stream, cudaStreamNonBlocking));
    HMGPU_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++)
    {
      HMGPU_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lane_done[i], cudaEventDisableTiming));
      HMGPU_CUDA(ctx, cudaEventCreateWithFlags(&ctx->scan_done[i], cudaEventDisableTiming));
    }
    HMGPU_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&ctx->d_scan, 2 * sizeof(ScanOut)));
    HMGPU_CUDA(ctx, hmgpu_hmalloc((void**)&ctx->h_scan, 2 * sizeof(ScanOut)));
  }
  ScanOut* d_scan = (ScanOut*)ctx->d_scan;
  ScanOut* h_scan = (ScanOut*)ctx->h_scan;
  uint32_t valid_slots = 0;
  for (int i = 0; i < ctx->max_refs; i++) if (ctx->refs[i].valid) valid_slots |= 1u << i;

  // bi-pred key patterns: the whole array goes up once, on lane 0, before the lanes fork
  const bool have_org = org_blocks && n_org_elems > 0;
  if (have_org)
  {
    const size_t ob = sizeof(int16_t) * (size_t)n_org_elems;
    if (ob > ctx->d_orgblk_bytes)
    {
      HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->d_orgblk) hmgpu_dfree(ctx->d_orgblk);
      ctx->d_orgblk = NULL; ctx->d_orgblk_bytes = 0;
      HMGPU_CUDA(ctx, hmgpu_dmalloc((void**)&ctx->d_orgblk, round_up(ob + ob / 4, 1 << 20)));
      ctx->d_orgblk_bytes = round_up(ob + ob / 4, 1 << 20);
    }
    if (is_pinned(org_blocks))
      HMGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_orgblk, org_blocks, ob, cudaMemcpyHostToDevice, ctx->stream));
    else
    {
      HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      int rc0 = hmgpu_reserve_pinned(ctx, ob);
      if (rc0) return rc0;
      memcpy(ctx->h_pin, org_blocks, ob);
      HMGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_orgblk, ctx->h_pin, ob, cudaMemcpyHostToDevice, ctx->stream));
      HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h_pin is reused for the job chunks below
    }
  }
  HMGPU_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_store[1].stream, ctx->fork_ev, 0));
  HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->fork_ev, 0));

  const bool jobs_pinned = is_pinned(jobs), res_pinned = is_pinned(results);
  const size_t jb = round_up(sizeof(hmgpu_me_job) * (size_t)chunk, 256), rb = round_up(sizeof(hmgpu_me_result) * (size_t)chunk, 256);
  int rc = HMGPU_OK;
  // both lanes' staging buffers up front (allocation synchronises the device)
  for (int l = 0; l < 2 && rc == HMGPU_OK; l++)
  {
    hmgpu_use_lane(ctx, l);
    rc = hmgpu_reserve_stage(ctx, jb + rb);
    if (rc == HMGPU_OK && !(jobs_pinned && res_pinned)) rc = hmgpu_reserve_pinned(ctx, jb + rb);
  }
  auto chunk_n = [&](int k) { return c_cnt[k]; };
  // copy stream: jobs of chunk k into its lane's staging buffer (once the lane's previous chunk is done with it), then the scan
  auto prefetch = [&](int k) -> int {
    const int l = k & 1, first = c_first[k], n = chunk_n(k);
    hmgpu_use_lane(ctx, l);
    char* dp = (char*)ctx->d_stage;
    if (k >= 2) HMGPU_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->lane_done[l], 0));
    const ScanOut init = { 0xffffffffu, 0u, 0u, 0u };
    h_scan[l] = init;                                       // chunk k-2's verdict was read long ago
    HMGPU_CUDA(ctx, cudaMemcpyAsync(&d_scan[l], &h_scan[l], sizeof(ScanOut), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (jobs_pinned)
      HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, jobs + first, sizeof(hmgpu_me_job) * (size_t)n, cudaMemcpyHostToDevice, ctx->copy_stream));
    else
    {
      memcpy(ctx->h_pin, jobs + first, sizeof(hmgpu_me_job) * (size_t)n);   // its previous H2D was awaited through scan_done
      HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, ctx->h_pin, sizeof(hmgpu_me_job) * (size_t)n, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    job_scan_kernel<<<(n + 255) / 256, 256, 0, ctx->copy_stream>>>((const hmgpu_me_job*)dp, n, ctx->pic_w, ctx->pic_h, valid_slots,
                                                                  have_org ? (unsigned long long)n_org_elems : 0ull, &d_scan[l]);
    ctx->launches++;
    HMGPU_CUDA(ctx, cudaMemcpyAsync(&h_scan[l], &d_scan[l], sizeof(ScanOut), cudaMemcpyDeviceToHost, ctx->copy_stream));
    HMGPU_CUDA(ctx, cudaEventRecord(ctx->scan_done[l], ctx->copy_stream));
    return HMGPU_OK;
  };
  // copy the results of chunk k out of its lane's pinned staging buffer (pageable `results` only)
  auto retire = [&](int k) -> int {
    hmgpu_use_lane(ctx, k & 1);
    const cudaError_t e = cudaEventSynchronize(ctx->lane_done[k & 1]);
    if (e != cudaSuccess) return hmgpu_fail(ctx, HMGPU_E_CUDA, "pipelined search, chunk %d: %s", k, cudaGetErrorString(e));
    if (!res_pinned) memcpy(results + (size_t)c_first[k], (char*)ctx->h_pin + jb, sizeof(hmgpu_me_result) * (size_t)chunk_n(k));
    return HMGPU_OK;
  };
  int next_retire = 0;
  if (rc == HMGPU_OK) rc = prefetch(0);
  for (int k = 0; k < n_chunks && rc == HMGPU_OK; k++)
  {
    const int l = k & 1, n = chunk_n(k);
    hmgpu_use_lane(ctx, l);
    cudaError_t e = cudaEventSynchronize(ctx->scan_done[l]);
    if (e != cudaSuccess) { rc = hmgpu_fail(ctx, HMGPU_E_CUDA, "pipelined search, scan of chunk %d: %s", k, cudaGetErrorString(e)); break; }
    const ScanOut so = h_scan[l];
    if (so.first_bad != 0xffffffffu)
    {
      rc = hmgpu_fail(ctx, (so.first_bad & 15u) == SCAN_SLOT ? HMGPU_E_STATE : HMGPU_E_INVALID, "job %d: %s",
                      c_first[k] + (int)(so.first_bad >> 4), k_scan_msg[so.first_bad & 15u]);
      break;
    }
    char* dp = (char*)ctx->d_stage;
    char* hp = (char*)ctx->h_pin;
    const bool any_org = (so.flags_any & HMGPU_F_ORG_BLOCK) != 0;
    if ((e = cudaStreamWaitEvent(ctx->stream, ctx->scan_done[l], 0)) != cudaSuccess) { rc = hmgpu_fail(ctx, HMGPU_E_CUDA, "%s", cudaGetErrorString(e)); break; }
    const bool any_sel = (so.int_kinds & 4u) != 0;
    if ((rc = hmgpu_launch_me(ctx, (const hmgpu_me_job*)dp, n, (any_org || any_sel) ? (const int16_t*)ctx->d_orgblk : NULL, (hmgpu_me_result*)(dp + jb),
                              any_org, (so.int_kinds & 2u) != 0, (so.int_kinds & 1u) != 0, (so.flags_any & HMGPU_F_FRAC) != 0, (int)so.max_win, any_sel))) break;
    e = cudaMemcpyAsync(res_pinned ? (void*)(results + (size_t)c_first[k]) : (void*)(hp + jb), dp + jb, sizeof(hmgpu_me_result) * (size_t)n,
                        cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->lane_done[l], ctx->stream);
    if (e != cudaSuccess) { rc = hmgpu_fail(ctx, HMGPU_E_CUDA, "pipelined search D2H: %s", cudaGetErrorString(e)); break; }
    // pageable `results`: chunk k-1 must have left its lane's pinned result buffer before chunk k+1 (same lane) is queued;
    // the wait happens with chunk k already queued, so the device does not idle
    if (k >= 1 && !res_pinned) { if ((rc = retire(k - 1))) break; next_retire = k; }
    if (k + 1 < n_chunks && (rc = prefetch(k + 1))) break;
  }
  if (rc == HMGPU_OK)
    for (int k = next_retire; k < n_chunks && rc == HMGPU_OK; k++) rc = retire(k);
  // leave everything idle and lane 0 current, whatever happened
  cudaStreamSynchronize(ctx->copy_stream);
  hmgpu_use_lane(ctx, 1); cudaStreamSynchronize(ctx->stream);
  hmgpu_use_lane(ctx, 0); cudaStreamSynchronize(ctx->stream);
  return rc;
}

// ---- mailbox server (host side) --------------------------------------------------------------------------
// The same code drives the server in both deployments: in-process (this process launched me_server_kernel on its own
// stream) and as the client of a broker daemon (ctx->remote: the daemon launched it, the mailbox is a shared-memory segment
// both processes map; starting / draining the kernel is one socket round trip, everything else is plain memory traffic).
static uint32_t line_check(const uint32_t* w)
{
  uint32_t h = 0x7F4A7C15u;
  for (int i = 0; i < 15; i++) h = (h ^ w[i]) * 0x85EBCA6Bu + (h >> 15);
  return h;
}

// One call = n_lines job lines (ME jobs first, prediction-error jobs after them).  All lines a server CTA may look at are
// (re)written: CTA b polls line b and, once it has seen the call's ticket there, fetches lines b + srv_ctas, b + 2 srv_ctas ...
// of the same call -- so the lines a CTA does NOT poll are written first (x86 stores become visible in order).
static void server_write_lines(const hmgpu_ctx* ctx, Mailbox* mb, const hmgpu_me_job* jobs, int n_jobs, const hmgpu_pred_job* pjobs,
                               const uint8_t* pfuncs, int n_pred, uint32_t ticket, uint32_t gen)
{
  const int n_lines = n_jobs + n_pred;
  const int n_write = n_lines > ctx->srv_ctas ? n_lines : ctx->srv_ctas;
  for (int b = n_write - 1; b >= 0; b--)
  {
    uint32_t w[16];
    memset(w, 0, sizeof w);
    uint32_t fl = (uint32_t)n_lines << 8;
    if (b < n_jobs) { memcpy(w, &jobs[b], sizeof(hmgpu_me_job)); fl |= 1u; }
    else if (b < n_lines) { memcpy(w, &pjobs[b - n_jobs], sizeof(hmgpu_pred_job)); fl |= 3u | ((uint32_t)pfuncs[b - n_jobs] << 16); }
    w[12] = ticket; w[13] = gen; w[14] = fl; w[15] = line_check(w);
    volatile uint32_t* dst = mb->lines[b];
    for (int i = 0; i < 16; i++) dst[i] = w[i];
  }
  __sync_synchronize();
}

// wait until the server kernel has left its stream
static int server_drain(hmgpu_ctx* ctx)
{
  if (ctx->remote) return hmgpu_remote_call(ctx, HMB_OP_SERVER_SYNC, NULL, NULL, NULL);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->srv_stream));
  return HMGPU_OK;
}

// make the server leave (new generation in every line) and wait for it.  Called before anything that rewrites the
// planes the server reads through the read-only path, and before work that wants the whole GPU.
static int hmgpu_server_stop(hmgpu_ctx* ctx)
{
  if (!ctx->srv_alive) return HMGPU_OK;
  Mailbox* mb = (Mailbox*)ctx->h_mail;
  ctx->srv_gen++;
  server_write_lines(ctx, mb, NULL, 0, NULL, NULL, 0, ctx->mail_ticket, ctx->srv_gen);
  const int rc = server_drain(ctx);
  if (rc) return rc;
  ctx->srv_alive = false;
  return HMGPU_OK;
}

// the mailbox of an in-process context: mapped pinned memory, allocated on first use
static int mailbox_ensure(hmgpu_ctx* ctx)
{
  if (ctx->h_mail) return HMGPU_OK;
  HMGPU_CUDA(ctx, cudaHostAlloc(&ctx->h_mail, sizeof(Mailbox), cudaHostAllocMapped));
  memset(ctx->h_mail, 0, sizeof(Mailbox));
  HMGPU_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_mail, ctx->h_mail, 0));
  return HMGPU_OK;
}

// launch generation `gen` of the server kernel of an in-process context (also what the broker daemon runs on behalf of a client)
static int hmgpu_server_launch(hmgpu_ctx* ctx, uint32_t gen, uint32_t last_ticket, int dyn_bytes)
{
  Mailbox* dm = (Mailbox*)ctx->d_mail;
  if (!dm) return hmgpu_fail(ctx, HMGPU_E_STATE, "no mailbox");
  if (!ctx->srv_stream) HMGPU_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->srv_stream, cudaStreamNonBlocking));
  // ordered after everything already queued on the context's stream (uploads are synchronous, this is belt and braces)
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return hmgpu_launch_server(ctx, ctx->srv_stream, &dm->lines[0][0], dm->org_blocks, dm->slots, dm->exited, gen, last_ticket,
                             (unsigned long long)ctx->tune.server_idle_us * 1000ull, ctx->srv_ctas, dyn_bytes);
}

static int server_start(hmgpu_ctx* ctx, int dyn_bytes)
{
  ctx->srv_dyn = dyn_bytes;
  int rc;
  if (ctx->remote)
  {
    const int32_t a[6] = { (int32_t)ctx->srv_gen, (int32_t)(ctx->mail_ticket - 1), dyn_bytes, 0, 0, 0 };
    rc = hmgpu_remote_call(ctx, HMB_OP_SERVER_START, a, NULL, NULL);
  }
  else rc = hmgpu_server_launch(ctx, ctx->srv_gen, ctx->mail_ticket - 1, dyn_bytes);
  if (rc) return rc;
  ctx->srv_alive = true;
  ctx->srv_starts++;
  return HMGPU_OK;
}

static inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();                                  // be kind to the sibling hyper-thread while spinning
#endif
}

static bool slot_ready(const volatile uint32_t* slot, uint32_t ticket, uint32_t* out6)
{
  uint32_t w[8];
  for (int k = 0; k < 8; k++) w[k] = slot[k];
  if (w[6] != ticket || w[7] != hmgpu_mail_check(w, ticket)) return false;
  memcpy(out6, w, 6 * sizeof(uint32_t));
  return true;
}

// dynamic shared memory a server generation needs: the staged TZ window, the largest full-search window of the call, and (for
// prediction-error lines) the scratch of the motion-compensation passes
#define SRV_DYN_MIN (10 * 1024)
#define SRV_DYN_PRED (28 * 1024)

// one call through the resident server, in two halves so that the caller can work while the device searches:
// submit = write the lines (start a generation if none is alive); wait = poll the result slots
static int server_submit(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs, bool any_org, const int16_t* org_blocks, int n_org_elems,
                         int max_win, const hmgpu_pred_job* pjobs, const uint8_t* pfuncs, int n_pred)
{
  Mailbox* mb = (Mailbox*)ctx->h_mail;
  int need_dyn = max_win > SRV_DYN_MIN ? max_win : SRV_DYN_MIN;
  if (n_pred && need_dyn < SRV_DYN_PRED) need_dyn = SRV_DYN_PRED;
  int rc;
  if (ctx->srv_alive && need_dyn > ctx->srv_dyn && (rc = hmgpu_server_stop(ctx))) return rc;
  if (ctx->srv_alive && ((volatile uint32_t*)mb->exited)[0] == ctx->srv_gen)
  {
    // the server left on its own (idle): its stream is drained by the next start
    ctx->srv_alive = false;
    ctx->srv_gen++;
  }
  if (any_org) memcpy(mb->org_blocks, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
  const uint32_t ticket = ++ctx->mail_ticket;
  // kept for a re-submit if the server must be restarted
  memcpy(ctx->pend_jobs, jobs, sizeof(hmgpu_me_job) * (size_t)n_jobs);
  if (n_pred) { memcpy(ctx->pend_pred, pjobs, sizeof(hmgpu_pred_job) * (size_t)n_pred); memcpy(ctx->pend_pfunc, pfuncs, (size_t)n_pred); }
  ctx->pend_n = n_jobs; ctx->pend_np = n_pred;
  server_write_lines(ctx, mb, jobs, n_jobs, pjobs, pfuncs, n_pred, ticket, ctx->srv_gen);
  if (!ctx->srv_alive && (rc = server_start(ctx, need_dyn > ctx->srv_dyn ? need_dyn : ctx->srv_dyn))) return rc;
  ctx->srv_calls++;
  return HMGPU_OK;
}

static int server_wait(hmgpu_ctx* ctx, hmgpu_me_result* results, uint32_t* pred_out)
{
  Mailbox* mb = (Mailbox*)ctx->h_mail;
  const uint32_t ticket = ctx->mail_ticket;
  const int n_jobs = ctx->pend_n, n_lines = ctx->pend_n + ctx->pend_np;
  int rc;
  for (int i = 0; i < n_lines; i++)
  {
    const volatile uint32_t* slot = (const volatile uint32_t*)&mb->slots[i];
    const int cta = i % ctx->srv_ctas;                     // the server CTA that serves line i
    uint32_t w6[6];
    unsigned spins = 0;
    while (!slot_ready(slot, ticket, w6))
    {
      cpu_relax();
      if ((++spins & 63u) == 0 && ((volatile uint32_t*)mb->exited)[cta] == ctx->srv_gen)
      {
        // this CTA stopped polling (idle exit racing with the call); anything it published is already visible
        if (slot_ready(slot, ticket, w6)) break;
        ctx->srv_alive = false;
        ctx->srv_gen++;
        server_write_lines(ctx, mb, ctx->pend_jobs, n_jobs, ctx->pend_pred, ctx->pend_pfunc, ctx->pend_np, ticket, ctx->srv_gen);
        if ((rc = server_start(ctx, ctx->srv_dyn))) return rc;
      }
      if (spins > 4000000u)
      {
        spins = 0;
        if (ctx->remote) { if ((rc = hmgpu_remote_call(ctx, HMB_OP_SERVER_QUERY, NULL, NULL, NULL))) return rc; }
        else
        {
          const cudaError_t e = cudaStreamQuery(ctx->srv_stream);
          if (e != cudaErrorNotReady && e != cudaSuccess)
            return hmgpu_fail(ctx, HMGPU_E_CUDA, "mailbox server failed: %s", cudaGetErrorString(e));
        }
      }
    }
    if (i < n_jobs) memcpy(&results[i], w6, sizeof(hmgpu_me_result));
    else pred_out[i - n_jobs] = w6[0];
  }
  ctx->pend_n = ctx->pend_np = 0;
  return HMGPU_OK;
}

static bool server_enabled(const hmgpu_ctx* ctx) { return ctx->tune.server && !ctx->tune.trace; }

// Asynchronous pair ("async submit/poll for the mailbox path").  hmgpu_pu_submit returns as soon as the jobs are
// visible to the device; the caller does host work that does not need the results and collects them with hmgpu_pu_wait.
// Batches the resident server cannot take are kept and run inside the wait.
static int validate_pred_jobs(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n, bool chroma, int n_dst, bool need_dst);

int hmgpu_pu_submit(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems,
                    const hmgpu_pred_job* pred_jobs, const uint8_t* pred_funcs, int n_pred)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (ctx->pend_n || ctx->pend_np || ctx->defer_n || ctx->defer_np) return hmgpu_fail(ctx, HMGPU_E_STATE, "submit: the previous submit has not been waited for");
  if (n_jobs == 0 && n_pred == 0) return HMGPU_OK;
  if (n_jobs < 0 || n_pred < 0 || n_jobs > MAIL_JOBS || n_pred > HMGPU_MAIL_LINES || (n_jobs && !jobs) || (n_pred && (!pred_jobs || !pred_funcs)))
    return hmgpu_fail(ctx, HMGPU_E_INVALID, "submit takes up to %d searches and up to %d prediction-error jobs", MAIL_JOBS, HMGPU_MAIL_LINES);
  if (org_blocks && n_org_elems < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "n_org_elems is negative");
  if (!ctx->remote) HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  bool any_org = false, any_full, any_tz, any_frac, any_sel = false;
  int max_win = 0;
  int rc = validate_jobs(ctx, jobs, n_jobs, org_blocks ? n_org_elems : 0, &any_org, &any_full, &any_tz, &any_frac, &max_win, 0, &any_sel);
  if (rc) return rc;
  for (int i = 0; i < n_pred; i++)
    if (pred_funcs[i] != HMGPU_DF_SAD && pred_funcs[i] != HMGPU_DF_SAD_GENERIC && pred_funcs[i] != HMGPU_DF_HADS)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "prediction-error job %d: func %d (only SAD and HADS prediction errors exist in the reference)", i, pred_funcs[i]);
  if ((rc = validate_pred_jobs(ctx, pred_jobs, n_pred, false, 0, false))) return rc;
  any_org = any_org || any_sel;                            // here: "the side array travels" (key patterns or MV predictors)
  if (server_enabled(ctx) && n_jobs + n_pred <= HMGPU_MAIL_LINES && max_win <= 180 * 1024 && (!any_org || n_org_elems <= MAIL_JOBS * 64 * 64))
  {
    if (!ctx->remote && (rc = mailbox_ensure(ctx))) return rc;
    return server_submit(ctx, jobs, n_jobs, any_org, org_blocks, n_org_elems, max_win, pred_jobs, pred_funcs, n_pred);
  }
  // deferred: the blocking calls run in the wait (the caller's buffers may be gone by then: copy)
  if (any_org)
  {
    if ((size_t)n_org_elems > ctx->defer_org_cap)
    {
      free(ctx->defer_org);
      ctx->defer_org = (int16_t*)malloc(sizeof(int16_t) * (size_t)n_org_elems);
      if (!ctx->defer_org) { ctx->defer_org_cap = 0; return hmgpu_fail(ctx, HMGPU_E_NOMEM, "out of host memory"); }
      ctx->defer_org_cap = (size_t)n_org_elems;
    }
    memcpy(ctx->defer_org, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
  }
  memcpy(ctx->pend_jobs, jobs, sizeof(hmgpu_me_job) * (size_t)n_jobs);
  if (n_pred) { memcpy(ctx->pend_pred, pred_jobs, sizeof(hmgpu_pred_job) * (size_t)n_pred); memcpy(ctx->pend_pfunc, pred_funcs, (size_t)n_pred); }
  ctx->defer_n = n_jobs; ctx->defer_np = n_pred; ctx->defer_org_n = any_org ? n_org_elems : 0;
  return HMGPU_OK;
}

int hmgpu_pu_wait(hmgpu_ctx* ctx, hmgpu_me_result* results, uint32_t* pred_out)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (ctx->defer_n || ctx->defer_np)
  {
    const int n = ctx->defer_n, np = ctx->defer_np;
    ctx->defer_n = ctx->defer_np = 0;
    if ((n && !results) || (np && !pred_out)) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL results");
    int rc = HMGPU_OK;
    if (n)
    {
      hmgpu_me_job jobs[MAIL_JOBS];
      memcpy(jobs, ctx->pend_jobs, sizeof(hmgpu_me_job) * (size_t)n);
      rc = hmgpu_me_search(ctx, jobs, n, ctx->defer_org_n ? ctx->defer_org : NULL, ctx->defer_org_n, results);
    }
    // prediction-error jobs of one call may mix SAD and HADS: one batched call per function
    for (int f = 0; f < 3 && rc == HMGPU_OK && np; f++)
    {
      const int func = f == 0 ? HMGPU_DF_SAD : (f == 1 ? HMGPU_DF_SAD_GENERIC : HMGPU_DF_HADS);
      hmgpu_pred_job pj[HMGPU_MAIL_LINES]; int at[HMGPU_MAIL_LINES]; uint32_t o[HMGPU_MAIL_LINES]; int m = 0;
      for (int i = 0; i < np; i++) if (ctx->pend_pfunc[i] == func) { pj[m] = ctx->pend_pred[i]; at[m++] = i; }
      if (!m) continue;
      rc = hmgpu_pred_error(ctx, pj, m, func, o);
      for (int i = 0; i < m && rc == HMGPU_OK; i++) pred_out[at[i]] = o[i];
    }
    return rc;
  }
  if (!ctx->pend_n && !ctx->pend_np) return HMGPU_OK;
  if ((ctx->pend_n && !results) || (ctx->pend_np && !pred_out)) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL results");
  return server_wait(ctx, results, pred_out);
}

int hmgpu_me_submit(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs != 0 && !jobs) return hmgpu_fail(ctx, HMGPU_E_INVALID, "hmgpu_me_submit takes 1..%d jobs", MAIL_JOBS);
  return hmgpu_pu_submit(ctx, jobs, n_jobs, org_blocks, n_org_elems, NULL, NULL, 0);
}

int hmgpu_me_wait(hmgpu_ctx* ctx, hmgpu_me_result* results) { return hmgpu_pu_wait(ctx, results, NULL); }

// a batch through the broker daemon: jobs, key blocks and results travel through the batch area of the shared segment
static int remote_me_batch(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems, hmgpu_me_result* results)
{
  size_t cap = 0;
  char* area = (char*)hmgpu_remote_area(ctx, 1, &cap);
  const size_t ob = round_up(org_blocks ? sizeof(int16_t) * (size_t)n_org_elems : 0, 256);
  if (ob + 4096 > cap) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "key-pattern blocks (%zu bytes) do not fit the broker's batch area", ob);
  const size_t per_job = sizeof(hmgpu_me_job) + sizeof(hmgpu_me_result);
  const int chunk = (int)((cap - ob - 512) / per_job);
  if (org_blocks && n_org_elems) memcpy(area, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
  for (int first = 0; first < n_jobs; first += chunk)
  {
    const int n = n_jobs - first < chunk ? n_jobs - first : chunk;
    const size_t jb = round_up(sizeof(hmgpu_me_job) * (size_t)n, 256);
    memcpy(area + ob, jobs + first, sizeof(hmgpu_me_job) * (size_t)n);
    const int32_t a[6] = { n, org_blocks ? n_org_elems : 0, first, 0, 0, 0 };
    const int rc = hmgpu_remote_call(ctx, HMB_OP_ME_BATCH, a, NULL, NULL);
    if (rc) return rc;
    memcpy(results + first, area + ob + jb, sizeof(hmgpu_me_result) * (size_t)n);
  }
  return HMGPU_OK;
}

int hmgpu_me_search(hmgpu_ctx* ctx, const hmgpu_me_job* jobs, int n_jobs,
                    const int16_t* org_blocks, int n_org_elems, hmgpu_me_result* results)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!jobs || !results || n_jobs < 0 || n_jobs > (1 << 26)) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad jobs/results/n_jobs");
  if (org_blocks && n_org_elems < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "n_org_elems is negative");
  if (ctx->pend_n || ctx->pend_np) return hmgpu_fail(ctx, HMGPU_E_STATE, "a submit is outstanding: wait for it first");
  if (!ctx->remote) HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_jobs > MAIL_JOBS) { const int rcs = hmgpu_server_stop(ctx); if (rcs) return rcs; }   // a batch wants the whole GPU
  if (n_jobs >= PIPE_MIN_JOBS && ctx->tune.pipeline && !ctx->remote) return me_search_pipelined(ctx, jobs, n_jobs, org_blocks, n_org_elems, results);
  bool any_org, any_full, any_tz, any_frac, any_sel;
  int max_win;
  int rc = validate_jobs(ctx, jobs, n_jobs, org_blocks ? n_org_elems : 0, &any_org, &any_full, &any_tz, &any_frac, &max_win, 0, &any_sel);
  if (rc) return rc;
  const bool key_blocks = any_org;                         // int16 key patterns present: selects the non-packed kernels
  any_org = any_org || any_sel;                            // below: "the side array travels" (key patterns or MV predictors)
  const bool mail_ok = n_jobs <= MAIL_JOBS && (!any_org || n_org_elems <= MAIL_JOBS * 64 * 64) && ctx->tune.fastpath;
  if (mail_ok && server_enabled(ctx) && max_win <= 180 * 1024)
  {
    // ---- low-latency path: the resident server polls the job lines, the host spins on the result slots ----
    if (!ctx->remote && (rc = mailbox_ensure(ctx))) return rc;
    rc = server_submit(ctx, jobs, n_jobs, any_org, org_blocks, n_org_elems, max_win, NULL, NULL, 0);
    return rc ? rc : server_wait(ctx, results, NULL);
  }
  if (ctx->remote) return remote_me_batch(ctx, jobs, n_jobs, any_org ? org_blocks : NULL, n_org_elems, results);
  if (mail_ok)
  {
    // ---- one fused launch per call (HMGPU_SERVER=0 / tracing): jobs as kernel parameters, results in the mailbox slots ----
    if ((rc = mailbox_ensure(ctx))) return rc;
    Mailbox* mb = (Mailbox*)ctx->h_mail;
    Mailbox* dm = (Mailbox*)ctx->d_mail;
    const bool s_trace = ctx->tune.trace != 0;
    const double t0 = s_trace ? now_us() : 0.0;
    HmgpuJobPack pack;
    memcpy(pack.jobs, jobs, sizeof(hmgpu_me_job) * (size_t)n_jobs);
    if (any_org) memcpy(mb->org_blocks, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
    const uint32_t ticket = ++ctx->mail_ticket;
    __sync_synchronize();
    const double t1 = s_trace ? now_us() : 0.0;
    rc = hmgpu_launch_single(ctx, pack, n_jobs, dm->org_blocks, dm->slots, ticket, key_blocks, max_win, s_trace ? dm->trace : NULL);
    if (rc) return rc;
    const double t2 = s_trace ? now_us() : 0.0;
    for (int i = 0; i < n_jobs; i++)
    {
      const volatile uint32_t* slot = (const volatile uint32_t*)&mb->slots[i];
      unsigned spins = 0;
      for (;;)
      {
        uint32_t w[8];
        for (int k = 0; k < 8; k++) w[k] = slot[k];
        if (w[6] == ticket && w[7] == hmgpu_mail_check(w, ticket)) { memcpy(&results[i], w, sizeof(hmgpu_me_result)); break; }
        if (++spins > 2000000u)
        {
          spins = 0;
          const cudaError_t e = cudaStreamQuery(ctx->stream);
          if (e != cudaErrorNotReady)
          {
            for (int k = 0; k < 8; k++) w[k] = slot[k];
            if (w[6] == ticket && w[7] == hmgpu_mail_check(w, ticket)) { memcpy(&results[i], w, sizeof(hmgpu_me_result)); break; }
            return hmgpu_fail(ctx, HMGPU_E_CUDA, "low-latency search kernel ended without publishing job %d: %s", i, cudaGetErrorString(e));
          }
        }
      }
    }
    const double t3 = s_trace ? now_us() : 0.0;
    if (s_trace)
    {
      cudaStreamSynchronize(ctx->stream);                   // the last stamp is written after the flag
      g_trace.host_prep += t1 - t0; g_trace.host_launch += t2 - t1; g_trace.host_wait += t3 - t2; g_trace.host_copy += now_us() - t3;
      for (int k = 0; k < 5; k++) g_trace.dev[k] += (double)(mb->trace[k + 1] - mb->trace[k]) * 1e-3;
      g_trace.n++;
    }
    return HMGPU_OK;
  }
  const size_t jb = round_up(sizeof(hmgpu_me_job) * (size_t)n_jobs, 256);
  const size_t ob = round_up(any_org ? sizeof(int16_t) * (size_t)n_org_elems : 0, 256);
  const size_t rb = round_up(sizeof(hmgpu_me_result) * (size_t)n_jobs, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = hmgpu_reserve_pinned(ctx, jb + ob + rb))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, jb + ob + rb))) return rc;
  char* hp = (char*)ctx->h_pin;
  char* dp = (char*)ctx->d_stage;
  // page-locked caller buffers are copied directly; pageable ones go through the pinned staging buffer
  const bool jobs_pinned = is_pinned(jobs), res_pinned = is_pinned(results);
  if (jobs_pinned)
    HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, jobs, sizeof(hmgpu_me_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, ctx->stream));
  else
  {
    memcpy(hp, jobs, sizeof(hmgpu_me_job) * (size_t)n_jobs);
    HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, jb, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (any_org)
  {
    memcpy(hp + jb, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
    HMGPU_CUDA(ctx, cudaMemcpyAsync(dp + jb, hp + jb, ob, cudaMemcpyHostToDevice, ctx->stream));
  }
  rc = hmgpu_launch_me(ctx, (const hmgpu_me_job*)dp, n_jobs, any_org ? (const int16_t*)(dp + jb) : NULL,
                       (hmgpu_me_result*)(dp + jb + ob), key_blocks, any_full, any_tz, any_frac, max_win, any_sel);
  if (rc) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(res_pinned ? (void*)results : (void*)(hp + jb + ob), dp + jb + ob,
                                  sizeof(hmgpu_me_result) * (size_t)n_jobs, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (!res_pinned) memcpy(results, hp + jb + ob, sizeof(hmgpu_me_result) * (size_t)n_jobs);
  return HMGPU_OK;
}

int hmgpu_me_search_device(hmgpu_ctx* ctx, const void* d_jobs, int n_jobs, const void* d_org_blocks, void* d_results,
                           int flags_any)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!d_jobs || !d_results || n_jobs < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad jobs/results/n_jobs");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_me_search_device");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  { const int rcs = hmgpu_server_stop(ctx); if (rcs) return rcs; }   // the server reads the planes through the read-only path
  // jobs are not visible to the host: worst-case full-search window (SR 64, 64x64 PU)
  const bool integer = (flags_any & HMGPU_F_INTEGER) != 0, full = (flags_any & HMGPU_F_FULL) != 0;
  return hmgpu_launch_me(ctx, (const hmgpu_me_job*)d_jobs, n_jobs, (const int16_t*)d_org_blocks,
                         (hmgpu_me_result*)d_results, (flags_any & HMGPU_F_ORG_BLOCK) != 0, integer && full,
                         integer, (flags_any & HMGPU_F_FRAC) != 0, 64 * 1024, false);
}

void hmgpu_clip_bounds_ctu(int pic_w, int pic_h, int cu_x, int cu_y, int max_cu_w, int max_cu_h, int16_t bounds[4])
{
  // TComDataCU::clipMv (TComDataCU.cpp:2917-2929): iOffset = 8, bounds relative to the CU origin, g_uiMaxCUWidth / Height on the
  // low side.  With pictures up to 8184 samples (hmgpu_create) every bound fits int16.
  bounds[0] = (int16_t)((-max_cu_w - 8 - cu_x + 1) * 4);
  bounds[1] = (int16_t)((pic_w + 8 - cu_x - 1) * 4);
  bounds[2] = (int16_t)((-max_cu_h - 8 - cu_y + 1) * 4);
  bounds[3] = (int16_t)((pic_h + 8 - cu_y - 1) * 4);
}

void hmgpu_clip_bounds(int pic_w, int pic_h, int cu_x, int cu_y, int16_t bounds[4])
{
  hmgpu_clip_bounds_ctu(pic_w, pic_h, cu_x, cu_y, 64, 64, bounds);      // g_uiMaxCUWidth = g_uiMaxCUHeight = 64 (every BASELINE cfg)
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void hmgpu_search_range(const int16_t bd[4], int pred_x, int pred_y, int srch_rng, int16_t ltrb[4])
{
  // TEncSearch::xSetSearchRange (TEncSearch.cpp:3911-3927); TComMv stores Short
  const int px = clampi(pred_x, bd[0], bd[1]), py = clampi(pred_y, bd[2], bd[3]);
  const int r4 = srch_rng << 2;
  ltrb[0] = (int16_t)(clampi((int16_t)(px - r4), bd[0], bd[1]) >> 2);
  ltrb[1] = (int16_t)(clampi((int16_t)(py - r4), bd[2], bd[3]) >> 2);
  ltrb[2] = (int16_t)(clampi((int16_t)(px + r4), bd[0], bd[1]) >> 2);
  ltrb[3] = (int16_t)(clampi((int16_t)(py + r4), bd[2], bd[3]) >> 2);
}

static uint32_t component_bits(int v)
{
  uint32_t t = (v <= 0) ? (uint32_t)((-v << 1) + 1) : (uint32_t)(v << 1);
  uint32_t len = 1;
  while (t != 1) { t >>= 1; len += 2; }
  return len;
}

uint32_t hmgpu_mv_bits(int pred_x, int pred_y, int scale, int x, int y)
{
  return component_bits((x << scale) - pred_x) + component_bits((y << scale) - pred_y);
}

uint32_t hmgpu_mv_cost(uint32_t ui_cost, int pred_x, int pred_y, int scale, int x, int y)
{
  return (uint32_t)(ui_cost * hmgpu_mv_bits(pred_x, pred_y, scale, x, y)) >> 16;
}

// ---- distortion batch --------------------------------------------------------------------------

int hmgpu_dist_batch(hmgpu_ctx* ctx, const int16_t* org, int n_org, const int16_t* cur, int n_cur,
                     const hmgpu_dist_item* items, int n_items, uint32_t* out)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_items == 0) return HMGPU_OK;
  if (!org || !cur || !items || !out || n_items < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_dist_batch");
  for (int i = 0; i < n_items; i++)
  {
    const hmgpu_dist_item& it = items[i];
    if (it.w < 1 || it.h < 1 || it.func > HMGPU_DF_SSE || it.sub_shift > 4)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "item %d: bad w/h/func/sub_shift", i);
    if (it.func == HMGPU_DF_HADS && ((it.w & 1) || (it.h & 1)))
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "item %d: HADS needs even dimensions", i);
    if (it.func == HMGPU_DF_SAD && (it.h & ((1 << it.sub_shift) - 1)))
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "item %d: rows not a multiple of the sub-sampling step", i);
    const long long oe = (long long)it.org_offset + (long long)(it.h - 1) * it.org_stride + it.w;
    const long long ce = (long long)it.cur_offset + (long long)(it.h - 1) * it.cur_stride + it.w;
    if (it.org_stride < 0 || it.cur_stride < 0 || oe > n_org || ce > n_cur)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "item %d: block outside its array", i);
  }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t b0 = round_up(sizeof(int16_t) * (size_t)n_org, 256), b1 = round_up(sizeof(int16_t) * (size_t)n_cur, 256);
  const size_t b2 = round_up(sizeof(hmgpu_dist_item) * (size_t)n_items, 256), b3 = round_up(sizeof(uint32_t) * (size_t)n_items, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1 + b2 + b3))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1 + b2 + b3))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, org, sizeof(int16_t) * (size_t)n_org);
  memcpy(hp + b0, cur, sizeof(int16_t) * (size_t)n_cur);
  memcpy(hp + b0 + b1, items, sizeof(hmgpu_dist_item) * (size_t)n_items);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0 + b1 + b2, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_dist(ctx, (const int16_t*)dp, (const int16_t*)(dp + b0), (const hmgpu_dist_item*)(dp + b0 + b1), n_items,
                              (uint32_t*)(dp + b0 + b1 + b2)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0 + b1 + b2, dp + b0 + b1 + b2, sizeof(uint32_t) * (size_t)n_items, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out, hp + b0 + b1 + b2, sizeof(uint32_t) * (size_t)n_items);
  return HMGPU_OK;
}

// ---- SAO statistics (sao.cu) ----------------------------------------------------------------------
int hmgpu_sao_stats(hmgpu_ctx* ctx, const int16_t* rec, int rec_stride, const int16_t* org, int org_stride, int width, int height,
                    int ctu_w, int ctu_h, const uint8_t* ctu_flags, const int32_t skip_r[5], const int32_t skip_b[5], int64_t* stats)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (!rec || !org || !skip_r || !skip_b || !stats) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_sao_stats");
  if (width < 1 || height < 1 || width > 16384 || height > 16384 || rec_stride < width || org_stride < width)
    return hmgpu_fail(ctx, HMGPU_E_INVALID, "component of %d x %d samples with strides %d / %d", width, height, rec_stride, org_stride);
  if (ctu_w < 8 || ctu_h < 8 || ctu_w > 64 || ctu_h > 64) return hmgpu_fail(ctx, HMGPU_E_INVALID, "CTU of %d x %d samples", ctu_w, ctu_h);
  for (int t = 0; t < 5; t++)
    if (skip_r[t] < 0 || skip_b[t] < 0 || skip_r[t] >= ctu_w || skip_b[t] >= ctu_h) return hmgpu_fail(ctx, HMGPU_E_INVALID, "skip lines of type %d outside the CTU", t);
  const int ctus_x = (width + ctu_w - 1) / ctu_w, n_ctus = ctus_x * ((height + ctu_h - 1) / ctu_h);
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t rb = round_up(sizeof(int16_t) * (size_t)width * height, 256), fb = round_up((size_t)n_ctus, 256);
  const size_t sb = round_up(sizeof(int64_t) * 5 * 64 * (size_t)n_ctus, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, 2 * rb + fb + sb))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, 2 * rb + fb + sb))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  for (int y = 0; y < height; y++)               // tight rows on the device
  {
    memcpy(hp + sizeof(int16_t) * (size_t)y * width, rec + (size_t)y * rec_stride, sizeof(int16_t) * (size_t)width);
    memcpy(hp + rb + sizeof(int16_t) * (size_t)y * width, org + (size_t)y * org_stride, sizeof(int16_t) * (size_t)width);
  }
  if (ctu_flags) memcpy(hp + 2 * rb, ctu_flags, (size_t)n_ctus);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, 2 * rb + fb, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_sao_stats(ctx, (const int16_t*)dp, width, (const int16_t*)(dp + rb), width, width, height, ctu_w, ctu_h,
                                   ctu_flags ? (const uint8_t*)(dp + 2 * rb) : NULL, skip_r, skip_b, (long long*)(dp + 2 * rb + fb)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + 2 * rb + fb, dp + 2 * rb + fb, sizeof(int64_t) * 5 * 64 * (size_t)n_ctus, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(stats, hp + 2 * rb + fb, sizeof(int64_t) * 5 * 64 * (size_t)n_ctus);
  return HMGPU_OK;
}

int hmgpu_sao_apply(hmgpu_ctx* ctx, const int16_t* rec, int rec_stride, int width, int height, int ctu_w, int ctu_h,
                    const uint8_t* ctu_flags, const int8_t* types, const int32_t* offsets, int16_t* out)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (!rec || !types || !offsets || !out) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_sao_apply");
  if (width < 1 || height < 1 || width > 16384 || height > 16384 || rec_stride < width)
    return hmgpu_fail(ctx, HMGPU_E_INVALID, "component of %d x %d samples with stride %d", width, height, rec_stride);
  if (ctu_w < 8 || ctu_h < 8 || ctu_w > 64 || ctu_h > 64) return hmgpu_fail(ctx, HMGPU_E_INVALID, "CTU of %d x %d samples", ctu_w, ctu_h);
  const int ctus_x = (width + ctu_w - 1) / ctu_w, n_ctus = ctus_x * ((height + ctu_h - 1) / ctu_h);
  for (int c = 0; c < n_ctus; c++)
    if (types[c] < -1 || types[c] > 4) return hmgpu_fail(ctx, HMGPU_E_INVALID, "CTU %d: SAO type %d", c, types[c]);
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t rb = round_up(sizeof(int16_t) * (size_t)width * height, 256), fb = round_up((size_t)n_ctus, 256);
  const size_t ob = round_up(sizeof(int32_t) * 32 * (size_t)n_ctus, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, 2 * rb + 2 * fb + ob))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, 2 * rb + 2 * fb + ob))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  for (int y = 0; y < height; y++) memcpy(hp + sizeof(int16_t) * (size_t)y * width, rec + (size_t)y * rec_stride, sizeof(int16_t) * (size_t)width);
  if (ctu_flags) memcpy(hp + rb, ctu_flags, (size_t)n_ctus);
  memcpy(hp + rb + fb, types, (size_t)n_ctus);
  memcpy(hp + rb + 2 * fb, offsets, sizeof(int32_t) * 32 * (size_t)n_ctus);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, rb + 2 * fb + ob, cudaMemcpyHostToDevice, ctx->stream));
  int16_t* d_out = (int16_t*)(dp + rb + 2 * fb + ob);
  if ((rc = hmgpu_launch_sao_apply(ctx, (const int16_t*)dp, width, width, height, ctu_w, ctu_h, ctu_flags ? (const uint8_t*)(dp + rb) : NULL,
                                   (const int8_t*)(dp + rb + fb), (const int32_t*)(dp + rb + 2 * fb), d_out))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + rb + 2 * fb + ob, d_out, sizeof(int16_t) * (size_t)width * height, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out, hp + rb + 2 * fb + ob, sizeof(int16_t) * (size_t)width * height);
  return HMGPU_OK;
}

// ---- deblocking (deblock.cu) -----------------------------------------------------------------------
int hmgpu_deblock(hmgpu_ctx* ctx, int16_t* y, int16_t* cb, int16_t* cr, int width, int height,
                  const uint8_t* bs_ver, const uint8_t* bs_hor, const int8_t* qp, const uint8_t* nofilter,
                  int beta_offset_div2, int tc_offset_div2, int cb_qp_offset, int cr_qp_offset)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (!y || !cb || !cr || !bs_ver || !bs_hor || !qp || !nofilter) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_deblock");
  if (width < 8 || height < 8 || width > 16384 || height > 16384 || (width & 1) || (height & 1))
    return hmgpu_fail(ctx, HMGPU_E_INVALID, "picture of %d x %d samples", width, height);
  if (beta_offset_div2 < -6 || beta_offset_div2 > 6 || tc_offset_div2 < -6 || tc_offset_div2 > 6 || cb_qp_offset < -12 || cb_qp_offset > 12 ||
      cr_qp_offset < -12 || cr_qp_offset > 12) return hmgpu_fail(ctx, HMGPU_E_INVALID, "deblocking / chroma QP offsets out of the ranges of the syntax");
  const size_t nu = (size_t)((width + 3) >> 2) * ((height + 3) >> 2);
  for (size_t i = 0; i < nu; i++)
    if (bs_ver[i] > 2 || bs_hor[i] > 2 || qp[i] < -48 || qp[i] > 51) return hmgpu_fail(ctx, HMGPU_E_INVALID, "unit %zu: boundary strength or QP out of range", i);
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t yb = round_up(sizeof(int16_t) * (size_t)width * height, 256), cbb = round_up(sizeof(int16_t) * (size_t)(width >> 1) * (height >> 1), 256);
  const size_t ub = round_up(nu, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, yb + 2 * cbb + 4 * ub))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, yb + 2 * cbb + 4 * ub))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, y, sizeof(int16_t) * (size_t)width * height);
  memcpy(hp + yb, cb, sizeof(int16_t) * (size_t)(width >> 1) * (height >> 1));
  memcpy(hp + yb + cbb, cr, sizeof(int16_t) * (size_t)(width >> 1) * (height >> 1));
  memcpy(hp + yb + 2 * cbb, bs_ver, nu); memcpy(hp + yb + 2 * cbb + ub, bs_hor, nu);
  memcpy(hp + yb + 2 * cbb + 2 * ub, qp, nu); memcpy(hp + yb + 2 * cbb + 3 * ub, nofilter, nu);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, yb + 2 * cbb + 4 * ub, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_deblock(ctx, (int16_t*)dp, (int16_t*)(dp + yb), (int16_t*)(dp + yb + cbb), width, height, ctx->bit_depth, ctx->bit_depth,
                                 (const uint8_t*)(dp + yb + 2 * cbb), (const uint8_t*)(dp + yb + 2 * cbb + ub), (const int8_t*)(dp + yb + 2 * cbb + 2 * ub),
                                 (const uint8_t*)(dp + yb + 2 * cbb + 3 * ub), beta_offset_div2, tc_offset_div2, cb_qp_offset, cr_qp_offset))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp, dp, yb + 2 * cbb, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(y, hp, sizeof(int16_t) * (size_t)width * height);
  memcpy(cb, hp + yb, sizeof(int16_t) * (size_t)(width >> 1) * (height >> 1));
  memcpy(cr, hp + yb + cbb, sizeof(int16_t) * (size_t)(width >> 1) * (height >> 1));
  return HMGPU_OK;
}

// ---- intra mode pre-selection (intra.cu) ---------------------------------------------------------
int hmgpu_intra_costs(hmgpu_ctx* ctx, const hmgpu_intra_job* jobs, int n_jobs, const int16_t* org_blocks, int n_org_elems,
                      const int16_t* ref_lines, int n_ref_elems, uint32_t* dist)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!jobs || !org_blocks || !ref_lines || !dist || n_jobs < 0 || n_org_elems < 0 || n_ref_elems < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_intra_costs");
  for (int i = 0; i < n_jobs; i++)
  {
    const hmgpu_intra_job& j = jobs[i];
    const int n = j.size;
    if (n != 4 && n != 8 && n != 16 && n != 32 && n != 64) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: size %d not in {4,8,16,32,64}", i, n);
    if (j.flags & ~31u) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: unknown flag bits", i);
    if ((unsigned long long)j.org_offset + (unsigned long long)n * n > (unsigned long long)n_org_elems) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: source block outside org_blocks", i);
    if ((unsigned long long)j.ref_offset + 2ull * (4 * n + 1) > (unsigned long long)n_ref_elems) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: reference lines outside ref_lines", i);
  }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t b0 = round_up(sizeof(int16_t) * (size_t)n_org_elems, 256), b1 = round_up(sizeof(int16_t) * (size_t)n_ref_elems, 256);
  const size_t b2 = round_up(sizeof(hmgpu_intra_job) * (size_t)n_jobs, 256), b3 = round_up(sizeof(uint32_t) * 35 * (size_t)n_jobs, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1 + b2 + b3))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1 + b2 + b3))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
  memcpy(hp + b0, ref_lines, sizeof(int16_t) * (size_t)n_ref_elems);
  memcpy(hp + b0 + b1, jobs, sizeof(hmgpu_intra_job) * (size_t)n_jobs);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0 + b1 + b2, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_intra_costs(ctx, (const hmgpu_intra_job*)(dp + b0 + b1), n_jobs, (const int16_t*)dp, (const int16_t*)(dp + b0),
                                     (uint32_t*)(dp + b0 + b1 + b2)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0 + b1 + b2, dp + b0 + b1 + b2, sizeof(uint32_t) * 35 * (size_t)n_jobs, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(dist, hp + b0 + b1 + b2, sizeof(uint32_t) * 35 * (size_t)n_jobs);
  return HMGPU_OK;
}

// ---- motion compensation -----------------------------------------------------------------------

int hmgpu_mc_luma(hmgpu_ctx* ctx, const hmgpu_mc_job* jobs, int n_jobs, int16_t* dst, int n_dst)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  HMGPU_NOT_REMOTE(ctx, "hmgpu_mc_luma");
  if (!jobs || !dst || n_jobs < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  for (int i = 0; i < n_jobs; i++)
  {
    const hmgpu_mc_job& j = jobs[i];
    if (j.ref_slot >= ctx->max_refs || !ctx->refs[j.ref_slot].valid) return hmgpu_fail(ctx, HMGPU_E_STATE, "job %d: slot %d not uploaded", i, j.ref_slot);
    const int x0 = j.pu_x + (j.mv_x >> 2), y0 = j.pu_y + (j.mv_y >> 2);
    if (x0 < -HMGPU_MARGIN + 4 || y0 < -HMGPU_MARGIN + 4 || x0 + j.pu_w > ctx->pic_w + HMGPU_MARGIN - 4 || y0 + j.pu_h > ctx->pic_h + HMGPU_MARGIN - 4)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: MV reaches outside the padded reference", i);
    if ((size_t)j.dst_offset + (size_t)j.pu_w * j.pu_h > (size_t)n_dst) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: dst block outside dst", i);
  }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t b0 = round_up(sizeof(hmgpu_mc_job) * (size_t)n_jobs, 256), b1 = round_up(sizeof(int16_t) * (size_t)n_dst, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, jobs, sizeof(hmgpu_mc_job) * (size_t)n_jobs);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  HMGPU_CUDA(ctx, cudaMemsetAsync(dp + b0, 0, b1, ctx->stream));
  if ((rc = hmgpu_launch_mc_luma(ctx, (const hmgpu_mc_job*)dp, n_jobs, (int16_t*)(dp + b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(int16_t) * (size_t)n_dst, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(dst, hp + b0, sizeof(int16_t) * (size_t)n_dst);
  return HMGPU_OK;
}

// PU prediction (luma + chroma, uni / bi) and prediction-error costs -------------------------------------------------

static int validate_pred_jobs(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n, bool chroma, int n_dst, bool need_dst)
{
  for (int i = 0; i < n; i++)
  {
    const hmgpu_pred_job& j = jobs[i];
    if (j.pu_w < 4 || j.pu_w > 64 || j.pu_h < 4 || j.pu_h > 64 || (j.pu_w & 3) || (j.pu_h & 3))
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: PU %dx%d unsupported", i, j.pu_w, j.pu_h);
    if (j.pu_x < 0 || j.pu_y < 0 || j.pu_x + j.pu_w > ctx->pic_w || j.pu_y + j.pu_h > ctx->pic_h || (j.pu_x & 1) || (j.pu_y & 1))
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: PU at (%d,%d) outside the picture or odd", i, j.pu_x, j.pu_y);
    if (j.ref_slot[0] < 0 && j.ref_slot[1] < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: no reference list used", i);
    for (int l = 0; l < 2; l++)
    {
      if (j.ref_slot[l] < 0) continue;
      if (j.ref_slot[l] >= ctx->max_refs || !ctx->refs[j.ref_slot[l]].valid)
        return hmgpu_fail(ctx, HMGPU_E_STATE, "job %d: reference slot %d not uploaded", i, j.ref_slot[l]);
      if (chroma && !ctx->refs[j.ref_slot[l]].has_chroma)
        return hmgpu_fail(ctx, HMGPU_E_STATE, "job %d: reference slot %d was uploaded without chroma", i, j.ref_slot[l]);
      // 8-tap luma support: 3 samples before, 4 after; 4-tap chroma: 1 before, 2 after
      const int x0 = j.pu_x + (j.mv_x[l] >> 2) - 3, x1 = j.pu_x + j.pu_w + (j.mv_x[l] >> 2) + 4;
      const int y0 = j.pu_y + (j.mv_y[l] >> 2) - 3, y1 = j.pu_y + j.pu_h + (j.mv_y[l] >> 2) + 4;
      if (x0 < -HMGPU_MARGIN || y0 < -HMGPU_MARGIN || x1 > ctx->pic_w + HMGPU_MARGIN || y1 > ctx->pic_h + HMGPU_MARGIN)
        return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: MV reaches outside the padded reference", i);
      if (chroma)
      {
        const int cx0 = (j.pu_x >> 1) + (j.mv_x[l] >> 3) - 1, cx1 = ((j.pu_x + j.pu_w) >> 1) + (j.mv_x[l] >> 3) + 2;
        const int cy0 = (j.pu_y >> 1) + (j.mv_y[l] >> 3) - 1, cy1 = ((j.pu_y + j.pu_h) >> 1) + (j.mv_y[l] >> 3) + 2;
        if (cx0 < -HMGPU_CMARGIN || cy0 < -HMGPU_CMARGIN || cx1 > ctx->pic_w / 2 + HMGPU_CMARGIN || cy1 > ctx->pic_h / 2 + HMGPU_CMARGIN)
          return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: MV reaches outside the padded chroma reference", i);
      }
    }
    const size_t need = (size_t)j.pu_w * j.pu_h * (chroma ? 3 : 2) / 2;
    if (need_dst && (size_t)j.dst_offset + need > (size_t)n_dst) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: output block outside dst", i);
  }
  return HMGPU_OK;
}

int hmgpu_predict(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n_jobs, int with_chroma, int16_t* dst, int n_dst)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!jobs || !dst || n_jobs < 0 || n_dst < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  int rc = validate_pred_jobs(ctx, jobs, n_jobs, with_chroma != 0, n_dst, true);
  if (rc) return rc;
  if (ctx->remote)
  {
    if ((rc = hmgpu_server_stop(ctx))) return rc;
    size_t cap = 0;
    char* area = (char*)hmgpu_remote_area(ctx, 1, &cap);
    const size_t jb = round_up(sizeof(hmgpu_pred_job) * (size_t)n_jobs, 256);
    if (jb + sizeof(int16_t) * (size_t)n_dst > cap) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "prediction batch does not fit the broker's batch area");
    memcpy(area, jobs, sizeof(hmgpu_pred_job) * (size_t)n_jobs);
    const int32_t a[6] = { n_jobs, with_chroma, n_dst, 0, 0, 0 };
    if ((rc = hmgpu_remote_call(ctx, HMB_OP_PREDICT, a, NULL, NULL))) return rc;
    memcpy(dst, area + jb, sizeof(int16_t) * (size_t)n_dst);
    return HMGPU_OK;
  }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t b0 = round_up(sizeof(hmgpu_pred_job) * (size_t)n_jobs, 256), b1 = round_up(sizeof(int16_t) * (size_t)n_dst, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, jobs, sizeof(hmgpu_pred_job) * (size_t)n_jobs);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  HMGPU_CUDA(ctx, cudaMemsetAsync(dp + b0, 0, b1, ctx->stream));
  if ((rc = hmgpu_launch_predict(ctx, (const hmgpu_pred_job*)dp, n_jobs, with_chroma, (int16_t*)(dp + b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(int16_t) * (size_t)n_dst, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(dst, hp + b0, sizeof(int16_t) * (size_t)n_dst);
  return HMGPU_OK;
}

// f2: motion compensation of every merge candidate + SSE of the skip reconstruction per component: the prediction kernel into
// device scratch, then the distortion-table kernel over (source block, predicted block) pairs -- nothing leaves the device in
// between
int hmgpu_merge_skip_dist(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n_jobs, const uint32_t* org_offset,
                          const int16_t* org_blocks, int n_org_elems, int16_t* pred, int n_pred, uint32_t* sse)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!jobs || !org_offset || !org_blocks || !sse || n_jobs < 0 || n_org_elems < 0 || n_pred < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_merge_skip_dist");
  int rc = validate_pred_jobs(ctx, jobs, n_jobs, true, n_pred, true);
  if (rc) return rc;
  for (int i = 0; i < n_jobs; i++)
  {
    const size_t need = (size_t)jobs[i].pu_w * jobs[i].pu_h * 3 / 2;
    if ((jobs[i].pu_w & 7) || (jobs[i].pu_h & 7)) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: a CU is a multiple of 8 samples wide and high", i);
    if ((size_t)org_offset[i] + need > (size_t)n_org_elems) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: source CU outside org_blocks", i);
  }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n_items = (size_t)n_jobs * 3;
  const size_t b0 = round_up(sizeof(hmgpu_pred_job) * (size_t)n_jobs, 256), b1 = round_up(sizeof(int16_t) * (size_t)n_pred, 256);
  const size_t b2 = round_up(sizeof(int16_t) * (size_t)n_org_elems, 256), b3 = round_up(sizeof(hmgpu_dist_item) * n_items, 256);
  const size_t b4 = round_up(sizeof(uint32_t) * n_items, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1 + b2 + b3 + b4))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1 + b2 + b3 + b4))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, jobs, sizeof(hmgpu_pred_job) * (size_t)n_jobs);
  memcpy(hp + b0 + b1, org_blocks, sizeof(int16_t) * (size_t)n_org_elems);
  hmgpu_dist_item* items = (hmgpu_dist_item*)(hp + b0 + b1 + b2);
  for (int i = 0; i < n_jobs; i++)
  {
    const int w = jobs[i].pu_w, h = jobs[i].pu_h;
    uint32_t oo = org_offset[i], po = jobs[i].dst_offset;
    for (int c = 0; c < 3; c++)
    {
      const int cw = c ? w >> 1 : w, ch = c ? h >> 1 : h;
      hmgpu_dist_item it;
      it.org_offset = oo; it.cur_offset = po; it.org_stride = cw; it.cur_stride = cw;
      it.w = (uint8_t)cw; it.h = (uint8_t)ch; it.func = HMGPU_DF_SSE; it.sub_shift = 0;
      items[i * 3 + c] = it;
      oo += (uint32_t)(cw * ch); po += (uint32_t)(cw * ch);
    }
  }
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp + b0 + b1, hp + b0 + b1, b2 + b3, cudaMemcpyHostToDevice, ctx->stream));
  HMGPU_CUDA(ctx, cudaMemsetAsync(dp + b0, 0, b1, ctx->stream));
  if ((rc = hmgpu_launch_predict(ctx, (const hmgpu_pred_job*)dp, n_jobs, 1, (int16_t*)(dp + b0)))) return rc;
  if ((rc = hmgpu_launch_dist(ctx, (const int16_t*)(dp + b0 + b1), (const int16_t*)(dp + b0), (const hmgpu_dist_item*)(dp + b0 + b1 + b2),
                              (int)n_items, (uint32_t*)(dp + b0 + b1 + b2 + b3)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0 + b1 + b2 + b3, dp + b0 + b1 + b2 + b3, sizeof(uint32_t) * n_items, cudaMemcpyDeviceToHost, ctx->stream));
  if (pred) HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(int16_t) * (size_t)n_pred, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(sse, hp + b0 + b1 + b2 + b3, sizeof(uint32_t) * n_items);
  if (pred) memcpy(pred, hp + b0, sizeof(int16_t) * (size_t)n_pred);
  return HMGPU_OK;
}

int hmgpu_pred_error(hmgpu_ctx* ctx, const hmgpu_pred_job* jobs, int n_jobs, int func, uint32_t* out)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!jobs || !out || n_jobs < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  if (func != HMGPU_DF_SAD && func != HMGPU_DF_SAD_GENERIC && func != HMGPU_DF_HADS)
    return hmgpu_fail(ctx, HMGPU_E_INVALID, "func %d: only SAD and HADS prediction errors exist in the reference", func);
  int rc = validate_pred_jobs(ctx, jobs, n_jobs, false, 0, false);
  if (rc) return rc;
  if (ctx->remote)
  {
    size_t cap = 0;
    char* area = (char*)hmgpu_remote_area(ctx, 1, &cap);
    const size_t jb = round_up(sizeof(hmgpu_pred_job) * (size_t)n_jobs, 256);
    if (jb + sizeof(uint32_t) * (size_t)n_jobs > cap) return hmgpu_fail(ctx, HMGPU_E_NOMEM, "prediction batch does not fit the broker's batch area");
    memcpy(area, jobs, sizeof(hmgpu_pred_job) * (size_t)n_jobs);
    const int32_t a[6] = { n_jobs, func, 0, 0, 0, 0 };
    if ((rc = hmgpu_remote_call(ctx, HMB_OP_PRED_ERROR, a, NULL, NULL))) return rc;
    memcpy(out, area + jb, sizeof(uint32_t) * (size_t)n_jobs);
    return HMGPU_OK;
  }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t b0 = round_up(sizeof(hmgpu_pred_job) * (size_t)n_jobs, 256), b1 = round_up(sizeof(uint32_t) * (size_t)n_jobs, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, jobs, sizeof(hmgpu_pred_job) * (size_t)n_jobs);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_pred_error(ctx, (const hmgpu_pred_job*)dp, n_jobs, func == HMGPU_DF_HADS ? 1 : 0, (uint32_t*)(dp + b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(uint32_t) * (size_t)n_jobs, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out, hp + b0, sizeof(uint32_t) * (size_t)n_jobs);
  return HMGPU_OK;
}

// ---- transform / quant -------------------------------------------------------------------------

int hmgpu_fwd_transform(hmgpu_ctx* ctx, const int16_t* resi, int n_tus, int n, int use_dst, int32_t* coeff)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_tus == 0) return HMGPU_OK;
  if (!resi || !coeff || n_tus < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  if (n != 4 && n != 8 && n != 16 && n != 32) return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  HMGPU_NOT_REMOTE(ctx, "hmgpu_fwd_transform");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t elems = (size_t)n_tus * n * n;
  const size_t b0 = round_up(sizeof(int16_t) * elems, 256), b1 = round_up(sizeof(int32_t) * elems, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, resi, sizeof(int16_t) * elems);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_fwd_transform(ctx, (const int16_t*)dp, n_tus, n, use_dst, (int32_t*)(dp + b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(int32_t) * elems, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(coeff, hp + b0, sizeof(int32_t) * elems);
  return HMGPU_OK;
}

int hmgpu_inv_transform(hmgpu_ctx* ctx, const int32_t* coeff, int n_tus, int n, int use_dst, int16_t* resi)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_tus == 0) return HMGPU_OK;
  if (!resi || !coeff || n_tus < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  if (n != 4 && n != 8 && n != 16 && n != 32) return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  HMGPU_NOT_REMOTE(ctx, "hmgpu_inv_transform");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t elems = (size_t)n_tus * n * n;
  const size_t b0 = round_up(sizeof(int32_t) * elems, 256), b1 = round_up(sizeof(int16_t) * elems, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, b0 + b1))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, b0 + b1))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, coeff, sizeof(int32_t) * elems);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_inv_transform(ctx, (const int32_t*)dp, n_tus, n, use_dst, (int16_t*)(dp + b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(int16_t) * elems, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(resi, hp + b0, sizeof(int16_t) * elems);
  return HMGPU_OK;
}

int hmgpu_quant(hmgpu_ctx* ctx, const int32_t* coeff, int n_tus, int n, int qp_per, int qp_rem,
                int is_intra_slice, int32_t* level, int32_t* delta_u, uint32_t* abs_sum)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_tus == 0) return HMGPU_OK;
  if (!coeff || !level || !abs_sum || n_tus < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  if (n != 4 && n != 8 && n != 16 && n != 32) return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  if (qp_rem < 0 || qp_rem > 5 || qp_per < 0 || qp_per > 12) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad qp per/rem");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_quant");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t elems = (size_t)n_tus * n * n;
  const size_t b0 = round_up(sizeof(int32_t) * elems, 256), b3 = round_up(sizeof(uint32_t) * (size_t)n_tus, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, 3 * b0 + b3))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, 3 * b0 + b3))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, coeff, sizeof(int32_t) * elems);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_quant(ctx, (const int32_t*)dp, n_tus, n, qp_per, qp_rem, is_intra_slice, (int32_t*)(dp + b0),
                               (int32_t*)(dp + 2 * b0), (uint32_t*)(dp + 3 * b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, 2 * b0 + sizeof(uint32_t) * (size_t)n_tus, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(level, hp + b0, sizeof(int32_t) * elems);
  if (delta_u) memcpy(delta_u, hp + 2 * b0, sizeof(int32_t) * elems);
  memcpy(abs_sum, hp + 3 * b0, sizeof(uint32_t) * (size_t)n_tus);
  return HMGPU_OK;
}

// ---- rate-distortion optimised quantisation (rdoq.cu) ----------------------------------------------------------------------
static int rdoq_check_jobs(hmgpu_ctx* ctx, const hmgpu_rdoq_job* jobs, int n_jobs, int n_bits, long long n_coef, int n_class[4])
{
  for (int i = 0; i < n_jobs; i++)
  {
    const hmgpu_rdoq_job& j = jobs[i];
    if (j.log2_size < 2 || j.log2_size > 5) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: log2 size %d not in 2..5", i, j.log2_size);
    if ((unsigned)j.channel > 1u || (unsigned)j.scan > 2u || (j.flags & ~HMGPU_RDOQ_SIGN_HIDE)) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: bad channel / scan / flags", i);
    if (j.qbits < 9 || j.qbits > 30 || (unsigned)j.qp_rem > 5u || (unsigned)j.qp_per > 12u) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: bad qbits / qp per / rem", i);
    if ((unsigned)j.go_rice_init > 4u || j.bit_depth < 8 || j.bit_depth > 16) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: bad Rice parameter / bit depth", i);
    if ((unsigned)j.bits_index >= (unsigned)n_bits) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: bits_index %d outside bits", i, j.bits_index);
    if (!(j.lambda > 0.0) || !(j.err_scale > 0.0)) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: lambda / err_scale not positive", i);
    if ((unsigned long long)j.coef_offset + (1ull << (2 * j.log2_size)) > (unsigned long long)n_coef) return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: coefficients outside coef", i);
    n_class[j.log2_size - 2]++;
  }
  return HMGPU_OK;
}
static const uint16_t* rdoq_scan_host()
{
  static uint16_t s_scan[4336];
  static bool s_scan_done = false;                       // (written with the same values by whoever comes first)
  if (!s_scan_done) { hmgpu_rdoq_scan_table(s_scan); s_scan_done = true; }
  return s_scan;
}

int hmgpu_rdoq(hmgpu_ctx* ctx, const hmgpu_rdoq_job* jobs, int n_jobs, const hmgpu_rdoq_bits* bits, int n_bits,
               const int32_t* coef, int n_coef, int32_t* level, int32_t* abs_sum)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_jobs == 0) return HMGPU_OK;
  if (!jobs || !bits || !coef || !level || !abs_sum || n_jobs < 0 || n_bits <= 0 || n_coef < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_rdoq");
  if (ctx->pend_n || ctx->pend_np || ctx->defer_n || ctx->defer_np) return hmgpu_fail(ctx, HMGPU_E_STATE, "a submitted search is outstanding: wait for it first");
  int n_class[4] = { 0, 0, 0, 0 };
  { const int rcv = rdoq_check_jobs(ctx, jobs, n_jobs, n_bits, n_coef, n_class); if (rcv) return rcv; }
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint16_t* s_scan = rdoq_scan_host();
  // staging: [coef | jobs | bits | scan table | index list] in, [level | abs_sum] out
  const size_t b_coef = round_up(sizeof(int32_t) * (size_t)n_coef, 256), b_jobs = round_up(sizeof(hmgpu_rdoq_job) * (size_t)n_jobs, 256);
  const size_t b_bits = round_up(sizeof(hmgpu_rdoq_bits) * (size_t)n_bits, 256), b_scan = round_up(sizeof(uint16_t) * 4336, 256);
  const size_t b_list = round_up(sizeof(int) * (size_t)n_jobs, 256);
  const size_t in_bytes = b_coef + b_jobs + b_bits + b_scan + b_list, out_bytes = b_coef + b_list;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, in_bytes + out_bytes))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, in_bytes + out_bytes))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  // page-locked caller buffers (hmgpu_host_alloc) are copied directly; pageable ones go through the pinned staging buffer
  const bool coef_pinned = is_pinned(coef), level_pinned = is_pinned(level);
  if (coef_pinned) HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, coef, sizeof(int32_t) * (size_t)n_coef, cudaMemcpyHostToDevice, ctx->stream));
  else memcpy(hp, coef, sizeof(int32_t) * (size_t)n_coef);
  memcpy(hp + b_coef, jobs, sizeof(hmgpu_rdoq_job) * (size_t)n_jobs);
  memcpy(hp + b_coef + b_jobs, bits, sizeof(hmgpu_rdoq_bits) * (size_t)n_bits);
  memcpy(hp + b_coef + b_jobs + b_bits, s_scan, sizeof(uint16_t) * 4336);
  int* list = (int*)(hp + b_coef + b_jobs + b_bits + b_scan);
  int at[4] = { 0, n_class[0], n_class[0] + n_class[1], n_class[0] + n_class[1] + n_class[2] };
  for (int i = 0; i < n_jobs; i++) list[at[jobs[i].log2_size - 2]++] = i;
  if (coef_pinned) HMGPU_CUDA(ctx, cudaMemcpyAsync(dp + b_coef, hp + b_coef, in_bytes - b_coef, cudaMemcpyHostToDevice, ctx->stream));
  else HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  // coefficients no job covers come back as zero levels
  HMGPU_CUDA(ctx, cudaMemsetAsync(dp + in_bytes, 0, out_bytes, ctx->stream));
  if ((rc = hmgpu_launch_rdoq(ctx, (const hmgpu_rdoq_job*)(dp + b_coef), (const int*)(dp + b_coef + b_jobs + b_bits + b_scan), n_class,
                              (const hmgpu_rdoq_bits*)(dp + b_coef + b_jobs), (const uint16_t*)(dp + b_coef + b_jobs + b_bits),
                              (const int32_t*)dp, (int32_t*)(dp + in_bytes), (int32_t*)(dp + in_bytes + b_coef)))) return rc;
  if (level_pinned)
  {
    HMGPU_CUDA(ctx, cudaMemcpyAsync(level, dp + in_bytes, sizeof(int32_t) * (size_t)n_coef, cudaMemcpyDeviceToHost, ctx->stream));
    HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + in_bytes + b_coef, dp + in_bytes + b_coef, b_list, cudaMemcpyDeviceToHost, ctx->stream));
  }
  else HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + in_bytes, dp + in_bytes, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (!level_pinned) memcpy(level, hp + in_bytes, sizeof(int32_t) * (size_t)n_coef);
  memcpy(abs_sum, hp + in_bytes + b_coef, sizeof(int32_t) * (size_t)n_jobs);
  return HMGPU_OK;
}

// ---- dequantiser (transform.cu) and the residual-costing loop of a batch of TUs without leaving the device -----------------------
int hmgpu_dequant(hmgpu_ctx* ctx, const int32_t* level, int n_tus, int n, int qp_per, int qp_rem, int32_t* coef)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_tus == 0) return HMGPU_OK;
  if (!level || !coef || n_tus < 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  if (n != 4 && n != 8 && n != 16 && n != 32) return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  if (qp_rem < 0 || qp_rem > 5 || qp_per < 0 || qp_per > 12) return hmgpu_fail(ctx, HMGPU_E_INVALID, "bad qp per/rem");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_dequant");
  if (ctx->pend_n || ctx->pend_np || ctx->defer_n || ctx->defer_np) return hmgpu_fail(ctx, HMGPU_E_STATE, "a submitted search is outstanding: wait for it first");
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t elems = (size_t)n_tus * n * n, b0 = round_up(sizeof(int32_t) * elems, 256);
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, 2 * b0))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, 2 * b0))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, level, sizeof(int32_t) * elems);
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, b0, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = hmgpu_launch_dequant(ctx, (const int32_t*)dp, n_tus, n, NULL, qp_per, qp_rem, (int32_t*)(dp + b0)))) return rc;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(hp + b0, dp + b0, sizeof(int32_t) * elems, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(coef, hp + b0, sizeof(int32_t) * elems);
  return HMGPU_OK;
}

int hmgpu_residual_tus(hmgpu_ctx* ctx, const int16_t* resi, int n_tus, int n, int use_dst, const hmgpu_rdoq_job* jobs,
                       const hmgpu_rdoq_bits* bits, int n_bits, int32_t* level, int32_t* abs_sum, int16_t* rec_resi, uint32_t* dist)
{
  if (!ctx) return HMGPU_E_INVALID;
  if (n_tus == 0) return HMGPU_OK;
  if (!resi || !jobs || !bits || !level || !abs_sum || !rec_resi || !dist || n_tus < 0 || n_bits <= 0) return hmgpu_fail(ctx, HMGPU_E_INVALID, "NULL argument");
  if (n != 4 && n != 8 && n != 16 && n != 32) return hmgpu_fail(ctx, HMGPU_E_INVALID, "transform size %d not in {4,8,16,32}", n);
  if (use_dst && n != 4) return hmgpu_fail(ctx, HMGPU_E_INVALID, "the DST exists for 4x4 TUs only");
  HMGPU_NOT_REMOTE(ctx, "hmgpu_residual_tus");
  if (ctx->pend_n || ctx->pend_np || ctx->defer_n || ctx->defer_np) return hmgpu_fail(ctx, HMGPU_E_STATE, "a submitted search is outstanding: wait for it first");
  const int nn = n * n, log2n = n == 4 ? 2 : n == 8 ? 3 : n == 16 ? 4 : 5;
  const size_t elems = (size_t)n_tus * nn;
  if (elems > 0x7fffffffu) return hmgpu_fail(ctx, HMGPU_E_INVALID, "%d TUs of %d x %d: more than 2^31 coefficients in one call", n_tus, n, n);
  int n_class[4] = { 0, 0, 0, 0 };
  { const int rcv = rdoq_check_jobs(ctx, jobs, n_tus, n_bits, (long long)elems, n_class); if (rcv) return rcv; }
  for (int i = 0; i < n_tus; i++)
    if (jobs[i].log2_size != log2n || jobs[i].coef_offset != (uint32_t)((size_t)i * nn) || jobs[i].bit_depth != ctx->bit_depth)
      return hmgpu_fail(ctx, HMGPU_E_INVALID, "job %d: log2_size / coef_offset / bit_depth do not describe TU %d of %d x %d in a %d-bit context", i, i, n, n, ctx->bit_depth);
  HMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint16_t* s_scan = rdoq_scan_host();
  // staging, in:  [resi | jobs | bits | scan table | index list | distortion items]
  //          work:[coefficients | dequantised coefficients]     out: [levels | abs_sum | reconstructed residual + one zero block | distortions]
  const size_t b_resi = round_up(sizeof(int16_t) * elems, 256), b_jobs = round_up(sizeof(hmgpu_rdoq_job) * (size_t)n_tus, 256);
  const size_t b_bits = round_up(sizeof(hmgpu_rdoq_bits) * (size_t)n_bits, 256), b_scan = round_up(sizeof(uint16_t) * 4336, 256);
  const size_t b_list = round_up(sizeof(int) * (size_t)n_tus, 256), b_items = round_up(sizeof(hmgpu_dist_item) * 2 * (size_t)n_tus, 256);
  const size_t b_c32 = round_up(sizeof(int32_t) * elems, 256), b_rec = round_up(sizeof(int16_t) * (elems + nn), 256), b_dist = round_up(sizeof(uint32_t) * 2 * (size_t)n_tus, 256);
  const size_t in_bytes = b_resi + b_jobs + b_bits + b_scan + b_list + b_items, work_bytes = 2 * b_c32, out_bytes = b_c32 + b_list + b_rec + b_dist;
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  if ((rc = hmgpu_reserve_pinned(ctx, in_bytes + out_bytes))) return rc;
  if ((rc = hmgpu_reserve_stage(ctx, in_bytes + work_bytes + out_bytes))) return rc;
  char* hp = (char*)ctx->h_pin; char* dp = (char*)ctx->d_stage;
  memcpy(hp, resi, sizeof(int16_t) * elems);
  size_t at = b_resi;
  memcpy(hp + at, jobs, sizeof(hmgpu_rdoq_job) * (size_t)n_tus); const size_t o_jobs = at; at += b_jobs;
  memcpy(hp + at, bits, sizeof(hmgpu_rdoq_bits) * (size_t)n_bits); const size_t o_bits = at; at += b_bits;
  memcpy(hp + at, s_scan, sizeof(uint16_t) * 4336); const size_t o_scan = at; at += b_scan;
  int* list = (int*)(hp + at); const size_t o_list = at; at += b_list;
  for (int i = 0; i < n_tus; i++) list[i] = i;
  hmgpu_dist_item* items = (hmgpu_dist_item*)(hp + at); const size_t o_items = at; at += b_items;
  for (int i = 0; i < n_tus; i++)
  {
    // the residual against its reconstruction, and against nothing coded (a block of zeros behind the reconstructions)
    hmgpu_dist_item a = { (uint32_t)((size_t)i * nn), (uint32_t)((size_t)i * nn), n, n, (uint8_t)n, (uint8_t)n, HMGPU_DF_SSE, 0 };
    hmgpu_dist_item z = { (uint32_t)((size_t)i * nn), (uint32_t)elems, n, n, (uint8_t)n, (uint8_t)n, HMGPU_DF_SSE, 0 };
    items[2 * i] = a; items[2 * i + 1] = z;
  }
  const size_t o_coef = in_bytes, o_deq = in_bytes + b_c32, o_level = in_bytes + work_bytes, o_sum = o_level + b_c32, o_rec = o_sum + b_list, o_dist = o_rec + b_rec;
  HMGPU_CUDA(ctx, cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  HMGPU_CUDA(ctx, cudaMemsetAsync(dp + o_level, 0, b_c32 + b_list, ctx->stream));          // (the RDOQ kernel stores the non-zero levels only)
  HMGPU_CUDA(ctx, cudaMemsetAsync(dp + o_rec + sizeof(int16_t) * elems, 0, sizeof(int16_t) * nn, ctx->stream));
  const int16_t* d_resi = (const int16_t*)dp;
  // transformNxN: xT, then the RDOQ branch of xQuant
  if ((rc = hmgpu_launch_fwd_transform(ctx, d_resi, n_tus, n, use_dst, (int32_t*)(dp + o_coef)))) return rc;
  if ((rc = hmgpu_launch_rdoq(ctx, (const hmgpu_rdoq_job*)(dp + o_jobs), (const int*)(dp + o_list), n_class, (const hmgpu_rdoq_bits*)(dp + o_bits),
                              (const uint16_t*)(dp + o_scan), (const int32_t*)(dp + o_coef), (int32_t*)(dp + o_level), (int32_t*)(dp + o_sum)))) return rc;
  // invTransformNxN: xDeQuant, then xIT (all-zero TUs come out as zero residuals, as the reference leaves them)
  if ((rc = hmgpu_launch_dequant(ctx, (const int32_t*)(dp + o_level), n_tus, n, (const hmgpu_rdoq_job*)(dp + o_jobs), 0, 0, (int32_t*)(dp + o_deq)))) return rc;
  if ((rc = hmgpu_launch_inv_transform(ctx, (const int32_t*)(dp + o_deq), n_tus, n, use_dst, (int16_t*)(dp + o_rec)))) return rc;
  // the two distortions xEstimateResidualQT compares
  if ((rc = hmgpu_launch_dist(ctx, d_resi, (const int16_t*)(dp + o_rec), (const hmgpu_dist_item*)(dp + o_items), 2 * n_tus, (uint32_t*)(dp + o_dist)))) return rc;
  char* ho = hp + in_bytes;                                         // the pinned copy of the out region
  HMGPU_CUDA(ctx, cudaMemcpyAsync(ho, dp + o_level, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(level, ho, sizeof(int32_t) * elems);
  memcpy(abs_sum, ho + (o_sum - o_level), sizeof(int32_t) * (size_t)n_tus);
  memcpy(rec_resi, ho + (o_rec - o_level), sizeof(int16_t) * elems);
  memcpy(dist, ho + (o_dist - o_level), sizeof(uint32_t) * 2 * (size_t)n_tus);
  return HMGPU_OK;
}

} // extern "C"
