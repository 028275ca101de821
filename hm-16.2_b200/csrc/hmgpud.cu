// hmgpud.cu -- the per-GPU broker daemon: ONE process owns the CUDA context of a B200, any number of HM encoder processes
// attach to it (libhmgpu in client mode, remote.cu) and share the GPU without creating a context of their own.
//
//   hmgpud --device 0 --socket /tmp/hmgpud.0.sock [--ctas 4] [--idle-us 2000] [--batch-mb 16]
//
// Why a daemon (BASELINE.json north_star: "per-PU calls ... batched into a persistent kernel fed through a pinned-memory
// mailbox ... several encoder instances share one GPU"; HM is one single-threaded process per encoder, TEncTop.h:78-98):
//   * no per-process CUDA set-up (0.3 - 4 s each, serialised when several processes start together) and no MPS;
//   * the server kernels of all clients live in one context, so they run CONCURRENTLY on disjoint SMs -- kernels of different
//     processes would time-slice the whole GPU;
//   * per-PU searches still cost no system call: each client owns a shared-memory segment, registered here with
//     cudaHostRegisterMapped, whose job lines the client's resident server kernel polls over PCIe (broker_proto.h).
//
// One thread per client, blocked in recv() between the client's rare control operations.  The daemon never frees device or
// pinned memory while it runs (hmgpu_internal_pool_mode: cudaFree synchronises the device and would wait for the other
// clients' resident kernels); segments of departed clients are reused, too.
#include "hmgpu_internal.cuh"

#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct Segment { std::string name; void* ptr; size_t bytes; bool in_use; };

static int         g_device = 0, g_ctas = 4, g_idle_us = 2000, g_batch_mb = 16;
static std::string g_socket;
static int         g_listen = -1;
static std::mutex  g_mu;
static std::vector<Segment> g_segments;
static std::atomic<int>     g_clients(0), g_served(0);

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool io_all(int fd, void* buf, size_t n, bool wr)
{
  char* p = (char*)buf;
  while (n)
  {
    const ssize_t k = wr ? send(fd, p, n, MSG_NOSIGNAL) : recv(fd, p, n, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return false;
    p += k; n -= (size_t)k;
  }
  return true;
}

// a segment of at least `bytes`: a parked one, or a new POSIX shm object registered with CUDA
static Segment* segment_acquire(size_t bytes, char* err, size_t err_len)
{
  std::lock_guard<std::mutex> g(g_mu);
  for (Segment& s : g_segments)
    if (!s.in_use && s.bytes >= bytes && s.bytes <= 2 * bytes) { s.in_use = true; return &s; }
  Segment s;
  char name[64];
  snprintf(name, sizeof name, "/hmgpud.%d.%zu", (int)getpid(), g_segments.size());
  s.name = name; s.bytes = round_up(bytes, 1 << 16); s.in_use = true;
  const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) { snprintf(err, err_len, "shm_open(%s): %s", name, strerror(errno)); return NULL; }
  if (ftruncate(fd, (off_t)s.bytes) != 0) { snprintf(err, err_len, "ftruncate(%s, %zu): %s", name, s.bytes, strerror(errno)); close(fd); shm_unlink(name); return NULL; }
  s.ptr = mmap(NULL, s.bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
  close(fd);
  if (s.ptr == MAP_FAILED) { snprintf(err, err_len, "mmap(%s): %s", name, strerror(errno)); shm_unlink(name); return NULL; }
  const cudaError_t e = cudaHostRegister(s.ptr, s.bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e != cudaSuccess)
  {
    snprintf(err, err_len, "cudaHostRegister(%zu bytes): %s", s.bytes, cudaGetErrorString(e));
    munmap(s.ptr, s.bytes); shm_unlink(name);
    return NULL;
  }
  g_segments.push_back(s);                                  // the vector is reserved up front: pointers into it stay valid
  return &g_segments.back();
}

static void segment_release(Segment* s)
{
  std::lock_guard<std::mutex> g(g_mu);
  s->in_use = false;
}

static void reply_rc(BrokerReply& r, hmgpu_ctx* ctx, int rc)
{
  r.rc = rc;
  if (rc != HMGPU_OK) { strncpy(r.text, hmgpu_last_error(ctx), sizeof r.text - 1); }
}

static void serve(int fd)
{
  cudaSetDevice(g_device);
  hmgpu_ctx* ctx = NULL;
  Segment* seg = NULL;
  BrokerShmHeader* hdr = NULL;
  int pic_w = 0, pic_h = 0;
  for (;;)
  {
    BrokerMsg m;
    if (!io_all(fd, &m, sizeof m, false)) break;            // client gone (or never spoke)
    BrokerReply r;
    memset(&r, 0, sizeof r);
    bool bye = false;
    if (m.magic != HMGPU_BROKER_MAGIC) { r.rc = HMGPU_E_INVALID; snprintf(r.text, sizeof r.text, "bad magic"); bye = true; }
    else if (m.op == HMB_OP_CREATE)
    {
      if (ctx) { r.rc = HMGPU_E_STATE; snprintf(r.text, sizeof r.text, "context already created on this connection"); }
      else if (m.a[4] != HMGPU_BROKER_PROTO) { r.rc = HMGPU_E_INVALID; snprintf(r.text, sizeof r.text, "protocol %d, this daemon speaks %d", m.a[4], HMGPU_BROKER_PROTO); }
      else
      {
        pic_w = m.a[0]; pic_h = m.a[1];
        const size_t mail = round_up(hmgpu_internal_mailbox_bytes(), 4096);
        const size_t upload = pic_w > 0 && pic_h > 0 && pic_w <= 8184 && pic_h <= 8184 ? round_up((size_t)pic_w * pic_h * 3, 4096) : 4096;
        const size_t batch = (size_t)g_batch_mb << 20;
        const size_t total = 4096 + mail + upload + batch;
        seg = segment_acquire(total, r.text, sizeof r.text);
        if (!seg) r.rc = HMGPU_E_NOMEM;
        else
        {
          memset(seg->ptr, 0, 4096 + mail);                 // header + mailbox of a previous client
          hdr = (BrokerShmHeader*)seg->ptr;
          hdr->proto = HMGPU_BROKER_PROTO; hdr->total_bytes = seg->bytes;
          hdr->mail_off = 4096; hdr->mail_bytes = mail;
          hdr->upload_off = 4096 + mail; hdr->upload_bytes = upload;
          hdr->batch_off = 4096 + mail + upload; hdr->batch_bytes = seg->bytes - hdr->batch_off;
          hdr->magic = HMGPU_BROKER_MAGIC;
          const int rc = hmgpu_internal_create_shared(g_device, m.a[0], m.a[1], m.a[2], m.a[3], (char*)seg->ptr + hdr->mail_off, g_ctas, &ctx);
          if (rc != HMGPU_OK) { reply_rc(r, NULL, rc); segment_release(seg); seg = NULL; ctx = NULL; }
          else
          {
            hmgpu_set_option(ctx, "server_idle_us", g_idle_us);
            r.v[0] = g_ctas; r.v[1] = g_idle_us; r.v64 = seg->bytes;
            strncpy(r.text, seg->name.c_str(), sizeof r.text - 1);
            g_clients++; g_served++;
          }
        }
      }
    }
    else if (!ctx) { r.rc = HMGPU_E_STATE; snprintf(r.text, sizeof r.text, "no context on this connection"); }
    else
    {
      char* up = (char*)seg->ptr + hdr->upload_off;
      char* ba = (char*)seg->ptr + hdr->batch_off;
      switch (m.op)
      {
      case HMB_OP_SERVER_START: reply_rc(r, ctx, hmgpu_internal_server_launch(ctx, (uint32_t)m.a[0], (uint32_t)m.a[1], m.a[2])); break;
      case HMB_OP_SERVER_SYNC:  reply_rc(r, ctx, hmgpu_internal_server_sync(ctx)); break;
      case HMB_OP_SERVER_QUERY: reply_rc(r, ctx, hmgpu_internal_server_query(ctx)); break;
      case HMB_OP_REF_UPLOAD:
      {
        const int16_t* y = (const int16_t*)up;
        const int16_t* cb = y + (size_t)pic_w * pic_h;
        const int16_t* cr = cb + (size_t)(pic_w / 2) * (pic_h / 2);
        reply_rc(r, ctx, hmgpu_ref_upload(ctx, m.a[0], y, pic_w, m.a[1] ? cb : NULL, m.a[1] ? cr : NULL, pic_w / 2));
        break;
      }
      case HMB_OP_ORG_UPLOAD:   reply_rc(r, ctx, hmgpu_org_upload(ctx, (const int16_t*)up, pic_w)); break;
      case HMB_OP_REF_RELEASE:  reply_rc(r, ctx, hmgpu_ref_release(ctx, m.a[0])); break;
      case HMB_OP_ME_BATCH:
      {
        const int n = m.a[0], n_org = m.a[1];
        const size_t ob = round_up(n_org > 0 ? sizeof(int16_t) * (size_t)n_org : 0, 256);
        const size_t jb = round_up(sizeof(hmgpu_me_job) * (size_t)(n > 0 ? n : 0), 256);
        if (n < 0 || n_org < 0 || ob + jb + sizeof(hmgpu_me_result) * (size_t)n > hdr->batch_bytes) { r.rc = HMGPU_E_INVALID; snprintf(r.text, sizeof r.text, "batch does not fit the batch area"); break; }
        reply_rc(r, ctx, hmgpu_me_search(ctx, (const hmgpu_me_job*)(ba + ob), n, n_org ? (const int16_t*)ba : NULL, n_org, (hmgpu_me_result*)(ba + ob + jb)));
        break;
      }
      case HMB_OP_PRED_ERROR:
      {
        const int n = m.a[0];
        const size_t jb = round_up(sizeof(hmgpu_pred_job) * (size_t)(n > 0 ? n : 0), 256);
        if (n < 0 || jb + sizeof(uint32_t) * (size_t)n > hdr->batch_bytes) { r.rc = HMGPU_E_INVALID; snprintf(r.text, sizeof r.text, "batch does not fit the batch area"); break; }
        reply_rc(r, ctx, hmgpu_pred_error(ctx, (const hmgpu_pred_job*)ba, n, m.a[1], (uint32_t*)(ba + jb)));
        break;
      }
      case HMB_OP_PREDICT:
      {
        const int n = m.a[0], n_dst = m.a[2];
        const size_t jb = round_up(sizeof(hmgpu_pred_job) * (size_t)(n > 0 ? n : 0), 256);
        if (n < 0 || n_dst < 0 || jb + sizeof(int16_t) * (size_t)n_dst > hdr->batch_bytes) { r.rc = HMGPU_E_INVALID; snprintf(r.text, sizeof r.text, "batch does not fit the batch area"); break; }
        reply_rc(r, ctx, hmgpu_predict(ctx, (const hmgpu_pred_job*)ba, n, m.a[1], (int16_t*)(ba + jb), n_dst));
        break;
      }
      case HMB_OP_SET_OPTION:   m.text[sizeof m.text - 1] = 0; reply_rc(r, ctx, hmgpu_set_option(ctx, m.text, m.a[0])); break;
      case HMB_OP_LAUNCH_COUNT: r.v64 = hmgpu_launch_count(ctx); break;
      case HMB_OP_DESTROY:      bye = true; break;
      default: r.rc = HMGPU_E_INVALID; snprintf(r.text, sizeof r.text, "unknown operation %u", m.op);
      }
    }
    if (!io_all(fd, &r, sizeof r, true) || bye) break;
  }
  if (ctx)
  {
    // whatever state the client left its server kernel in: make it leave, then take the context apart (pooled memory)
    hmgpu_internal_server_kill(ctx);
    hmgpu_destroy(ctx);
    g_clients--;
  }
  if (seg) segment_release(seg);
  close(fd);
}

static void cleanup_and_exit(int)
{
  if (!g_socket.empty()) unlink(g_socket.c_str());
  for (const Segment& s : g_segments) shm_unlink(s.name.c_str());
  _exit(0);
}

int main(int argc, char** argv)
{
  for (int i = 1; i < argc; i++)
  {
    const std::string a = argv[i];
    const char* v = i + 1 < argc ? argv[i + 1] : NULL;
    if (a == "--device" && v) { g_device = atoi(v); i++; }
    else if (a == "--socket" && v) { g_socket = v; i++; }
    else if (a == "--ctas" && v) { g_ctas = atoi(v); i++; }
    else if (a == "--idle-us" && v) { g_idle_us = atoi(v); i++; }
    else if (a == "--batch-mb" && v) { g_batch_mb = atoi(v); i++; }
    else { fprintf(stderr, "usage: hmgpud --device D --socket PATH [--ctas N (1..%d)] [--idle-us U] [--batch-mb M]\n", HMGPU_SERVER_CTAS); return 2; }
  }
  if (g_socket.empty()) { char p[64]; snprintf(p, sizeof p, "/tmp/hmgpud.%d.sock", g_device); g_socket = p; }
  if (g_ctas < 1 || g_ctas > HMGPU_SERVER_CTAS || g_batch_mb < 1) { fprintf(stderr, "hmgpud: bad --ctas / --batch-mb\n"); return 2; }
  g_segments.reserve(4096);
  cudaError_t e = cudaSetDevice(g_device);
  if (e == cudaSuccess) e = cudaFree(0);                    // the ONE context creation of this GPU's encoders
  if (e != cudaSuccess) { fprintf(stderr, "hmgpud: CUDA device %d: %s (no CPU fallback)\n", g_device, cudaGetErrorString(e)); return 1; }
  hmgpu_internal_pool_mode(1);

  signal(SIGPIPE, SIG_IGN);
  signal(SIGTERM, cleanup_and_exit);
  signal(SIGINT, cleanup_and_exit);
  g_listen = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  sockaddr_un sa;
  memset(&sa, 0, sizeof sa);
  sa.sun_family = AF_UNIX;
  if (g_socket.size() >= sizeof sa.sun_path) { fprintf(stderr, "hmgpud: socket path too long\n"); return 2; }
  strcpy(sa.sun_path, g_socket.c_str());
  unlink(g_socket.c_str());
  if (g_listen < 0 || bind(g_listen, (sockaddr*)&sa, sizeof sa) != 0 || listen(g_listen, 256) != 0)
  {
    fprintf(stderr, "hmgpud: cannot listen on %s: %s\n", g_socket.c_str(), strerror(errno));
    return 1;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, g_device);
  printf("hmgpud ready: device %d (%s, %d SMs), socket %s, %d server CTAs per client, idle exit %d us\n", g_device, prop.name,
         prop.multiProcessorCount, g_socket.c_str(), g_ctas, g_idle_us);
  fflush(stdout);
  for (;;)
  {
    const int fd = accept(g_listen, NULL, NULL);
    if (fd < 0) { if (errno == EINTR) continue; break; }
    std::thread(serve, fd).detach();
  }
  cleanup_and_exit(0);
  return 0;
}
