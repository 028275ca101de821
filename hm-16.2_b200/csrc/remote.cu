// remote.cu -- libhmgpu as the CLIENT of a per-GPU broker daemon (hmgpud.cu).  Pure host code: a process in this mode
// (HMGPU_BROKER=<socket path> in its environment) never initialises CUDA.  See broker_proto.h for the division of labour:
// the control socket carries rare operations (one round trip each), the shared-memory segment carries the mailbox the
// daemon's resident server kernel polls, picture uploads and large batches.
#include "hmgpu_internal.cuh"

#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>
#include <new>

struct HmgpuRemote
{
  int    fd;            // control socket
  void*  shm;           // the segment, mapped read/write
  size_t shm_bytes;
};

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

static int raw_call(int fd, int op, const int32_t a[6], const char* text, BrokerReply* reply)
{
  BrokerMsg m;
  memset(&m, 0, sizeof m);
  m.magic = HMGPU_BROKER_MAGIC; m.op = (uint32_t)op;
  if (a) memcpy(m.a, a, sizeof m.a);
  if (text) strncpy(m.text, text, sizeof m.text - 1);
  if (!io_all(fd, &m, sizeof m, true) || !io_all(fd, reply, sizeof *reply, false)) return 1;
  return 0;
}

int hmgpu_remote_call(hmgpu_ctx* ctx, int op, const int32_t a[6], const char* text, BrokerReply* reply)
{
  BrokerReply local;
  if (!reply) reply = &local;
  HmgpuRemote* r = ctx->remote;
  if (r->fd < 0) return hmgpu_fail(ctx, HMGPU_E_STATE, "the broker daemon is gone");
  if (raw_call(r->fd, op, a, text, reply))
  {
    close(r->fd); r->fd = -1;
    return hmgpu_fail(ctx, HMGPU_E_STATE, "lost the broker daemon (control socket: %s)", strerror(errno));
  }
  if (reply->rc != HMGPU_OK)
  {
    reply->text[sizeof reply->text - 1] = 0;
    return hmgpu_fail(ctx, reply->rc, "%s", reply->text);
  }
  return HMGPU_OK;
}

void* hmgpu_remote_area(hmgpu_ctx* ctx, int which, size_t* bytes)
{
  const BrokerShmHeader* h = (const BrokerShmHeader*)ctx->remote->shm;
  *bytes = (size_t)(which == 0 ? h->upload_bytes : h->batch_bytes);
  return (char*)ctx->remote->shm + (which == 0 ? h->upload_off : h->batch_off);
}

int hmgpu_remote_create(const char* socket_path, int pic_w, int pic_h, int bit_depth, int max_refs, hmgpu_ctx** out)
{
  if (pic_w < 8 || pic_h < 8 || (pic_w & 3) || (pic_h & 3) || pic_w > 8184 || pic_h > 8184)
    return hmgpu_fail(NULL, HMGPU_E_INVALID, "picture size %dx%d unsupported (multiples of 4, 8..8184: quarter-pel clipMv bounds are int16)", pic_w, pic_h);
  if (bit_depth < 8 || bit_depth > 12) return hmgpu_fail(NULL, HMGPU_E_INVALID, "bit depth %d unsupported (8..12)", bit_depth);
  if (max_refs < 1 || max_refs > HMGPU_MAX_REFS) return hmgpu_fail(NULL, HMGPU_E_INVALID, "max_refs %d not in 1..%d", max_refs, HMGPU_MAX_REFS);
  sockaddr_un sa;
  memset(&sa, 0, sizeof sa);
  sa.sun_family = AF_UNIX;
  if (strlen(socket_path) >= sizeof sa.sun_path) return hmgpu_fail(NULL, HMGPU_E_INVALID, "HMGPU_BROKER path too long");
  strcpy(sa.sun_path, socket_path);
  // a daemon that is still starting is waited for (HMGPU_BROKER_WAIT_MS, default 10 s); there is no fallback to a private context
  const char* w = getenv("HMGPU_BROKER_WAIT_MS");
  const long wait_ms = w && *w ? atol(w) : 10000;
  int fd = -1;
  for (long waited = 0;; waited += 20)
  {
    fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return hmgpu_fail(NULL, HMGPU_E_STATE, "socket: %s", strerror(errno));
    if (connect(fd, (sockaddr*)&sa, sizeof sa) == 0) break;
    const int e = errno;
    close(fd); fd = -1;
    if ((e != ENOENT && e != ECONNREFUSED) || waited >= wait_ms)
      return hmgpu_fail(NULL, HMGPU_E_STATE, "cannot reach the broker daemon at %s (%s); libhmgpu has no CPU fallback", socket_path, strerror(e));
    struct timespec ts = { 0, 20 * 1000 * 1000 };
    nanosleep(&ts, NULL);
  }
  const int32_t a[6] = { pic_w, pic_h, bit_depth, max_refs, HMGPU_BROKER_PROTO, 0 };
  BrokerReply rep;
  if (raw_call(fd, HMB_OP_CREATE, a, NULL, &rep)) { close(fd); return hmgpu_fail(NULL, HMGPU_E_STATE, "the broker daemon closed the connection"); }
  rep.text[sizeof rep.text - 1] = 0;
  if (rep.rc != HMGPU_OK) { close(fd); return hmgpu_fail(NULL, rep.rc, "broker: %s", rep.text); }
  const int sfd = shm_open(rep.text, O_RDWR, 0);
  if (sfd < 0) { close(fd); return hmgpu_fail(NULL, HMGPU_E_STATE, "shm_open(%s): %s", rep.text, strerror(errno)); }
  void* shm = mmap(NULL, (size_t)rep.v64, PROT_READ | PROT_WRITE, MAP_SHARED, sfd, 0);
  close(sfd);
  if (shm == MAP_FAILED) { close(fd); return hmgpu_fail(NULL, HMGPU_E_NOMEM, "mmap of the broker segment (%llu bytes): %s", (unsigned long long)rep.v64, strerror(errno)); }
  const BrokerShmHeader* h = (const BrokerShmHeader*)shm;
  if (h->magic != HMGPU_BROKER_MAGIC || h->proto != HMGPU_BROKER_PROTO || h->total_bytes != rep.v64 || h->mail_bytes < sizeof(Mailbox))
  {
    munmap(shm, (size_t)rep.v64); close(fd);
    return hmgpu_fail(NULL, HMGPU_E_STATE, "broker segment %s has an unexpected layout (protocol %u, this library speaks %u)", rep.text, h->proto, HMGPU_BROKER_PROTO);
  }
  hmgpu_ctx* ctx = new (std::nothrow) hmgpu_ctx;
  HmgpuRemote* r = new (std::nothrow) HmgpuRemote;
  if (!ctx || !r) { delete ctx; delete r; munmap(shm, (size_t)rep.v64); close(fd); return hmgpu_fail(NULL, HMGPU_E_NOMEM, "out of host memory"); }
  memset(ctx, 0, sizeof *ctx);
  r->fd = fd; r->shm = shm; r->shm_bytes = (size_t)rep.v64;
  ctx->remote = r;
  ctx->device = -1; ctx->pic_w = pic_w; ctx->pic_h = pic_h; ctx->bit_depth = bit_depth; ctx->max_refs = max_refs;
  ctx->px_bytes = bit_depth == 8 ? 1 : 2;
  ctx->tune.server = 1; ctx->tune.fastpath = 1; ctx->tune.pipeline = 0;
  ctx->tune.server_stats = getenv("HMGPU_SERVER_STATS") != NULL;
  ctx->srv_ctas = rep.v[0];
  ctx->tune.server_idle_us = rep.v[1];
  ctx->h_mail = (char*)shm + h->mail_off;
  ctx->mail_external = true;
  *out = ctx;
  return HMGPU_OK;
}

void hmgpu_remote_destroy(hmgpu_ctx* ctx)
{
  HmgpuRemote* r = ctx->remote;
  if (r->fd >= 0)
  {
    BrokerReply rep;
    raw_call(r->fd, HMB_OP_DESTROY, NULL, NULL, &rep);
    close(r->fd);
  }
  munmap(r->shm, r->shm_bytes);
  delete r;
  delete ctx;
}
