// broker_proto.h -- wire format between an encoder process (libhmgpu in client mode, remote.cu) and the per-GPU broker
// daemon (hmgpud.cu).  Plain C structs, no CUDA types: this header is all a client needs.
//
// One GPU, many encoders (BASELINE.json north_star: "several encoder instances share one GPU"): HM is single-threaded with
// process-global state (TComRom.cpp:245,542), so several encoder instances = several processes.  Giving every process its own
// CUDA context costs 2-4 s of set-up each and time-slices the GPU between their kernels (or needs MPS).  Instead ONE daemon per
// GPU owns the only CUDA context; an encoder attaches over a UNIX socket and gets a POSIX shared-memory segment that the daemon
// has registered with CUDA (cudaHostRegisterMapped):
//
//   segment = [ BrokerShmHeader | Mailbox (job lines, result slots, key-pattern blocks) | upload area | batch area ]
//
// * per-PU searches never touch the socket: the encoder writes its job lines into the segment, the resident server kernel
//   the daemon launched for this client (me_server_kernel, me_single.cu) polls them over PCIe and publishes the results into
//   the segment, where the encoder polls them -- exactly the mailbox protocol of the in-process path, with the encoder as the
//   host side.  All clients' server kernels live in the daemon's context and run concurrently.
// * the socket carries the rare control operations (one round trip each): start / drain the client's server kernel, picture
//   uploads (pixels travel through the upload area), batches too large for the mailbox (through the batch area).
#pragma once
#include <stdint.h>

#define HMGPU_BROKER_MAGIC 0x484d4742u   /* "HMGB" */
#define HMGPU_BROKER_PROTO 1

enum
{
  HMB_OP_CREATE = 1,      // a: pic_w, pic_h, bit_depth, max_refs                         -> v: shm bytes, server CTAs, idle us; text: shm name
  HMB_OP_SERVER_START,    // a: generation, last ticket, dynamic smem bytes               launches the client's resident server kernel
  HMB_OP_SERVER_SYNC,     // wait until the client's server kernel has left (the client has bumped the generation in its lines)
  HMB_OP_SERVER_QUERY,    // -> rc != 0 when the server kernel died with a CUDA error
  HMB_OP_REF_UPLOAD,      // a: slot, has_chroma; upload area: luma [cb cr] tight int16       hmgpu_ref_upload
  HMB_OP_ORG_UPLOAD,      // upload area: luma tight int16                                 hmgpu_org_upload
  HMB_OP_REF_RELEASE,     // a: slot
  HMB_OP_ME_BATCH,        // a: n_jobs, n_org_elems; batch area: jobs | org blocks | results  hmgpu_me_search with the batch kernels
  HMB_OP_PRED_ERROR,      // a: n_jobs, func; batch area: pred jobs | uint32 out           hmgpu_pred_error
  HMB_OP_PREDICT,         // a: n_jobs, with_chroma, n_dst; batch area: pred jobs | dst    hmgpu_predict
  HMB_OP_SET_OPTION,      // a: value; text: option name
  HMB_OP_LAUNCH_COUNT,    // -> v64: kernels launched for this client
  HMB_OP_DESTROY
};

typedef struct BrokerMsg
{
  uint32_t magic, op;
  int32_t  a[6];
  char     text[32];
} BrokerMsg;                /* 64 bytes */

typedef struct BrokerReply
{
  int32_t  rc;              /* HMGPU_OK or HMGPU_E_* */
  int32_t  v[3];
  uint64_t v64;
  char     text[232];       /* error text, or the shm name */
} BrokerReply;              /* 256 bytes */

typedef struct BrokerShmHeader
{
  uint32_t magic, proto;
  uint64_t total_bytes;
  uint64_t mail_off, mail_bytes;       /* struct Mailbox (hmgpu_internal.cuh) */
  uint64_t upload_off, upload_bytes;   /* picture uploads */
  uint64_t batch_off, batch_bytes;     /* large job batches, prediction jobs */
  uint64_t pad[7];
} BrokerShmHeader;          /* 128 bytes */
