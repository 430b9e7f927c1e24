// comm.cu -- multi-GPU plumbing inside the library: one process per GPU, every context a rank of one NCCL communicator
// (replicated keys).  The data path it serves is the level-sharded match of regex_api.cu: every PBS level of
// the plan is cut into `world` contiguous slices, rank r bootstraps slice r, and the slices are exchanged device to device
// over NVLink (one grouped ncclBroadcast per rank, in place in the ciphertext arena) -- no host hop between levels, and the
// final bitor fold of the reference (/root/reference/src/regex/engine.rs:22-35) is just the last level of the same plan.
//
// NCCL is resolved at run time (dlopen): the library has no link-time dependency on it, a single-GPU user never loads it,
// and inside a process that already carries an NCCL (e.g. PyTorch's) the same instance is shared.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <string>
#include "context.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);   // an instance the process already carries
    if (api.handle) break;
  }
  for (const char* n : names) {
    if (api.handle) break;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!api.handle) {
    api.err = "libnccl.so.2 not found";
    return &api;
  }
#define FB_SYM(field, name)                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
  if (!api.field) api.err = std::string("missing NCCL symbol ") + name;
  FB_SYM(GetUniqueId, "ncclGetUniqueId")
  FB_SYM(CommInitRank, "ncclCommInitRank")
  FB_SYM(CommDestroy, "ncclCommDestroy")
  FB_SYM(Broadcast, "ncclBroadcast")
  FB_SYM(AllGather, "ncclAllGather")
  FB_SYM(GroupStart, "ncclGroupStart")
  FB_SYM(GroupEnd, "ncclGroupEnd")
  FB_SYM(GetErrorString, "ncclGetErrorString")
#undef FB_SYM
  return &api;
}

int nccl_fail(fb_ctx* ctx, ncclResult_t r, const char* what) {
  NcclApi* a = nccl_api();
  return fb_fail(ctx, FB_ERR_CUDA, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}

}  // namespace

static_assert(FB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "fb_comm id is an ncclUniqueId");

extern "C" int fb_comm_unique_id(uint8_t* id) {
  if (!id) return FB_ERR_ARG;
  NcclApi* a = nccl_api();
  if (!a->handle || !a->err.empty()) return FB_ERR_NO_DEVICE;
  ncclUniqueId u;
  if (a->GetUniqueId(&u) != ncclSuccess) return FB_ERR_CUDA;
  std::memcpy(id, u.internal, FB_COMM_ID_BYTES);
  return FB_OK;
}

extern "C" int fb_comm_init(fb_ctx* ctx, const uint8_t* id, int rank, int world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fb_fail(ctx, FB_ERR_ARG, "bad communicator arguments");
  if (ctx->comm) return fb_fail(ctx, FB_ERR_ARG, "context already has a communicator");
  NcclApi* a = nccl_api();
  if (!a->handle || !a->err.empty()) return fb_fail(ctx, FB_ERR_NO_DEVICE, "NCCL unavailable: " + a->err);
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId u;
  std::memcpy(u.internal, id, FB_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  ncclResult_t r = a->CommInitRank(&comm, world, u, rank);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclCommInitRank");
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return FB_OK;
}

extern "C" int fb_comm_destroy(fb_ctx* ctx) {
  if (!ctx) return FB_ERR_ARG;
  if (!ctx->comm) return FB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nccl_api()->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  return FB_OK;
}

extern "C" int fb_comm_info(fb_ctx* ctx, int* rank, int* world) {
  if (!ctx) return FB_ERR_ARG;
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm_world;
  return ctx->comm ? FB_OK : FB_ERR_ARG;
}

// slice r of `world` contiguous slices of n items: [lo, hi)
void fb_comm_slice(size_t n, int r, int world, size_t* lo, size_t* hi) {
  *lo = n * (size_t)r / (size_t)world;
  *hi = n * (size_t)(r + 1) / (size_t)world;
}

// Every rank holds slice r of d_rows[n][row_words] (computed locally, in place); afterwards every rank holds all n rows.
// One grouped broadcast per non-empty slice, on the context stream.
int fb_comm_exchange_rows(fb_ctx* ctx, uint64_t* d_rows, size_t n, size_t row_words) {
  if (!ctx->comm || ctx->comm_world == 1 || n == 0) return FB_OK;
  NcclApi* a = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
  ncclResult_t r = a->GroupStart();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupStart");
  for (int root = 0; root < ctx->comm_world; root++) {
    size_t lo, hi;
    fb_comm_slice(n, root, ctx->comm_world, &lo, &hi);
    if (hi == lo) continue;
    uint64_t* p = d_rows + lo * row_words;
    r = a->Broadcast(p, p, (hi - lo) * row_words, ncclUint64, root, comm, ctx->stream);
    if (r != ncclSuccess) {
      a->GroupEnd();
      return nccl_fail(ctx, r, "ncclBroadcast");
    }
  }
  r = a->GroupEnd();
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclGroupEnd");
  ctx->comm_exchanges++;
  ctx->comm_bytes += n * row_words * 8;
  return FB_OK;
}
