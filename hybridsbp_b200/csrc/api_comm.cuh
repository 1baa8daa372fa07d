// Multi-GPU plumbing inside the library: one NCCL communicator per context (one process per GPU, or several
// contexts of one process), created from a unique id the host passes around by whatever means it has
// (torch.distributed in bench.py, sockets / MPI in a Julia host).
//
// The reference has no parallel code at all; what shards is the block structure: blocks are
// independent given lambda (global_curved.jl:732-737), faces couple exactly two blocks (:525-554).  The data path
// therefore has one exchange step -- the partial F-bar^T contributions of cut faces, point to point -- and the
// reductions of the CG scalars.
//
// NCCL is loaded with dlopen when the first communicator is made, so that libhsbp.so itself has no link-time
// dependency on it (single-GPU hosts never touch it) and shares the copy a host framework may already have loaded.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "hsbp_internal.h"

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;   // optional
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi g_nccl;

// 0 on success; fills g_nccl.err otherwise
int nccl_load() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { g_nccl.err = std::string("cannot load NCCL: ") + dlerror(); return 1; }
  auto sym = [&](const char *n) { return dlsym(h, n); };
#define HSBP_NCCL_SYM(field, name)                                            \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(sym(name));          \
  if (!g_nccl.field) { g_nccl.err = std::string("NCCL symbol missing: ") + name; return 1; }
  HSBP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  HSBP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  HSBP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  HSBP_NCCL_SYM(GroupStart, "ncclGroupStart")
  HSBP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  HSBP_NCCL_SYM(Send, "ncclSend")
  HSBP_NCCL_SYM(Recv, "ncclRecv")
  HSBP_NCCL_SYM(AllReduce, "ncclAllReduce")
  HSBP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef HSBP_NCCL_SYM
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
  g_nccl.lib = h;
  return 0;
}

#define HSBP_NCCL(ctx, call)                                                              \
  do {                                                                                    \
    ncclResult_t _r = (call);                                                             \
    if (_r != ncclSuccess) {                                                              \
      (ctx)->err = std::string(#call) + ": " + g_nccl.GetErrorString(_r);                 \
      return HSBP_ERR_NCCL;                                                               \
    }                                                                                     \
  } while (0)

// sum over all ranks of n doubles, on the context's stream (in == out allowed); a single-rank context copies
int comm_allreduce(hsbp_ctx *ctx, const double *in, double *out, size_t n) {
  if (ctx->world <= 1) {
    if (in != out) HSBP_CUDA(ctx, cudaMemcpyAsync(out, in, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return HSBP_OK;
  }
  HSBP_NCCL(ctx, g_nccl.AllReduce(in, out, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return HSBP_OK;
}

// one message per peer in both directions: send[off[i] .. off[i]+cnt[i]) to peers[i], the same range of recv from it
int comm_exchange(hsbp_ctx *ctx, const std::vector<int> &peers, const std::vector<int64_t> &off,
                  const std::vector<int64_t> &cnt, const double *send, double *recv) {
  if (ctx->world <= 1 || peers.empty()) return HSBP_OK;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  HSBP_NCCL(ctx, g_nccl.GroupStart());
  for (size_t i = 0; i < peers.size(); ++i) {
    ncclResult_t r1 = g_nccl.Send(send + off[i], (size_t)cnt[i], ncclDouble, peers[i], comm, ctx->stream);
    ncclResult_t r2 = g_nccl.Recv(recv + off[i], (size_t)cnt[i], ncclDouble, peers[i], comm, ctx->stream);
    if (r1 != ncclSuccess || r2 != ncclSuccess) {
      g_nccl.GroupEnd();
      ctx->err = std::string("ncclSend / ncclRecv: ") + g_nccl.GetErrorString(r1 != ncclSuccess ? r1 : r2);
      return HSBP_ERR_NCCL;
    }
  }
  HSBP_NCCL(ctx, g_nccl.GroupEnd());
  return HSBP_OK;
}

}  // namespace

extern "C" {

int hsbp_comm_unique_id(void *id128) {
  if (!id128) return HSBP_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (nccl_load()) return HSBP_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return HSBP_ERR_NCCL;
  memcpy(id128, &id, sizeof(id));
  return HSBP_OK;
}

int hsbp_comm_init(hsbp_ctx *ctx, const void *id128, int rank, int world) {
  if (!ctx) return HSBP_ERR_ARG;
  if (!id128 || world < 1 || rank < 0 || rank >= world) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_comm_init: bad arguments");
  if (ctx->comm) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_comm_init: the context already has a communicator");
  if (nccl_load()) HSBP_FAIL(ctx, HSBP_ERR_NCCL, g_nccl.err);
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  HSBP_NCCL(ctx, g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->comm = comm; ctx->rank = rank; ctx->world = world;
  // NCCL sets its channels up lazily, at the first collective (seconds on 8 GPUs): pay that here, not inside the first solve
  if (world > 1) {
    double *d = nullptr;
    HSBP_CUDA(ctx, cudaMalloc((void **)&d, 2 * sizeof(double)));
    HSBP_CUDA(ctx, cudaMemsetAsync(d, 0, 2 * sizeof(double), ctx->stream));
    int rc = comm_allreduce(ctx, d, d + 1, 1);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (rc) return rc;
    if (e != cudaSuccess) { ctx->err = std::string("hsbp_comm_init: ") + cudaGetErrorString(e); return HSBP_ERR_CUDA; }
  }
  return HSBP_OK;
}

int hsbp_comm_destroy(hsbp_ctx *ctx) {
  if (!ctx) return HSBP_ERR_ARG;
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->rank = 0; ctx->world = 1;
  return HSBP_OK;
}

int hsbp_comm_rank(const hsbp_ctx *ctx) { return ctx ? ctx->rank : -1; }
int hsbp_comm_world(const hsbp_ctx *ctx) { return ctx ? ctx->world : -1; }

/* sum over the ranks of the context's communicator, in place, of n doubles in device memory (blocking) */
int hsbp_comm_allreduce_sum(hsbp_ctx *ctx, double *x_dev, int64_t n) {
  if (!ctx) return HSBP_ERR_ARG;
  if (n < 0 || (n > 0 && !x_dev)) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_comm_allreduce_sum: bad arguments");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = comm_allreduce(ctx, x_dev, x_dev, (size_t)n);
  if (rc) return rc;
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HSBP_OK;
}

}  // extern "C"
