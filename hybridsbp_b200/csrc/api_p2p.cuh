// Peer-memory exchange for the iteration loop of the trace CG (K4): the cut-face parts of q and the partial sums of the CG
// scalars are written straight into the partners' mailboxes over NVLink by a tiny kernel that follows the producer, and a
// second tiny kernel on the consumer side waits for the partners' flags and adds the partials up in rank order (every rank gets
// the same bits).  Per iteration this replaces one grouped ncclSend / ncclRecv and two ncclAllReduce calls -- 25 - 40 us of
// launch and protocol latency at 8 GPUs, which is most of an iteration on small blocks -- by four kernels of 2 - 3 us that live
// in the same CUDA graph as the rest of the chunk.
//
//   mailbox of a rank (doubles):  flagA[16] flagB[16] | red1[2][16] | red2[2][16][nred2] | recv[2][nrecv of that rank]
//   phase A (after k_cg_q):       push: my cut-face parts -> partner.recv[parity] at the partner's offset for me, my part of p.q ->
//                                 red1[parity][me] of every rank, then flagA[me] = epoch everywhere;
//                                 wait: flagA[r] >= epoch for all r, recv[parity] -> the trace's receive buffer, sum of red1 in rank order
//   phase B (after k_cg_coarse):  the same for the reduction vector [r.z1, r.r, b_I.t, y] of the preconditioner
// Slots are double-buffered by the parity of the epoch; the two phases of an iteration order the reuse anyway (a rank can only
// reach phase A of iteration k+1 after every rank has pushed phase B of iteration k, i.e. has consumed phase A of k).
//
// The mailboxes are ordinary device allocations mapped into the partner processes with cudaIpc handles, exchanged once through the
// NCCL communicator (setup is collective).  If any rank cannot map a partner (no peer access, several ranks in one process) all
// ranks agree to keep the NCCL path.  Setup-time exchanges (face blocks of cut faces, the coarse Schur complement) and the two
// exchanges outside the loop stay on NCCL.  The reference has no parallel code.
#pragma once
#include <unistd.h>

#include "api_comm.cuh"
#include "k_cg.cuh"

namespace hsbp {

constexpr int P2P_MAXW = 16;

struct P2PDev {
  int world, me, npeers, error;
  int peer_rank[P2P_MAXW];
  long long peer_soff[P2P_MAXW], peer_cnt[P2P_MAXW], peer_roff[P2P_MAXW];   // my send range for peer i; where it lands in its recv area
  long long nrecv_of[P2P_MAXW];                                             // recv-area length of every rank
  double *base[P2P_MAXW];                                                   // mailbox of every rank as mapped here
  long long nred2;
  unsigned long long epochA, epochB;
};

__host__ __device__ inline long long p2p_off_red1() { return 2 * P2P_MAXW; }
__host__ __device__ inline long long p2p_off_red2() { return 4 * P2P_MAXW; }
__host__ __device__ inline long long p2p_off_recv(long long nred2) { return 4 * P2P_MAXW + 2 * P2P_MAXW * nred2; }

__device__ __forceinline__ double p2p_ld(const double *p) { return *reinterpret_cast<const volatile double *>(p); }

// spin until the flags of all ranks have reached epoch e (one thread per rank); gives up after about half a minute (a rank that
// is merely late -- its host was descheduled between two chunks -- must not be mistaken for a lost one)
__device__ __forceinline__ void p2p_wait_flags(P2PDev *pp, const double *mb, int flag0, unsigned long long e) {
  if ((int)threadIdx.x < pp->world) {
    const volatile unsigned long long *f = reinterpret_cast<const volatile unsigned long long *>(mb) + flag0 + threadIdx.x;
    const long long t0 = clock64();
    while (*f < e) {
      if (clock64() - t0 > 60000000000ll) { pp->error = 1; break; }
    }
  }
  __threadfence_system();
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
k_p2p_push_a(const CgState *__restrict__ st, int force, P2PDev *pp, const double *__restrict__ send, const double *__restrict__ red1) {
  if (!force && st->done) return;
  const unsigned long long e = pp->epochA + 1;
  const int par = (int)(e & 1);
  const long long recv0 = p2p_off_recv(pp->nred2);
  for (int i = 0; i < pp->npeers; ++i) {
    const int pr = pp->peer_rank[i];
    double *dst = pp->base[pr] + recv0 + par * pp->nrecv_of[pr] + pp->peer_roff[i];
    const double *src = send + pp->peer_soff[i];
    for (long long k = threadIdx.x; k < pp->peer_cnt[i]; k += blockDim.x) dst[k] = src[k];
  }
  if ((int)threadIdx.x < pp->world) pp->base[threadIdx.x][p2p_off_red1() + par * P2P_MAXW + pp->me] = red1[0];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pp->world)
    *(reinterpret_cast<volatile unsigned long long *>(pp->base[threadIdx.x]) + pp->me) = e;          // flagA[me] of rank threadIdx.x
  if (threadIdx.x == 0) pp->epochA = e;
}

__global__ void __launch_bounds__(1024)
k_p2p_wait_a(const CgState *__restrict__ st, int force, P2PDev *pp, double *__restrict__ recv, double *__restrict__ red1) {
  if (!force && st->done) return;
  const unsigned long long e = pp->epochA;
  const int par = (int)(e & 1);
  const double *mb = pp->base[pp->me];
  p2p_wait_flags(pp, mb, 0, e);
  const long long nrecv = pp->nrecv_of[pp->me];
  const double *src = mb + p2p_off_recv(pp->nred2) + par * nrecv;
  for (long long k = threadIdx.x; k < nrecv; k += blockDim.x) recv[k] = p2p_ld(src + k);
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int r = 0; r < pp->world; ++r) s += p2p_ld(mb + p2p_off_red1() + par * P2P_MAXW + r);
    red1[1] = s;
  }
}

__global__ void __launch_bounds__(1024)
k_p2p_push_b(const CgState *__restrict__ st, int force, P2PDev *pp, const double *__restrict__ in, int n) {
  if (!force && st->done) return;
  const unsigned long long e = pp->epochB + 1;
  const int par = (int)(e & 1);
  for (int idx = threadIdx.x; idx < pp->world * n; idx += blockDim.x) {
    const int r = idx / n, k = idx - r * n;
    pp->base[r][p2p_off_red2() + ((long long)par * P2P_MAXW + pp->me) * pp->nred2 + k] = in[k];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pp->world)
    *(reinterpret_cast<volatile unsigned long long *>(pp->base[threadIdx.x]) + P2P_MAXW + pp->me) = e;   // flagB[me]
  if (threadIdx.x == 0) pp->epochB = e;
}

__global__ void __launch_bounds__(1024)
k_p2p_wait_b(const CgState *__restrict__ st, int force, P2PDev *pp, double *__restrict__ out, int n) {
  if (!force && st->done) return;
  const unsigned long long e = pp->epochB;
  const int par = (int)(e & 1);
  const double *mb = pp->base[pp->me];
  p2p_wait_flags(pp, mb, P2P_MAXW, e);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < pp->world; ++r) s += p2p_ld(mb + p2p_off_red2() + ((long long)par * P2P_MAXW + r) * pp->nred2 + k);
    out[k] = s;
  }
}

}  // namespace hsbp

namespace {

using namespace hsbp;

struct P2PHost {
  P2PDev *d_dev = nullptr;
  double *mailbox = nullptr;
  size_t mailbox_bytes = 0;
  int device = 0;
  std::vector<void *> opened;          // mappings of the partners' mailboxes (cudaIpcCloseMemHandle)
  long long nred2 = 0, nrecv = 0;
};

// A mailbox that partner processes have mapped must not be freed while they may still have it open, and the ranks do not
// destroy their traces in step: a released mailbox goes back to this process-wide pool (reused by the next trace that needs
// one of at most its size on the same device) and is returned to the driver only when the process ends.
struct MailboxPoolEntry { int device; size_t bytes; double *ptr; };
std::vector<MailboxPoolEntry> g_mailbox_pool;

void p2p_free(hsbp_trace *t) {
  P2PHost *h = (P2PHost *)t->p2p;
  if (!h) return;
  if (t->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)t->graph_exec); t->graph_exec = nullptr; }   // it holds the mailbox pointers
  for (void *p : h->opened) cudaIpcCloseMemHandle(p);
  cudaFree(h->d_dev);
  if (h->mailbox) g_mailbox_pool.push_back({h->device, h->mailbox_bytes, h->mailbox});
  delete h;
  t->p2p = nullptr;
}

// collective over the communicator; leaves t->p2p == nullptr (NCCL path) when any rank cannot map its partners
int p2p_setup(hsbp_trace *t) {
  hsbp_ctx *ctx = t->blocks->ctx;
  const int world = ctx->world, me = ctx->rank;
  const long long nred2 = 3 + (t->cmodes > 0 ? t->nGt : 0), nrecv = (long long)t->msg_len;
  if (t->p2p) {
    P2PHost *h = (P2PHost *)t->p2p;
    if (h->nred2 == nred2 && h->nrecv == nrecv) return HSBP_OK;
    p2p_free(t);
  }
  if (world < 2 || world > P2P_MAXW || !g_nccl.AllGather) return HSBP_OK;
  struct Rec { long long pid; unsigned char handle[64]; long long nrecv; long long roff[P2P_MAXW]; };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  P2PHost *h = new (std::nothrow) P2PHost();
  if (!h) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
  h->nred2 = nred2; h->nrecv = nrecv;
  const size_t mb_doubles = (size_t)(p2p_off_recv(nred2) + 2 * nrecv + 2);
  int bad = 0;
  Rec mine;
  memset(&mine, 0, sizeof(mine));
  h->device = ctx->device;
  for (size_t i = 0; i < g_mailbox_pool.size(); ++i)
    if (g_mailbox_pool[i].device == ctx->device && g_mailbox_pool[i].bytes >= mb_doubles * sizeof(double)) {
      h->mailbox = g_mailbox_pool[i].ptr; h->mailbox_bytes = g_mailbox_pool[i].bytes;
      g_mailbox_pool.erase(g_mailbox_pool.begin() + i);
      break;
    }
  if (!h->mailbox) {
    h->mailbox_bytes = mb_doubles * sizeof(double);
    if (cudaMalloc((void **)&h->mailbox, h->mailbox_bytes) != cudaSuccess) { h->mailbox = nullptr; bad = 1; cudaGetLastError(); }
  }
  if (bad || cudaMemset(h->mailbox, 0, h->mailbox_bytes) != cudaSuccess ||
      cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(mine.handle), h->mailbox) != cudaSuccess) {
    bad = 1;
    cudaGetLastError();
  }
  mine.pid = (long long)getpid(); mine.nrecv = nrecv;
  for (int q = 0; q < P2P_MAXW; ++q) mine.roff[q] = -1;
  for (size_t j = 0; j < t->peers.size(); ++j) mine.roff[t->peers[j]] = (long long)t->peer_off[j];
  // all-gather of the records
  std::vector<Rec> recs((size_t)world);
  unsigned char *d_in = nullptr, *d_out = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_in, sizeof(Rec)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_out, sizeof(Rec) * world));
  HSBP_CUDA(ctx, cudaMemcpyAsync(d_in, &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->stream));
  HSBP_NCCL(ctx, g_nccl.AllGather(d_in, d_out, sizeof(Rec), ncclInt8, (ncclComm_t)ctx->comm, ctx->stream));
  HSBP_CUDA(ctx, cudaMemcpyAsync(recs.data(), d_out, sizeof(Rec) * world, cudaMemcpyDeviceToHost, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_in); cudaFree(d_out);
  P2PDev dev;
  memset(&dev, 0, sizeof(dev));
  dev.world = world; dev.me = me; dev.npeers = (int)t->peers.size(); dev.nred2 = nred2;
  for (int r = 0; r < world && !bad; ++r) {
    dev.nrecv_of[r] = recs[r].nrecv;
    if (r == me) { dev.base[r] = h->mailbox; continue; }
    if (recs[r].pid == mine.pid) { bad = 1; break; }              // several ranks in one process: IPC handles cannot be opened there
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, *reinterpret_cast<cudaIpcMemHandle_t *>(recs[r].handle), cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      bad = 1;
      cudaGetLastError();
      break;
    }
    h->opened.push_back(p);
    dev.base[r] = (double *)p;
  }
  for (size_t j = 0; j < t->peers.size() && !bad; ++j) {
    const int pr = t->peers[j];
    dev.peer_rank[j] = pr; dev.peer_soff[j] = (long long)t->peer_off[j]; dev.peer_cnt[j] = (long long)t->peer_cnt[j];
    dev.peer_roff[j] = recs[pr].roff[me];
    if (dev.peer_roff[j] < 0) bad = 1;
  }
  // every rank keeps the peer path or none does
  double *d_flag = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_flag, 2 * sizeof(double)));
  const double fl = bad ? 1.0 : 0.0;
  HSBP_CUDA(ctx, cudaMemcpyAsync(d_flag, &fl, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int rc = comm_allreduce(ctx, d_flag, d_flag + 1, 1);
  double total = 1.0;
  if (rc == HSBP_OK && cudaMemcpyAsync(&total, d_flag + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess)
    cudaStreamSynchronize(ctx->stream);
  cudaFree(d_flag);
  t->p2p = h;
  if (rc != HSBP_OK || total != 0.0) { p2p_free(t); return rc; }
  if (cudaMalloc((void **)&h->d_dev, sizeof(P2PDev)) != cudaSuccess ||
      cudaMemcpy(h->d_dev, &dev, sizeof(P2PDev), cudaMemcpyHostToDevice) != cudaSuccess) {
    // (cannot happen in practice; a rank that fails here would leave the others waiting, so make it loud)
    p2p_free(t);
    HSBP_FAIL(ctx, HSBP_ERR_CUDA, "p2p_setup: out of device memory");
  }
  return HSBP_OK;
}

}  // namespace
