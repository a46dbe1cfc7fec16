// comm.cu -- the two NCCL exchanges of the multi-GPU path behind the C ABI, so that a rank-per-GPU host in
// any language (the reference's host is Julia) can drive them without torch:
//   1. cut-plane exchange of the assembly: every slab r > 0 sends the packed bottom layer of its condensed cells
//      (ghb_pack_cut_plane_f64) to slab r-1, which treats them as ghost cells (one grouped ncclSend/ncclRecv per cut);
//   2. all-gather of the owned lambda ranges before the backward step (src/HybridAffineFEOperators.jl:113-118); the
//      ranges differ in length, so it is a group of ncclBroadcast, one per rank.
// NCCL is loaded at run time (dlopen libnccl.so.2): the library has no link-time dependency on it and the single-GPU
// entry points work on a box without NCCL.  The unique id travels through whatever the host already has (MPI,
// torch.distributed's store, a file): ghb_comm_unique_id on rank 0, ghb_comm_init on every rank.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace ghb {

namespace {

// the subset of the NCCL API used here (signatures of nccl.h 2.x; enums passed as int)
struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void** comm, int nranks, NcclId id, int rank);
typedef int (*fn_comm_destroy)(void* comm);
typedef int (*fn_send)(const void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t s);
typedef int (*fn_recv)(void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t s);
typedef int (*fn_bcast)(const void* send, void* recv, size_t count, int dtype, int root, void* comm, cudaStream_t s);
typedef int (*fn_group)(void);
typedef const char* (*fn_errstr)(int);
constexpr int kNcclFloat64 = 8;   // ncclDouble

struct Nccl {
  void* h = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_send send = nullptr;
  fn_recv recv = nullptr;
  fn_bcast bcast = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_errstr errstr = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  if (n.h) return n;
  // RTLD_NOLOAD first: reuse the copy the process already has (torch bundles its own libnccl.so.2)
  n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!n.h) return n;
  n.get_unique_id = (fn_get_unique_id)dlsym(n.h, "ncclGetUniqueId");
  n.comm_init_rank = (fn_comm_init_rank)dlsym(n.h, "ncclCommInitRank");
  n.comm_destroy = (fn_comm_destroy)dlsym(n.h, "ncclCommDestroy");
  n.send = (fn_send)dlsym(n.h, "ncclSend");
  n.recv = (fn_recv)dlsym(n.h, "ncclRecv");
  n.bcast = (fn_bcast)dlsym(n.h, "ncclBroadcast");
  n.group_start = (fn_group)dlsym(n.h, "ncclGroupStart");
  n.group_end = (fn_group)dlsym(n.h, "ncclGroupEnd");
  n.errstr = (fn_errstr)dlsym(n.h, "ncclGetErrorString");
  n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.send && n.recv && n.bcast && n.group_start &&
         n.group_end;
  return n;
}

int nccl_fail(ghb_ctx* ctx, const char* what, int rc) {
  Nccl& n = nccl();
  return fail(ctx, GHB_ECUDA, std::string(what) + ": " + (n.errstr ? n.errstr(rc) : "NCCL error " + std::to_string(rc)));
}

}  // namespace

void comm_free(ghb_ctx* ctx) {
  if (ctx->comm) { nccl().comm_destroy(ctx->comm); ctx->comm = nullptr; }
  ctx->comm_rank = 0; ctx->comm_size = 1;
}

}  // namespace ghb

using namespace ghb;

extern "C" {

int ghb_comm_unique_id(ghb_ctx* ctx, void* id128) {
  if (!ctx || !id128) return GHB_EINVAL;
  Nccl& n = nccl();
  if (!n.ok) return fail(ctx, GHB_EUNSUPPORTED, "ghb_comm_unique_id: libnccl.so.2 cannot be loaded");
  NcclId id;
  int rc = n.get_unique_id(&id);
  if (rc != 0) return nccl_fail(ctx, "ncclGetUniqueId", rc);
  memcpy(id128, id.internal, 128);
  return GHB_OK;
}

int ghb_comm_init(ghb_ctx* ctx, int nranks, int rank, const void* id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, GHB_EINVAL, "ghb_comm_init: bad argument");
  Nccl& n = nccl();
  if (!n.ok) return fail(ctx, GHB_EUNSUPPORTED, "ghb_comm_init: libnccl.so.2 cannot be loaded");
  cudaSetDevice(ctx->device);
  comm_free(ctx);
  NcclId id;
  memcpy(id.internal, id128, 128);
  int rc = n.comm_init_rank(&ctx->comm, nranks, id, rank);
  if (rc != 0) { ctx->comm = nullptr; return nccl_fail(ctx, "ncclCommInitRank", rc); }
  ctx->comm_rank = rank; ctx->comm_size = nranks;
  return GHB_OK;
}

int ghb_comm_destroy(ghb_ctx* ctx) {
  if (!ctx) return GHB_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_free(ctx);
  return GHB_OK;
}

int ghb_exchange_cut_plane_f64(ghb_ctx* ctx, const double* send_down, int64_t nsend, double* recv_from_up, int64_t nrecv) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->comm) return fail(ctx, GHB_ESTATE, "ghb_exchange_cut_plane_f64: call ghb_comm_init first");
  if (nsend < 0 || nrecv < 0 || (nsend > 0 && !send_down) || (nrecv > 0 && !recv_from_up))
    return fail(ctx, GHB_EINVAL, "ghb_exchange_cut_plane_f64: bad argument");
  const bool snd = ctx->comm_rank > 0 && nsend > 0, rcv = ctx->comm_rank < ctx->comm_size - 1 && nrecv > 0;
  if ((snd && !is_device_ptr(send_down)) || (rcv && !is_device_ptr(recv_from_up)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_exchange_cut_plane_f64: device pointers required");
  cudaSetDevice(ctx->device);
  Nccl& n = nccl();
  int rc = n.group_start();
  if (rc == 0 && snd) rc = n.send(send_down, (size_t)nsend, kNcclFloat64, ctx->comm_rank - 1, ctx->comm, ctx->stream);
  if (rc == 0 && rcv) rc = n.recv(recv_from_up, (size_t)nrecv, kNcclFloat64, ctx->comm_rank + 1, ctx->comm, ctx->stream);
  int rc2 = n.group_end();
  if (rc != 0 || rc2 != 0) return nccl_fail(ctx, "cut-plane send/recv", rc != 0 ? rc : rc2);
  return GHB_OK;
}

int ghb_allgather_lambda_f64(ghb_ctx* ctx, const double* owned, const int64_t* counts, double* all) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->comm) return fail(ctx, GHB_ESTATE, "ghb_allgather_lambda_f64: call ghb_comm_init first");
  if (!counts || !all) return fail(ctx, GHB_EINVAL, "ghb_allgather_lambda_f64: null argument");
  if (!is_device_ptr(all) || (owned && !is_device_ptr(owned)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_allgather_lambda_f64: device pointers required");
  cudaSetDevice(ctx->device);
  Nccl& n = nccl();
  int64_t off = 0;
  int rc = n.group_start();
  for (int r = 0; r < ctx->comm_size && rc == 0; ++r) {
    if (counts[r] < 0) { n.group_end(); return fail(ctx, GHB_EINVAL, "ghb_allgather_lambda_f64: negative count"); }
    if (counts[r] > 0) {
      const double* src = (r == ctx->comm_rank) ? (owned ? owned : all + off) : all + off;
      rc = n.bcast(src, all + off, (size_t)counts[r], kNcclFloat64, r, ctx->comm, ctx->stream);
    }
    off += counts[r];
  }
  int rc2 = n.group_end();
  if (rc != 0 || rc2 != 0) return nccl_fail(ctx, "lambda all-gather", rc != 0 ? rc : rc2);
  return GHB_OK;
}

}  // extern "C"
