// Neighbour exchange of the cuda_c backend.
// Replaces sendrecv_fields (src/backend/omp/sendrecv.f90:10-36, src/backend/cuda/sendrecv.f90:10-100):
//   nproc_dir == 1 : f_recv_s = f_send_e ; f_recv_e = f_send_s            (device-to-device copies)
//   nproc_dir  > 1 : send_s -> prev, recv_e <- next, send_e -> next, recv_s <- prev   as one NCCL group
//                    on the context's stream (no host synchronisation, unlike cuda/sendrecv.f90:28).
// NCCL is bound with dlopen so that single-rank runs have no NCCL dependency; when the process already
// loaded torch's bundled libnccl.so.2 the same library instance is reused.
#include <dlfcn.h>

#include "common.cuh"

namespace x3d2c {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;

struct NcclApi {
  void* handle = nullptr;
  int (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_p) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static const int kNcclFloat64 = 8;  // ncclDouble
static const int kNcclSum = 0, kNcclMax = 2;

#define X3D2C_CHECK_NCCL(ctx, expr)                                                                 \
  do {                                                                                              \
    int r__ = (expr);                                                                               \
    if (r__ != 0) {                                                                                 \
      set_error(std::string(#expr) + ": " +                                                         \
                ((ctx)->nccl->GetErrorString ? (ctx)->nccl->GetErrorString(r__) : "nccl error"));   \
      return X3D2C_ENCCL;                                                                           \
    }                                                                                               \
  } while (0)

int nccl_init(x3d2c_ctx* ctx) {
  auto* api = new NcclApi;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api->handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api->handle) break;
  }
  if (!api->handle) {
    set_error(std::string("x3d2c_create: cannot dlopen libnccl.so.2: ") + dlerror());
    delete api;
    return X3D2C_ENCCL;
  }
#define LOAD(field, sym)                                              \
  *(void**)(&api->field) = dlsym(api->handle, sym);                   \
  if (!api->field) {                                                  \
    set_error(std::string("x3d2c_create: missing NCCL symbol ") + sym); \
    return X3D2C_ENCCL;                                               \
  }
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(Send, "ncclSend")
  LOAD(Recv, "ncclRecv")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  ctx->nccl = api;
  ncclUniqueId_t id;
  std::memcpy(&id, ctx->cfg.nccl_unique_id, sizeof id);
  ncclComm_p comm = nullptr;
  X3D2C_CHECK_NCCL(ctx, api->CommInitRank(&comm, ctx->cfg.nproc, id, ctx->cfg.rank));
  ctx->nccl_comm = comm;
  return X3D2C_OK;
}

void nccl_finalize(x3d2c_ctx* ctx) {
  if (ctx->nccl) {
    if (ctx->nccl_comm) ctx->nccl->CommDestroy(ctx->nccl_comm);
    delete ctx->nccl;
    ctx->nccl = nullptr;
    ctx->nccl_comm = nullptr;
  }
}

int sendrecv_fields(x3d2c_ctx* ctx, int dir, double* recv_s, double* recv_e, const double* send_s,
                    const double* send_e, size_t count) {
  const int P = ctx->cfg.nproc_dir[dir - 1];
  if (P == 1) {
    X3D2C_CHECK_CUDA(cudaMemcpyAsync(recv_s, send_e, sizeof(double) * count, cudaMemcpyDeviceToDevice, ctx->stream));
    X3D2C_CHECK_CUDA(cudaMemcpyAsync(recv_e, send_s, sizeof(double) * count, cudaMemcpyDeviceToDevice, ctx->stream));
    return X3D2C_OK;
  }
  X3D2C_REQUIRE(ctx->nccl && ctx->nccl_comm, "sendrecv_fields: multi-rank direction without an NCCL communicator");
  const int prev = ctx->cfg.pprev[dir - 1], next = ctx->cfg.pnext[dir - 1];
  NcclApi* a = ctx->nccl;
  X3D2C_CHECK_NCCL(ctx, a->GroupStart());
  X3D2C_CHECK_NCCL(ctx, a->Send(send_s, count, kNcclFloat64, prev, ctx->nccl_comm, ctx->stream));
  X3D2C_CHECK_NCCL(ctx, a->Recv(recv_e, count, kNcclFloat64, next, ctx->nccl_comm, ctx->stream));
  X3D2C_CHECK_NCCL(ctx, a->Send(send_e, count, kNcclFloat64, next, ctx->nccl_comm, ctx->stream));
  X3D2C_CHECK_NCCL(ctx, a->Recv(recv_s, count, kNcclFloat64, prev, ctx->nccl_comm, ctx->stream));
  X3D2C_CHECK_NCCL(ctx, a->GroupEnd());
  ctx->launches++;
  return X3D2C_OK;
}

// in-place all-reduce of `count` doubles on the device; op: 0 = sum, 1 = max
int allreduce(x3d2c_ctx* ctx, double* dev, size_t count, int op) {
  if (ctx->cfg.nproc == 1) return X3D2C_OK;
  X3D2C_REQUIRE(ctx->nccl && ctx->nccl_comm, "allreduce: no NCCL communicator");
  X3D2C_CHECK_NCCL(ctx, ctx->nccl->AllReduce(dev, dev, count, kNcclFloat64, op == 0 ? kNcclSum : kNcclMax,
                                            ctx->nccl_comm, ctx->stream));
  ctx->launches++;
  return X3D2C_OK;
}

// all-to-all of equal blocks (used by the distributed FFT transposes): block r of `send` goes to rank r
int alltoall(x3d2c_ctx* ctx, double* recv, const double* send, size_t block_doubles) {
  const int P = ctx->cfg.nproc;
  if (P == 1) {
    X3D2C_CHECK_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * block_doubles, cudaMemcpyDeviceToDevice, ctx->stream));
    return X3D2C_OK;
  }
  NcclApi* a = ctx->nccl;
  X3D2C_CHECK_NCCL(ctx, a->GroupStart());
  for (int r = 0; r < P; ++r) {
    X3D2C_CHECK_NCCL(ctx, a->Send(send + (size_t)r * block_doubles, block_doubles, kNcclFloat64, r, ctx->nccl_comm, ctx->stream));
    X3D2C_CHECK_NCCL(ctx, a->Recv(recv + (size_t)r * block_doubles, block_doubles, kNcclFloat64, r, ctx->nccl_comm, ctx->stream));
  }
  X3D2C_CHECK_NCCL(ctx, a->GroupEnd());
  ctx->launches++;
  return X3D2C_OK;
}

// Maps every rank's exchange buffer (ctx->halo + flags) into this process with CUDA IPC so that the halo / carry rows
// of rank-split directions can be stored straight into the neighbours' buffers (m3_edge.cu). All ranks agree on the
// outcome; on any failure (or X3D2C_NO_P2P) the grouped ncclSend / ncclRecv exchange stays in use.
int setup_peer_halo(x3d2c_ctx* ctx) {
  const int P = ctx->cfg.nproc, me = ctx->cfg.rank;
  if (P <= 1 || P > 8 || std::getenv("X3D2C_NO_P2P") || std::getenv("X3D2C_NO_P2P_HALO")) return X3D2C_OK;
  constexpr int HD = sizeof(cudaIpcMemHandle_t) / sizeof(double);
  cudaIpcMemHandle_t h;
  std::memset(&h, 0, sizeof h);
  bool ok = cudaIpcGetMemHandle(&h, ctx->halo) == cudaSuccess;
  if (!ok) cudaGetLastError();
  std::vector<double> rep((size_t)HD * P), all((size_t)HD * P);
  for (int r = 0; r < P; ++r) std::memcpy(&rep[(size_t)r * HD], &h, sizeof h);
  double *d_send = nullptr, *d_recv = nullptr, *d_flag = nullptr;
  X3D2C_CHECK_CUDA(cudaMalloc(&d_send, sizeof(double) * HD * P));
  X3D2C_CHECK_CUDA(cudaMalloc(&d_recv, sizeof(double) * HD * P));
  X3D2C_CHECK_CUDA(cudaMalloc(&d_flag, sizeof(double)));
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(d_send, rep.data(), sizeof(double) * HD * P, cudaMemcpyHostToDevice, ctx->stream));
  int rc = alltoall(ctx, d_recv, d_send, HD);
  if (rc) return rc;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(all.data(), d_recv, sizeof(double) * HD * P, cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < P && ok; ++r) {
    if (r == me) { ctx->peer_halo[r] = ctx->halo; continue; }
    cudaIpcMemHandle_t hr;
    std::memcpy(&hr, &all[(size_t)r * HD], sizeof hr);
    void* pa = nullptr;
    ok = cudaIpcOpenMemHandle(&pa, hr, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (!ok) cudaGetLastError();
    ctx->peer_halo[r] = (double*)pa;
  }
  const double flag = ok ? 0.0 : 1.0;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = allreduce(ctx, d_flag, 1, 0))) return rc;
  double failed = 1.0;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(&failed, d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->halo_p2p = failed == 0.0;
  if (ctx->halo_p2p)
    for (int r = 0; r < P; ++r)
      ctx->peer_halo_flags[r] = reinterpret_cast<unsigned long long*>(ctx->peer_halo[r] + ctx->halo_doubles);
  cudaFree(d_send); cudaFree(d_recv); cudaFree(d_flag);
  if (std::getenv("X3D2C_TRACE"))
    std::fprintf(stderr, "[x3d2c] rank %d halo / carry exchange: %s\n", me,
                 ctx->halo_p2p ? "peer stores from the pack / edge kernels + flags (CUDA IPC)" : "ncclSend/Recv");
  return X3D2C_OK;
}

void release_peer_halo(x3d2c_ctx* ctx) {
  for (int r = 0; r < 8; ++r)
    if (r != ctx->cfg.rank && ctx->peer_halo[r]) { cudaIpcCloseMemHandle(ctx->peer_halo[r]); ctx->peer_halo[r] = nullptr; }
}

}  // namespace x3d2c

extern "C" int x3d2c_nccl_unique_id(void* out128) {
  using namespace x3d2c;
  X3D2C_REQUIRE(out128, "x3d2c_nccl_unique_id: null argument");
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror()); return X3D2C_ENCCL; }
  int (*get_id)(ncclUniqueId_t*) = nullptr;
  *(void**)(&get_id) = dlsym(h, "ncclGetUniqueId");
  if (!get_id) { set_error("missing NCCL symbol ncclGetUniqueId"); return X3D2C_ENCCL; }
  ncclUniqueId_t id;
  if (get_id(&id) != 0) { set_error("ncclGetUniqueId failed"); return X3D2C_ENCCL; }
  std::memcpy(out128, &id, sizeof id);
  return X3D2C_OK;
}
