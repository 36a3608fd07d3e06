// Context, field storage and tdsops upload of the cuda_c backend.
// Mirrors cuda_backend_t%init (src/backend/cuda/backend.f90:95-152), cuda_allocator_t
// (src/backend/cuda/allocator.f90:44-90), cuda_tdsops_t upload (src/backend/cuda/tdsops.f90:31-90).
#include "common.cuh"

namespace x3d2c {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// src/mesh.f90:196-261
int get_dims_dataloc(const x3d2c_ctx* ctx, int data_loc, int dims[3], bool global) {
  const int* v = global ? ctx->cfg.dims_vert_global : ctx->cfg.dims_vert;
  const int* c = global ? ctx->cfg.dims_cell_global : ctx->cfg.dims_cell;
  switch (data_loc) {
    case X3D2C_VERT: dims[0] = v[0]; dims[1] = v[1]; dims[2] = v[2]; break;
    case X3D2C_CELL: dims[0] = c[0]; dims[1] = c[1]; dims[2] = c[2]; break;
    case X3D2C_X_FACE: dims[0] = v[0]; dims[1] = c[1]; dims[2] = c[2]; break;
    case X3D2C_Y_FACE: dims[0] = c[0]; dims[1] = v[1]; dims[2] = c[2]; break;
    case X3D2C_Z_FACE: dims[0] = c[0]; dims[1] = c[1]; dims[2] = v[2]; break;
    case X3D2C_X_EDGE: dims[0] = c[0]; dims[1] = v[1]; dims[2] = v[2]; break;
    case X3D2C_Y_EDGE: dims[0] = v[0]; dims[1] = c[1]; dims[2] = v[2]; break;
    case X3D2C_Z_EDGE: dims[0] = v[0]; dims[1] = v[1]; dims[2] = c[2]; break;
    default: set_error("unknown data_loc " + std::to_string(data_loc)); return X3D2C_EINVAL;
  }
  return X3D2C_OK;
}

int ensure_scratch(x3d2c_ctx* ctx, int count) {
  for (int i = 0; i < count; ++i)
    if (!ctx->scratch[i]) X3D2C_CHECK_CUDA(cudaMalloc(&ctx->scratch[i], sizeof(double) * ctx->ngrid));
  return X3D2C_OK;
}
int ensure_scratch_slot(x3d2c_ctx* ctx, int i) {
  if (!ctx->scratch[i]) X3D2C_CHECK_CUDA(cudaMalloc(&ctx->scratch[i], sizeof(double) * ctx->ngrid));
  return X3D2C_OK;
}
int nccl_init(x3d2c_ctx* ctx);      // nccl.cu
void nccl_finalize(x3d2c_ctx* ctx);  // nccl.cu
}  // namespace x3d2c

using namespace x3d2c;

namespace {
__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
}  // namespace

extern "C" {

const char* x3d2c_last_error(void) { return g_last_error.c_str(); }
int x3d2c_version(void) { return 100; }

int x3d2c_create(const x3d2c_config* cfg, x3d2c_ctx** out) {
  X3D2C_REQUIRE(cfg && out, "x3d2c_create: null argument");
  X3D2C_REQUIRE(cfg->sz == SZ, "x3d2c_create: sz must be X3D2C_SZ (32)");
  for (int d = 0; d < 3; ++d) X3D2C_REQUIRE(cfg->dims_vert[d] > 0, "x3d2c_create: non-positive dims");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("x3d2c_create: no CUDA device available (") + cudaGetErrorString(e) +
              "). The cuda_c backend has no CPU fallback.");
    return X3D2C_ECUDA;
  }
  // released by the guard on every early return (X3D2C_CHECK_* / REQUIRE below), handed over at the end
  struct Guard {
    x3d2c_ctx* c;
    ~Guard() { if (c) x3d2c_destroy(c); }
  } guard{new x3d2c_ctx};
  x3d2c_ctx* ctx = guard.c;
  ctx->cfg = *cfg;
  ctx->cfg.nccl_unique_id = nullptr;
  if (cfg->device >= 0) X3D2C_CHECK_CUDA(cudaSetDevice(cfg->device));
  X3D2C_CHECK_CUDA(cudaGetDevice(&ctx->device));
  X3D2C_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->strict = (cfg->flags & X3D2C_FLAG_STRICT) ? 1 : 0;
  ctx->force_dist = std::getenv("X3D2C_FORCE_DIST") ? 1 : 0;
  // src/allocator.f90:72-83
  const int nx = cfg->dims_vert[0], ny = cfg->dims_vert[1], nz = cfg->dims_vert[2];
  ctx->nx_pad = nx - 1 + (-(nx - 1)) % SZ + SZ;
  ctx->ny_pad = ny - 1 + (-(ny - 1)) % SZ + SZ;
  ctx->nz_pad = nz;
  ctx->ngrid = (long long)ctx->nx_pad * ctx->ny_pad * ctx->nz_pad;
  ctx->n_groups[1] = ctx->ny_pad * ctx->nz_pad / SZ;
  ctx->n_groups[2] = ctx->nx_pad * ctx->nz_pad / SZ;
  ctx->n_groups[3] = ctx->nx_pad * ctx->ny_pad / SZ;
  // halo / reduced-row exchange buffers: 3 fields x (send_s, send_e, recv_s, recv_e) x 4 rows, plus
  // 9 quantities x 4 buffers x 1 row, sized for the largest cross-section (omp/backend.f90:84-112)
  int ng = ctx->n_groups[1] > ctx->n_groups[2] ? ctx->n_groups[1] : ctx->n_groups[2];
  if (ctx->n_groups[3] > ng) ng = ctx->n_groups[3];
  const int halo_rows = kHaloRows;  // common.cuh: both carvings fit (checked where they are made)
  ctx->halo_doubles = (size_t)SZ * ng * halo_rows;
  X3D2C_CHECK_CUDA(cudaMalloc(&ctx->halo, sizeof(double) * ctx->halo_doubles + 1024));  // + flags of the peer exchange
  X3D2C_CHECK_CUDA(cudaMemsetAsync(ctx->halo, 0, sizeof(double) * ctx->halo_doubles + 1024, ctx->stream));
  ctx->halo_flags = reinterpret_cast<unsigned long long*>(ctx->halo + ctx->halo_doubles);
  {  // the in-kernel carry exchange expects its receive slots at the sentinel
    const size_t n_inl = (size_t)SZ * ng * kHaloRowsInline;
    fill_u64_kernel<<<1184, 256, 0, ctx->stream>>>(
        reinterpret_cast<unsigned long long*>(ctx->halo + ctx->halo_doubles - n_inl), n_inl, kCarrySentinel);
    X3D2C_CHECK_CUDA(cudaGetLastError());
  }
  ctx->red_blocks = 1184;  // 8 CTAs per SM on 148 SMs
  X3D2C_CHECK_CUDA(cudaMalloc(&ctx->red, sizeof(double) * (2 * ctx->red_blocks + 8)));
  X3D2C_CHECK_CUDA(cudaMallocHost(&ctx->red_host, sizeof(double) * 8));
  if (cfg->nproc > 1) {
    X3D2C_REQUIRE(cfg->nccl_unique_id, "x3d2c_create: nproc > 1 needs an ncclUniqueId");
    ctx->cfg.nccl_unique_id = cfg->nccl_unique_id;
    int rc = nccl_init(ctx);
    ctx->cfg.nccl_unique_id = nullptr;
    if (rc) return rc;
    if ((rc = setup_peer_halo(ctx))) return rc;
  }
  guard.c = nullptr;
  *out = ctx;
  return X3D2C_OK;
}

int x3d2c_destroy(x3d2c_ctx* ctx) {
  X3D2C_ENTER(ctx);
  if (!ctx) return X3D2C_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int l = 1; l <= 2; ++l)
    if (ctx->lane[l]) { cudaStreamSynchronize(ctx->lane[l]); cudaStreamDestroy(ctx->lane[l]); }
  for (auto& e : ctx->lane_ev)
    if (e) cudaEventDestroy(e);
  release_peer_halo(ctx);
  nccl_finalize(ctx);
  for (int i = 0; i < 6; ++i)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->halo) cudaFree(ctx->halo);
  if (ctx->red) cudaFree(ctx->red);
  if (ctx->red_host) cudaFreeHost(ctx->red_host);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return X3D2C_OK;
}

int x3d2c_sync(x3d2c_ctx* ctx) {
  X3D2C_ENTER(ctx);
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return X3D2C_OK;
}

int x3d2c_get_padded_dims(const x3d2c_ctx* ctx, int dims_padded[3], int n_groups[3], long long* ngrid) {
  X3D2C_ENTER(ctx);
  dims_padded[0] = ctx->nx_pad; dims_padded[1] = ctx->ny_pad; dims_padded[2] = ctx->nz_pad;
  n_groups[0] = ctx->n_groups[1]; n_groups[1] = ctx->n_groups[2]; n_groups[2] = ctx->n_groups[3];
  *ngrid = ctx->ngrid;
  return X3D2C_OK;
}

long long x3d2c_launch_count(const x3d2c_ctx* ctx) { return ctx->launches; }
void* x3d2c_stream(const x3d2c_ctx* ctx) { return (void*)ctx->stream; }

int x3d2c_field_alloc(x3d2c_ctx* ctx, double** dev) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dev, "x3d2c_field_alloc: null argument");
  cudaError_t e = cudaMalloc(dev, sizeof(double) * ctx->ngrid);
  if (e != cudaSuccess) {
    set_error(std::string("x3d2c_field_alloc: ") + cudaGetErrorString(e));
    return X3D2C_ENOMEM;
  }
  X3D2C_CHECK_CUDA(cudaMemsetAsync(*dev, 0, sizeof(double) * ctx->ngrid, ctx->stream));
  return X3D2C_OK;
}

int x3d2c_field_free(x3d2c_ctx* ctx, double* dev) {
  X3D2C_ENTER(ctx);
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  X3D2C_CHECK_CUDA(cudaFree(dev));
  return X3D2C_OK;
}

int x3d2c_copy_data_to_f(x3d2c_ctx* ctx, double* dev, const double* host_data) {
  X3D2C_ENTER(ctx);
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(dev, host_data, sizeof(double) * ctx->ngrid, cudaMemcpyHostToDevice, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return X3D2C_OK;
}

int x3d2c_copy_f_to_data(x3d2c_ctx* ctx, double* host_data, const double* dev) {
  X3D2C_ENTER(ctx);
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(host_data, dev, sizeof(double) * ctx->ngrid, cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return X3D2C_OK;
}

// ---- I/O lanes (x3d2c.h)
static int lane_stream(x3d2c_ctx* ctx, int lane, cudaStream_t* s) {
  X3D2C_REQUIRE(lane >= 0 && lane <= 2, "x3d2c lane: lane must be 0 (compute), 1 (upload) or 2 (download)");
  if (lane == 0) { *s = ctx->stream; return X3D2C_OK; }
  if (!ctx->lane[lane]) X3D2C_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->lane[lane], cudaStreamNonBlocking));
  *s = ctx->lane[lane];
  return X3D2C_OK;
}
int x3d2c_copy_data_to_f_async(x3d2c_ctx* ctx, double* dev, const double* host_pinned, int lane) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dev && host_pinned, "x3d2c_copy_data_to_f_async: null argument");
  cudaStream_t s;
  int rc = lane_stream(ctx, lane, &s);
  if (rc) return rc;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(dev, host_pinned, sizeof(double) * ctx->ngrid, cudaMemcpyHostToDevice, s));
  return X3D2C_OK;
}
int x3d2c_copy_f_to_data_async(x3d2c_ctx* ctx, double* host_pinned, const double* dev, int lane) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dev && host_pinned, "x3d2c_copy_f_to_data_async: null argument");
  cudaStream_t s;
  int rc = lane_stream(ctx, lane, &s);
  if (rc) return rc;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(host_pinned, dev, sizeof(double) * ctx->ngrid, cudaMemcpyDeviceToHost, s));
  return X3D2C_OK;
}
int x3d2c_lane_record(x3d2c_ctx* ctx, int lane, int ev) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && ev >= 0 && ev < 16, "x3d2c_lane_record: ev must be 0..15");
  cudaStream_t s;
  int rc = lane_stream(ctx, lane, &s);
  if (rc) return rc;
  if (!ctx->lane_ev[ev]) X3D2C_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->lane_ev[ev], cudaEventDisableTiming));
  X3D2C_CHECK_CUDA(cudaEventRecord(ctx->lane_ev[ev], s));
  return X3D2C_OK;
}
int x3d2c_lane_wait(x3d2c_ctx* ctx, int lane, int ev) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && ev >= 0 && ev < 16, "x3d2c_lane_wait: ev must be 0..15");
  cudaStream_t s;
  int rc = lane_stream(ctx, lane, &s);
  if (rc) return rc;
  if (!ctx->lane_ev[ev]) return X3D2C_OK;  // never recorded: nothing to wait for
  X3D2C_CHECK_CUDA(cudaStreamWaitEvent(s, ctx->lane_ev[ev], 0));
  return X3D2C_OK;
}
int x3d2c_lane_sync(x3d2c_ctx* ctx, int lane) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx, "x3d2c_lane_sync: null argument");
  cudaStream_t s;
  int rc = lane_stream(ctx, lane, &s);
  if (rc) return rc;
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(s));
  return X3D2C_OK;
}

int x3d2c_tdsops_create(x3d2c_ctx* ctx, int n_tds, int n_rhs, int move, int periodic, const double* coeffs,
                        const double* coeffs_s, const double* coeffs_e, const double* dist_fw,
                        const double* dist_bw, const double* dist_sa, const double* dist_sc,
                        const double* dist_af, const double* stretch, const double* stretch_correct,
                        x3d2c_tdsops** out) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out && coeffs && coeffs_s && coeffs_e && dist_fw && dist_bw && dist_sa && dist_sc &&
                    dist_af && stretch && stretch_correct, "x3d2c_tdsops_create: null argument");
  X3D2C_REQUIRE(n_tds >= 9 && (n_rhs == n_tds || n_rhs == n_tds + 1), "x3d2c_tdsops_create: bad n_tds / n_rhs");
  auto* t = new x3d2c_tdsops;
  t->n_tds = n_tds; t->n_rhs = n_rhs; t->move = move; t->periodic = periodic;
  t->h_fw.assign(dist_fw, dist_fw + n_rhs); t->h_bw.assign(dist_bw, dist_bw + n_rhs);
  t->h_sa.assign(dist_sa, dist_sa + n_rhs); t->h_sc.assign(dist_sc, dist_sc + n_rhs);
  t->h_af.assign(dist_af, dist_af + n_rhs);
  t->h_stretch.assign(stretch, stretch + n_tds);
  t->h_stretch_correct.assign(stretch_correct, stretch_correct + n_tds);
  for (int i = 0; i < n_tds; ++i) {
    if (stretch[i] != 1.0) t->has_stretch = 1;
    if (stretch_correct[i] != 0.0) t->has_stretch_correct = 1;
  }
  t->tap_mask = 0;
  for (int k = 0; k < 9; ++k)
    if (coeffs[k] != 0.0) t->tap_mask |= 1u << k;
  const size_t n = (size_t)n_rhs;
  std::vector<double> blk(7 * n, 0.0);
  std::memcpy(&blk[0 * n], dist_fw, sizeof(double) * n);
  std::memcpy(&blk[1 * n], dist_bw, sizeof(double) * n);
  std::memcpy(&blk[2 * n], dist_sa, sizeof(double) * n);
  std::memcpy(&blk[3 * n], dist_sc, sizeof(double) * n);
  std::memcpy(&blk[4 * n], dist_af, sizeof(double) * n);
  std::memcpy(&blk[5 * n], stretch, sizeof(double) * n_tds);
  std::memcpy(&blk[6 * n], stretch_correct, sizeof(double) * n_tds);
  X3D2C_CHECK_CUDA(cudaMalloc(&t->d_block, sizeof(double) * blk.size()));
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(t->d_block, blk.data(), sizeof(double) * blk.size(), cudaMemcpyHostToDevice, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  TdsDev& d = t->dev;
  d.n_tds = n_tds; d.n_rhs = n_rhs;
  std::memcpy(d.coeffs, coeffs, sizeof(double) * 9);
  std::memcpy(d.coeffs_s, coeffs_s, sizeof(double) * 36);
  std::memcpy(d.coeffs_e, coeffs_e, sizeof(double) * 36);
  d.fw = t->d_block + 0 * n; d.bw = t->d_block + 1 * n; d.sa = t->d_block + 2 * n; d.sc = t->d_block + 3 * n;
  d.af = t->d_block + 4 * n; d.stretch = t->d_block + 5 * n; d.stretch_correct = t->d_block + 6 * n;
  *out = t;
  return X3D2C_OK;
}

int x3d2c_tdsops_destroy(x3d2c_ctx* ctx, x3d2c_tdsops* ops) {
  X3D2C_ENTER(ctx);
  if (!ops) return X3D2C_OK;
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ops->d_block) cudaFree(ops->d_block);
  if (ops->d_m3) cudaFree(ops->d_m3);
  if (ops->d_stc) cudaFree(ops->d_stc);
  delete ops;
  return X3D2C_OK;
}

}  // extern "C"
