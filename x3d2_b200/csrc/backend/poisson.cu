// FFT Poisson solver (periodic 000 case) of the cuda_c backend.
// Replaces omp_poisson_fft_t / cuda_poisson_fft_t: fft_forward, fft_postprocess_000, fft_backward
//   src/backend/omp/poisson_fft.f90:49-97,129-137,169-181
//   src/backend/omp/kernels/spectral_processing.f90:7-106   (process_spectral_000)
//   src/backend/cuda/poisson_fft.f90:97-260,640-720         (prior art: cuFFT 3-D plans / cuFFTMp)
//
// cuFFT is used for batched 1-D transforms only; everything else is hand written:
//   forward : [strip padding] -> D2Z along x (batch ny*nz_loc) -> A(nxh, ny, nz_loc)
//             -> tile transpose A -> B(ny, nxh, nz_loc) -> Z2Z along y (contiguous, batch nxh*nz_loc)
//             -> [P > 1: pack + NCCL all-to-all: z-slabs -> y-slabs] -> C(ny_loc, nxh, nz)
//             -> Z2Z along z (stride ny_loc*nxh, dist 1, batch ny_loc*nxh)
//   spectral: one fused kernel on C: normalise, three half-cell phase rotations, division by the modified
//             wavenumber, three inverse rotations (arithmetic order of spectral_processing.f90:36-100)
//   backward: the mirror image, ending with Z2D along x into the DIR_C block.
// The spectrum is therefore stored as C(j_loc, i, k) (y fastest); the reference's waves(i, j, k) table is
// permuted to that order once at creation.
#include "common.cuh"

namespace x3d2c {
int alltoall(x3d2c_ctx* ctx, double* recv, const double* send, size_t block_doubles);  // nccl.cu
int allreduce(x3d2c_ctx* ctx, double* dev, size_t count, int op);                        // nccl.cu
}

namespace {

// A(i, j, k) [nxh, ny, nz] <-> B(j, i, k) [ny, nxh, nz]; one 32x32 tile of complex numbers per CTA
__global__ void __launch_bounds__(256)
cplx_transpose_kernel(double2* __restrict__ dst, const double2* __restrict__ src, const int n0, const int n1) {
  // src has fastest extent n0 and second extent n1; dst has fastest extent n1 and second extent n0
  __shared__ double2 tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const size_t plane = (size_t)n0 * n1 * blockIdx.z;
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + tx, j = j0 + ty + 8 * r;
    if (i < n0 && j < n1) tile[ty + 8 * r][tx] = src[plane + (size_t)j * n0 + i];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + tx, i = i0 + ty + 8 * r;
    if (i < n0 && j < n1) dst[plane + (size_t)i * n1 + j] = tile[tx][ty + 8 * r];
  }
}

// strip / restore the allocator padding of a DIR_C block (role of memcpy3D, cuda/kernels/spectral_processing.f90:10-60)
template <bool TO_COMPACT>
__global__ void __launch_bounds__(256)
pad_copy_kernel(double* __restrict__ compact, double* __restrict__ padded, const int nx, const int ny, const int nz,
                const int nx_pad, const int ny_pad) {
  const size_t n = (size_t)nx * ny * nz;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % nx);
    const size_t r = t / nx;
    const int j = (int)(r % ny), k = (int)(r / ny);
    const size_t p = i + (size_t)nx_pad * (j + (size_t)ny_pad * k);
    if (TO_COMPACT) compact[t] = padded[p];
    else padded[p] = compact[t];
  }
}

// z-slab -> y-slab packing: B(ny, nxh, nz_loc) -> send[r][k_loc][i][j_loc]  (and its inverse)
template <bool PACK>
__global__ void __launch_bounds__(256)
slab_pack_kernel(double2* __restrict__ packed, double2* __restrict__ b, const int ny, const int ny_loc,
                 const int nxh, const int nz_loc) {
  const size_t n = (size_t)ny * nxh * nz_loc;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % ny);
    const size_t q = t / ny;  // i + nxh * k_loc
    const int r = j / ny_loc, jl = j - r * ny_loc;
    const size_t p = ((size_t)r * nz_loc * nxh + q) * ny_loc + jl;
    if (PACK) packed[p] = b[t];
    else b[t] = packed[p];
  }
}

// The same exchange with the pack / unpack fused into the transfer: every element is written straight into the
// destination rank's buffer through its CUDA-IPC mapping (NVLink peer stores; the own block is an ordinary store).
//   FWD: local B(ny, nxh, nz_loc)  ->  rank r = j / ny_loc:  A_r[(rank * nz_loc * nxh + q) * ny_loc + jl]
//   BWD: local A = C(j_loc, i, k)  ->  rank s = k / nz_loc:  B_s[(rank * ny_loc + jl) + ny * q]
struct PeerPtrs { double2* p[8]; };
template <bool FWD>
__global__ void __launch_bounds__(256)
slab_exchange_kernel(const PeerPtrs dst, const double2* __restrict__ src, const int ny, const int ny_loc, const int nxh,
                     const int nz_loc, const int rank) {
  const size_t n = (size_t)ny * nxh * nz_loc;
  const size_t blk = (size_t)nz_loc * nxh;  // rows (i, k_loc) per rank block
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    if (FWD) {
      const int j = (int)(t % ny);
      const size_t q = t / ny;  // i + nxh * k_loc
      const int r = j / ny_loc, jl = j - r * ny_loc;
      dst.p[r][((size_t)rank * blk + q) * ny_loc + jl] = src[t];
    } else {
      const int jl = (int)(t % ny_loc);
      const size_t row = t / ny_loc;  // q + blk * s
      const int s = (int)(row / blk);
      const size_t q = row - (size_t)s * blk;
      dst.p[s][(size_t)rank * ny_loc + jl + (size_t)ny * q] = src[t];
    }
  }
}

// One chunk of planes of the pipelined exchange (rows q0 .. q0 + rows - 1 of (i, k_loc)), every element stored straight
// into the destination rank's buffer (peer stores over NVLink; the own block is an ordinary store):
//   FWD: local B(j, q)           ->  rank r = j / ny_loc:  Cx_r[(rank * blk + q) * ny_loc + jl]
//   BWD: local Cx(jl, q, s)      ->  rank s:               B_s[(rank * ny_loc + jl) + ny * q]
template <bool FWD>
__global__ void __launch_bounds__(256)
chunk_exchange_kernel(const PeerPtrs dst, const double2* __restrict__ src, const int ny, const int ny_loc, const int P,
                      const size_t blk, const size_t q0, const size_t rows, const int rank) {
  const size_t n = (size_t)ny * rows;  // == ny_loc * rows * P
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    if (FWD) {
      const int j = (int)(t % ny);
      const size_t q = q0 + t / ny;
      const int r = j / ny_loc, jl = j - r * ny_loc;
      dst.p[r][((size_t)rank * blk + q) * ny_loc + jl] = src[(size_t)ny * q + j];
    } else {
      const int jl = (int)(t % ny_loc);
      const size_t u = t / ny_loc;
      const int s = (int)(u / rows);
      const size_t q = q0 + (u - (size_t)s * rows);
      dst.p[s][(size_t)rank * ny_loc + jl + (size_t)ny * q] = src[((size_t)s * blk + q) * ny_loc + jl];
    }
  }
}

// cross-rank signalling of the pipelined exchange: flags in peer memory (monotonic solve counter)
struct FlagPtrs { unsigned long long* p[8]; };
__global__ void set_flags_kernel(const FlagPtrs dst, const int P, const int slot, const unsigned long long v) {
  const int r = threadIdx.x;
  if (r < P) {
    __threadfence_system();  // the copies issued before this kernel on the same stream are complete
    *reinterpret_cast<volatile unsigned long long*>(dst.p[r] + slot) = v;
    __threadfence_system();
  }
}
__global__ void wait_flags_kernel(const unsigned long long* flags, const int P, const int base, const int stride,
                                  const unsigned long long v) {
  const int r = threadIdx.x;
  if (r < P) {
    const volatile unsigned long long* f = flags + base + r * stride;
    while (*f < v) __nanosleep(200);
    __threadfence_system();
  }
}

struct SpecParams {
  int ny_loc, nxh, nz;  // extents of C(j_loc, i, k)
  int nx_g, ny_g, nz_g; // global cell dims
  int y_off;            // sp_st(2)
  int pow2;             // nx*ny*nz is a power of two => the normalisation is an exact scaling
  double inv_n;
};

// spectral_processing.f90:36-100, one thread per mode of C(j_loc, i, k)
template <bool STRICT>
__global__ void __launch_bounds__(256)
process_spectral_000_kernel(double2* __restrict__ c, const double2* __restrict__ waves,
                            const double* __restrict__ ax, const double* __restrict__ bx,
                            const double* __restrict__ ay, const double* __restrict__ by,
                            const double* __restrict__ az, const double* __restrict__ bz, const SpecParams p) {
  const int jl = blockIdx.x * blockDim.x + threadIdx.x;
  if (jl >= p.ny_loc) return;
  const int i = blockIdx.y, k = blockIdx.z;
  const size_t idx = jl + (size_t)p.ny_loc * (i + (size_t)p.nxh * k);
  const int ix = i, iy = jl + p.y_off, iz = k;  // 0-based
  double2 v = c[idx];
  double div_r, div_c;
  if (p.pow2) {
    div_r = v.x * p.inv_n;
    div_c = v.y * p.inv_n;
  } else {
    div_r = v.x / p.nx_g / p.ny_g / p.nz_g;
    div_c = v.y / p.nx_g / p.ny_g / p.nz_g;
  }
  const double axv = ax[ix], bxv = bx[ix], ayv = ay[iy], byv = by[iy], azv = az[iz], bzv = bz[iz];
  const bool fz = (iz + 1) > p.nz_g / 2 + 1, fy = (iy + 1) > p.ny_g / 2 + 1;
  double tr, tc;
#define ROT(A, B, SGN_R, SGN_C)                                  \
  tr = div_r; tc = div_c;                                        \
  if (STRICT) {                                                  \
    div_r = __dadd_rn(__dmul_rn(tr, B), SGN_R __dmul_rn(tc, A)); \
    div_c = __dadd_rn(SGN_C __dmul_rn(tc, B), -__dmul_rn(tr, A)); \
  } else {                                                       \
    div_r = tr * B SGN_R tc * A;                                 \
    div_c = SGN_C tc * B - tr * A;                               \
  }
  // forward: z, y, x
  ROT(azv, bzv, +, +)
  if (fz) { div_r = -div_r; div_c = -div_c; }
  ROT(ayv, byv, +, +)
  if (fy) { div_r = -div_r; div_c = -div_c; }
  ROT(axv, bxv, +, +)
  // solve
  const double2 w = waves[idx];  // fast mode: -1 / waves, or 0 (poisson_create_common)
  if (!STRICT) {
    div_r *= w.x;
    div_c *= w.y;
  } else if ((w.x < 1.e-16) || (w.y < 1.e-16)) {
    div_r = 0.0; div_c = 0.0;
  } else {
    div_r = -div_r / w.x;
    div_c = -div_c / w.y;
  }
  // backward: z (r*b - c*a, -c*b - r*a), y, x (r*b + c*a, -c*b + r*a)
  ROT(azv, bzv, -, -)
  if (fz) { div_r = -div_r; div_c = -div_c; }
  ROT(ayv, byv, +, +)
  if (fy) { div_r = -div_r; div_c = -div_c; }
  tr = div_r; tc = div_c;
  if (STRICT) {
    div_r = __dadd_rn(__dmul_rn(tr, bxv), __dmul_rn(tc, axv));
    div_c = __dadd_rn(-__dmul_rn(tc, bxv), __dmul_rn(tr, axv));
  } else {
    div_r = tr * bxv + tc * axv;
    div_c = -tc * bxv + tr * axv;
  }
#undef ROT
  c[idx] = make_double2(div_r, div_c);
}

int upload(double** dst, const double* src, size_t n, cudaStream_t s) {
  X3D2C_CHECK_CUDA(cudaMalloc(dst, sizeof(double) * n));
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(*dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  return X3D2C_OK;
}

bool slab_z(const x3d2c_ctx* ctx) { return ctx->cfg.nproc_dir[0] == 1 && ctx->cfg.nproc_dir[1] == 1; }

}  // namespace

using namespace x3d2c;

extern "C" {

int x3d2c_poisson_spec_layout(const x3d2c_ctx* ctx, int n_spec[3], int n_sp_st[3]) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && n_spec && n_sp_st, "x3d2c_poisson_spec_layout: null argument");
  X3D2C_REQUIRE(slab_z(ctx), "the cuda_c FFT Poisson solver needs nproc_dir = (1, 1, P)");
  const int P = ctx->cfg.nproc, ny = ctx->cfg.dims_cell_global[1];
  X3D2C_REQUIRE(ny % P == 0, "ny must be divisible by the number of ranks");
  n_spec[0] = ctx->cfg.dims_cell_global[0] / 2 + 1;
  n_spec[1] = ny / P;
  n_spec[2] = ctx->cfg.dims_cell_global[2];
  n_sp_st[0] = 0; n_sp_st[1] = (ny / P) * ctx->cfg.rank; n_sp_st[2] = 0;
  return X3D2C_OK;
}

// Maps every peer's A and B buffers into this process (CUDA IPC). All ranks agree on the outcome; when any mapping
// fails the NCCL send/recv exchange stays in use.
static int setup_peer_buffers(x3d2c_ctx* ctx, x3d2c_poisson* p) {
  const int P = ctx->cfg.nproc, me = ctx->cfg.rank;
  constexpr int HD = 3 * sizeof(cudaIpcMemHandle_t) / sizeof(double);  // doubles per rank: handles of A, B and Cx
  static_assert(sizeof(cudaIpcMemHandle_t) % sizeof(double) == 0, "handle size");
  std::vector<double> mine(HD), all((size_t)HD * P), rep((size_t)HD * P);
  cudaIpcMemHandle_t h[3];
  std::memset(h, 0, sizeof h);
  bool ok = cudaIpcGetMemHandle(&h[0], p->A) == cudaSuccess && cudaIpcGetMemHandle(&h[1], p->B) == cudaSuccess;
  if (ok && p->Cx) ok = cudaIpcGetMemHandle(&h[2], p->Cx) == cudaSuccess;
  if (!ok) cudaGetLastError();
  std::memcpy(mine.data(), h, sizeof h);
  for (int r = 0; r < P; ++r) std::memcpy(&rep[(size_t)r * HD], mine.data(), sizeof h);
  double *d_send = nullptr, *d_recv = nullptr;
  X3D2C_CHECK_CUDA(cudaMalloc(&d_send, sizeof(double) * HD * P));
  X3D2C_CHECK_CUDA(cudaMalloc(&d_recv, sizeof(double) * HD * P));
  X3D2C_CHECK_CUDA(cudaMalloc(&p->bar_word, sizeof(double) * 2));
  X3D2C_CHECK_CUDA(cudaMemsetAsync(p->bar_word, 0, sizeof(double) * 2, ctx->stream));
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(d_send, rep.data(), sizeof(double) * HD * P, cudaMemcpyHostToDevice, ctx->stream));
  int rc = alltoall(ctx, d_recv, d_send, HD);
  if (rc) return rc;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(all.data(), d_recv, sizeof(double) * HD * P, cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < P && ok; ++r) {
    if (r == me) { p->peerA[r] = p->A; p->peerB[r] = p->B; p->peerC[r] = p->Cx; continue; }
    cudaIpcMemHandle_t hr[3];
    std::memcpy(hr, &all[(size_t)r * HD], sizeof hr);
    void *pa = nullptr, *pb = nullptr, *pc = nullptr;
    ok = cudaIpcOpenMemHandle(&pa, hr[0], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
         cudaIpcOpenMemHandle(&pb, hr[1], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (ok && p->Cx) ok = cudaIpcOpenMemHandle(&pc, hr[2], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (!ok) cudaGetLastError();
    p->peerA[r] = (cufftDoubleComplex*)pa;
    p->peerB[r] = (cufftDoubleComplex*)pb;
    p->peerC[r] = (cufftDoubleComplex*)pc;
  }
  // agreement: minimum of the success flags
  const double flag = ok ? 0.0 : 1.0;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(p->bar_word, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = allreduce(ctx, p->bar_word, 1, 0))) return rc;
  double failed = 1.0;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(&failed, p->bar_word, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  p->p2p = failed == 0.0;
  cudaFree(d_send);
  cudaFree(d_recv);
  if (std::getenv("X3D2C_TRACE"))
    std::fprintf(stderr, "[x3d2c] rank %d poisson slab exchange: %s\n", me,
                 p->p2p ? "peer stores over NVLink (CUDA IPC)" : "ncclSend/Recv");
  return X3D2C_OK;
}

// barrier across ranks on the context's stream (tiny all-reduce)
static int stream_barrier(x3d2c_ctx* ctx, x3d2c_poisson* p) { return allreduce(ctx, p->bar_word + 1, 1, 0); }

int x3d2c_poisson_create(x3d2c_ctx* ctx, const double* waves, const double* ax, const double* bx,
                         const double* ay, const double* by, const double* az, const double* bz,
                         x3d2c_poisson** out) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && waves && ax && bx && ay && by && az && bz && out, "x3d2c_poisson_create: null argument");
  X3D2C_REQUIRE(ctx->cfg.periodic[0] && ctx->cfg.periodic[1] && ctx->cfg.periodic[2],
                "x3d2c_poisson_create: the fully periodic (000) solver; walls in y: x3d2c_poisson_create_010");
  return x3d2c::poisson_create_common(ctx, 0, waves, ax, bx, ay, by, az, bz, out);
}

}  // extern "C"

int x3d2c::poisson_create_common(x3d2c_ctx* ctx, int bc_case, const double* waves, const double* ax, const double* bx,
                                 const double* ay, const double* by, const double* az, const double* bz,
                                 x3d2c_poisson** out) {
  X3D2C_REQUIRE(slab_z(ctx), "x3d2c_poisson_create: needs nproc_dir = (1, 1, P)");
  struct Guard {  // no leak on the early returns below
    x3d2c_ctx* c;
    x3d2c_poisson* p;
    ~Guard() { if (p) x3d2c_poisson_destroy(c, p); }
  } guard{ctx, new x3d2c_poisson};
  x3d2c_poisson* p = guard.p;
  p->bc_case = bc_case;
  const int P = ctx->cfg.nproc;
  p->nx = ctx->cfg.dims_cell_global[0]; p->ny = ctx->cfg.dims_cell_global[1]; p->nz = ctx->cfg.dims_cell_global[2];
  p->nxh = p->nx / 2 + 1;
  p->nz_loc = ctx->cfg.dims_cell[2];
  X3D2C_REQUIRE(p->ny % P == 0 && p->nz_loc * P == p->nz, "x3d2c_poisson_create: ny, nz must divide by nproc");
  p->ny_loc = p->ny / P;
  X3D2C_REQUIRE(p->nz <= 65535 && p->nxh <= 65535, "x3d2c_poisson_create: grid limit exceeded");
  const size_t n_spec = (size_t)p->nxh * p->ny_loc * p->nz;  // == nxh * ny * nz_loc
  X3D2C_CHECK_CUDA(cudaMalloc(&p->A, sizeof(cufftDoubleComplex) * n_spec));
  X3D2C_CHECK_CUDA(cudaMalloc(&p->B, sizeof(cufftDoubleComplex) * n_spec));
  // waves(i, j, k) -> C order (j, i, k)
  std::vector<double> wc(2 * n_spec);
  for (int k = 0; k < p->nz; ++k)
    for (int j = 0; j < p->ny_loc; ++j)
      for (int i = 0; i < p->nxh; ++i) {
        const size_t s = i + (size_t)p->nxh * (j + (size_t)p->ny_loc * k);
        const size_t d = j + (size_t)p->ny_loc * (i + (size_t)p->nxh * k);
        wc[2 * d] = waves[2 * s];
        wc[2 * d + 1] = waves[2 * s + 1];
        if (bc_case == 0 && !ctx->strict) {
          // fast mode: the kernel multiplies by -1 / waves (0 for the modes the reference sets to zero) instead of
          // dividing twice per mode - the two FP64 divisions were a third of process_spectral_000's issue slots
          const bool zero = waves[2 * s] < 1.e-16 || waves[2 * s + 1] < 1.e-16;
          wc[2 * d] = zero ? 0.0 : -1.0 / waves[2 * s];
          wc[2 * d + 1] = zero ? 0.0 : -1.0 / waves[2 * s + 1];
        }
      }
  int rc;
  if ((rc = upload(&p->waves, wc.data(), wc.size(), ctx->stream))) return rc;
  if ((rc = upload(&p->ax, ax, p->nx, ctx->stream))) return rc;
  if ((rc = upload(&p->bx, bx, p->nx, ctx->stream))) return rc;
  if ((rc = upload(&p->ay, ay, p->ny, ctx->stream))) return rc;
  if ((rc = upload(&p->by, by, p->ny, ctx->stream))) return rc;
  if ((rc = upload(&p->az, az, p->nz, ctx->stream))) return rc;
  if ((rc = upload(&p->bz, bz, p->nz, ctx->stream))) return rc;
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  const bool padded = ctx->nx_pad != p->nx || ctx->ny_pad != p->ny;
  if (padded) X3D2C_CHECK_CUDA(cudaMalloc(&p->compact, sizeof(double) * (size_t)p->nx * p->ny * p->nz_loc));
  // batched 1-D plans
  int n_x[1] = {p->nx}, n_y[1] = {p->ny}, n_z[1] = {p->nz};
  X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_r2c, 1, n_x, n_x, 1, p->nx, n_x, 1, p->nxh, CUFFT_D2Z, p->ny * p->nz_loc));
  X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_c2r, 1, n_x, n_x, 1, p->nxh, n_x, 1, p->nx, CUFFT_Z2D, p->ny * p->nz_loc));
  X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_y, 1, n_y, n_y, 1, p->ny, n_y, 1, p->ny, CUFFT_Z2Z, p->nxh * p->nz_loc));
  const int zs = p->ny_loc * p->nxh;
  X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_z, 1, n_z, n_z, zs, 1, n_z, zs, 1, CUFFT_Z2Z, zs));
  X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_r2c, ctx->stream));
  X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_c2r, ctx->stream));
  X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_y, ctx->stream));
  X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_z, ctx->stream));
  p->have_plans = true;
  if (P > 1 && P <= 8 && !std::getenv("X3D2C_NO_P2P")) {
    // pipelined exchange: a dedicated destination buffer (+ 1 KB of flags behind it), chunked plans, a copy stream
    int nch = 4;  // X3D2C_PIPE_CHUNKS: 1..16
    if (const char* e = std::getenv("X3D2C_PIPE_CHUNKS")) nch = std::atoi(e) > 0 && std::atoi(e) <= 16 ? std::atoi(e) : nch;
    while (nch > 1 && p->nz_loc % nch) nch >>= 1;
    const bool want_pipe = !std::getenv("X3D2C_NO_PIPE");
    if (want_pipe) {
      X3D2C_CHECK_CUDA(cudaMalloc(&p->Cx, sizeof(cufftDoubleComplex) * n_spec + 2048));
      X3D2C_CHECK_CUDA(cudaMemset(p->Cx, 0, sizeof(cufftDoubleComplex) * n_spec + 2048));
      p->flags = reinterpret_cast<unsigned long long*>(p->Cx + n_spec);
    }
    rc = setup_peer_buffers(ctx, p);
    if (rc) return rc;
    if (p->p2p && want_pipe) {
      p->nch = nch;
      const int pl = p->nz_loc / nch;
      for (int r = 0; r < P; ++r) p->peerFlags[r] = reinterpret_cast<unsigned long long*>(p->peerC[r] + n_spec);
      X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_r2c_c, 1, n_x, n_x, 1, p->nx, n_x, 1, p->nxh, CUFFT_D2Z, p->ny * pl));
      X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_c2r_c, 1, n_x, n_x, 1, p->nxh, n_x, 1, p->nx, CUFFT_Z2D, p->ny * pl));
      X3D2C_CHECK_CUFFT(cufftPlanMany(&p->plan_y_c, 1, n_y, n_y, 1, p->ny, n_y, 1, p->ny, CUFFT_Z2Z, p->nxh * pl));
      X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_r2c_c, ctx->stream));
      X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_c2r_c, ctx->stream));
      X3D2C_CHECK_CUFFT(cufftSetStream(p->plan_y_c, ctx->stream));
      // the exchange stream outranks the transforms: its blocks are placed first whenever an SM frees slots, so the
      // NVLink stores start as soon as a chunk is ready instead of behind the queued blocks of the next transform
      int prio_lo = 0, prio_hi = 0;
      X3D2C_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      X3D2C_CHECK_CUDA(cudaStreamCreateWithPriority(&p->s2, cudaStreamNonBlocking, getenv("X3D2C_PIPE_NO_PRIO") ? prio_lo : prio_hi));
      for (int c = 0; c < nch; ++c) X3D2C_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_chunk[c], cudaEventDisableTiming));
      X3D2C_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_z, cudaEventDisableTiming));
      p->pipe = true;
      if (std::getenv("X3D2C_TRACE"))
        std::fprintf(stderr, "[x3d2c] rank %d poisson exchange pipelined: %d chunks of %d planes, peer stores on a second stream + flags\n",
                     ctx->cfg.rank, nch, pl);
    }
  }
  guard.p = nullptr;
  *out = p;
  return X3D2C_OK;
}

extern "C" {

int x3d2c_poisson_destroy(x3d2c_ctx* ctx, x3d2c_poisson* p) {
  X3D2C_ENTER(ctx);
  if (!p) return X3D2C_OK;
  cudaStreamSynchronize(ctx->stream);
  if (p->have_plans) { cufftDestroy(p->plan_r2c); cufftDestroy(p->plan_c2r); cufftDestroy(p->plan_y); cufftDestroy(p->plan_z); }
  for (int r = 0; r < 8; ++r) {
    if (r == ctx->cfg.rank) continue;
    if (p->peerA[r]) cudaIpcCloseMemHandle(p->peerA[r]);
    if (p->peerB[r]) cudaIpcCloseMemHandle(p->peerB[r]);
  }
  for (int r = 0; r < 8; ++r)
    if (r != ctx->cfg.rank && p->peerC[r]) cudaIpcCloseMemHandle(p->peerC[r]);
  if (p->plan_r2c_c) { cufftDestroy(p->plan_r2c_c); cufftDestroy(p->plan_c2r_c); cufftDestroy(p->plan_y_c); }
  for (int c = 0; c < 16; ++c)
    if (p->ev_chunk[c]) cudaEventDestroy(p->ev_chunk[c]);
  if (p->ev_z) cudaEventDestroy(p->ev_z);
  if (p->s2) cudaStreamDestroy(p->s2);
  if (p->Cx) cudaFree(p->Cx);
  if (p->bar_word) cudaFree(p->bar_word);
  if (p->fac_im && p->fac_im != p->fac_re) cudaFree(p->fac_im);
  if (p->fac_re) cudaFree(p->fac_re);
  for (void* q : {(void*)p->A, (void*)p->B, (void*)p->waves, (void*)p->ax, (void*)p->bx, (void*)p->ay, (void*)p->by,
                  (void*)p->az, (void*)p->bz, (void*)p->compact})
    if (q) cudaFree(q);
  delete p;
  return X3D2C_OK;
}

// the spectrum C(j_loc, i, k) lives in p->B, or in p->A when the peers write it there directly
static cufftDoubleComplex* spec_buf(const x3d2c_ctx*, x3d2c_poisson* p) { return p->pipe ? p->Cx : (p->p2p ? p->A : p->B); }

// Pipelined forward transform (P > 1): the slab is processed in chunks of planes; x transform, transpose and y
// transform of chunk c + 1 run on the context's stream while an exchange kernel moves chunk c to the y slabs of all
// ranks with peer stores (second stream). No all-reduce barrier: every rank raises a flag on every peer once its copies are complete.
// blocks of the chunk exchange: 4 per SM keep half of every SM's thread slots free for the transforms of the next chunk
static int exchange_blocks() {
  static const int n = [] {
    const char* e = getenv("X3D2C_PIPE_BLOCKS");
    return e && atoi(e) > 0 ? atoi(e) : 148 * 4;
  }();
  return n;
}

// Diagnostic timeline (X3D2C_POISSON_TIMELINE=1): events recorded on both streams during the sixth solve, printed as
// milliseconds since the start of its forward transform.
struct Timeline {
  static constexpr int kMax = 96;
  cudaEvent_t ev[kMax];
  const char* label[kMax];
  int chunk[kMax];
  int n = 0, calls = 0;
  bool enabled = false, armed = false, created = false;
  void begin(cudaStream_t s) {
    static const bool on = getenv("X3D2C_POISSON_TIMELINE") != nullptr;
    enabled = on;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (on) cudaStreamIsCapturing(s, &st);
    armed = enabled && st == cudaStreamCaptureStatusNone && ++calls == 6;
    if (armed && !created) {
      for (auto& e : ev) cudaEventCreate(&e);
      created = true;
    }
    n = 0;
  }
  void mark(const char* what, int c, cudaStream_t s) {
    if (!armed || n >= kMax) return;
    label[n] = what;
    chunk[n] = c;
    cudaEventRecord(ev[n++], s);
  }
  void report(int rank, cudaStream_t s1, cudaStream_t s2) {
    if (!armed) return;
    cudaStreamSynchronize(s1);
    cudaStreamSynchronize(s2);
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[0], ev[i]);
      fprintf(stderr, "[x3d2c] rank %d poisson timeline %7.3f ms  %s %d\n", rank, ms, label[i], chunk[i]);
    }
    armed = false;
  }
};
static Timeline g_tl;

static int forward_pipelined(x3d2c_ctx* ctx, x3d2c_poisson* p, const double* in) {
  const int P = ctx->cfg.nproc, me = ctx->cfg.rank, nch = p->nch, pl = p->nz_loc / nch;
  const size_t plane_r = (size_t)p->nx * p->ny, plane_c = (size_t)p->nxh * p->ny;
  const size_t blk = (size_t)p->nz_loc * p->nxh;  // rows (i, k_loc) of one rank block in Cx
  p->epoch++;
  FlagPtrs fp;
  PeerPtrs dstC;
  for (int r = 0; r < 8; ++r) { fp.p[r] = p->peerFlags[r]; dstC.p[r] = (double2*)p->peerC[r]; }
  g_tl.begin(ctx->stream);
  g_tl.mark("start", 0, ctx->stream);
  for (int c = 0; c < nch; ++c) {
    cufftDoubleComplex* a = p->A + (size_t)c * pl * plane_c;
    cufftDoubleComplex* b = p->B + (size_t)c * pl * plane_c;
    X3D2C_CHECK_CUFFT(cufftExecD2Z(p->plan_r2c_c, const_cast<double*>(in) + (size_t)c * pl * plane_r, a));
    ctx->launches++;
    const dim3 grid((p->nxh + 31) / 32, (p->ny + 31) / 32, pl), block(32, 8);
    cplx_transpose_kernel<<<grid, block, 0, ctx->stream>>>((double2*)b, (const double2*)a, p->nxh, p->ny);
    X3D2C_CHECK_LAUNCH(ctx);
    X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_y_c, b, b, CUFFT_FORWARD));
    ctx->launches++;
    X3D2C_CHECK_CUDA(cudaEventRecord(p->ev_chunk[c], ctx->stream));
    X3D2C_CHECK_CUDA(cudaStreamWaitEvent(p->s2, p->ev_chunk[c], 0));
    const size_t q0 = (size_t)p->nxh * pl * c;  // first row (i + nxh k_loc) of the chunk
    chunk_exchange_kernel<true><<<exchange_blocks(), 256, 0, p->s2>>>(dstC, (const double2*)p->B, p->ny, p->ny_loc, P, blk, q0,
                                                                   (size_t)p->nxh * pl, me);
    X3D2C_CHECK_LAUNCH(ctx);
    g_tl.mark("fwd xy transforms done (main)", c, ctx->stream);
    g_tl.mark("fwd exchange done (s2)", c, p->s2);
  }
  set_flags_kernel<<<1, 32, 0, p->s2>>>(fp, P, me, p->epoch);
  X3D2C_CHECK_LAUNCH(ctx);
  wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(p->flags, P, 0, 1, p->epoch);
  X3D2C_CHECK_LAUNCH(ctx);
  g_tl.mark("fwd all ranks delivered (main)", 0, ctx->stream);
  X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_z, p->Cx, p->Cx, CUFFT_FORWARD));
  ctx->launches++;
  g_tl.mark("fwd z transform done (main)", 0, ctx->stream);
  return X3D2C_OK;
}

// Pipelined backward transform: after the z transform exchange kernels deliver the planes chunk by chunk to their
// z slabs; a rank starts the y transform, transpose and x transform of a chunk as soon as all ranks have flagged it.
static int backward_pipelined(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_c) {
  const int P = ctx->cfg.nproc, me = ctx->cfg.rank, nch = p->nch, pl = p->nz_loc / nch;
  const size_t plane_r = (size_t)p->nx * p->ny, plane_c = (size_t)p->nxh * p->ny;
  const size_t blk = (size_t)p->nz_loc * p->nxh;
  FlagPtrs fp;
  PeerPtrs dstB;
  for (int r = 0; r < 8; ++r) { fp.p[r] = p->peerFlags[r]; dstB.p[r] = (double2*)p->peerB[r]; }
  g_tl.mark("bwd start (main)", 0, ctx->stream);
  X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_z, p->Cx, p->Cx, CUFFT_INVERSE));
  ctx->launches++;
  g_tl.mark("bwd z transform done (main)", 0, ctx->stream);
  X3D2C_CHECK_CUDA(cudaEventRecord(p->ev_z, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamWaitEvent(p->s2, p->ev_z, 0));
  for (int c = 0; c < nch; ++c) {
    const size_t q0 = (size_t)p->nxh * pl * c;
    chunk_exchange_kernel<false><<<exchange_blocks(), 256, 0, p->s2>>>(dstB, (const double2*)p->Cx, p->ny, p->ny_loc, P, blk, q0,
                                                                    (size_t)p->nxh * pl, me);
    X3D2C_CHECK_LAUNCH(ctx);
    set_flags_kernel<<<1, 32, 0, p->s2>>>(fp, P, 8 + 16 * me + c, p->epoch);
    X3D2C_CHECK_LAUNCH(ctx);
    g_tl.mark("bwd exchange done (s2)", c, p->s2);
  }
  double* outp = p->compact ? p->compact : f_c;
  for (int c = 0; c < nch; ++c) {
    wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(p->flags, P, 8 + c, 16, p->epoch);
    X3D2C_CHECK_LAUNCH(ctx);
    g_tl.mark("bwd chunk delivered by all ranks (main)", c, ctx->stream);
    cufftDoubleComplex* a = p->A + (size_t)c * pl * plane_c;
    cufftDoubleComplex* b = p->B + (size_t)c * pl * plane_c;
    X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_y_c, b, b, CUFFT_INVERSE));
    ctx->launches++;
    const dim3 grid((p->ny + 31) / 32, (p->nxh + 31) / 32, pl), block(32, 8);
    cplx_transpose_kernel<<<grid, block, 0, ctx->stream>>>((double2*)a, (const double2*)b, p->ny, p->nxh);
    X3D2C_CHECK_LAUNCH(ctx);
    X3D2C_CHECK_CUFFT(cufftExecZ2D(p->plan_c2r_c, a, outp + (size_t)c * pl * plane_r));
    ctx->launches++;
    g_tl.mark("bwd yx transforms done (main)", c, ctx->stream);
  }
  if (p->compact) {
    pad_copy_kernel<false><<<1184, 256, 0, ctx->stream>>>(p->compact, f_c, p->nx, p->ny, p->nz_loc, ctx->nx_pad, ctx->ny_pad);
    X3D2C_CHECK_LAUNCH(ctx);
  }
  g_tl.report(ctx->cfg.rank, ctx->stream, p->s2);
  return X3D2C_OK;
}

int x3d2c_fft_forward(x3d2c_ctx* ctx, x3d2c_poisson* p, const double* f_c) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p && f_c, "x3d2c_fft_forward: null argument");
  const double* in = f_c;
  if (p->compact) {
    pad_copy_kernel<true><<<1184, 256, 0, ctx->stream>>>(p->compact, const_cast<double*>(f_c), p->nx, p->ny,
                                                         p->nz_loc, ctx->nx_pad, ctx->ny_pad);
    X3D2C_CHECK_LAUNCH(ctx);
    in = p->compact;
  }
  if (p->pipe) return forward_pipelined(ctx, p, in);
  g_tl.begin(ctx->stream);
  g_tl.mark("start", 0, ctx->stream);
  X3D2C_CHECK_CUFFT(cufftExecD2Z(p->plan_r2c, const_cast<double*>(in), p->A));
  ctx->launches++;
  g_tl.mark("fwd x transform done", 0, ctx->stream);
  const dim3 grid((p->nxh + 31) / 32, (p->ny + 31) / 32, p->nz_loc), block(32, 8);
  cplx_transpose_kernel<<<grid, block, 0, ctx->stream>>>((double2*)p->B, (const double2*)p->A, p->nxh, p->ny);
  X3D2C_CHECK_LAUNCH(ctx);
  g_tl.mark("fwd transpose done", 0, ctx->stream);
  X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_y, p->B, p->B, CUFFT_FORWARD));
  ctx->launches++;
  g_tl.mark("fwd y transform done", 0, ctx->stream);
  cufftDoubleComplex* c = p->B;
  if (p->p2p) {
    // every rank has finished reading its A (transpose above) -> A may be written by the peers; afterwards: all
    // peer stores have landed
    int rc = stream_barrier(ctx, p);
    if (rc) return rc;
    PeerPtrs dst;
    for (int r = 0; r < 8; ++r) dst.p[r] = (double2*)p->peerA[r];
    slab_exchange_kernel<true><<<1184, 256, 0, ctx->stream>>>(dst, (const double2*)p->B, p->ny, p->ny_loc, p->nxh,
                                                              p->nz_loc, ctx->cfg.rank);
    X3D2C_CHECK_LAUNCH(ctx);
    g_tl.mark("fwd barrier + exchange kernel done", 0, ctx->stream);
    if ((rc = stream_barrier(ctx, p))) return rc;
    g_tl.mark("fwd second barrier done", 0, ctx->stream);
    c = p->A;
  } else if (ctx->cfg.nproc > 1) {
    slab_pack_kernel<true><<<1184, 256, 0, ctx->stream>>>((double2*)p->A, (double2*)p->B, p->ny, p->ny_loc, p->nxh, p->nz_loc);
    X3D2C_CHECK_LAUNCH(ctx);
    int rc = alltoall(ctx, (double*)p->B, (const double*)p->A, 2 * (size_t)p->ny_loc * p->nxh * p->nz_loc);
    if (rc) return rc;
    // received blocks [s][k_loc][i][j_loc] are exactly C(j_loc, i, k = s*nz_loc + k_loc)
  }
  X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_z, c, c, CUFFT_FORWARD));
  ctx->launches++;
  g_tl.mark("fwd z transform done", 0, ctx->stream);
  return X3D2C_OK;
}

int x3d2c_fft_postprocess_000(x3d2c_ctx* ctx, x3d2c_poisson* p) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p, "x3d2c_fft_postprocess_000: null argument");
  SpecParams sp;
  sp.ny_loc = p->ny_loc; sp.nxh = p->nxh; sp.nz = p->nz;
  sp.nx_g = p->nx; sp.ny_g = p->ny; sp.nz_g = p->nz;
  sp.y_off = p->ny_loc * ctx->cfg.rank;
  const long long N = (long long)p->nx * p->ny * p->nz;
  sp.pow2 = (N & (N - 1)) == 0;
  sp.inv_n = 1.0 / (double)N;
  const int bx_ = p->ny_loc >= 256 ? 256 : ((p->ny_loc + 31) / 32) * 32;
  const dim3 grid((p->ny_loc + bx_ - 1) / bx_, p->nxh, p->nz), block(bx_);
  cufftDoubleComplex* c = spec_buf(ctx, p);
  if (ctx->strict)
    process_spectral_000_kernel<true><<<grid, block, 0, ctx->stream>>>((double2*)c, (const double2*)p->waves, p->ax,
                                                                      p->bx, p->ay, p->by, p->az, p->bz, sp);
  else
    process_spectral_000_kernel<false><<<grid, block, 0, ctx->stream>>>((double2*)c, (const double2*)p->waves, p->ax,
                                                                       p->bx, p->ay, p->by, p->az, p->bz, sp);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_fft_backward(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_c) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p && f_c, "x3d2c_fft_backward: null argument");
  if (p->pipe) return backward_pipelined(ctx, p, f_c);
  cufftDoubleComplex* c = spec_buf(ctx, p);
  g_tl.mark("bwd start", 0, ctx->stream);
  X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_z, c, c, CUFFT_INVERSE));
  ctx->launches++;
  g_tl.mark("bwd z transform done", 0, ctx->stream);
  if (p->p2p) {
    // B is free on every rank since the barrier that followed the forward exchange
    PeerPtrs dst;
    for (int r = 0; r < 8; ++r) dst.p[r] = (double2*)p->peerB[r];
    slab_exchange_kernel<false><<<1184, 256, 0, ctx->stream>>>(dst, (const double2*)p->A, p->ny, p->ny_loc, p->nxh,
                                                               p->nz_loc, ctx->cfg.rank);
    X3D2C_CHECK_LAUNCH(ctx);
    int rc = stream_barrier(ctx, p);
    if (rc) return rc;
    g_tl.mark("bwd exchange kernel + barrier done", 0, ctx->stream);
  } else if (ctx->cfg.nproc > 1) {
    // y-slabs -> z-slabs: block s of C (k range of rank s) goes back to rank s
    int rc = alltoall(ctx, (double*)p->A, (const double*)p->B, 2 * (size_t)p->ny_loc * p->nxh * p->nz_loc);
    if (rc) return rc;
    slab_pack_kernel<false><<<1184, 256, 0, ctx->stream>>>((double2*)p->A, (double2*)p->B, p->ny, p->ny_loc, p->nxh, p->nz_loc);
    X3D2C_CHECK_LAUNCH(ctx);
  }
  X3D2C_CHECK_CUFFT(cufftExecZ2Z(p->plan_y, p->B, p->B, CUFFT_INVERSE));
  ctx->launches++;
  g_tl.mark("bwd y transform done", 0, ctx->stream);
  const dim3 grid((p->ny + 31) / 32, (p->nxh + 31) / 32, p->nz_loc), block(32, 8);
  cplx_transpose_kernel<<<grid, block, 0, ctx->stream>>>((double2*)p->A, (const double2*)p->B, p->ny, p->nxh);
  X3D2C_CHECK_LAUNCH(ctx);
  g_tl.mark("bwd transpose done", 0, ctx->stream);
  double* outp = p->compact ? p->compact : f_c;
  X3D2C_CHECK_CUFFT(cufftExecZ2D(p->plan_c2r, p->A, outp));
  ctx->launches++;
  g_tl.mark("bwd x transform done", 0, ctx->stream);
  if (p->compact) {
    pad_copy_kernel<false><<<1184, 256, 0, ctx->stream>>>(p->compact, f_c, p->nx, p->ny, p->nz_loc, ctx->nx_pad, ctx->ny_pad);
    X3D2C_CHECK_LAUNCH(ctx);
  }
  g_tl.report(ctx->cfg.rank, ctx->stream, ctx->stream);
  return X3D2C_OK;
}

int x3d2c_poisson_get_spectrum(x3d2c_ctx* ctx, x3d2c_poisson* p, double* host_spec) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p && host_spec, "x3d2c_poisson_get_spectrum: null argument");
  const size_t n = (size_t)p->nxh * p->ny_loc * p->nz;
  std::vector<double> tmp(2 * n);
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(tmp.data(), spec_buf(ctx, p), sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < p->nz; ++k)
    for (int j = 0; j < p->ny_loc; ++j)
      for (int i = 0; i < p->nxh; ++i) {
        const size_t d = i + (size_t)p->nxh * (j + (size_t)p->ny_loc * k);
        const size_t s = j + (size_t)p->ny_loc * (i + (size_t)p->nxh * k);
        host_spec[2 * d] = tmp[2 * s];
        host_spec[2 * d + 1] = tmp[2 * s + 1];
      }
  return X3D2C_OK;
}

}  // extern "C"
