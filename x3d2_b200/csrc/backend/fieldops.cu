// Elementwise field operations and reductions of the cuda_c backend.
// Replaces veccopy/vecadd/vecmult/field_scale/field_shift (src/backend/omp/backend.f90:529-614,883-901),
// scalar_product (:651-712), field_max_mean (:739-810), field_volume_integral (:1023-1066) and the
// CUDA-Fortran kernels of src/backend/cuda/kernels/fieldops.f90.
// Streaming ops move 128-bit vectors with a grid of 148 SMs x 8 CTAs; reductions are a deterministic
// two-stage tree (per-CTA partials in a fixed order, then one CTA), not per-thread atomics.
#include <algorithm>

#include "common.cuh"

namespace x3d2c {
int allreduce(x3d2c_ctx* ctx, double* dev, size_t count, int op);  // nccl.cu
}

namespace {

constexpr int kBlocks = 148 * 8;
constexpr int kThreads = 256;

enum { OP_COPY, OP_AXPBY, OP_MULT, OP_SCALE, OP_SHIFT, OP_FILL };

template <int OP>
__global__ void __launch_bounds__(kThreads)
stream_kernel(double* __restrict__ y, const double* __restrict__ x, const double a, const double b,
              const long long n2) {  // n2 = number of double2 elements
  const long long stride = (long long)gridDim.x * blockDim.x;
  double2* y2 = reinterpret_cast<double2*>(y);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 r;
    if (OP == OP_COPY) {
      r = x2[i];
    } else if (OP == OP_AXPBY) {
      const double2 xv = x2[i], yv = y2[i];
      r.x = a * xv.x + b * yv.x;
      r.y = a * xv.y + b * yv.y;
    } else if (OP == OP_MULT) {
      const double2 xv = x2[i], yv = y2[i];
      r.x = yv.x * xv.x;
      r.y = yv.y * xv.y;
    } else if (OP == OP_SCALE) {
      const double2 yv = y2[i];
      r.x = a * yv.x;
      r.y = a * yv.y;
    } else if (OP == OP_SHIFT) {
      const double2 yv = y2[i];
      r.x = yv.x + a;
      r.y = yv.y + a;
    } else {
      r.x = a;
      r.y = a;
    }
    y2[i] = r;
  }
}

// strict variant of axpby: a*x + b*y with separately rounded products (omp/backend.f90:578)
__global__ void __launch_bounds__(kThreads)
axpby_strict_kernel(double* __restrict__ y, const double* __restrict__ x, const double a, const double b,
                    const long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
}

// out = base; out = c_k x_k + out, k = 0..n-1: the vecadd(c_k, x_k, 1.0, out) chain of the time integrators
// (src/time_integrator.f90:166-231) in one pass, each term rounded exactly like OP_AXPBY / axpby_strict_kernel.
struct LinComb {
  const double* base;
  const double* x[4];
  double c[4];
  int n;
};
template <bool STRICT>
__global__ void __launch_bounds__(kThreads)
lincomb_kernel(double* out, const __grid_constant__ LinComb q, const long long n2) {  // out may alias q.base
  const long long stride = (long long)gridDim.x * blockDim.x;
  double2* o2 = reinterpret_cast<double2*>(out);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 r = reinterpret_cast<const double2*>(q.base)[i];
    double2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < q.n) v[k] = reinterpret_cast<const double2*>(q.x[k])[i];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < q.n) {
        if (STRICT) {
          r.x = __dadd_rn(__dmul_rn(q.c[k], v[k].x), __dmul_rn(1.0, r.x));
          r.y = __dadd_rn(__dmul_rn(q.c[k], v[k].y), __dmul_rn(1.0, r.y));
        } else {
          r.x = q.c[k] * v[k].x + 1.0 * r.x;
          r.y = q.c[k] * v[k].y + 1.0 * r.y;
        }
      }
    o2[i] = r;
  }
}

// field_set_face / field_set_face_from_field on DIR_X fields (omp/backend.f90:903-1021). One thread per face point.
// Y faces: points (x, z), rows y = 0 and y = ny - 1; X faces: points (y, z), columns x = 0 and x = nx - 1.
struct FaceGeom {
  int nx, ny, nz;  // extents of the data location
  int nx_pad, nyb;
};
__device__ __forceinline__ size_t dirx_index(const FaceGeom& g, int x, int y, int z) {
  return (size_t)(y % SZ) + (size_t)SZ * (x + (size_t)g.nx_pad * ((y / SZ) + (size_t)g.nyb * z));
}
template <bool FROM_FIELD>
__global__ void __launch_bounds__(256)
set_yface_kernel(double* __restrict__ f, const double* __restrict__ f_start, const double c_start, const double c_end,
                 const FaceGeom g) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
  if (x >= g.nx) return;
  const size_t lo = dirx_index(g, x, 0, z), hi = dirx_index(g, x, g.ny - 1, z);
  f[lo] = FROM_FIELD ? f_start[lo] : c_start;
  f[hi] = FROM_FIELD ? f_start[hi] : c_end;
}
// inlet from f_start, convective outflow: f(nx-1) -= c_end (f(nx-1) - f(nx-2)) - flow_rate_diff
__global__ void __launch_bounds__(256)
set_xface_from_field_kernel(double* __restrict__ f, const double* __restrict__ f_start, const double c_end,
                            const double flow_rate_diff, const FaceGeom g) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
  if (y >= g.ny) return;
  const size_t i0 = dirx_index(g, 0, y, z), i1 = dirx_index(g, g.nx - 1, y, z), i2 = dirx_index(g, g.nx - 2, y, z);
  f[i0] = f_start[i0];
  const double fd = f[i1], fd1 = f[i2];
  f[i1] = __dadd_rn(__dadd_rn(fd, -__dmul_rn(c_end, __dadd_rn(fd, -fd1))), flow_rate_diff);
}

struct RedGeom {
  int dir;
  int n_pad;        // padded line length
  int n_groups;     // groups of 32 lines
  int nblk;         // groups per plane index (nyb for X, nxb for Y and Z)
  int n_line;       // valid points along the line
  int n_lane_dim;   // valid extent of the dimension mapped on (lane, block)
  int n_plane_dim;  // valid extent of the remaining dimension
};

enum { RED_DOT, RED_ABS, RED_SUM };

// one warp per (j, g) row of 32 lanes; rows are dealt to warps in a fixed grid-stride order
template <int MODE>
__global__ void __launch_bounds__(kThreads)
reduce_stage1(const double* __restrict__ x, const double* __restrict__ y, const RedGeom q,
              double* __restrict__ part_sum, double* __restrict__ part_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_rows = (long long)q.n_pad * q.n_groups;
  const long long warps_total = (long long)gridDim.x * (kThreads / 32);
  double s = 0.0, m = 0.0;
  for (long long row = (long long)blockIdx.x * (kThreads / 32) + warp; row < n_rows; row += warps_total) {
    const int g = (int)(row / q.n_pad);
    const int j = (int)(row - (long long)g * q.n_pad);
    const int blk = g % q.nblk, pl = g / q.nblk;
    const bool ok = j < q.n_line && (blk * SZ + lane) < q.n_lane_dim && pl < q.n_plane_dim;
    if (ok) {
      const double xv = x[row * SZ + lane];
      if (MODE == RED_DOT) {
        s += xv * y[row * SZ + lane];
      } else if (MODE == RED_ABS) {
        const double av = fabs(xv);
        s += av;
        m = fmax(m, av);
      } else {
        s += xv;
      }
    }
  }
  __shared__ double sh_s[kThreads / 32], sh_m[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  if (lane == 0) { sh_s[warp] = s; sh_m[warp] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tm = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { ts += sh_s[w]; tm = fmax(tm, sh_m[w]); }
    part_sum[blockIdx.x] = ts;
    part_max[blockIdx.x] = tm;
  }
}

__global__ void __launch_bounds__(256)
reduce_stage2(const double* __restrict__ part_sum, const double* __restrict__ part_max, const int n,
              double* __restrict__ out) {
  __shared__ double sh_s[256], sh_m[256];
  double s = 0.0, m = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { s += part_sum[i]; m = fmax(m, part_max[i]); }
  sh_s[threadIdx.x] = s; sh_m[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh_s[threadIdx.x] += sh_s[threadIdx.x + o];
      sh_m[threadIdx.x] = fmax(sh_m[threadIdx.x], sh_m[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh_s[0]; out[1] = sh_m[0]; }
}

// slice_max_sum (omp/backend.f90:816-881): signed maximum and signed sum over the plane j = i_slice of a directional
// field, rank-local. One warp per group row; same two-stage deterministic tree as the other reductions.
__global__ void __launch_bounds__(kThreads)
slice_stage1(const double* __restrict__ x, const RedGeom q, const int j, double* __restrict__ part_sum,
             double* __restrict__ part_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_total = gridDim.x * (kThreads / 32);
  double s = 0.0, m = -1.7976931348623157e308;  // -huge(1._dp)
  for (int g = blockIdx.x * (kThreads / 32) + warp; g < q.n_groups; g += warps_total) {
    const int blk = g % q.nblk, pl = g / q.nblk;
    if ((blk * SZ + lane) < q.n_lane_dim && pl < q.n_plane_dim) {
      const double v = x[((size_t)g * q.n_pad + j) * SZ + lane];
      s += v;
      m = fmax(m, v);
    }
  }
  __shared__ double sh_s[kThreads / 32], sh_m[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  if (lane == 0) { sh_s[warp] = s; sh_m[warp] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tm = -1.7976931348623157e308;
    for (int w = 0; w < kThreads / 32; ++w) { ts += sh_s[w]; tm = fmax(tm, sh_m[w]); }
    part_sum[blockIdx.x] = ts;
    part_max[blockIdx.x] = tm;
  }
}
__global__ void __launch_bounds__(256)
slice_stage2(const double* __restrict__ part_sum, const double* __restrict__ part_max, const int n,
             double* __restrict__ out) {
  __shared__ double sh_s[256], sh_m[256];
  double s = 0.0, m = -1.7976931348623157e308;
  for (int i = threadIdx.x; i < n; i += 256) { s += part_sum[i]; m = fmax(m, part_max[i]); }
  sh_s[threadIdx.x] = s; sh_m[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh_s[threadIdx.x] += sh_s[threadIdx.x + o];
      sh_m[threadIdx.x] = fmax(sh_m[threadIdx.x], sh_m[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh_s[0]; out[1] = sh_m[0]; }
}

// compute_vorticity / compute_qcriterion (omp/backend.f90:616-649): pointwise over the whole padded block, the
// expressions in the reference's order with separately rounded operations
struct Grad9 { const double* g[9]; };  // dudx dudy dudz dvdx dvdy dvdz dwdx dwdy dwdz
template <bool QCRIT>
__global__ void __launch_bounds__(kThreads)
derive_kernel(double* __restrict__ out, const __grid_constant__ Grad9 q, const long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double dudx = q.g[0][i], dudy = q.g[1][i], dudz = q.g[2][i], dvdx = q.g[3][i], dvdy = q.g[4][i],
                 dvdz = q.g[5][i], dwdx = q.g[6][i], dwdy = q.g[7][i], dwdz = q.g[8][i];
    double r;
    if (QCRIT) {
      const double tr = __dadd_rn(__dadd_rn(__dmul_rn(dudx, dudx), __dmul_rn(dvdy, dvdy)), __dmul_rn(dwdz, dwdz));
      r = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(-0.5, tr), -__dmul_rn(dudy, dvdx)), -__dmul_rn(dudz, dwdx)),
                    -__dmul_rn(dvdz, dwdy));
    } else {
      const double a = __dadd_rn(dwdy, -dvdz), b = __dadd_rn(dudz, -dwdx), c = __dadd_rn(dvdx, -dudy);
      r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    }
    out[i] = r;
  }
}

int red_geom(const x3d2c_ctx* ctx, int dir, int data_loc, RedGeom* q) {
  int dims[3];
  int rc = x3d2c::get_dims_dataloc(ctx, data_loc, dims, false);
  if (rc) return rc;
  q->dir = dir;
  q->n_pad = ctx->n_pad(dir);
  q->n_groups = ctx->n_groups[dir];
  if (dir == X3D2C_DIR_X) { q->nblk = ctx->ny_pad / SZ; q->n_line = dims[0]; q->n_lane_dim = dims[1]; q->n_plane_dim = dims[2]; }
  else if (dir == X3D2C_DIR_Y) { q->nblk = ctx->nx_pad / SZ; q->n_line = dims[1]; q->n_lane_dim = dims[0]; q->n_plane_dim = dims[2]; }
  else if (dir == X3D2C_DIR_Z) { q->nblk = ctx->nx_pad / SZ; q->n_line = dims[2]; q->n_lane_dim = dims[0]; q->n_plane_dim = dims[1]; }
  else { x3d2c::set_error("reductions support DIR_X/Y/Z fields only"); return X3D2C_EINVAL; }
  return X3D2C_OK;
}

template <int MODE>
int run_reduce(x3d2c_ctx* ctx, const double* x, const double* y, const RedGeom& q) {
  double* ps = ctx->red;
  double* pm = ctx->red + ctx->red_blocks;
  double* out = ctx->red + 2 * ctx->red_blocks;
  reduce_stage1<MODE><<<ctx->red_blocks, kThreads, 0, ctx->stream>>>(x, y, q, ps, pm);
  X3D2C_CHECK_LAUNCH(ctx);
  reduce_stage2<<<1, 256, 0, ctx->stream>>>(ps, pm, ctx->red_blocks, out);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int fetch2(x3d2c_ctx* ctx, double* s, double* m) {
  double* out = ctx->red + 2 * ctx->red_blocks;
  X3D2C_CHECK_CUDA(cudaMemcpyAsync(ctx->red_host, out, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  X3D2C_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (s) *s = ctx->red_host[0];
  if (m) *m = ctx->red_host[1];
  return X3D2C_OK;
}

template <int OP>
int run_stream(x3d2c_ctx* ctx, double* y, const double* x, double a, double b) {
  stream_kernel<OP><<<kBlocks, kThreads, 0, ctx->stream>>>(y, x, a, b, ctx->ngrid / 2);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

}  // namespace

using namespace x3d2c;

extern "C" {

int x3d2c_field_fill(x3d2c_ctx* ctx, double* dev, double c) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dev, "x3d2c_field_fill: null argument");
  return run_stream<OP_FILL>(ctx, dev, nullptr, c, 0.0);
}
int x3d2c_veccopy(x3d2c_ctx* ctx, double* dst, const double* src) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dst && src, "x3d2c_veccopy: null argument");
  return run_stream<OP_COPY>(ctx, dst, src, 0.0, 0.0);
}
int x3d2c_vecadd(x3d2c_ctx* ctx, double a, const double* x, double b, double* y) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && x && y, "x3d2c_vecadd: null argument");
  if (ctx->strict) {
    axpby_strict_kernel<<<kBlocks, kThreads, 0, ctx->stream>>>(y, x, a, b, ctx->ngrid);
    X3D2C_CHECK_LAUNCH(ctx);
    return X3D2C_OK;
  }
  return run_stream<OP_AXPBY>(ctx, y, x, a, b);
}
int x3d2c_veclincomb(x3d2c_ctx* ctx, double* out, const double* base, int n, const double* coef,
                     const double* const* x) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out && base && n >= 1 && n <= 4 && coef && x, "x3d2c_veclincomb: bad argument");
  LinComb q{};
  q.base = base;
  q.n = n;
  for (int k = 0; k < n; ++k) {
    X3D2C_REQUIRE(x[k] && x[k] != out, "x3d2c_veclincomb: a term is null or aliases out");
    q.x[k] = x[k];
    q.c[k] = coef[k];
  }
  if (ctx->strict)
    lincomb_kernel<true><<<kBlocks, kThreads, 0, ctx->stream>>>(out, q, ctx->ngrid / 2);
  else
    lincomb_kernel<false><<<kBlocks, kThreads, 0, ctx->stream>>>(out, q, ctx->ngrid / 2);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}
int x3d2c_vecmult(x3d2c_ctx* ctx, double* y, const double* x) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && x && y, "x3d2c_vecmult: null argument");
  return run_stream<OP_MULT>(ctx, y, x, 0.0, 0.0);
}
int x3d2c_field_scale(x3d2c_ctx* ctx, double* f, double a) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f, "x3d2c_field_scale: null argument");
  return run_stream<OP_SCALE>(ctx, f, nullptr, a, 0.0);
}
int x3d2c_field_shift(x3d2c_ctx* ctx, double* f, double a) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f, "x3d2c_field_shift: null argument");
  return run_stream<OP_SHIFT>(ctx, f, nullptr, a, 0.0);
}

static int face_geom(const x3d2c_ctx* ctx, int data_loc, FaceGeom* g) {
  int dims[3];
  int rc = x3d2c::get_dims_dataloc(ctx, data_loc, dims, false);
  if (rc) return rc;
  g->nx = dims[0]; g->ny = dims[1]; g->nz = dims[2];
  g->nx_pad = ctx->nx_pad;
  g->nyb = ctx->ny_pad / SZ;
  return X3D2C_OK;
}

int x3d2c_field_set_face(x3d2c_ctx* ctx, double* f, int data_loc, double c_start, double c_end, int face) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f, "x3d2c_field_set_face: null argument");
  X3D2C_REQUIRE(face == X3D2C_X_FACE || face == X3D2C_Y_FACE || face == X3D2C_Z_FACE, "face is undefined.");
  X3D2C_REQUIRE(face != X3D2C_X_FACE, "Setting X_FACE is not yet supported.");  // omp/backend.f90:930-931
  X3D2C_REQUIRE(face != X3D2C_Z_FACE, "Setting Z_FACE is not yet supported.");  // :946-947
  FaceGeom g;
  int rc = face_geom(ctx, data_loc, &g);
  if (rc) return rc;
  set_yface_kernel<false><<<dim3((g.nx + 255) / 256, g.nz), 256, 0, ctx->stream>>>(f, nullptr, c_start, c_end, g);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_field_set_face_from_field(x3d2c_ctx* ctx, double* f, const double* f_start, int data_loc, double c_end,
                                    int face, double flow_rate_diff) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f && f_start, "x3d2c_field_set_face_from_field: null argument");
  X3D2C_REQUIRE(face == X3D2C_X_FACE || face == X3D2C_Y_FACE,
                "field_set_face_from_field: only X_FACE and Y_FACE supported.");  // omp/backend.f90:1017-1018
  FaceGeom g;
  int rc = face_geom(ctx, data_loc, &g);
  if (rc) return rc;
  if (face == X3D2C_Y_FACE)
    set_yface_kernel<true><<<dim3((g.nx + 255) / 256, g.nz), 256, 0, ctx->stream>>>(f, f_start, 0.0, 0.0, g);
  else
    set_xface_from_field_kernel<<<dim3((g.ny + 255) / 256, g.nz), 256, 0, ctx->stream>>>(f, f_start, c_end,
                                                                                          flow_rate_diff, g);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_scalar_product(x3d2c_ctx* ctx, int dir, int data_loc, const double* x, const double* y, double* s) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && x && y && s, "x3d2c_scalar_product: null argument");
  RedGeom q;
  int rc = red_geom(ctx, dir, data_loc, &q);
  if (rc) return rc;
  rc = run_reduce<RED_DOT>(ctx, x, y, q);
  if (rc) return rc;
  rc = allreduce(ctx, ctx->red + 2 * ctx->red_blocks, 1, 0);  // MPI_Allreduce SUM (omp/backend.f90:708)
  if (rc) return rc;
  return fetch2(ctx, s, nullptr);
}

int x3d2c_field_max_mean(x3d2c_ctx* ctx, int dir, int data_loc, const double* f, double* max_val, double* mean_val) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f && max_val && mean_val, "x3d2c_field_max_mean: null argument");
  RedGeom q;
  int rc = red_geom(ctx, dir, data_loc, &q);
  if (rc) return rc;
  rc = run_reduce<RED_ABS>(ctx, f, nullptr, q);
  if (rc) return rc;
  double* out = ctx->red + 2 * ctx->red_blocks;
  rc = allreduce(ctx, out, 1, 0);
  if (rc) return rc;
  rc = allreduce(ctx, out + 1, 1, 1);
  if (rc) return rc;
  double s = 0, m = 0;
  rc = fetch2(ctx, &s, &m);
  if (rc) return rc;
  int g[3];
  rc = get_dims_dataloc(ctx, data_loc, g, true);
  if (rc) return rc;
  *max_val = m;
  *mean_val = s / ((double)g[0] * g[1] * g[2]);  // omp/backend.f90:802
  return X3D2C_OK;
}

int x3d2c_slice_max_sum(x3d2c_ctx* ctx, int dir, int data_loc, const double* f, int i_slice, double* max_val,
                        double* sum_val) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f && max_val && sum_val, "x3d2c_slice_max_sum: null argument");
  X3D2C_REQUIRE(dir != X3D2C_DIR_C, "slice_max_sum does not support DIR_C fields!");
  RedGeom q;
  int rc = red_geom(ctx, dir, data_loc, &q);
  if (rc) return rc;
  X3D2C_REQUIRE(i_slice >= 1 && i_slice <= q.n_line, "slice_max_sum: i_slice out of range");
  double* ps = ctx->red;
  double* pm = ctx->red + ctx->red_blocks;
  double* out = ctx->red + 2 * ctx->red_blocks;
  const int blocks = std::min(ctx->red_blocks, (q.n_groups + kThreads / 32 - 1) / (kThreads / 32));
  slice_stage1<<<blocks, kThreads, 0, ctx->stream>>>(f, q, i_slice - 1, ps, pm);
  X3D2C_CHECK_LAUNCH(ctx);
  slice_stage2<<<1, 256, 0, ctx->stream>>>(ps, pm, blocks, out);
  X3D2C_CHECK_LAUNCH(ctx);
  return fetch2(ctx, sum_val, max_val);  // rank-local: the caller reduces across ranks (omp/backend.f90:877-879)
}

static int derive(x3d2c_ctx* ctx, bool qcrit, double* out, const double* const* grads) {
  X3D2C_REQUIRE(ctx && out && grads, "compute_vorticity / compute_qcriterion: null argument");
  Grad9 q;
  for (int k = 0; k < 9; ++k) {
    X3D2C_REQUIRE(grads[k], "compute_vorticity / compute_qcriterion: null gradient field");
    q.g[k] = grads[k];
  }
  if (qcrit) derive_kernel<true><<<kBlocks, kThreads, 0, ctx->stream>>>(out, q, ctx->ngrid);
  else derive_kernel<false><<<kBlocks, kThreads, 0, ctx->stream>>>(out, q, ctx->ngrid);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}
int x3d2c_compute_vorticity(x3d2c_ctx* ctx, double* field_out, const double* dudx, const double* dudy,
                            const double* dudz, const double* dvdx, const double* dvdy, const double* dvdz,
                            const double* dwdx, const double* dwdy, const double* dwdz) {
  X3D2C_ENTER(ctx);
  const double* g[9] = {dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz};
  return derive(ctx, false, field_out, g);
}
int x3d2c_compute_qcriterion(x3d2c_ctx* ctx, double* field_out, const double* dudx, const double* dudy,
                             const double* dudz, const double* dvdx, const double* dvdy, const double* dvdz,
                             const double* dwdx, const double* dwdy, const double* dwdz) {
  X3D2C_ENTER(ctx);
  const double* g[9] = {dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz};
  return derive(ctx, true, field_out, g);
}

int x3d2c_field_volume_integral(x3d2c_ctx* ctx, int data_loc, const double* f, double* s) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && f && s, "x3d2c_field_volume_integral: null argument");
  RedGeom q;
  int rc = red_geom(ctx, X3D2C_DIR_X, data_loc, &q);
  if (rc) return rc;
  rc = run_reduce<RED_SUM>(ctx, f, nullptr, q);
  if (rc) return rc;
  rc = allreduce(ctx, ctx->red + 2 * ctx->red_blocks, 1, 0);
  if (rc) return rc;
  return fetch2(ctx, s, nullptr);
}

}  // extern "C"
