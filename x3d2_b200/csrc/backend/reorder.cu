// Pencil reorders and sum_{y,z}intox of the cuda_c backend.
// Replaces reorder_omp / sum_intox_omp (src/backend/omp/backend.f90:393-527; index maps
// src/ordering.f90:13-87) and the CUDA-Fortran tile kernels (src/backend/cuda/kernels/reorder.f90:9-314).
//
// Every layout of common.cuh is affine in the tile coordinates (x_l, xb, y_l, yb, z) with
// x = xb*32 + x_l, y = yb*32 + y_l, so ONE kernel serves all twelve RDR_* codes: a CTA moves one
// 32x32 (x_l, y_l) tile, reading with the source's unit-stride index across lanes and writing with the
// destination's unit-stride index across lanes (through a padded shared-memory tile when they differ).
#include "common.cuh"

namespace {

struct Lay {
  long long sxl, sxb, syl, syb, sz;
  int fast;  // 0: x_l has unit stride, 1: y_l has unit stride
};

Lay layout_of(const x3d2c_ctx* ctx, int dir) {
  const long long nxp = ctx->nx_pad, nyp = ctx->ny_pad, nz = ctx->nz_pad;
  const long long nxb = nxp / SZ, nyb = nyp / SZ;
  Lay l{};
  switch (dir) {
    case X3D2C_DIR_X: l = {SZ, (long long)SZ * SZ, 1, SZ * nxp, SZ * nxp * nyb, 1}; break;
    case X3D2C_DIR_Y: l = {1, SZ * nyp, SZ, (long long)SZ * SZ, SZ * nyp * nxb, 0}; break;
    case X3D2C_DIR_Z: l = {1, SZ * nz, SZ * nz * nxb, SZ * SZ * nz * nxb, SZ, 0}; break;
    default: l = {1, SZ, nxp, SZ * nxp, nxp * nyp, 0}; break;
  }
  return l;
}

template <bool ACC>
__global__ void __launch_bounds__(256)
reorder_tile_kernel(double* __restrict__ dst, const double* __restrict__ src, const Lay ls, const Lay ld) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long bs = blockIdx.x * ls.sxb + blockIdx.y * ls.syb + blockIdx.z * ls.sz;
  const long long bd = blockIdx.x * ld.sxb + blockIdx.y * ld.syb + blockIdx.z * ld.sz;
  const long long s_fast = ls.fast ? ls.syl : ls.sxl, s_other = ls.fast ? ls.sxl : ls.syl;
  const long long d_fast = ld.fast ? ld.syl : ld.sxl, d_other = ld.fast ? ld.sxl : ld.syl;
  double v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) v[r] = src[bs + tx * s_fast + (ty + 8 * r) * s_other];
  if (ls.fast == ld.fast) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double* p = dst + bd + tx * d_fast + (ty + 8 * r) * d_other;
      *p = ACC ? *p + v[r] : v[r];
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) tile[ty + 8 * r][tx] = v[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double* p = dst + bd + tx * d_fast + (ty + 8 * r) * d_other;
    const double t = tile[tx][ty + 8 * r];
    *p = ACC ? *p + t : t;
  }
}

// dst_y = reorder_x2y(src), dst_z = reorder_x2z(src): the source tile is read once (transeq_default reorders u, v, w
// into both pencil layouts, src/solver.f90:325-327,355-357). Both destinations have x_l fastest.
__global__ void __launch_bounds__(256)
reorder_x2yz_kernel(double* __restrict__ dst_y, double* __restrict__ dst_z, const double* __restrict__ src, const Lay ls,
                    const Lay ly, const Lay lz) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long bs = blockIdx.x * ls.sxb + blockIdx.y * ls.syb + blockIdx.z * ls.sz;
  const long long by = blockIdx.x * ly.sxb + blockIdx.y * ly.syb + blockIdx.z * ly.sz;
  const long long bz = blockIdx.x * lz.sxb + blockIdx.y * lz.syb + blockIdx.z * lz.sz;
#pragma unroll
  for (int r = 0; r < 4; ++r) tile[ty + 8 * r][tx] = src[bs + tx * ls.syl + (ty + 8 * r) * ls.sxl];  // [x_l][y_l]
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double t = tile[tx][ty + 8 * r];  // x_l = tx, y_l = ty + 8 r
    dst_y[by + tx * ly.sxl + (ty + 8 * r) * ly.syl] = t;
    dst_z[bz + tx * lz.sxl + (ty + 8 * r) * lz.syl] = t;
  }
}

// u (DIR_X) = (u + reorder(a)) + reorder(b): sum_yintox followed by sum_zintox in one pass over u
// (same order of additions as the two reference calls, src/solver.f90:340-372).
__global__ void __launch_bounds__(256)
sum2_tile_kernel(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b, const Lay la,
                 const Lay lb, const Lay ld) {
  __shared__ double ta[32][33], tb[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long ba = blockIdx.x * la.sxb + blockIdx.y * la.syb + blockIdx.z * la.sz;
  const long long bb = blockIdx.x * lb.sxb + blockIdx.y * lb.syb + blockIdx.z * lb.sz;
  const long long bd = blockIdx.x * ld.sxb + blockIdx.y * ld.syb + blockIdx.z * ld.sz;
  // both sources have x_l fastest (DIR_Y / DIR_Z), the destination y_l (DIR_X)
  double va[4], vb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    va[r] = a[ba + tx * la.sxl + (ty + 8 * r) * la.syl];
    vb[r] = b[bb + tx * lb.sxl + (ty + 8 * r) * lb.syl];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) { ta[ty + 8 * r][tx] = va[r]; tb[ty + 8 * r][tx] = vb[r]; }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double* p = dst + bd + tx * ld.syl + (ty + 8 * r) * ld.sxl;
    *p = (*p + ta[tx][ty + 8 * r]) + tb[tx][ty + 8 * r];
  }
}

// sum2_tile_kernel followed by a veclincomb whose last term is the finished sum, in the same pass:
//   s = (u + reorder(a)) + reorder(b);  [u = s];  out = base + sum_k c[k] x[k] + c_s s      (all of them DIR_X)
struct SumLin {
  const double* base;
  const double* x[3];
  double c[3], c_s;
  int n, store;
};
__global__ void __launch_bounds__(256)
sum2_lincomb_kernel(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b, double* out,
                    const Lay la, const Lay lb, const Lay ld, const __grid_constant__ SumLin q) {
  __shared__ double ta[32][33], tb[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long ba = blockIdx.x * la.sxb + blockIdx.y * la.syb + blockIdx.z * la.sz;
  const long long bb = blockIdx.x * lb.sxb + blockIdx.y * lb.syb + blockIdx.z * lb.sz;
  const long long bd = blockIdx.x * ld.sxb + blockIdx.y * ld.syb + blockIdx.z * ld.sz;
  double va[4], vb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    va[r] = a[ba + tx * la.sxl + (ty + 8 * r) * la.syl];
    vb[r] = b[bb + tx * lb.sxl + (ty + 8 * r) * lb.syl];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) { ta[ty + 8 * r][tx] = va[r]; tb[ty + 8 * r][tx] = vb[r]; }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long i = bd + tx * ld.syl + (ty + 8 * r) * ld.sxl;
    const double s = (dst[i] + ta[tx][ty + 8 * r]) + tb[tx][ty + 8 * r];
    if (q.store) dst[i] = s;
    double o = q.base[i];  // out may alias base
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (k < q.n) o = q.c[k] * q.x[k][i] + 1.0 * o;
    out[i] = q.c_s * s + 1.0 * o;
  }
}

}  // namespace

namespace x3d2c {
int launch_reorder(x3d2c_ctx* ctx, int dir_from, int dir_to, double* dst, const double* src, bool accumulate) {
  const Lay ls = layout_of(ctx, dir_from), ld = layout_of(ctx, dir_to);
  const dim3 grid(ctx->nx_pad / SZ, ctx->ny_pad / SZ, ctx->nz_pad), block(32, 8);
  if (accumulate)
    reorder_tile_kernel<true><<<grid, block, 0, ctx->stream>>>(dst, src, ls, ld);
  else
    reorder_tile_kernel<false><<<grid, block, 0, ctx->stream>>>(dst, src, ls, ld);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}
}  // namespace x3d2c

using namespace x3d2c;

extern "C" {

int x3d2c_reorder(x3d2c_ctx* ctx, int rdr, double* dst, const double* src) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dst && src, "x3d2c_reorder: null argument");
  X3D2C_REQUIRE(dst != src, "x3d2c_reorder: in-place reorder is not supported");
  const int from = rdr / 10, to = rdr % 10;  // src/common.f90:23-26,44-53
  X3D2C_REQUIRE(from >= 1 && from <= 4 && to >= 1 && to <= 4 && from != to, "x3d2c_reorder: unknown RDR code");
  X3D2C_REQUIRE(ctx->nz_pad <= 65535, "x3d2c_reorder: nz exceeds the grid limit");
  return launch_reorder(ctx, from, to, dst, src, false);
}

int x3d2c_reorder_x2yz(x3d2c_ctx* ctx, double* dst_y, double* dst_z, const double* src) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dst_y && dst_z && src, "x3d2c_reorder_x2yz: null argument");
  X3D2C_REQUIRE(dst_y != src && dst_z != src && dst_y != dst_z, "x3d2c_reorder_x2yz: fields must be distinct");
  X3D2C_REQUIRE(ctx->nz_pad <= 65535, "x3d2c_reorder_x2yz: nz exceeds the grid limit");
  const dim3 grid(ctx->nx_pad / SZ, ctx->ny_pad / SZ, ctx->nz_pad), block(32, 8);
  reorder_x2yz_kernel<<<grid, block, 0, ctx->stream>>>(dst_y, dst_z, src, layout_of(ctx, X3D2C_DIR_X),
                                                       layout_of(ctx, X3D2C_DIR_Y), layout_of(ctx, X3D2C_DIR_Z));
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_sum_yintox(x3d2c_ctx* ctx, double* u, const double* u_y) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && u && u_y, "x3d2c_sum_yintox: null argument");
  return launch_reorder(ctx, X3D2C_DIR_Y, X3D2C_DIR_X, u, u_y, true);
}

int x3d2c_sum_yzintox(x3d2c_ctx* ctx, double* u, const double* u_y, const double* u_z) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && u && u_y && u_z, "x3d2c_sum_yzintox: null argument");
  X3D2C_REQUIRE(ctx->nz_pad <= 65535, "x3d2c_sum_yzintox: nz exceeds the grid limit");
  const dim3 grid(ctx->nx_pad / SZ, ctx->ny_pad / SZ, ctx->nz_pad), block(32, 8);
  sum2_tile_kernel<<<grid, block, 0, ctx->stream>>>(u, u_y, u_z, layout_of(ctx, X3D2C_DIR_Y), layout_of(ctx, X3D2C_DIR_Z),
                                                    layout_of(ctx, X3D2C_DIR_X));
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_sum_yzintox_lincomb(x3d2c_ctx* ctx, double* u, const double* u_y, const double* u_z, int store_u, double* out,
                              const double* base, int n, const double* coef, const double* const* x, double c_u) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && u && u_y && u_z && out && base && n >= 0 && n <= 3 && (n == 0 || (coef && x)),
                "x3d2c_sum_yzintox_lincomb: bad argument");
  X3D2C_REQUIRE(out != u, "x3d2c_sum_yzintox_lincomb: out aliases u");
  for (int k = 0; k < n; ++k)
    X3D2C_REQUIRE(x[k] && x[k] != out && x[k] != u, "x3d2c_sum_yzintox_lincomb: a term is null or aliases out / u");
  if (ctx->strict) {  // the defining sequence, bit for bit
    int rc = x3d2c_sum_yzintox(ctx, u, u_y, u_z);
    if (rc) return rc;
    double c[4];
    const double* t[4];
    for (int k = 0; k < n; ++k) { c[k] = coef[k]; t[k] = x[k]; }
    c[n] = c_u;
    t[n] = u;
    return x3d2c_veclincomb(ctx, out, base, n + 1, c, t);
  }
  X3D2C_REQUIRE(ctx->nz_pad <= 65535, "x3d2c_sum_yzintox_lincomb: nz exceeds the grid limit");
  SumLin q{};
  q.base = base;
  q.n = n;
  q.c_s = c_u;
  q.store = store_u ? 1 : 0;
  for (int k = 0; k < n; ++k) { q.x[k] = x[k]; q.c[k] = coef[k]; }
  const dim3 grid(ctx->nx_pad / SZ, ctx->ny_pad / SZ, ctx->nz_pad), block(32, 8);
  sum2_lincomb_kernel<<<grid, block, 0, ctx->stream>>>(u, u_y, u_z, out, layout_of(ctx, X3D2C_DIR_Y),
                                                       layout_of(ctx, X3D2C_DIR_Z), layout_of(ctx, X3D2C_DIR_X), q);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_sum_zintox(x3d2c_ctx* ctx, double* u, const double* u_z) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && u && u_z, "x3d2c_sum_zintox: null argument");
  return launch_reorder(ctx, X3D2C_DIR_Z, X3D2C_DIR_X, u, u_z, true);
}

}  // extern "C"
