// Fast path of the DistD2-TDS operators for periodic, single-rank directions ("m3" kernels).
//
// What is computed is what the reference computes (compact-scheme solve A x = r with the periodic
// tridiagonal A = tridiag(alpha, 1, alpha), r = 9-point stencil of the input; omp/kernels/distributed.f90:11-337),
// but organised for a B200 SM instead of a CPU core:
//
//  * one CTA owns whole lines: a tile of L lanes x n points per field is staged in shared memory with
//    cp.async (LDGSTS, 16-byte chunks), double buffered across tiles of a persistent CTA, so HBM sees exactly
//    one read of every input and one write of every output (48 B/pt for transeq, 16 B/pt for tds_solve);
//  * the line is cut into 16-point segments, one thread per (lane, segment); each thread keeps its three
//    recurrences (du, d(u conv), d2u) in registers: stencil -> local forward sweep -> local backward sweep;
//  * the sweeps use the converged (Toeplitz) factors fw, bw, alpha of the tdsops tables. For a periodic line
//    A = L U holds exactly with these constant factors, so no substitution phase (dist_sa/dist_sc) is needed;
//    segments are coupled through carries that decay like (alpha fw)^16 per segment, exchanged once per
//    component through shared memory and summed over D <= 3 neighbouring segments (wrapping periodically).
//    Truncation is below 1e-18 relative; differences to the sequential reference order are rounding only
//    (measured 2-6e-16 relative, tests/test_gpu_fast_path.py).
//  * -1/2 and nu of the transeq combination are folded into the stencil coefficients, FMA everywhere.
// Shapes that do not qualify (non-periodic operators, stretched meshes, multi-rank directions, strict mode)
// use the reference-order kernels of tds_m1.cu.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int S = 16;     // points per segment
constexpr int DMAX = 3;   // neighbouring segments that contribute to a carry

struct M3Op {
  double cfw[9];          // scale * fw * coeffs
  double a, cb;           // forward / backward propagators: a = -fw*alpha, cb = -bw
  double zw[DMAX], yw[DMAX];
  double om[2 * DMAX - 1];  // index m + DMAX - 1, m = d - d'
  double W[S], Cp[S];
  unsigned mask;
};

__device__ __forceinline__ int prow(int r) { return r + (r >> 4); }  // one pad row per segment: no bank conflicts

template <unsigned M>
__device__ __forceinline__ double sten(const double (&c)[9], const double (&w)[9]) {
  double t = 0.0;
  bool first = true;
#pragma unroll
  for (int k = 0; k < 9; ++k)
    if (M & (1u << k)) {
      t = first ? c[k] * w[k] : fma(c[k], w[k], t);
      first = false;
    }
  return t;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct TileGeom {
  int n, n_pad, L, lshift, nseg, tiles, tiles_per_group, field_doubles;
};

// global (32 lanes, n_pad rows, G groups) -> smem tile [prow(j)][L]
__device__ __forceinline__ void tile_load(double* sm, const double* __restrict__ g, const TileGeom& q, int tile) {
  const int grp = tile / q.tiles_per_group, l0 = (tile - grp * q.tiles_per_group) * q.L;
  const int cpr = q.L >> 1, cshift = q.lshift - 1, total = q.n * cpr;
  const double* base = g + (size_t)grp * q.n_pad * SZ + l0;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int j = idx >> cshift, c = idx & (cpr - 1);
    cp_async16(sm + prow(j) * q.L + 2 * c, base + (size_t)j * SZ + 2 * c);
  }
}
__device__ __forceinline__ void tile_store(double* __restrict__ g, const double* sm, const TileGeom& q, int tile) {
  const int grp = tile / q.tiles_per_group, l0 = (tile - grp * q.tiles_per_group) * q.L;
  const int cpr = q.L >> 1, cshift = q.lshift - 1, total = q.n * cpr;
  double* base = g + (size_t)grp * q.n_pad * SZ + l0;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int j = idx >> cshift, c = idx & (cpr - 1);
    const double2 v = *reinterpret_cast<const double2*>(sm + prow(j) * q.L + 2 * c);
    *reinterpret_cast<double2*>(base + (size_t)j * SZ + 2 * c) = v;
  }
}

// carries of one recurrence: zin (from the left), yin (from the right)
__device__ __forceinline__ void carries(const double* ze, const double* ys, const M3Op& o, int q, int nseg, int L,
                                        int l, double& zin, double& yin) {
  double zv[2 * DMAX];  // ze(q - DMAX .. q + DMAX - 1)
#pragma unroll
  for (int t = 0; t < 2 * DMAX; ++t) {
    int s = q - DMAX + t;
    if (s < 0) s += nseg;
    if (s >= nseg) s -= nseg;
    zv[t] = ze[s * L + l];
  }
  zin = 0.0;
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) zin = fma(o.zw[d - 1], zv[DMAX - d], zin);
  yin = 0.0;
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) {
    int s = q + d;
    if (s >= nseg) s -= nseg;
    yin = fma(o.yw[d - 1], ys[s * L + l], yin);
  }
#pragma unroll
  for (int m = -(DMAX - 1); m <= DMAX - 1; ++m) yin = fma(o.om[m + DMAX - 1], zv[DMAX + m], yin);
}

// ---------------------------------------------------------------------------------------------- transeq
struct TranseqParams {
  const double* in[3];  // in[0] is the line-aligned velocity (conv)
  double* out[3];
  TileGeom g;
  M3Op o_du, o_dud, o_d2u;  // scaled by -1/2, -1/2, nu
};

// one velocity component of one tile: F (in/out, in place), Cv = conv tile
template <unsigned M1, unsigned M2>
__device__ __forceinline__ void transeq_component(double* F, const double* Cv, double* carr, const TranseqParams& p,
                                                  const int q, const int l) {
  const int L = p.g.L, n = p.g.n, nseg = p.g.nseg;
  const int j0 = q * S;
  double z1[S], z2[S], z3[S];
  {
    double wf[9], wp[9];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      int r = j0 - 4 + t;
      if (r < 0) r += n;
      const int o = prow(r) * L + l;
      wf[t] = F[o];
      wp[t] = wf[t] * Cv[o];
    }
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      int r = j0 + 4 + k;
      if (r >= n) r -= n;
      const int o = prow(r) * L + l;
      wf[8] = F[o];
      wp[8] = wf[8] * Cv[o];
      p1 = fma(p.o_du.a, p1, sten<M1>(p.o_du.cfw, wf));
      p2 = fma(p.o_dud.a, p2, sten<M1>(p.o_dud.cfw, wp));
      p3 = fma(p.o_d2u.a, p3, sten<M2>(p.o_d2u.cfw, wf));
      z1[k] = p1; z2[k] = p2; z3[k] = p3;
#pragma unroll
      for (int t = 0; t < 8; ++t) { wf[t] = wf[t + 1]; wp[t] = wp[t + 1]; }
    }
  }
  double* ze = carr;
  double* ys = carr + 3 * nseg * L;
  ze[(0 * nseg + q) * L + l] = z1[S - 1];
  ze[(1 * nseg + q) * L + l] = z2[S - 1];
  ze[(2 * nseg + q) * L + l] = z3[S - 1];
  {
    double y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
    for (int k = S - 1; k >= 0; --k) {
      y1 = fma(p.o_du.cb, y1, z1[k]);
      y2 = fma(p.o_dud.cb, y2, z2[k]);
      y3 = fma(p.o_d2u.cb, y3, z3[k]);
      z1[k] = y1; z2[k] = y2; z3[k] = y3;
    }
  }
  ys[(0 * nseg + q) * L + l] = z1[0];
  ys[(1 * nseg + q) * L + l] = z2[0];
  ys[(2 * nseg + q) * L + l] = z3[0];
  __syncthreads();
  double zi1, yi1, zi2, yi2, zi3, yi3;
  carries(ze + 0 * nseg * L, ys + 0 * nseg * L, p.o_du, q, nseg, L, l, zi1, yi1);
  carries(ze + 1 * nseg * L, ys + 1 * nseg * L, p.o_dud, q, nseg, L, l, zi2, yi2);
  carries(ze + 2 * nseg * L, ys + 2 * nseg * L, p.o_d2u, q, nseg, L, l, zi3, yi3);
  const int ob = (j0 + q) * L + l;  // prow(j0 + k) = j0 + k + q
#pragma unroll
  for (int k = 0; k < S; ++k) {
    const double du = fma(p.o_du.Cp[k], yi1, fma(p.o_du.W[k], zi1, z1[k]));     // -1/2 du
    const double dud = fma(p.o_dud.Cp[k], yi2, fma(p.o_dud.W[k], zi2, z2[k]));  // -1/2 d(u conv)
    const double d2u = fma(p.o_d2u.Cp[k], yi3, fma(p.o_d2u.W[k], zi3, z3[k]));  // nu d2u
    const double cv = Cv[ob + k * L];
    F[ob + k * L] = fma(cv, du, dud + d2u);
  }
  __syncthreads();  // carries may be overwritten by the next component; F is complete
}

template <unsigned M1, unsigned M2>
__global__ void __launch_bounds__(256, 1) transeq_m3_kernel(const __grid_constant__ TranseqParams p) {
  extern __shared__ __align__(16) double smem[];
  const TileGeom& g = p.g;
  double* carr = smem + 6 * g.field_doubles;
  const int l = threadIdx.x & (g.L - 1), q = threadIdx.x >> g.lshift;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
    if (it == 0) {
#pragma unroll
      for (int f = 0; f < 3; ++f) tile_load(smem + f * g.field_doubles, p.in[f], g, tile);
      cp_async_commit();
      const int nx = tile + gridDim.x;
      if (nx < g.tiles) {
#pragma unroll
        for (int f = 0; f < 3; ++f) tile_load(smem + (3 + f) * g.field_doubles, p.in[f], g, nx);
      }
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncthreads();
    double* b = smem + (it & 1) * 3 * g.field_doubles;
    // components 1 and 2 first: they read the aligned velocity (field 0) as conv; field 0 is overwritten last
    transeq_component<M1, M2>(b + 1 * g.field_doubles, b, carr, p, q, l);
    transeq_component<M1, M2>(b + 2 * g.field_doubles, b, carr, p, q, l);
    transeq_component<M1, M2>(b, b, carr, p, q, l);
#pragma unroll
    for (int f = 0; f < 3; ++f) tile_store(p.out[f], b + f * g.field_doubles, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) {
#pragma unroll
      for (int f = 0; f < 3; ++f) tile_load(b + f * g.field_doubles, p.in[f], g, nn);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- tds_solve
struct TdsParams {
  const double* in;
  double* out;
  TileGeom g;
  M3Op o;
};

template <unsigned M>
__global__ void __launch_bounds__(256, 3) tds_m3_kernel(const __grid_constant__ TdsParams p) {
  extern __shared__ __align__(16) double smem[];
  const TileGeom& g = p.g;
  double* carr = smem + 2 * g.field_doubles;
  const int L = g.L, n = g.n, nseg = g.nseg;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x >> g.lshift;
  const int j0 = q * S;
  double* ze = carr;
  double* ys = carr + nseg * L;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
    if (it == 0) {
      tile_load(smem, p.in, g, tile);
      cp_async_commit();
      const int nx = tile + gridDim.x;
      if (nx < g.tiles) tile_load(smem + g.field_doubles, p.in, g, nx);
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncthreads();
    double* F = smem + (it & 1) * g.field_doubles;
    double z[S];
    {
      double wf[9];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        int r = j0 - 4 + t;
        if (r < 0) r += n;
        wf[t] = F[prow(r) * L + l];
      }
      double pz = 0.0;
#pragma unroll
      for (int k = 0; k < S; ++k) {
        int r = j0 + 4 + k;
        if (r >= n) r -= n;
        wf[8] = F[prow(r) * L + l];
        pz = fma(p.o.a, pz, sten<M>(p.o.cfw, wf));
        z[k] = pz;
#pragma unroll
        for (int t = 0; t < 8; ++t) wf[t] = wf[t + 1];
      }
    }
    ze[q * L + l] = z[S - 1];
    {
      double y = 0.0;
#pragma unroll
      for (int k = S - 1; k >= 0; --k) {
        y = fma(p.o.cb, y, z[k]);
        z[k] = y;
      }
    }
    ys[q * L + l] = z[0];
    __syncthreads();
    double zin, yin;
    carries(ze, ys, p.o, q, nseg, L, l, zin, yin);
    const int ob = (j0 + q) * L + l;
#pragma unroll
    for (int k = 0; k < S; ++k) F[ob + k * L] = fma(p.o.Cp[k], yin, fma(p.o.W[k], zin, z[k]));
    __syncthreads();
    tile_store(p.out, F, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) tile_load(F, p.in, g, nn);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
bool same_tables(const x3d2c_tdsops* a, const x3d2c_tdsops* b) {
  if (a->n_tds != b->n_tds || a->n_rhs != b->n_rhs) return false;
  if (std::memcmp(a->dev.coeffs, b->dev.coeffs, sizeof a->dev.coeffs)) return false;
  const int m = a->n_tds / 2;
  return a->h_fw[m] == b->h_fw[m] && a->h_bw[m] == b->h_bw[m] && a->h_af[m] == b->h_af[m];
}

// builds the constant set of one operator; false when the operator does not qualify for the fast path
bool make_m3op(const x3d2c_tdsops* t, double scale, M3Op* o) {
  const int n = t->n_tds;
  if (!t->periodic || t->n_rhs != n || n < 4 * S || n % S) return false;
  if (t->has_stretch || t->has_stretch_correct) return false;
  const int m = n / 2;
  const double fw = t->h_fw[m], bw = t->h_bw[m], al = t->h_af[m];
  // the factors must have converged to their Toeplitz limit over the whole central region
  for (int j = 40; j < n - 40; ++j) {
    if (std::fabs(t->h_fw[j] - fw) > 4e-16 * std::fabs(fw) || std::fabs(t->h_bw[j] - bw) > 4e-16 * std::fabs(bw) ||
        t->h_af[j] != al)
      return false;
  }
  for (int k = 0; k < 9; ++k) o->cfw[k] = scale * fw * t->dev.coeffs[k];
  o->mask = t->tap_mask;
  o->a = -fw * al;
  o->cb = -bw;
  if (std::pow(std::fabs(o->a), S * DMAX) > 1e-18 || std::pow(std::fabs(o->cb), S * DMAX) > 1e-18) return false;
  for (int d = 0; d < DMAX; ++d) { o->zw[d] = std::pow(o->a, S * d); o->yw[d] = std::pow(o->cb, S * d); }
  double W[S + 1];
  W[S] = 0.0;
  for (int k = S - 1; k >= 0; --k) W[k] = std::pow(o->a, k + 1) + o->cb * W[k + 1];
  for (int k = 0; k < S; ++k) { o->W[k] = W[k]; o->Cp[k] = std::pow(o->cb, S - k); }
  for (int m2 = 0; m2 < 2 * DMAX - 1; ++m2) o->om[m2] = 0.0;
  for (int d = 1; d <= DMAX; ++d)
    for (int dp = 1; dp <= DMAX; ++dp) o->om[d - dp + DMAX - 1] += W[0] * o->yw[d - 1] * o->zw[dp - 1];
  return true;
}

bool make_geom(const x3d2c_ctx* ctx, int dir, int n, int max_threads, size_t fields_in_smem, size_t smem_limit,
               TileGeom* g, size_t* smem_bytes, int carr_recs) {
  const int n_pad = ctx->n_pad(dir);
  int L = 32;
  while (L > 2 && (L * (n / S) > max_threads)) L >>= 1;
  for (;; L >>= 1) {
    if (L < 2) return false;
    const int field = (n + n / S) * L;
    const size_t bytes = sizeof(double) * (2 * fields_in_smem * field + 2 * (size_t)carr_recs * (n / S) * L);
    if (bytes <= smem_limit) {
      g->field_doubles = field;
      *smem_bytes = bytes;
      break;
    }
  }
  g->n = n; g->n_pad = n_pad; g->L = L;
  g->lshift = 0;
  while ((1 << g->lshift) < L) g->lshift++;
  g->nseg = n / S;
  g->tiles_per_group = SZ / L;
  g->tiles = ctx->n_groups[dir] * g->tiles_per_group;
  return L * g->nseg >= 32;
}

int num_sms(const x3d2c_ctx* ctx) {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  return sms > 0 ? sms : 148;
}

}  // namespace

namespace x3d2c {

int tds_solve_m3(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops) {
  TdsParams p;
  if (!make_m3op(ops, 1.0, &p.o)) return X3D2C_EUNSUPPORTED;
  size_t smem = 0;
  if (!make_geom(ctx, dir, ops->n_tds, 256, 1, 72 * 1024, &p.g, &smem, 1)) return X3D2C_EUNSUPPORTED;
  p.in = u;
  p.out = du;
  const int threads = p.g.L * p.g.nseg;
  int grid = num_sms(ctx) * 3;
  if (grid > p.g.tiles) grid = p.g.tiles;
#define LAUNCH_TDS(MASK)                                                                                        \
  do {                                                                                                          \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      X3D2C_CHECK_CUDA(cudaFuncSetAttribute(tds_m3_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)); \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    tds_m3_kernel<MASK><<<grid, threads, smem, ctx->stream>>>(p);                                               \
  } while (0)
  switch (ops->tap_mask) {
    case 0x6Cu: LAUNCH_TDS(0x6Cu); break;  // first derivative
    case 0x7Cu: LAUNCH_TDS(0x7Cu); break;  // second derivative
    case 0x78u: LAUNCH_TDS(0x78u); break;  // staggered derivative / interpolation v2p
    case 0x3Cu: LAUNCH_TDS(0x3Cu); break;  // staggered derivative / interpolation p2v
    default: LAUNCH_TDS(0x1FFu); break;
  }
#undef LAUNCH_TDS
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int transeq_m3(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym) {
  // periodic operators do not distinguish the symmetric variants (src/tdsops.f90:277-396 only edits BC rows)
  if (!same_tables(der1st, der1st_sym) || !same_tables(der2nd, der2nd_sym)) return X3D2C_EUNSUPPORTED;
  TranseqParams p;
  if (!make_m3op(der1st, -0.5, &p.o_du) || !make_m3op(der1st, -0.5, &p.o_dud) || !make_m3op(der2nd, nu, &p.o_d2u))
    return X3D2C_EUNSUPPORTED;
  size_t smem = 0;
  if (!make_geom(ctx, dir, der1st->n_tds, 256, 3, 227 * 1024, &p.g, &smem, 3)) return X3D2C_EUNSUPPORTED;
  if (dir == X3D2C_DIR_X) { p.out[0] = du; p.out[1] = dv; p.out[2] = dw; p.in[0] = u; p.in[1] = v; p.in[2] = w; }
  else if (dir == X3D2C_DIR_Y) { p.out[0] = dv; p.out[1] = du; p.out[2] = dw; p.in[0] = v; p.in[1] = u; p.in[2] = w; }
  else { p.out[0] = dw; p.out[1] = du; p.out[2] = dv; p.in[0] = w; p.in[1] = u; p.in[2] = v; }
  const int threads = p.g.L * p.g.nseg;
  int grid = num_sms(ctx);
  if (grid > p.g.tiles) grid = p.g.tiles;
  const bool compact = der1st->tap_mask == 0x6Cu && der2nd->tap_mask == 0x7Cu;
  static bool attr_set = false;
  if (!attr_set) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m3_kernel<0x6Cu, 0x7Cu>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m3_kernel<0x1FFu, 0x1FFu>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (compact)
    transeq_m3_kernel<0x6Cu, 0x7Cu><<<grid, threads, smem, ctx->stream>>>(p);
  else
    transeq_m3_kernel<0x1FFu, 0x1FFu><<<grid, threads, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

}  // namespace x3d2c
