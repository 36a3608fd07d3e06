// Fast path of the DistD2-TDS operators for periodic, single-rank directions ("m3" kernels).
//
// What is computed is what the reference computes (compact-scheme solve A x = r with the periodic
// tridiagonal A = tridiag(alpha, 1, alpha), r = 9-point stencil of the input; omp/kernels/distributed.f90:11-337),
// but organised for a B200 SM instead of a CPU core:
//
//  * one CTA owns whole lines: a tile of L lanes x n points per field is staged in shared memory with
//    cp.async (LDGSTS, 16-byte chunks), double buffered across the tiles of a persistent CTA, so HBM sees
//    exactly one read of every input and one write of every output (48 B/pt for transeq, 16 B/pt for
//    tds_solve; ncu: 6.39 GB moved for 6.44 GB algorithmic at 512^3). Several CTAs share an SM so that the
//    copy phases of one overlap the FP64 phases of another;
//  * the line is cut into 16-point segments, one thread per (lane, segment); each thread keeps its three
//    recurrences (du, d(u conv), d2u) in registers: stencil -> local forward sweep -> local backward sweep;
//  * the sweeps use the converged (Toeplitz) factors fw, bw, alpha of the tdsops tables. For a periodic line
//    A = L U holds exactly with these constant factors, so no substitution phase (dist_sa/dist_sc) is needed;
//    segments are coupled through carries that decay like (alpha fw)^16 per segment, exchanged once per
//    component through shared memory and summed over D <= 3 neighbouring segments (wrapping periodically).
//    Truncation is below 1e-18 relative; differences to the sequential reference order are rounding only
//    (measured 2-7e-16 relative, tests/test_gpu_fast_path.py);
//  * -1/2 and nu of the transeq combination are folded into the stencil coefficients, FMA everywhere.
// Shapes that do not qualify (non-periodic operators, stretched meshes, multi-rank directions, strict mode)
// use the reference-order kernels of tds_m1.cu.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int S = 16;     // points per segment
constexpr int SP = S + 1; // rows per segment in shared memory (one pad row: conflict-free column access)
constexpr int DMAX = 3;   // neighbouring segments that contribute to a carry

struct M3Op {
  double cfw[9];            // scale * fw * coeffs
  double a, cb;             // forward / backward propagators: a = -fw*alpha, cb = -bw
  double zw[DMAX], yw[DMAX];
  double om[2 * DMAX - 1];  // index m + DMAX - 1, m = d - d'
  double W[S], Cp[S];
  unsigned mask;
};

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
  int n, n_pad, nseg, tiles, field_doubles;
};

// Tile copies. Global: (32 lanes, n_pad rows, G groups); shared: [segment][SP rows][L lanes].
// A thread owns chunk c (2 lanes) of rows j_t, j_t + R, j_t + 2R, ... with R = blockDim / (L/2) a multiple of 16,
// so both addresses advance by constants.
template <int L>
struct Copier {
  int c2, g_off, j, rows_per_pass;
  __device__ __forceinline__ Copier() {
    constexpr int cpr = L / 2;
    j = threadIdx.x / cpr;
    c2 = 2 * (threadIdx.x - j * cpr);
    rows_per_pass = blockDim.x / cpr;
    g_off = j * SZ + c2;
  }
  __device__ __forceinline__ const double* tile_base(const double* g, const TileGeom& q, int tile) const {
    constexpr int tpg = SZ / L;
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    return g + (size_t)grp * q.n_pad * SZ + l0 + g_off;
  }
  __device__ __forceinline__ void load(double* sm, const double* g, const TileGeom& q, int tile) const {
    const double* src = tile_base(g, q, tile);
    for (int r = j; r < q.n; r += rows_per_pass) {
      cp_async16(sm + (r + (r >> 4)) * L + c2, src);
      src += (size_t)rows_per_pass * SZ;
    }
  }
  __device__ __forceinline__ void store(double* g, const double* sm, const TileGeom& q, int tile) const {
    double* dst = const_cast<double*>(tile_base(g, q, tile));
    for (int r0 = j; r0 < q.n; r0 += 4 * rows_per_pass) {  // four chunks in flight per thread
      double2 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i * rows_per_pass;
        if (r < q.n) v[i] = *reinterpret_cast<const double2*>(sm + (r + (r >> 4)) * L + c2);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i * rows_per_pass;
        if (r < q.n) __stcs(reinterpret_cast<double2*>(dst + (size_t)i * rows_per_pass * SZ), v[i]);
      }
      dst += (size_t)4 * rows_per_pass * SZ;
    }
  }
};

// carries of one recurrence: zin (from the left), yin (from the right); ze/ys are [segment][L]
template <int L>
__device__ __forceinline__ void carries(const double* ze, const double* ys, const M3Op& o, int q, int nseg, int l,
                                        double& zin, double& yin) {
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

// window element t (row j0 - 4 + t, t = 0..23) of a [segment][SP][L] tile given the three segment bases
template <int L>
__device__ __forceinline__ int woff(int t, int bm, int b0, int bp) {
  return t < 4 ? bm + (12 + t) * L : (t < 20 ? b0 + (t - 4) * L : bp + (t - 20) * L);
}

// ---------------------------------------------------------------------------------------------- tds_solve
struct TdsParams {
  const double* in;
  double* out;
  TileGeom g;
  M3Op o;
};

template <int L, unsigned M>
__global__ void __launch_bounds__(256, 3) tds_m3_kernel(const __grid_constant__ TdsParams p) {
  extern __shared__ __align__(16) double smem[];
  const TileGeom& g = p.g;
  const int fd = g.field_doubles, nseg = g.nseg;
  double* carr = smem + 2 * fd;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  const int qm = q == 0 ? nseg - 1 : q - 1, qp = q == nseg - 1 ? 0 : q + 1;
  const int bm = qm * SP * L + l, b0 = q * SP * L + l, bp = qp * SP * L + l;
  double* ze = carr;
  double* ys = carr + nseg * L;
  const Copier<L> cp;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
    if (it == 0) {
      cp.load(smem, p.in, g, tile);
      cp_async_commit();
      const int nx = tile + gridDim.x;
      if (nx < g.tiles) cp.load(smem + fd, p.in, g, nx);
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncthreads();
    double* F = smem + (it & 1) * fd;
    double z[S];
    {
      double wf[9];
#pragma unroll
      for (int t = 0; t < 8; ++t) wf[t] = F[woff<L>(t, bm, b0, bp)];
      double pz = 0.0;
#pragma unroll
      for (int k = 0; k < S; ++k) {
        wf[8] = F[woff<L>(k + 8, bm, b0, bp)];
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
    carries<L>(ze, ys, p.o, q, nseg, l, zin, yin);
#pragma unroll
    for (int k = 0; k < S; ++k) F[b0 + k * L] = fma(p.o.Cp[k], yin, fma(p.o.W[k], zin, z[k]));
    __syncthreads();
    cp.store(p.out, F, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) cp.load(F, p.in, g, nn);
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

// tile width: the largest power of two L <= 32 with L * nseg <= max_threads
int pick_lanes(int n, int max_threads) {
  int L = 32;
  while (L >= 2 && L * (n / S) > max_threads) L >>= 1;
  if (L < 2) return 0;
  const int threads = L * (n / S);
  if (threads < 32 || threads % 32) return 0;
  return L;
}

void fill_geom(const x3d2c_ctx* ctx, int dir, int n, int L, TileGeom* g) {
  g->n = n;
  g->n_pad = ctx->n_pad(dir);
  g->nseg = n / S;
  g->tiles = ctx->n_groups[dir] * (SZ / L);
  g->field_doubles = g->nseg * SP * L;
}

int num_sms(const x3d2c_ctx* ctx) {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  return sms > 0 ? sms : 148;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  X3D2C_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return X3D2C_OK;
}

template <int L, unsigned M>
int launch_tds(x3d2c_ctx* ctx, const TdsParams& p, int threads, size_t smem) {
  static bool attr_set = false;
  if (!attr_set) {
    int rc = set_smem(tds_m3_kernel<L, M>, 75 * 1024);
    if (rc) return rc;
    attr_set = true;
  }
  int grid = num_sms(ctx) * 3;
  if (grid > p.g.tiles) grid = p.g.tiles;
  tds_m3_kernel<L, M><<<grid, threads, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L>
int dispatch_tds(x3d2c_ctx* ctx, const TdsParams& p, unsigned mask, int threads, size_t smem) {
  switch (mask) {
    case 0x6Cu: return launch_tds<L, 0x6Cu>(ctx, p, threads, smem);  // first derivative
    case 0x7Cu: return launch_tds<L, 0x7Cu>(ctx, p, threads, smem);  // second derivative
    case 0x78u: return launch_tds<L, 0x78u>(ctx, p, threads, smem);  // staggered derivative / interpolation v2p
    case 0x3Cu: return launch_tds<L, 0x3Cu>(ctx, p, threads, smem);  // staggered derivative / interpolation p2v
    default: return launch_tds<L, 0x1FFu>(ctx, p, threads, smem);
  }
}

}  // namespace

namespace x3d2c {

int tds_solve_m3(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops) {
  TdsParams p;
  if (!make_m3op(ops, 1.0, &p.o)) return X3D2C_EUNSUPPORTED;
  const int n = ops->n_tds, L = pick_lanes(n, 256);
  if (!L) return X3D2C_EUNSUPPORTED;
  fill_geom(ctx, dir, n, L, &p.g);
  const size_t smem = sizeof(double) * (2 * (size_t)p.g.field_doubles + 2 * (size_t)p.g.nseg * L);
  if (smem > 75 * 1024) return X3D2C_EUNSUPPORTED;
  p.in = u;
  p.out = du;
  const int threads = L * p.g.nseg;
  switch (L) {
    case 2: return dispatch_tds<2>(ctx, p, ops->tap_mask, threads, smem);
    case 4: return dispatch_tds<4>(ctx, p, ops->tap_mask, threads, smem);
    case 8: return dispatch_tds<8>(ctx, p, ops->tap_mask, threads, smem);
    case 16: return dispatch_tds<16>(ctx, p, ops->tap_mask, threads, smem);
    default: return dispatch_tds<32>(ctx, p, ops->tap_mask, threads, smem);
  }
}

}  // namespace x3d2c
