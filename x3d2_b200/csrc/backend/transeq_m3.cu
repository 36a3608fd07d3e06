// Fused transeq fast path (periodic, single-rank direction): one launch computes all three components
//     rhs_c = -1/2 (conv * d f_c + d(f_c conv)) + nu d2 f_c ,   c = 1..3,
// reading u, v, w once and writing du, dv, dw once (48 B per grid point; replaces transeq_{x,y,z}_omp ->
// transeq_omp_dist -> exec_dist_transeq_compact, src/backend/omp/backend.f90:145-338, exec_dist.f90:67-186,
// whose kernels move 16 doubles per point per component, perf_cuda_transeq.f90:16).
//
// Same method as tds_m3.cu (constant Toeplitz factors, local sweeps in registers, decaying carries between
// segments), tuned for the FP64-issue-heavy fused operator:
//  * 16-point segments: 3 recurrences x 16 points live in 96 registers (an 8-point variant with 128 registers
//    and twice the warps was measured slower: more carry work and spills);
//  * carries from the DMAX = 3 nearest segments on either side ((alpha fw)^48 < 1e-18);
//  * the tile is [segment][17 rows][L lanes]: the 17th (pad) row removes shared-memory bank conflicts of the
//    column accesses and stores the carries (3 fields x 2 buffers = the 6 carry values per segment and lane),
//    so two CTAs, each double buffering a tile of three fields, fit in one SM (2 x 102 KB) and the copy
//    phases of one CTA overlap the FP64 phases of the other;
//  * copies are batched (four 16-byte chunks in flight per thread) so the shared->global stores pipeline.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int S = 16;      // points per segment
constexpr int SP = S + 1;  // rows per segment in shared memory
constexpr int DMAX = 3;    // neighbouring segments that contribute to a carry
constexpr int LOG2S = 4;
constexpr int kMaxThreads = 256;  // 255 registers per thread: at most 256 threads per SM, in 1, 2 or 4 CTAs

struct Op {
  double cfw[9];            // scale * fw * coeffs
  double a, cb;             // forward / backward propagators
  double zw[DMAX], yw[DMAX];
  double om[2 * DMAX - 1];  // index m + DMAX - 1
  double W[S], Cp[S];
};

struct Geom {
  int n, n_pad, nseg, tiles, field_doubles;
};

struct Params {
  const double* in[3];  // in[0] is the line-aligned velocity (conv)
  double* out[3];
  Geom g;
  Op o_du, o_dud, o_d2u;  // scaled by -1/2, -1/2, nu
};

extern __shared__ __align__(16) double smem[];

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

// Tile copies. Global: (32 lanes, n_pad rows, G groups); shared: [segment][SP rows][L lanes].
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
  __device__ __forceinline__ const double* tile_base(const double* g, const Geom& q, int tile) const {
    constexpr int tpg = SZ / L;
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    return g + (size_t)grp * q.n_pad * SZ + l0 + g_off;
  }
  __device__ __forceinline__ void load(double* sm, const double* g, const Geom& q, int tile) const {
    const double* src = tile_base(g, q, tile);
    for (int r = j; r < q.n; r += rows_per_pass) {
      cp_async16(sm + (r + (r >> LOG2S)) * L + c2, src);
      src += (size_t)rows_per_pass * SZ;
    }
  }
  __device__ __forceinline__ void store(double* g, const double* sm, const Geom& q, int tile) const {
    double* dst = const_cast<double*>(tile_base(g, q, tile));
    for (int r0 = j; r0 < q.n; r0 += 4 * rows_per_pass) {  // four chunks in flight per thread
      double2 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i * rows_per_pass;
        if (r < q.n) v[i] = *reinterpret_cast<const double2*>(sm + (r + (r >> LOG2S)) * L + c2);
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

// ze / ys: shared-memory offsets of the pad row of segment 0 (lane offset applied); segment s is at + s * SP * L
template <int L>
__device__ __forceinline__ void carries(const int ze, const int ys, const Op& o, int q, int nseg, double& zin,
                                        double& yin) {
  double zv[2 * DMAX];  // ze(q - DMAX .. q + DMAX - 1)
#pragma unroll
  for (int t = 0; t < 2 * DMAX; ++t) {
    int s = q - DMAX + t;
    if (s < 0) s += nseg;
    if (s >= nseg) s -= nseg;
    zv[t] = smem[ze + s * (SP * L)];
  }
  zin = 0.0;
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) zin = fma(o.zw[d - 1], zv[DMAX - d], zin);
  yin = 0.0;
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) {
    int s = q + d;
    if (s >= nseg) s -= nseg;
    yin = fma(o.yw[d - 1], smem[ys + s * (SP * L)], yin);
  }
#pragma unroll
  for (int m = -(DMAX - 1); m <= DMAX - 1; ++m) yin = fma(o.om[m + DMAX - 1], zv[DMAX + m], yin);
}

// window element t (row j0 - 4 + t, t = 0..S+7) given the bases of the previous, own and next segment
template <int L>
__device__ __forceinline__ int woff(int t, int bm, int b0, int bp) {
  return t < 4 ? bm + (S - 4 + t) * L : (t < S + 4 ? b0 + (t - 4) * L : bp + (t - S - 4) * L);
}

// One velocity component of one tile, everything addressed by offsets into smem[]. fF: field tile (in/out, in
// place); fC: conv tile; cur / oth: the two three-field buffers whose pad rows hold the carries (ze in cur,
// ys in oth).
template <int L, unsigned M1, unsigned M2, bool SELF>  // SELF: the field is its own conv (aligned velocity)
__device__ __forceinline__ void component(const int fF, const int fC, const int cur, const int oth, const Params& p,
                                          const int q, const int l, const int bm, const int b0, const int bp) {
  const int nseg = p.g.nseg, fd = p.g.field_doubles;
  double z1[S], z2[S], z3[S];
  {
    double wf[9], wp[9];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int o = woff<L>(t, bm, b0, bp);
      wf[t] = smem[fF + o];
      wp[t] = wf[t] * (SELF ? wf[t] : smem[fC + o]);
    }
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const int o = woff<L>(k + 8, bm, b0, bp);
      wf[8] = smem[fF + o];
      wp[8] = wf[8] * (SELF ? wf[8] : smem[fC + o]);
      p1 = fma(p.o_du.a, p1, sten<M1>(p.o_du.cfw, wf));
      p2 = fma(p.o_dud.a, p2, sten<M1>(p.o_dud.cfw, wp));
      p3 = fma(p.o_d2u.a, p3, sten<M2>(p.o_d2u.cfw, wf));
      z1[k] = p1; z2[k] = p2; z3[k] = p3;
#pragma unroll
      for (int t = 0; t < 8; ++t) { wf[t] = wf[t + 1]; wp[t] = wp[t + 1]; }
    }
  }
  const int pad = b0 + S * L;  // pad row of this thread's segment
  smem[cur + 0 * fd + pad] = z1[S - 1];
  smem[cur + 1 * fd + pad] = z2[S - 1];
  smem[cur + 2 * fd + pad] = z3[S - 1];
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
  smem[oth + 0 * fd + pad] = z1[0];
  smem[oth + 1 * fd + pad] = z2[0];
  smem[oth + 2 * fd + pad] = z3[0];
  __syncthreads();
  const int pad0 = S * L + l;  // pad row of segment 0
  {  // nu d2 f, then -1/2 d(f conv): folded into z2
    double zi, yi;
    carries<L>(cur + 2 * fd + pad0, oth + 2 * fd + pad0, p.o_d2u, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z3[k] = fma(p.o_d2u.Cp[k], yi, fma(p.o_d2u.W[k], zi, z3[k]));
    carries<L>(cur + 1 * fd + pad0, oth + 1 * fd + pad0, p.o_dud, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z2[k] = fma(p.o_dud.Cp[k], yi, fma(p.o_dud.W[k], zi, z2[k])) + z3[k];
  }
  {  // -1/2 d f, combined with conv
    double zi, yi;
    carries<L>(cur + 0 * fd + pad0, oth + 0 * fd + pad0, p.o_du, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const double du = fma(p.o_du.Cp[k], yi, fma(p.o_du.W[k], zi, z1[k]));
      smem[fF + b0 + k * L] = fma(smem[fC + b0 + k * L], du, z2[k]);
    }
  }
  __syncthreads();  // the carries are overwritten by the next component; F is complete
}

template <int L, unsigned M1, unsigned M2>
__global__ void __launch_bounds__(kMaxThreads, 1) transeq_m3_kernel(const __grid_constant__ Params p) {
  const Geom& g = p.g;
  const int fd = g.field_doubles;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  const int qm = q == 0 ? g.nseg - 1 : q - 1, qp = q == g.nseg - 1 ? 0 : q + 1;
  const int bm = qm * SP * L + l, b0 = q * SP * L + l, bp = qp * SP * L + l;
  const Copier<L> cp;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
    if (it == 0) {
#pragma unroll
      for (int f = 0; f < 3; ++f) cp.load(smem + f * fd, p.in[f], g, tile);
      cp_async_commit();
      const int nx = tile + gridDim.x;
      if (nx < g.tiles) {
#pragma unroll
        for (int f = 0; f < 3; ++f) cp.load(smem + (3 + f) * fd, p.in[f], g, nx);
      }
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncthreads();
    const int bo = (it & 1) * 3 * fd, oo = ((it & 1) ^ 1) * 3 * fd;
    double* b = smem + bo;
    // components 1 and 2 first: they read the aligned velocity (field 0) as conv; field 0 is overwritten last
    component<L, M1, M2, false>(bo + 1 * fd, bo, bo, oo, p, q, l, bm, b0, bp);
    component<L, M1, M2, false>(bo + 2 * fd, bo, bo, oo, p, q, l, bm, b0, bp);
    component<L, M1, M2, true>(bo, bo, bo, oo, p, q, l, bm, b0, bp);
#pragma unroll
    for (int f = 0; f < 3; ++f) cp.store(p.out[f], b + f * fd, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) {
#pragma unroll
      for (int f = 0; f < 3; ++f) cp.load(b + f * fd, p.in[f], g, nn);
    }
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

// constant set of one operator; false when the operator does not qualify for the fast path
bool make_op(const x3d2c_tdsops* t, double scale, Op* o) {
  const int n = t->n_tds;
  if (!t->periodic || t->n_rhs != n || n < 64 || n % 16) return false;
  if (t->has_stretch || t->has_stretch_correct) return false;
  const int m = n / 2;
  const double fw = t->h_fw[m], bw = t->h_bw[m], al = t->h_af[m];
  for (int j = 40; j < n - 40; ++j) {  // Toeplitz limit reached over the central region
    if (std::fabs(t->h_fw[j] - fw) > 4e-16 * std::fabs(fw) || std::fabs(t->h_bw[j] - bw) > 4e-16 * std::fabs(bw) ||
        t->h_af[j] != al)
      return false;
  }
  for (int k = 0; k < 9; ++k) o->cfw[k] = scale * fw * t->dev.coeffs[k];
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

int num_sms(const x3d2c_ctx* ctx) {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  return sms > 0 ? sms : 148;
}

constexpr size_t kSmemMax = 113 * 1024;

template <int L, unsigned M1, unsigned M2>
int launch(x3d2c_ctx* ctx, const Params& p, int threads, size_t smem) {
  static bool attr_set = false;
  if (!attr_set) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m3_kernel<L, M1, M2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kSmemMax));
    attr_set = true;
  }
  int grid = num_sms(ctx) * (kMaxThreads / threads);  // CTAs that fit one SM's register file
  if (grid > p.g.tiles) grid = p.g.tiles;
  transeq_m3_kernel<L, M1, M2><<<grid, threads, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L>
int dispatch(x3d2c_ctx* ctx, const Params& p, bool compact, int threads, size_t smem) {
  if (compact) return launch<L, 0x6Cu, 0x7Cu>(ctx, p, threads, smem);
  return launch<L, 0x1FFu, 0x1FFu>(ctx, p, threads, smem);
}

}  // namespace

namespace x3d2c {

int transeq_m3(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym) {
  // periodic operators do not distinguish the symmetric variants (src/tdsops.f90:277-396 only edits BC rows)
  if (!same_tables(der1st, der1st_sym) || !same_tables(der2nd, der2nd_sym)) return X3D2C_EUNSUPPORTED;
  Params p;
  if (!make_op(der1st, -0.5, &p.o_du) || !make_op(der1st, -0.5, &p.o_dud) || !make_op(der2nd, nu, &p.o_d2u))
    return X3D2C_EUNSUPPORTED;
  const int n = der1st->n_tds, nseg = n / S;
  // threads per CTA: 128 (two CTAs per SM interleave their copy and FP64 phases); X3D2C_TRANSEQ_THREADS overrides
  static int target_threads = 0;
  if (!target_threads) {
    const char* e = std::getenv("X3D2C_TRANSEQ_THREADS");
    target_threads = e ? std::atoi(e) : 128;
    if (target_threads != 64 && target_threads != 128 && target_threads != 256) target_threads = 128;
  }
  int L = 32;
  while (L >= 2 && L * nseg > target_threads) L >>= 1;
  if (L < 2 || (L * nseg) % 32) return X3D2C_EUNSUPPORTED;
  p.g.n = n;
  p.g.n_pad = ctx->n_pad(dir);
  p.g.nseg = nseg;
  p.g.tiles = ctx->n_groups[dir] * (SZ / L);
  p.g.field_doubles = nseg * SP * L;
  const size_t smem = sizeof(double) * 6 * (size_t)p.g.field_doubles;
  if (smem > kSmemMax) return X3D2C_EUNSUPPORTED;
  if (dir == X3D2C_DIR_X) { p.out[0] = du; p.out[1] = dv; p.out[2] = dw; p.in[0] = u; p.in[1] = v; p.in[2] = w; }
  else if (dir == X3D2C_DIR_Y) { p.out[0] = dv; p.out[1] = du; p.out[2] = dw; p.in[0] = v; p.in[1] = u; p.in[2] = w; }
  else { p.out[0] = dw; p.out[1] = du; p.out[2] = dv; p.in[0] = w; p.in[1] = u; p.in[2] = v; }
  const int threads = L * nseg;
  const bool compact = der1st->tap_mask == 0x6Cu && der2nd->tap_mask == 0x7Cu;
  switch (L) {
    case 2: return dispatch<2>(ctx, p, compact, threads, smem);
    case 4: return dispatch<4>(ctx, p, compact, threads, smem);
    case 8: return dispatch<8>(ctx, p, compact, threads, smem);
    case 16: return dispatch<16>(ctx, p, compact, threads, smem);
    default: return dispatch<32>(ctx, p, compact, threads, smem);
  }
}

}  // namespace x3d2c
