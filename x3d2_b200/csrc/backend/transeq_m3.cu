// Fused transeq fast path (periodic, uniform direction): one launch computes all three components
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
//  * copies are batched (four 16-byte chunks in flight per thread) so the shared->global stores pipeline;
//  * rank-split directions (DIST): the halo rows of each field sit behind its segments, the neighbouring ranks'
//    carries of all nine recurrences are staged next to the tiles (m3_common.cuh, m3_edge.cu).
#include "m3_common.cuh"

using namespace m3;

namespace {

constexpr int kMaxThreads = 256;  // 255 registers per thread: at most 256 threads per SM, in 1 or 2 CTAs
constexpr int NS = 9;             // recurrences per tile: 3 fields x (d f, d(f conv), d2 f)

struct Params {
  const double* in[3];  // in[0] is the line-aligned velocity (conv)
  double* out[3];
  Geom g;
  Op o_du, o_dud, o_d2u;  // du, dud scaled by -1/2; d2u unscaled (reference-order stencil)
  double d2u_scale;       // nu * fw of the second derivative
  // rank-split direction only
  const double *halo_s, *halo_e, *from_prev, *from_next;
};

// One velocity component of one tile, everything addressed by offsets into smem[]. fF: field tile (in/out, in
// place); fC: conv tile; cur / oth: the two three-field buffers whose pad rows hold the carries (ze in cur,
// ys in oth). xp / xn: neighbour carries of this field's three recurrences (DIST), lane applied.
template <int L, unsigned M1, unsigned M2, bool SELF, bool DIST>  // SELF: the field is its own conv
__device__ __forceinline__ void component(const int fF, const int fC, const int cur, const int oth, const int xp,
                                          const int xn, const Params& p, const int q, const int l, const int bm,
                                          const int b0, const int bp) {
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
      p3 = fma(p.o_d2u.a, p3, sten_exact_sym<M2>(p.o_d2u.cfw, wf));  // unscaled: nu * fw is applied below
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
  constexpr int st = SP * L, xs = EXP_ROWS * L;
  {  // nu d2 f, then -1/2 d(f conv): folded into z2
    double zi, yi;
    carries<L, DIST>(cur + 2 * fd + pad0, oth + 2 * fd + pad0, st, xp + 2 * xs, xn + 2 * xs, p.o_d2u, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z3[k] = fma(p.o_d2u.Cp[k], yi, fma(p.o_d2u.W[k], zi, z3[k]));
    carries<L, DIST>(cur + 1 * fd + pad0, oth + 1 * fd + pad0, st, xp + 1 * xs, xn + 1 * xs, p.o_dud, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z2[k] = fma(p.d2u_scale, z3[k], fma(p.o_dud.Cp[k], yi, fma(p.o_dud.W[k], zi, z2[k])));
  }
  {  // -1/2 d f, combined with conv
    double zi, yi;
    carries<L, DIST>(cur + 0 * fd + pad0, oth + 0 * fd + pad0, st, xp, xn, p.o_du, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const double du = fma(p.o_du.Cp[k], yi, fma(p.o_du.W[k], zi, z1[k]));
      smem[fF + b0 + k * L] = fma(smem[fC + b0 + k * L], du, z2[k]);
    }
  }
  __syncthreads();  // the carries are overwritten by the next component; F is complete
}

// shared memory: [2 buffers][3 fields][field_doubles], DIST: + [2 buffers][from prev: NS x 5 rows][from next: NS x 5 rows]
template <int L, unsigned M1, unsigned M2, bool DIST>
__global__ void __launch_bounds__(kMaxThreads, 1) transeq_m3_kernel(const __grid_constant__ Params p) {
  const Geom& g = p.g;
  const int fd = g.field_doubles;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  int bm, b0, bp;
  segment_bases<L, DIST>(q, l, g.nseg, bm, b0, bp);
  const Copier<L> cp;
  constexpr int xbuf = 2 * NS * EXP_ROWS * L;  // neighbour carries per buffer
  auto load_tile = [&](int buf, int tile) {
#pragma unroll
    for (int f = 0; f < 3; ++f) cp.load(smem + (3 * buf + f) * fd, p.in[f], g, tile);
    if (DIST) {
      cp.load_halos(smem + 3 * buf * fd, 3, fd, g.nseg * SP * L, p.halo_s, p.halo_e, tile);
      cp.load_rows2(smem + 6 * fd + buf * xbuf, p.from_prev, p.from_next, NS * EXP_ROWS, tile);
    }
  };
  int it = 0;
  for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
    if (it == 0) {
      load_tile(0, tile);
      cp_async_commit();
      const int nx = tile + gridDim.x;
      if (nx < g.tiles) load_tile(1, nx);
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncthreads();
    const int bo = (it & 1) * 3 * fd, oo = ((it & 1) ^ 1) * 3 * fd;
    const int xp = 6 * fd + (it & 1) * xbuf + l, xn = xp + NS * EXP_ROWS * L;
    constexpr int xf = 3 * EXP_ROWS * L;  // three recurrences per field
    double* b = smem + bo;
    // components 1 and 2 first: they read the aligned velocity (field 0) as conv; field 0 is overwritten last
    component<L, M1, M2, false, DIST>(bo + 1 * fd, bo, bo, oo, xp + 1 * xf, xn + 1 * xf, p, q, l, bm, b0, bp);
    component<L, M1, M2, false, DIST>(bo + 2 * fd, bo, bo, oo, xp + 2 * xf, xn + 2 * xf, p, q, l, bm, b0, bp);
    component<L, M1, M2, true, DIST>(bo, bo, bo, oo, xp, xn, p, q, l, bm, b0, bp);
#pragma unroll
    for (int f = 0; f < 3; ++f) cp.store(p.out[f], b + f * fd, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) load_tile(it & 1, nn);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
constexpr size_t kSmemSm = 227 * 1024;  // shared memory of one SM available to CTAs; each CTA reserves 1 KB more

template <int L, unsigned M1, unsigned M2, bool DIST>
int launch(x3d2c_ctx* ctx, const Params& p, int threads, size_t smem) {
  static bool attr_set[x3d2c::kMaxDevices] = {};  // per device: function attributes belong to the device's context
  if (!attr_set[ctx->device]) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m3_kernel<L, M1, M2, DIST>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemSm - 1024)));
    attr_set[ctx->device] = true;
  }
  int per_sm = (int)((kSmemSm + 1024) / (smem + 1024));  // CTAs that fit one SM's shared memory ...
  if (per_sm > kMaxThreads / threads) per_sm = kMaxThreads / threads;  // ... and register file
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.g.tiles) grid = p.g.tiles;
  transeq_m3_kernel<L, M1, M2, DIST><<<grid, threads, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L, bool DIST>
int dispatch_mask(x3d2c_ctx* ctx, const Params& p, bool compact, int threads, size_t smem) {
  if (compact) return launch<L, 0x6Cu, 0x7Cu, DIST>(ctx, p, threads, smem);
  return launch<L, 0x1FFu, 0x1FFu, DIST>(ctx, p, threads, smem);
}

template <bool DIST>
int dispatch_lanes(x3d2c_ctx* ctx, const Params& p, int L, bool compact, int threads, size_t smem) {
  switch (L) {
    case 2: return dispatch_mask<2, DIST>(ctx, p, compact, threads, smem);
    case 4: return dispatch_mask<4, DIST>(ctx, p, compact, threads, smem);
    case 8: return dispatch_mask<8, DIST>(ctx, p, compact, threads, smem);
    case 16: return dispatch_mask<16, DIST>(ctx, p, compact, threads, smem);
    default: return dispatch_mask<32, DIST>(ctx, p, compact, threads, smem);
  }
}

size_t smem_bytes(int nseg, int L, bool split) {
  const size_t fd = (size_t)nseg * SP * L + (split ? HALO_ROWS * L : 0);
  return sizeof(double) * (6 * fd + (split ? 2 * 2 * NS * EXP_ROWS * L : 0));
}

}  // namespace

namespace x3d2c {

int transeq_m3(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym) {
  // periodic operators do not distinguish the symmetric variants (src/tdsops.f90:277-396 only edits BC rows)
  if (!same_tables(der1st, der1st_sym) || !same_tables(der2nd, der2nd_sym)) return X3D2C_EUNSUPPORTED;
  const int n = der1st->n_tds, nseg = n / S;
  const bool split = ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist;
  if (split && !dist_supported(ctx, dir, n)) return X3D2C_EUNSUPPORTED;
  Params p{};
  // du, d(u conv): FMA stencils with -1/2 and fw folded in; d2u: the reference's summation order (sten_exact), its
  // recurrence runs unscaled (the edge kernel uses the same Op) and nu * fw multiplies the finished second derivative
  if (!make_op(der1st, -0.5, split, &p.o_du, false) || !make_op(der1st, -0.5, split, &p.o_dud, false) ||
      !make_op(der2nd, nu, split, &p.o_d2u, true))
    return X3D2C_EUNSUPPORTED;
  for (int k = 0; k < 4; ++k)  // sten_exact_sym shares the products of mirror taps
    if (p.o_d2u.cfw[k] != p.o_d2u.cfw[8 - k]) return X3D2C_EUNSUPPORTED;
  p.d2u_scale = p.o_d2u.fs;
  p.o_d2u.fs = 1.0;
  // Tile width L (lanes), threads = L * nseg per CTA: the choice that keeps most threads resident per SM (shared
  // memory and the 255-register budget both limit it), rows of at least one 32-byte DRAM sector (L >= 4) when
  // possible, and on a tie more CTAs per SM (their copy and FP64 phases interleave). X3D2C_TRANSEQ_THREADS
  // restricts the CTA size.
  static int forced_threads = -1;
  if (forced_threads < 0) {
    const char* e = std::getenv("X3D2C_TRANSEQ_THREADS");
    forced_threads = e ? std::atoi(e) : 0;
    if (forced_threads != 64 && forced_threads != 128 && forced_threads != 256) forced_threads = 0;
  }
  int L = 0, best = -1;
  for (int cand = 32; cand >= 2; cand >>= 1) {
    const int threads = cand * nseg;
    if (threads > kMaxThreads || threads % 32 || (forced_threads && threads != forced_threads)) continue;
    const size_t sm = smem_bytes(nseg, cand, split);
    if (sm > kSmemSm - 1024) continue;
    int per_sm = (int)((kSmemSm + 1024) / (sm + 1024));
    if (per_sm > kMaxThreads / threads) per_sm = kMaxThreads / threads;
    const int score = (cand >= 4 ? 100000 : 0) + per_sm * threads * 8 + per_sm;
    if (score > best) { best = score; L = cand; }
  }
  if (!L) return X3D2C_EUNSUPPORTED;
  const size_t smem = smem_bytes(nseg, L, split);
  p.g.n = n;
  p.g.n_pad = ctx->n_pad(dir);
  p.g.nseg = nseg;
  p.g.tiles = ctx->n_groups[dir] * (SZ / L);
  p.g.field_doubles = nseg * SP * L + (split ? HALO_ROWS * L : 0);
  if (dir == X3D2C_DIR_X) { p.out[0] = du; p.out[1] = dv; p.out[2] = dw; p.in[0] = u; p.in[1] = v; p.in[2] = w; }
  else if (dir == X3D2C_DIR_Y) { p.out[0] = dv; p.out[1] = du; p.out[2] = dw; p.in[0] = v; p.in[1] = u; p.in[2] = w; }
  else { p.out[0] = dw; p.out[1] = du; p.out[2] = dv; p.in[0] = w; p.in[1] = u; p.in[2] = v; }
  const int threads = L * nseg;
  const bool compact = der1st->tap_mask == 0x6Cu && der2nd->tap_mask == 0x7Cu;
  if (!split) return dispatch_lanes<false>(ctx, p, L, compact, threads, smem);
  DistBufs b = carve_dist(ctx);
  EdgeParams ep{};
  ep.n = n;
  ep.n_pad = p.g.n_pad;
  ep.nseg = nseg;
  ep.ns = NS;
  ep.transeq = 1;
  ep.ops[0] = p.o_du;
  ep.ops[1] = p.o_dud;
  ep.ops[2] = p.o_d2u;
  for (int f = 0; f < 3; ++f) ep.f[f] = p.in[f];  // recurrence f*3 + k: k = 0 d f, 1 d(f conv), 2 d2 f
  int rc = exchange_edges(ctx, dir, p.in, 3, ep, b);
  if (rc) return rc;
  p.halo_s = b.halo_recv_s;
  p.halo_e = b.halo_recv_e;
  p.from_prev = b.carr_from_prev;
  p.from_next = b.carr_from_next;
  return dispatch_lanes<true>(ctx, p, L, compact, threads, smem);
}

}  // namespace x3d2c
