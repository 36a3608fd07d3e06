// Two tridiagonal operators of one direction in a single pass over memory (fast path of x3d2c_tds_solve_sum,
// x3d2c_tds_solve_dual and x3d2c_tds_solve_axpy; extensions of the operator API, include/x3d2c.h).
//
// The reference's vector calculus applies two operators and combines them with a separate vecadd, or applies two
// operators to the same field in two calls (divergence_v2c / gradient_c2v, src/vector_calculus.f90:142-332):
//     SUM :  out   = A(in_a) + B(in_b)        reference: 2 tds_solve + vecadd = 56 B/pt, here 24 B/pt
//     DUAL:  out_a = A(in), out_b = B(in)     reference: 2 tds_solve          = 32 B/pt, here 24 B/pt
//     AXPY:  y     = y + a A(in)              reference: tds_solve + vecadd   = 40 B/pt, here 24 B/pt
// Same method and shared-memory layout as tds_m3.cu with two field slots per buffer; rank-split directions exchange
// the halos and carries of both recurrences in one pair of messages (m3_edge.cu).
#include "m3_common.cuh"

using namespace m3;

namespace {

enum Mode { SUM = 0, DUAL = 1, AXPY = 2 };

struct PairParams {
  const double *in_a, *in_b;  // SUM: two inputs; DUAL: in_a only; AXPY: in_a, and in_b = y
  double *out_a, *out_b;      // SUM: out_a; DUAL: both; AXPY: out_a = y
  Geom g;
  Op oa, ob;
  const double *halo_s, *halo_e, *from_prev, *from_next;  // rank-split direction only
};

template <int L, unsigned M>
__device__ __forceinline__ void local_sweeps(const double* F, const Op& o, int bm, int b0, int bp, double (&z)[S]) {
  double wf[9];
#pragma unroll
  for (int t = 0; t < 8; ++t) wf[t] = F[woff<L>(t, bm, b0, bp)];
  double pz = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    wf[8] = F[woff<L>(k + 8, bm, b0, bp)];
    pz = fma(o.a, pz, o.fs * sten_exact<M>(o.cfw, wf));
    z[k] = pz;
#pragma unroll
    for (int t = 0; t < 8; ++t) wf[t] = wf[t + 1];
  }
}

__device__ __forceinline__ void backward(const Op& o, double (&z)[S]) {
  double y = 0.0;
#pragma unroll
  for (int k = S - 1; k >= 0; --k) {
    y = fma(o.cb, y, z[k]);
    z[k] = y;
  }
}

// shared memory: [2 buffers][2 slots][field_doubles] [NR x (ze, ys): nseg*L each] [DIST: 2 buffers x NR x EXT_ROWS*L]
template <int L, unsigned M, int MODE, bool DIST>
__global__ void __launch_bounds__(256, 1) tds_pair_kernel(const __grid_constant__ PairParams p) {
  constexpr int NR = MODE == AXPY ? 1 : 2;
  const Geom& g = p.g;
  const int fd = g.field_doubles, nseg = g.nseg;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  int bm, b0, bp;
  segment_bases<L, DIST>(q, l, nseg, bm, b0, bp);
  const int carr = 4 * fd + l, ext0 = 4 * fd + 2 * NR * nseg * L;
  const Copier<L> cp;
  auto load_tile = [&](int buf, int tile) {
    cp.load(smem + 2 * buf * fd, p.in_a, g, tile);
    if (MODE != DUAL) cp.load(smem + (2 * buf + 1) * fd, p.in_b, g, tile);
    if (DIST) {
      cp.load_halos(smem + 2 * buf * fd, MODE == SUM ? 2 : 1, fd, nseg * SP * L, p.halo_s, p.halo_e, tile);
      cp.load_rows2(smem + ext0 + buf * NR * EXT_ROWS * L, p.from_prev, p.from_next, NR * EXP_ROWS, tile);
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
    double* F0 = smem + 2 * (it & 1) * fd;
    double* F1 = F0 + fd;
    double za[S], zb[S];
    local_sweeps<L, M>(F0, p.oa, bm, b0, bp, za);
    smem[carr + q * L] = za[S - 1];
    backward(p.oa, za);
    smem[carr + (nseg + q) * L] = za[0];
    if (NR == 2) {
      double(&zb2)[S] = zb;
      local_sweeps<L, M>(MODE == SUM ? F1 : F0, p.ob, bm, b0, bp, zb2);
      smem[carr + (2 * nseg + q) * L] = zb2[S - 1];
      backward(p.ob, zb2);
      smem[carr + (3 * nseg + q) * L] = zb2[0];
    }
    __syncthreads();
    const int xp = ext0 + (it & 1) * NR * EXT_ROWS * L + l, xn = xp + NR * EXP_ROWS * L;
    double zin, yin;
    carries<L, DIST>(carr, carr + nseg * L, L, xp, xn, p.oa, q, nseg, zin, yin);
#pragma unroll
    for (int k = 0; k < S; ++k) za[k] = fma(p.oa.Cp[k], yin, fma(p.oa.W[k], zin, za[k]));
    if (NR == 2) {
      double(&zb2)[S] = zb;
      carries<L, DIST>(carr + 2 * nseg * L, carr + 3 * nseg * L, L, xp + EXP_ROWS * L, xn + EXP_ROWS * L, p.ob, q, nseg,
                       zin, yin);
#pragma unroll
      for (int k = 0; k < S; ++k) zb2[k] = fma(p.ob.Cp[k], yin, fma(p.ob.W[k], zin, zb2[k]));
      if (MODE == SUM) {
#pragma unroll
        for (int k = 0; k < S; ++k) F0[b0 + k * L] = za[k] + zb2[k];
      } else {
#pragma unroll
        for (int k = 0; k < S; ++k) { F0[b0 + k * L] = za[k]; F1[b0 + k * L] = zb2[k]; }
      }
    } else {  // AXPY: the scale is folded into oa.cfw
#pragma unroll
      for (int k = 0; k < S; ++k) F1[b0 + k * L] += za[k];
    }
    __syncthreads();
    if (MODE != AXPY) cp.store(p.out_a, F0, g, tile);
    if (MODE == DUAL) cp.store(p.out_b, F1, g, tile);
    if (MODE == AXPY) cp.store(p.out_a, F1, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) load_tile(it & 1, nn);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
constexpr size_t kSmemSm = 227 * 1024;

size_t smem_bytes(int nseg, int L, int nr, bool split) {
  const size_t fd = (size_t)nseg * SP * L + (split ? HALO_ROWS * L : 0);
  return sizeof(double) * (4 * fd + 2 * (size_t)nr * nseg * L + (split ? 2 * (size_t)nr * EXT_ROWS * L : 0));
}

template <int L, unsigned M, int MODE, bool DIST>
int launch(x3d2c_ctx* ctx, const PairParams& p, int threads, size_t smem) {
  static bool attr_set[x3d2c::kMaxDevices] = {};  // per device: function attributes belong to the device's context
  if (!attr_set[ctx->device]) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(tds_pair_kernel<L, M, MODE, DIST>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemSm - 1024)));
    attr_set[ctx->device] = true;
  }
  int per_sm = (int)((kSmemSm + 1024) / (smem + 1024));
  if (per_sm > 256 / threads) per_sm = 256 / threads;  // register file
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.g.tiles) grid = p.g.tiles;
  tds_pair_kernel<L, M, MODE, DIST><<<grid, threads, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L, int MODE, bool DIST>
int dispatch_mask(x3d2c_ctx* ctx, const PairParams& p, unsigned mask, int threads, size_t smem) {
  switch (mask) {
    case 0x78u: return launch<L, 0x78u, MODE, DIST>(ctx, p, threads, smem);  // v2p staggered derivative / interpolation
    case 0x3Cu: return launch<L, 0x3Cu, MODE, DIST>(ctx, p, threads, smem);  // p2v
    default: return launch<L, 0x1FFu, MODE, DIST>(ctx, p, threads, smem);
  }
}

template <int MODE, bool DIST>
int dispatch_lanes(x3d2c_ctx* ctx, const PairParams& p, int L, unsigned mask, int threads, size_t smem) {
  switch (L) {
    case 4: return dispatch_mask<4, MODE, DIST>(ctx, p, mask, threads, smem);
    case 8: return dispatch_mask<8, MODE, DIST>(ctx, p, mask, threads, smem);
    case 16: return dispatch_mask<16, MODE, DIST>(ctx, p, mask, threads, smem);
    default: return dispatch_mask<32, MODE, DIST>(ctx, p, mask, threads, smem);
  }
}

template <int MODE>
int run(x3d2c_ctx* ctx, int dir, PairParams& p, const x3d2c_tdsops* ta, const x3d2c_tdsops* tb, double scale_a) {
  constexpr int NR = MODE == AXPY ? 1 : 2;
  const int n = ta->n_tds, nseg = n / S;
  if (NR == 2 && (tb->n_tds != n || tb->n_rhs != ta->n_rhs)) return X3D2C_EUNSUPPORTED;
  const bool split = ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist;
  if (split && !dist_supported(ctx, dir, n)) return X3D2C_EUNSUPPORTED;
  if (!make_op(ta, scale_a, split, &p.oa)) return X3D2C_EUNSUPPORTED;
  if (NR == 2 && !make_op(tb, 1.0, split, &p.ob)) return X3D2C_EUNSUPPORTED;
  const unsigned mask = NR == 2 ? (ta->tap_mask | tb->tap_mask) : ta->tap_mask;
  // widest tile (rows of L x 8 bytes) whose two double-buffered slots fit one SM, at most 256 threads
  int L = 0;
  for (int cand = 32; cand >= 4; cand >>= 1) {
    const int threads = cand * nseg;
    if (threads > 256 || threads % 32) continue;
    if (smem_bytes(nseg, cand, NR, split) > kSmemSm - 1024) continue;
    L = cand;
    break;
  }
  if (!L) return X3D2C_EUNSUPPORTED;
  const size_t smem = smem_bytes(nseg, L, NR, split);
  p.g.n = n;
  p.g.n_pad = ctx->n_pad(dir);
  p.g.nseg = nseg;
  p.g.tiles = ctx->n_groups[dir] * (SZ / L);
  p.g.field_doubles = nseg * SP * L + (split ? HALO_ROWS * L : 0);
  const int threads = L * nseg;
  if (!split) return dispatch_lanes<MODE, false>(ctx, p, L, mask, threads, smem);
  DistBufs b = carve_dist(ctx);
  EdgeParams ep{};
  ep.n = n;
  ep.n_pad = p.g.n_pad;
  ep.nseg = nseg;
  ep.ns = NR;
  ep.ops[0] = p.oa;
  ep.ops[1] = p.ob;
  ep.f[0] = p.in_a;
  ep.f[1] = p.in_b;
  const double* fields[2] = {p.in_a, p.in_b};
  int rc = exchange_edges(ctx, dir, fields, MODE == SUM ? 2 : 1, ep, b);
  if (rc) return rc;
  p.halo_s = b.halo_recv_s;
  p.halo_e = b.halo_recv_e;
  p.from_prev = b.carr_from_prev;
  p.from_next = b.carr_from_next;
  return dispatch_lanes<MODE, true>(ctx, p, L, mask, threads, smem);
}

}  // namespace

namespace x3d2c {

int tds_pair_m3(x3d2c_ctx* ctx, int dir, int mode, double* out_a, double* out_b, const double* in_a,
                const double* in_b, const x3d2c_tdsops* ta, const x3d2c_tdsops* tb, double scale_a) {
  PairParams p{};
  p.in_a = in_a;
  p.in_b = in_b;
  p.out_a = out_a;
  p.out_b = out_b;
  switch (mode) {
    case SUM: return run<SUM>(ctx, dir, p, ta, tb, 1.0);
    case DUAL: return run<DUAL>(ctx, dir, p, ta, tb, 1.0);
    default: return run<AXPY>(ctx, dir, p, ta, ta, scale_a);
  }
}

}  // namespace x3d2c
