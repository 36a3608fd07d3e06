// Fast path of the DistD2-TDS operators for periodic, uniform directions ("m3" kernels).
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
//  * the line is cut into 16-point segments, one thread per (lane, segment); each thread keeps its
//    recurrences in registers: stencil -> local forward sweep -> local backward sweep;
//  * the sweeps use the converged (Toeplitz) factors fw, bw, alpha of the tdsops tables. For a periodic line
//    A = L U holds exactly with these constant factors, so no substitution phase (dist_sa/dist_sc) is needed;
//    segments are coupled through carries that decay like (alpha fw)^16 per segment, exchanged once per
//    component through shared memory and summed over D <= 3 neighbouring segments (wrapping periodically, or
//    reaching into the neighbouring ranks' segments when the direction is rank-split: m3_common.cuh, m3_edge.cu).
//    Truncation is below 1e-18 relative; differences to the sequential reference order are rounding only
//    (measured 2-7e-16 relative, tests/test_gpu_fast_path.py);
//  * -1/2 and nu of the transeq combination are folded into the stencil coefficients, FMA everywhere.
// Shapes that do not qualify (non-periodic operators, stretched meshes, strict mode) use the reference-order
// kernels of tds_m1.cu.
#include "m3_common.cuh"

using namespace m3;

namespace {

struct TdsParams {
  const double* in;
  double* out;
  Geom g;
  Op o;
  // rank-split direction only
  const double *halo_s, *halo_e, *from_prev, *from_next;
};

// shared memory: [2 field tiles][ze, ys: nseg*L each][DIST: 2 x EXT_ROWS*L neighbour carries]
template <int L, unsigned M, bool DIST>
__global__ void __launch_bounds__(256, 3) tds_m3_kernel(const __grid_constant__ TdsParams p) {
  const Geom& g = p.g;
  const int fd = g.field_doubles, nseg = g.nseg;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  int bm, b0, bp;
  segment_bases<L, DIST>(q, l, nseg, bm, b0, bp);
  const int ze = 2 * fd + l, ys = ze + nseg * L, ext0 = 2 * fd + 2 * nseg * L + l;
  const Copier<L> cp;
  auto load_tile = [&](int buf, int tile) {
    cp.load(smem + buf * fd, p.in, g, tile);
    if (DIST) {
      cp.load_rows2(smem + buf * fd + nseg * SP * L, p.halo_s, p.halo_e, 4, tile);
      cp.load_rows2(smem + 2 * fd + 2 * nseg * L + buf * EXT_ROWS * L, p.from_prev, p.from_next, EXP_ROWS, tile);
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
        pz = fma(p.o.a, pz, p.o.fs * sten_exact<M>(p.o.cfw, wf));
        z[k] = pz;
#pragma unroll
        for (int t = 0; t < 8; ++t) wf[t] = wf[t + 1];
      }
    }
    smem[ze + q * L] = z[S - 1];
    {
      double y = 0.0;
#pragma unroll
      for (int k = S - 1; k >= 0; --k) {
        y = fma(p.o.cb, y, z[k]);
        z[k] = y;
      }
    }
    smem[ys + q * L] = z[0];
    __syncthreads();
    double zin, yin;
    const int ext = ext0 + (it & 1) * EXT_ROWS * L;
    carries<L, DIST>(ze, ys, L, ext, ext + EXP_ROWS * L, p.o, q, nseg, zin, yin);
#pragma unroll
    for (int k = 0; k < S; ++k) F[b0 + k * L] = fma(p.o.Cp[k], yin, fma(p.o.W[k], zin, z[k]));
    __syncthreads();
    cp.store(p.out, F, g, tile);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) load_tile(it & 1, nn);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
// tile width: the largest power of two L <= 32 with L * nseg <= max_threads
int pick_lanes(int n, int max_threads) {
  int L = 32;
  while (L >= 2 && L * (n / S) > max_threads) L >>= 1;
  if (L < 2) return 0;
  const int threads = L * (n / S);
  if (threads < 32 || threads % 32) return 0;
  return L;
}

constexpr size_t kSmemMax = 75 * 1024;  // three CTAs per SM

template <int L, unsigned M, bool DIST>
int launch_tds(x3d2c_ctx* ctx, const TdsParams& p, int threads, size_t smem) {
  static bool attr_set[x3d2c::kMaxDevices] = {};  // per device: function attributes belong to the device's context
  if (!attr_set[ctx->device]) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(tds_m3_kernel<L, M, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kSmemMax));
    attr_set[ctx->device] = true;
  }
  int grid = num_sms(ctx) * 3;
  if (grid > p.g.tiles) grid = p.g.tiles;
  tds_m3_kernel<L, M, DIST><<<grid, threads, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L, bool DIST>
int dispatch_mask(x3d2c_ctx* ctx, const TdsParams& p, unsigned mask, int threads, size_t smem) {
  switch (mask) {
    case 0x6Cu: return launch_tds<L, 0x6Cu, DIST>(ctx, p, threads, smem);  // first derivative
    case 0x7Cu: return launch_tds<L, 0x7Cu, DIST>(ctx, p, threads, smem);  // second derivative
    case 0x78u: return launch_tds<L, 0x78u, DIST>(ctx, p, threads, smem);  // staggered derivative / interpolation v2p
    case 0x3Cu: return launch_tds<L, 0x3Cu, DIST>(ctx, p, threads, smem);  // staggered derivative / interpolation p2v
    default: return launch_tds<L, 0x1FFu, DIST>(ctx, p, threads, smem);
  }
}

template <bool DIST>
int dispatch_lanes(x3d2c_ctx* ctx, const TdsParams& p, int L, unsigned mask, int threads, size_t smem) {
  switch (L) {
    case 2: return dispatch_mask<2, DIST>(ctx, p, mask, threads, smem);
    case 4: return dispatch_mask<4, DIST>(ctx, p, mask, threads, smem);
    case 8: return dispatch_mask<8, DIST>(ctx, p, mask, threads, smem);
    case 16: return dispatch_mask<16, DIST>(ctx, p, mask, threads, smem);
    default: return dispatch_mask<32, DIST>(ctx, p, mask, threads, smem);
  }
}

}  // namespace

namespace x3d2c {

int tds_solve_m3(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops) {
  const int n = ops->n_tds;
  const bool split = ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist;
  if (split && !dist_supported(ctx, dir, n)) return X3D2C_EUNSUPPORTED;
  TdsParams p{};
  if (!make_op(ops, 1.0, split, &p.o)) return X3D2C_EUNSUPPORTED;
  auto smem_for = [&](int lanes) {
    const size_t fd = (size_t)(n / S) * SP * lanes + (split ? HALO_ROWS * lanes : 0);
    return sizeof(double) * (2 * fd + 2 * (size_t)(n / S) * lanes + (split ? 2 * EXT_ROWS * lanes : 0));
  };
  int L = pick_lanes(n, 256);
  while (L >= 4 && smem_for(L) > kSmemMax) L >>= 1;  // rank-split lines carry halo rows and neighbour carries
  if (!L || (L * (n / S)) % 32) return X3D2C_EUNSUPPORTED;
  p.g.n = n;
  p.g.n_pad = ctx->n_pad(dir);
  p.g.nseg = n / S;
  p.g.tiles = ctx->n_groups[dir] * (SZ / L);
  p.g.field_doubles = p.g.nseg * SP * L + (split ? HALO_ROWS * L : 0);
  const size_t smem = smem_for(L);
  if (smem > kSmemMax) return X3D2C_EUNSUPPORTED;
  p.in = u;
  p.out = du;
  const int threads = L * p.g.nseg;
  if (!split) return dispatch_lanes<false>(ctx, p, L, ops->tap_mask, threads, smem);
  DistBufs b = carve_dist(ctx);
  EdgeParams ep{};
  ep.n = n;
  ep.n_pad = p.g.n_pad;
  ep.nseg = p.g.nseg;
  ep.ns = 1;
  ep.f[0] = u;
  ep.ops[0] = p.o;
  const double* fields[1] = {u};
  int rc = exchange_edges(ctx, dir, fields, 1, ep, b);
  if (rc) return rc;
  p.halo_s = b.halo_recv_s;
  p.halo_e = b.halo_recv_e;
  p.from_prev = b.carr_from_prev;
  p.from_next = b.carr_from_next;
  return dispatch_lanes<true>(ctx, p, L, ops->tap_mask, threads, smem);
}

}  // namespace x3d2c
