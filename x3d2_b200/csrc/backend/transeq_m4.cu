// transeq fast path with the tile copies done by the TMA engine ("m4"; same arithmetic as transeq_m3.cu).
//
// transeq_m3.cu moves its tiles with per-thread cp.async / ld.shared + st.global: at 512^3 a quarter of the issued
// instructions and most long-scoreboard stalls are copy work (profiles/r01_ncu_full_transeq_m3_*.txt). Here one
// thread per CTA issues three tensor loads and three tensor stores per tile and the 128 threads only do FP64 work:
//
//  * the field is described to the TMA as a 4-D tensor (lane 0..31, segment q, row-in-segment k, line group) with
//    byte strides (8, 16*256, 256, n_pad*256); a box (L lanes, all nseg segments, 16 rows, 1 group) is one whole
//    tile, and because q precedes k in the dimension order it lands in shared memory as [k][q][L]: row k of every
//    (lane, segment) is the contiguous vector  k * NT + threadIdx  (NT = L * nseg threads), so the column accesses
//    of the sweeps are conflict-free without pad rows and every offset is an immediate;
//  * loads complete on an mbarrier (expect_tx = 3 tiles), stores are bulk-async groups; a buffer is reloaded as
//    soon as its stores have read it (cp.async.bulk.wait_group.read), while the other buffer is being computed;
//  * two CTAs x two buffers per SM as before (104 KB each).
// Periodic, uniform, single-rank directions with n a power of two (64..2048) and the compact6 tap masks; everything
// else uses transeq_m3.cu.
#include "m4_common.cuh"

using namespace m4;

namespace {

struct Params4 {
  CUtensorMap in[3];   // in[0] is the line-aligned velocity (conv)
  CUtensorMap out[3];
  int tiles;
  Op o_du, o_dud, o_d2u;  // scaled by -1/2, -1/2, nu
};

// One velocity component of one tile. fF / fC: offsets of the field and conv tiles ([16][NT] each); cz: offset of the
// carry arrays ze[3][NT], ys[3][NT].
template <int L, int NT, bool SELF>
__device__ __forceinline__ void component4(const int fF, const int fC, const int cz, const Params4& p, const int q,
                                           const int l, const int bm, const int b0, const int bp) {
  constexpr int nseg = NT / L;
  double z1[S], z2[S], z3[S];
  {
    double wf[9], wp[9];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int o = woff4<NT>(t, bm, b0, bp);
      wf[t] = smem4[fF + o];
      wp[t] = wf[t] * (SELF ? wf[t] : smem4[fC + o]);
    }
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const int o = woff4<NT>(k + 8, bm, b0, bp);
      wf[8] = smem4[fF + o];
      wp[8] = wf[8] * (SELF ? wf[8] : smem4[fC + o]);
      p1 = fma(p.o_du.a, p1, sten<0x6Cu>(p.o_du.cfw, wf));
      p2 = fma(p.o_dud.a, p2, sten<0x6Cu>(p.o_dud.cfw, wp));
      p3 = fma(p.o_d2u.a, p3, sten<0x7Cu>(p.o_d2u.cfw, wf));
      z1[k] = p1; z2[k] = p2; z3[k] = p3;
#pragma unroll
      for (int t = 0; t < 8; ++t) { wf[t] = wf[t + 1]; wp[t] = wp[t + 1]; }
    }
  }
  smem4[cz + 0 * NT + b0] = z1[S - 1];
  smem4[cz + 1 * NT + b0] = z2[S - 1];
  smem4[cz + 2 * NT + b0] = z3[S - 1];
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
  smem4[cz + 3 * NT + b0] = z1[0];
  smem4[cz + 4 * NT + b0] = z2[0];
  smem4[cz + 5 * NT + b0] = z3[0];
  __syncthreads();
  // carries(): the shared array of m3_common.cuh is the same dynamic shared memory as smem4
  {
    double zi, yi;
    carries<L, false>(cz + 2 * NT + l, cz + 5 * NT + l, L, 0, 0, p.o_d2u, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z3[k] = fma(p.o_d2u.Cp[k], yi, fma(p.o_d2u.W[k], zi, z3[k]));
    carries<L, false>(cz + 1 * NT + l, cz + 4 * NT + l, L, 0, 0, p.o_dud, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z2[k] = fma(p.o_dud.Cp[k], yi, fma(p.o_dud.W[k], zi, z2[k])) + z3[k];
  }
  {
    double zi, yi;
    carries<L, false>(cz + 0 * NT + l, cz + 3 * NT + l, L, 0, 0, p.o_du, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const double du = fma(p.o_du.Cp[k], yi, fma(p.o_du.W[k], zi, z1[k]));
      smem4[fF + b0 + k * NT] = fma(smem4[fC + b0 + k * NT], du, z2[k]);
    }
  }
  fence_async_smem();  // F was written through the generic proxy, the TMA store reads through the async proxy
  __syncthreads();     // the carries are overwritten by the next component; F is complete
}

template <int L, int NT>
__global__ void __launch_bounds__(NT, 1) transeq_m4_kernel(const __grid_constant__ Params4 p) {
  constexpr int nseg = NT / L, fd = S * NT, tpg = SZ / L;
  constexpr int cz = 6 * fd;                       // carries
  constexpr unsigned tile_bytes = fd * sizeof(double);
  const int tid = threadIdx.x, l = tid & (L - 1), q = tid / L;
  const int b0 = tid, bm = tid - L + (q == 0 ? NT : 0), bp = tid + L - (q == nseg - 1 ? NT : 0);
  const unsigned bar0 = saddr(smem4 + cz + 6 * NT), bar1 = bar0 + 8;
  auto issue_loads = [&](int buf, int tile) {
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    const unsigned bar = buf ? bar1 : bar0;
    mbar_expect_tx(bar, 3 * tile_bytes);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load_4d(saddr(smem4 + (3 * buf + f) * fd), &p.in[f], bar, l0, 0, 0, grp);
  };
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    issue_loads(0, blockIdx.x);
    if ((int)(blockIdx.x + gridDim.x) < p.tiles) issue_loads(1, blockIdx.x + gridDim.x);
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    mbar_wait(buf ? bar1 : bar0, (it >> 1) & 1);
    const int bo = buf * 3 * fd;
    // components 1 and 2 first: they read the aligned velocity (field 0) as conv; field 0 is overwritten last.
    // Each result is stored as soon as its component is complete.
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    auto store_field = [&](int f) {
      if (tid == 0) {
        tma_store_4d(&p.out[f], saddr(smem4 + bo + f * fd), l0, 0, 0, grp);
        tma_commit();
      }
    };
    component4<L, NT, false>(bo + 1 * fd, bo, cz, p, q, l, bm, b0, bp);
    store_field(1);
    component4<L, NT, false>(bo + 2 * fd, bo, cz, p, q, l, bm, b0, bp);
    store_field(2);
    component4<L, NT, true>(bo, bo, cz, p, q, l, bm, b0, bp);
    store_field(0);
    if (tid == 0) {
      const int nn = tile + 2 * gridDim.x;
      if (nn < p.tiles) {
        tma_wait_read();  // the stores have read this buffer
        issue_loads(buf, nn);
      }
    }
  }
  if (tid == 0) tma_wait_all();
}

// ---------------------------------------------------------------------------------------------- host side
template <int L, int NT>
int launch4(x3d2c_ctx* ctx, const Params4& p) {
  constexpr size_t smem = sizeof(double) * (6 * S * NT + 6 * NT) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m4_kernel<L, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    attr_set = true;
  }
  int grid = num_sms(ctx) * (256 / NT);
  if (grid > p.tiles) grid = p.tiles;
  transeq_m4_kernel<L, NT><<<grid, NT, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

}  // namespace

namespace x3d2c {

int transeq_m4(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym) {
  static const bool disabled = std::getenv("X3D2C_NO_TMA") != nullptr;
  if (disabled || ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist) return X3D2C_EUNSUPPORTED;
  if (!same_tables(der1st, der1st_sym) || !same_tables(der2nd, der2nd_sym)) return X3D2C_EUNSUPPORTED;
  if (der1st->tap_mask != 0x6Cu || der2nd->tap_mask != 0x7Cu) return X3D2C_EUNSUPPORTED;
  const int n = der1st->n_tds, nseg = n / S;
  int L = 0, NT = 0;
  if (!tile_shape(n, &L, &NT)) return X3D2C_EUNSUPPORTED;
  Params4 p{};
  if (!make_op(der1st, -0.5, false, &p.o_du) || !make_op(der1st, -0.5, false, &p.o_dud) ||
      !make_op(der2nd, nu, false, &p.o_d2u))
    return X3D2C_EUNSUPPORTED;
  const double* in[3];
  double* out[3];
  if (dir == X3D2C_DIR_X) { out[0] = du; out[1] = dv; out[2] = dw; in[0] = u; in[1] = v; in[2] = w; }
  else if (dir == X3D2C_DIR_Y) { out[0] = dv; out[1] = du; out[2] = dw; in[0] = v; in[1] = u; in[2] = w; }
  else { out[0] = dw; out[1] = du; out[2] = dv; in[0] = w; in[1] = u; in[2] = v; }
  const int G = ctx->n_groups[dir], n_pad = ctx->n_pad(dir);
  for (int f = 0; f < 3; ++f)
    if (!make_line_map(&p.in[f], in[f], L, nseg, n_pad, G) || !make_line_map(&p.out[f], out[f], L, nseg, n_pad, G))
      return X3D2C_EUNSUPPORTED;
  p.tiles = G * (SZ / L);
  if (L == 32) return launch4<32, 128>(ctx, p);
  if (L == 16) return launch4<16, 128>(ctx, p);
  if (L == 8) return launch4<8, 128>(ctx, p);
  if (NT == 128) return launch4<4, 128>(ctx, p);
  return launch4<4, 256>(ctx, p);
}

}  // namespace x3d2c
