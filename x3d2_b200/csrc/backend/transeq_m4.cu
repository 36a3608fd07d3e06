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
//  * rank-split directions (DIST): the halo rows are staged with cp.async by all threads and complete on the same
//    mbarrier as the tensor loads; the first / last segment's threads read the rows beyond the line from that staging
//    area. The boundary carries are exchanged with the neighbouring ranks from inside the kernel, component by
//    component (m3_common.cuh: InlineCarries); without peer-mapped buffers they come from the edge kernel
//    (m3_edge.cu) and are staged like the halos.
// Periodic, uniform directions with n a power of two (64..1024) and the compact6 tap masks; everything else uses
// transeq_m3.cu.
#include <type_traits>

#include "m4_common.cuh"

using namespace m4;

namespace {

constexpr int NS = 9;  // recurrences per tile: 3 fields x (d f, d(f conv), d2 f)

struct Params4 {
  CUtensorMap in[3];   // in[0] is the line-aligned velocity (conv)
  CUtensorMap out[3];
  int tiles, nb;       // nb: XTIN only, line groups per z plane (tile coordinates of the swizzled input maps)
  Op o_du, o_dud, o_d2u;  // du, dud scaled by -1/2; d2u unscaled (reference-order stencil)
  double d2u_scale;       // nu * fw of the second derivative
  // rank-split direction only: received halos (SZ, 4, 3, G) and carries (SZ, 3, 9, G)
  const double *halo_s, *halo_e, *from_prev, *from_next;
  InlineCarries inl;  // from_prev != null: the carries are exchanged inside this kernel
};

// Halo staging (DIST): group j = (buffer * 3 + field) * 2 + side holds 4 rows of L lanes with row stride NT, so that
// the window offsets of woff4 apply unchanged; nseg groups share one block of 4 x NT doubles.
template <int L, int NT>
__device__ __forceinline__ constexpr int stage_off(int j) {
  return (j / (NT / L)) * 4 * NT + (j % (NT / L)) * L;
}
template <int L, int NT>
struct Stage { static constexpr int doubles = ((12 + NT / L - 1) / (NT / L)) * 4 * NT; };

// One velocity component of one tile. fF / fC: offsets of the field and conv tiles ([16][NT] each); cz: offset of the
// carry arrays ze[3][NT], ys[3][NT]. oFm / oCm: base of the four rows before the own segment (field / conv), oFp / oCp:
// base of the four rows after it; xp / xn: neighbouring ranks' carries of this field's recurrences (DIST).
// XTIN: the input tiles are 128-byte-swizzled [segment][lane][16] boxes of fields kept in the x layout (y lines reading
// the velocity without an x2y reorder; see tds_m4.cu "XT"). Row r = threadIdx.x holds the thread's own points; element kk
// of row rr sits at rr * 16 + (((kk >> 1) ^ (rr & 7)) << 1) + (kk & 1). The results are written in the direction's own
// layout as usual.
template <int L, int NT>
__device__ __forceinline__ int sw_off(const int t, const int r, const int q) {  // window element t of thread r
  constexpr int nseg = NT / L;
  const int rr = t < 4 ? (q == 0 ? r + NT - L : r - L) : (t < S + 4 ? r : (q == nseg - 1 ? r - (NT - L) : r + L));
  const int kk = t < 4 ? S - 4 + t : (t < S + 4 ? t - 4 : t - S - 4);
  return rr * 16 + (((kk >> 1) ^ (rr & 7)) << 1) + (kk & 1);
}

template <int L, int NT, bool SELF, bool DIST, bool XTIN = false>
__device__ __forceinline__ void component4(const int fF, const int fC, const int cz, const Params4& p, const int q,
                                           const int l, const int b0, const int oFm, const int oCm, const int oFp,
                                           const int oCp, const int xp, const int xn) {
  constexpr int nseg = NT / L;
  double z1[S], z2[S], z3[S];
  {
    double wf[9], wp[9];
    auto load = [&](int t, double& f, double& pr) {  // window element t: row j0 - 4 + t
      if (XTIN) {
        const int o = sw_off<L, NT>(t, b0, q);
        f = smem4[fF + o];
        pr = f * (SELF ? f : smem4[fC + o]);
        return;
      }
      const int oF = t < 4 ? oFm + t * NT : (t < S + 4 ? fF + b0 + (t - 4) * NT : oFp + (t - S - 4) * NT);
      const int oC = t < 4 ? oCm + t * NT : (t < S + 4 ? fC + b0 + (t - 4) * NT : oCp + (t - S - 4) * NT);
      f = smem4[oF];
      pr = f * (SELF ? f : smem4[oC]);
    };
#pragma unroll
    for (int t = 0; t < 8; ++t) load(t, wf[t], wp[t]);
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      load(k + 8, wf[8], wp[8]);
      p1 = fma(p.o_du.a, p1, sten<0x6Cu>(p.o_du.cfw, wf));
      p2 = fma(p.o_dud.a, p2, sten<0x6Cu>(p.o_dud.cfw, wp));
      p3 = fma(p.o_d2u.a, p3, sten_exact_sym<0x7Cu>(p.o_d2u.cfw, wf));  // unscaled: nu * fw is applied below
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
  constexpr int xs = EXP_ROWS * L;

  {
    double zi, yi;
    carries<L, DIST>(cz + 2 * NT + l, cz + 5 * NT + l, L, xp + 2 * xs, xn + 2 * xs, p.o_d2u, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z3[k] = fma(p.o_d2u.Cp[k], yi, fma(p.o_d2u.W[k], zi, z3[k]));
    carries<L, DIST>(cz + 1 * NT + l, cz + 4 * NT + l, L, xp + 1 * xs, xn + 1 * xs, p.o_dud, q, nseg, zi, yi);
#pragma unroll
    for (int k = 0; k < S; ++k) z2[k] = fma(p.d2u_scale, z3[k], fma(p.o_dud.Cp[k], yi, fma(p.o_dud.W[k], zi, z2[k])));
  }
  {
    double zi, yi;
    carries<L, DIST>(cz + 0 * NT + l, cz + 3 * NT + l, L, xp, xn, p.o_du, q, nseg, zi, yi);
    if (XTIN) {
      // conv comes from the swizzled tile; the result goes to the same memory in the direction's own layout, so for the
      // line-aligned component (conv == the field itself) every thread must have read before anyone writes
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const double du = fma(p.o_du.Cp[k], yi, fma(p.o_du.W[k], zi, z1[k]));
        z2[k] = fma(smem4[fC + sw_off<L, NT>(k + 4, b0, q)], du, z2[k]);
      }
      if (SELF) __syncthreads();
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[fF + b0 + k * NT] = z2[k];
    } else {
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const double du = fma(p.o_du.Cp[k], yi, fma(p.o_du.W[k], zi, z1[k]));
        smem4[fF + b0 + k * NT] = fma(smem4[fC + b0 + k * NT], du, z2[k]);
      }
    }
  }
  fence_async_smem();  // F was written through the generic proxy, the TMA store reads through the async proxy
  __syncthreads();     // the carries are overwritten by the next component; F is complete
}

// shared memory: [2 buffers][3 fields][16][NT] | carries ze[3][NT], ys[3][NT] | 2 mbarriers (16 B)
//                DIST: | halo staging (Stage::doubles) | neighbour carries [2 buffers][prev 27 rows | next 27 rows][L]
template <int L, int NT, bool DIST, bool XTIN = false>
__global__ void __launch_bounds__(NT, 1) transeq_m4_kernel(const __grid_constant__ Params4 p) {
  static_assert(!(DIST && XTIN), "swizzled inputs are not combined with rank-split lines");
  constexpr int nseg = NT / L, fd = S * NT, tpg = SZ / L, cpr = L / 2;
  constexpr int cz = 6 * fd;                           // carries
  constexpr int hs0 = cz + 6 * NT + 2;                 // halo staging (after the two mbarriers)
  constexpr int xb0 = hs0 + Stage<L, NT>::doubles;    // neighbour carries
  constexpr int xbuf = 2 * NS * EXP_ROWS * L;
  constexpr unsigned tile_bytes = fd * sizeof(double);
  const int tid = threadIdx.x, l = tid & (L - 1), q = tid / L;
  const int b0 = tid, bm = tid - L + (q == 0 ? NT : 0), bp = tid + L - (q == nseg - 1 ? NT : 0);
  const unsigned bar0 = saddr(smem4 + cz + 6 * NT), bar1 = bar0 + 8;
  auto issue_loads = [&](int buf, int tile) {  // thread 0
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    const unsigned bar = buf ? bar1 : bar0;
    mbar_expect_tx(bar, 3 * tile_bytes);
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (XTIN) {  // (0, x of the first lane, 0, 0, z)
        const int c4 = grp / p.nb, c3 = grp - c4 * p.nb;
        tma_load_5d(saddr(smem4 + (3 * buf + f) * fd), &p.in[f], bar, 0, SZ * c3 + l0, 0, 0, c4);
      } else {
        tma_load_4d(saddr(smem4 + (3 * buf + f) * fd), &p.in[f], bar, l0, 0, 0, grp);
      }
    }
  };
  auto stage_neighbours = [&](int buf, int tile) {  // all threads: halo rows and neighbour carries of one tile
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    for (int idx = tid; idx < (24 + 2 * NS * EXP_ROWS) * cpr; idx += NT) {
      const int row = idx / cpr, c = 2 * (idx - row * cpr);
      if (row < 24) {
        const int f = row >> 3, side = (row >> 2) & 1, r = row & 3;
        const double* src = (side ? p.halo_e : p.halo_s) + ((size_t)(grp * 3 + f) * 4 + r) * SZ + l0 + c;
        cp_async16(smem4 + hs0 + stage_off<L, NT>((buf * 3 + f) * 2 + side) + r * NT + c, src);
      } else {
        const int e = row - 24, second = e >= NS * EXP_ROWS, rr = e - second * NS * EXP_ROWS;
        const double* src = (second ? p.from_next : p.from_prev) + ((size_t)grp * NS * EXP_ROWS + rr) * SZ + l0 + c;
        cp_async16(smem4 + xb0 + buf * xbuf + e * L + c, src);
      }
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(buf ? bar1 : bar0) : "memory");
  };
  if (tid == 0) {
    mbar_init(bar0, DIST ? 1 + NT : 1);
    mbar_init(bar1, DIST ? 1 + NT : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {
    const int t1 = blockIdx.x + gridDim.x;
    if (tid == 0) {
      issue_loads(0, blockIdx.x);
      if (t1 < p.tiles) issue_loads(1, t1);
    }
    if (DIST) {
      stage_neighbours(0, blockIdx.x);
      if (t1 < p.tiles) stage_neighbours(1, t1);
    }
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
    // rows before / after the own segment: the neighbouring segment, or (first / last segment of a rank-split line)
    // the staged halo rows
    auto before = [&](int f) {
      return (DIST && q == 0) ? hs0 + stage_off<L, NT>((buf * 3 + f) * 2) + l : bo + f * fd + bm + (S - 4) * NT;
    };
    auto after = [&](int f) {
      return (DIST && q == nseg - 1) ? hs0 + stage_off<L, NT>((buf * 3 + f) * 2 + 1) + l : bo + f * fd + bp;
    };
    const int xp = xb0 + buf * xbuf + l, xn = xp + NS * EXP_ROWS * L;
    constexpr int xf = 3 * EXP_ROWS * L;  // three recurrences per field
    const int m0 = before(0), a0 = after(0);
    component4<L, NT, false, DIST, XTIN>(bo + 1 * fd, bo, cz, p, q, l, b0, before(1), m0, after(1), a0, xp + xf, xn + xf);
    store_field(1);
    component4<L, NT, false, DIST, XTIN>(bo + 2 * fd, bo, cz, p, q, l, b0, before(2), m0, after(2), a0, xp + 2 * xf,
                                         xn + 2 * xf);
    store_field(2);
    component4<L, NT, true, DIST, XTIN>(bo, bo, cz, p, q, l, b0, m0, m0, a0, a0, xp, xn);
    store_field(0);
    const int nn = tile + 2 * gridDim.x;
    if (nn < p.tiles) {
      if (DIST) stage_neighbours(buf, nn);  // the staging areas of this buffer are free (barrier of component 3)
      if (tid == 0) {
        tma_wait_read();  // the stores have read this buffer
        issue_loads(buf, nn);
      }
    }
  }
  if (tid == 0) tma_wait_all();
}

// ---------------------------------------------------------------------------------------------- host side
template <int L, int NT, bool DIST, bool XTIN = false>
int launch4(x3d2c_ctx* ctx, const Params4& p) {
  constexpr size_t smem = sizeof(double) * (6 * S * NT + 6 * NT + 2 +
                                            (DIST ? Stage<L, NT>::doubles + 2 * 2 * NS * EXP_ROWS * L : 0));
  static int per_sm_dev[x3d2c::kMaxDevices] = {};  // once per device
  int& per_sm = per_sm_dev[ctx->device];
  if (!per_sm) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m4_kernel<L, NT, DIST, XTIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    X3D2C_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transeq_m4_kernel<L, NT, DIST, XTIN>, NT, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  transeq_m4_kernel<L, NT, DIST, XTIN><<<grid, NT, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}


// ---- rank-split direction with the carry exchange inside the kernel (m3_common.cuh: InlineCarries) -----------------
// Same tiles and arithmetic as above. A field is finished with the carries of this rank's own segments as soon as its
// sweeps are done, and its boundary carries are pushed to the neighbours at the same time; what the neighbouring ranks'
// segments add (a linear term that only reaches the first / last three segments) is applied
// one tile later, when those carries have had at least a field's worth of time to cross NVLink:
//   tile i:  field 1: sweeps, push, finish -> F1;  [tile i - 1: poll, correct its edge segments, store, reload];
//            field 2 -> F2;  field 0 -> F0 (the conv rows of the edge segments are kept aside for the correction).

// One field without the neighbouring ranks' terms: result in out[] (not stored: the caller decides where it lives)
template <int L, int NT, bool SELF>
__device__ __forceinline__ void component4i(const int fF, const int fC, const int cz, const Params4& p, const int q,
                                            const int l, const int b0, const int oFm, const int oCm, const int oFp,
                                            const int oCp, const InlineCarries& inl, const size_t slot, double (&out)[S]) {
  constexpr int nseg = NT / L;
  double z1[S], z3[S];
  {
    double wf[9], wp[9];
    auto load = [&](int t, double& f, double& pr) {  // window element t: row j0 - 4 + t
      const int oF = t < 4 ? oFm + t * NT : (t < S + 4 ? fF + b0 + (t - 4) * NT : oFp + (t - S - 4) * NT);
      const int oC = t < 4 ? oCm + t * NT : (t < S + 4 ? fC + b0 + (t - 4) * NT : oCp + (t - S - 4) * NT);
      f = smem4[oF];
      pr = f * (SELF ? f : smem4[oC]);
    };
#pragma unroll
    for (int t = 0; t < 8; ++t) load(t, wf[t], wp[t]);
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      load(k + 8, wf[8], wp[8]);
      p1 = fma(p.o_du.a, p1, sten<0x6Cu>(p.o_du.cfw, wf));
      p2 = fma(p.o_dud.a, p2, sten<0x6Cu>(p.o_dud.cfw, wp));
      p3 = fma(p.o_d2u.a, p3, sten_exact_sym<0x7Cu>(p.o_d2u.cfw, wf));
      z1[k] = p1; out[k] = p2; z3[k] = p3;
#pragma unroll
      for (int t = 0; t < 8; ++t) { wf[t] = wf[t + 1]; wp[t] = wp[t + 1]; }
    }
  }
  smem4[cz + 0 * NT + b0] = z1[S - 1];
  smem4[cz + 1 * NT + b0] = out[S - 1];
  smem4[cz + 2 * NT + b0] = z3[S - 1];
  {
    double y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
    for (int k = S - 1; k >= 0; --k) {
      y1 = fma(p.o_du.cb, y1, z1[k]);
      y2 = fma(p.o_dud.cb, y2, out[k]);
      y3 = fma(p.o_d2u.cb, y3, z3[k]);
      z1[k] = y1; out[k] = y2; z3[k] = y3;
    }
  }
  smem4[cz + 3 * NT + b0] = z1[0];
  smem4[cz + 4 * NT + b0] = out[0];
  smem4[cz + 5 * NT + b0] = z3[0];
  __syncthreads();
  constexpr size_t ss = (size_t)EXP_ROWS * SZ;
  push_carries_inline<L>(inl, slot, cz + l, cz + 3 * NT + l, p.o_du, q, nseg);
  push_carries_inline<L>(inl, slot + ss, cz + 1 * NT + l, cz + 4 * NT + l, p.o_dud, q, nseg);
  push_carries_inline<L>(inl, slot + 2 * ss, cz + 2 * NT + l, cz + 5 * NT + l, p.o_d2u, q, nseg);
  double zi, yi;
  carries<L, true, false>(cz + 2 * NT + l, cz + 5 * NT + l, L, 0, 0, p.o_d2u, q, nseg, zi, yi);
#pragma unroll
  for (int k = 0; k < S; ++k) z3[k] = fma(p.o_d2u.Cp[k], yi, fma(p.o_d2u.W[k], zi, z3[k]));
  carries<L, true, false>(cz + 1 * NT + l, cz + 4 * NT + l, L, 0, 0, p.o_dud, q, nseg, zi, yi);
#pragma unroll
  for (int k = 0; k < S; ++k) out[k] = fma(p.d2u_scale, z3[k], fma(p.o_dud.Cp[k], yi, fma(p.o_dud.W[k], zi, out[k])));
  carries<L, true, false>(cz + l, cz + 3 * NT + l, L, 0, 0, p.o_du, q, nseg, zi, yi);
#pragma unroll
  for (int k = 0; k < S; ++k) {
    const double du = fma(p.o_du.Cp[k], yi, fma(p.o_du.W[k], zi, z1[k]));
    out[k] = fma(smem4[fC + b0 + k * NT], du, out[k]);
  }
}

// what the neighbouring ranks' carries add to one field: v[k] += conv[k] * c_du[k] + c_dud[k] + nu c_d2u[k]
// (xp / xn: the field's rows of the polled carries, lane applied). Only the first / last DMAX segments are affected.
template <int L, int NT>
__device__ __forceinline__ void ext_terms4(const Params4& p, const int q, const int xp, const int xn, double (&zi)[3],
                                           double (&yi)[3]) {
  constexpr int nseg = NT / L, xs = EXP_ROWS * L;
  carries_ext<L>(xp, xn, p.o_du, q, nseg, zi[0], yi[0]);
  carries_ext<L>(xp + xs, xn + xs, p.o_dud, q, nseg, zi[1], yi[1]);
  carries_ext<L>(xp + 2 * xs, xn + 2 * xs, p.o_d2u, q, nseg, zi[2], yi[2]);
}
__device__ __forceinline__ double ext_term4(const Params4& p, const int k, const double conv, const double (&zi)[3],
                                            const double (&yi)[3]) {
  const double c1 = fma(p.o_du.Cp[k], yi[0], p.o_du.W[k] * zi[0]);
  const double c2 = fma(p.o_dud.Cp[k], yi[1], p.o_dud.W[k] * zi[1]);
  const double c3 = fma(p.o_d2u.Cp[k], yi[2], p.o_d2u.W[k] * zi[2]);
  return fma(conv, c1, fma(p.d2u_scale, c3, c2));
}

// shared memory: [2 buffers][3 fields][16][NT] | carries ze[3][NT], ys[3][NT] | 2 mbarriers |
//                halo staging (Stage::doubles) | neighbour carries [prev 27 rows | next 27 rows][L] |
//                conv rows of the six edge segments [16][6 L]
template <int L, int NT>
__global__ void __launch_bounds__(NT, 1) transeq_m4i_kernel(const __grid_constant__ Params4 p) {
  constexpr int nseg = NT / L, fd = S * NT, tpg = SZ / L, cpr = L / 2;
  constexpr int cz = 6 * fd;                           // carries
  constexpr int hs0 = cz + 6 * NT + 2;                 // halo staging (after the two mbarriers)
  constexpr int xb0 = hs0 + Stage<L, NT>::doubles;    // neighbour carries
  constexpr int xs = EXP_ROWS * L, xf = 3 * xs;        // per recurrence / per field
  constexpr int cv0 = xb0 + 2 * NS * xs;  // conv stash (one tile: written at the end of a tile, read in the middle of the next)
  constexpr unsigned tile_bytes = fd * sizeof(double);
  constexpr size_t ss = (size_t)EXP_ROWS * SZ;
  const int tid = threadIdx.x, l = tid & (L - 1), q = tid / L;
  const int b0 = tid, bm = tid - L + (q == 0 ? NT : 0), bp = tid + L - (q == nseg - 1 ? NT : 0);
  const bool edge = q < DMAX || q >= nseg - DMAX;
  const int es = (q < DMAX ? q : q - (nseg - 2 * DMAX)) * L + l;  // edge threads: column of the conv stash
  const int xp = xb0 + l, xn = xp + NS * xs;
  const unsigned bar0 = saddr(smem4 + cz + 6 * NT), bar1 = bar0 + 8;
  auto issue_loads = [&](int buf, int tile) {  // thread 0
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    const unsigned bar = buf ? bar1 : bar0;
    mbar_expect_tx(bar, 3 * tile_bytes);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load_4d(saddr(smem4 + (3 * buf + f) * fd), &p.in[f], bar, l0, 0, 0, grp);
  };
  auto stage_halos = [&](int buf, int tile) {  // all threads
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    for (int idx = tid; idx < 24 * cpr; idx += NT) {
      const int row = idx / cpr, c = 2 * (idx - row * cpr);
      const int f = row >> 3, side = (row >> 2) & 1, r = row & 3;
      const double* src = (side ? p.halo_e : p.halo_s) + ((size_t)(grp * 3 + f) * 4 + r) * SZ + l0 + c;
      cp_async16(smem4 + hs0 + stage_off<L, NT>((buf * 3 + f) * 2 + side) + r * NT + c, src);
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(buf ? bar1 : bar0) : "memory");
  };
  // The neighbouring ranks' share of a tile whose fields were finished without it (buffer pbuf), then store the tile
  // and reload the buffer with the tile after the next.
  auto finish = [&](int ptile, int pbuf) {
    const int grp = ptile / tpg, l0 = (ptile - grp * tpg) * L, bo = pbuf * 3 * fd;
    const size_t slot = (size_t)grp * NS * EXP_ROWS * SZ + l0 + l;
    poll_carries_inline<L, NS>(p.inl, slot, ss, q, nseg, xp, xn, xs);
    __syncthreads();
    if (edge) {
#pragma unroll 1
      for (int f = 0; f < 3; ++f) {
        double zi[3], yi[3];
        ext_terms4<L, NT>(p, q, xp + f * xf, xn + f * xf, zi, yi);
#pragma unroll
        for (int k = 0; k < S; ++k)
          smem4[bo + f * fd + b0 + k * NT] += ext_term4(p, k, smem4[cv0 + k * 2 * DMAX * L + es], zi, yi);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int f = 0; f < 3; ++f) tma_store_4d(&p.out[f], saddr(smem4 + bo + f * fd), l0, 0, 0, grp);
      tma_commit();
    }
    const int nn = ptile + 2 * gridDim.x;
    if (nn < p.tiles) {
      stage_halos(pbuf, nn);
      if (tid == 0) {
        tma_wait_read();
        issue_loads(pbuf, nn);
      }
    }
  };
  if (tid == 0) {
    mbar_init(bar0, 1 + NT);
    mbar_init(bar1, 1 + NT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {
    const int t1 = blockIdx.x + gridDim.x;
    if (tid == 0) {
      issue_loads(0, blockIdx.x);
      if (t1 < p.tiles) issue_loads(1, t1);
    }
    stage_halos(0, blockIdx.x);
    if (t1 < p.tiles) stage_halos(1, t1);
  }
  int it = 0, prev = -1;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    mbar_wait(buf ? bar1 : bar0, (it >> 1) & 1);
    const int bo = buf * 3 * fd;
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    auto before = [&](int f) { return q == 0 ? hs0 + stage_off<L, NT>((buf * 3 + f) * 2) + l : bo + f * fd + bm + (S - 4) * NT; };
    auto after = [&](int f) { return q == nseg - 1 ? hs0 + stage_off<L, NT>((buf * 3 + f) * 2 + 1) + l : bo + f * fd + bp; };
    const size_t slot = (size_t)grp * NS * EXP_ROWS * SZ + l0 + l;
    const int m0 = before(0), a0 = after(0);
#pragma unroll 1
    for (int f = 1; f <= 2; ++f) {
      double v[S];
      component4i<L, NT, false>(bo + f * fd, bo, cz, p, q, l, b0, before(f), m0, after(f), a0, p.inl, slot + 3 * f * ss, v);
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[bo + f * fd + b0 + k * NT] = v[k];
      __syncthreads();  // the carry arrays are free for the next field
      // the previous tile: its last carries left the neighbours a whole field ago
      if (f == 1 && prev >= 0) finish(prev, buf ^ 1);
    }
    {
      double v[S];
      component4i<L, NT, true>(bo, bo, cz, p, q, l, b0, m0, m0, a0, a0, p.inl, slot, v);
      if (edge) {
#pragma unroll
        for (int k = 0; k < S; ++k) smem4[cv0 + k * 2 * DMAX * L + es] = smem4[bo + b0 + k * NT];
      }
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[bo + b0 + k * NT] = v[k];
      __syncthreads();
    }
    prev = tile;
  }
  if (prev >= 0) finish(prev, (it - 1) & 1);
  if (tid == 0) tma_wait_all();
}

template <int L, int NT>
int launch4i(x3d2c_ctx* ctx, const Params4& p) {
  constexpr size_t smem = sizeof(double) * (6 * S * NT + 6 * NT + 2 + Stage<L, NT>::doubles + 2 * NS * EXP_ROWS * L +
                                            S * 2 * DMAX * L);
  static int per_sm_dev[x3d2c::kMaxDevices] = {};  // once per device
  int& per_sm = per_sm_dev[ctx->device];
  if (!per_sm) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(transeq_m4i_kernel<L, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    X3D2C_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transeq_m4i_kernel<L, NT>, NT, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  transeq_m4i_kernel<L, NT><<<grid, NT, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

// rank-split lines want smaller tiles: the staging areas must fit next to two (or three) CTAs' buffers
bool dist_shape(int n, int* L, int* NT) {
  switch (n) {
    case 128: *L = 8; *NT = 64; return true;
    case 256: *L = 4; *NT = 64; return true;
    case 512: *L = 4; *NT = 128; return true;
    case 1024: *L = 4; *NT = 256; return true;
    default: return false;
  }
}

}  // namespace

namespace x3d2c {

// lay_in: the layout u, v, w are stored in; != dir only for y lines reading the x layout (swizzled input tiles).
// dry_run: all checks, no launch.
int transeq_m4(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym, int lay_in, bool dry_run) {
  static const bool disabled = std::getenv("X3D2C_NO_TMA") != nullptr;
  if (disabled) return X3D2C_EUNSUPPORTED;
  if (!same_tables(der1st, der1st_sym) || !same_tables(der2nd, der2nd_sym)) return X3D2C_EUNSUPPORTED;
  if (der1st->tap_mask != 0x6Cu || der2nd->tap_mask != 0x7Cu) return X3D2C_EUNSUPPORTED;
  const int n = der1st->n_tds, nseg = n / S;
  const bool split = ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist;
  if (split && !dist_supported(ctx, dir, n)) return X3D2C_EUNSUPPORTED;
  const bool xtin = lay_in != dir;
  if (xtin && (dir != X3D2C_DIR_Y || lay_in != X3D2C_DIR_X || split)) return X3D2C_EUNSUPPORTED;
  int L = 0, NT = 0;
  if (!(split ? dist_shape(n, &L, &NT) : tile_shape(n, &L, &NT))) return X3D2C_EUNSUPPORTED;
  Params4 p{};
  // du, d(u conv): FMA stencils with -1/2 and fw folded in; d2u: the reference's summation order (sten_exact), its
  // recurrence runs unscaled (the edge kernel uses the same Op) and nu * fw multiplies the finished second derivative
  if (!make_op(der1st, -0.5, split, &p.o_du, false) || !make_op(der1st, -0.5, split, &p.o_dud, false) ||
      !make_op(der2nd, nu, split, &p.o_d2u, true))
    return X3D2C_EUNSUPPORTED;
  for (int k = 0; k < 4; ++k)  // sten_exact_sym shares the products of mirror taps
    if (p.o_d2u.cfw[k] != p.o_d2u.cfw[8 - k]) return X3D2C_EUNSUPPORTED;
  p.d2u_scale = p.o_d2u.fs;
  p.o_d2u.fs = 1.0;
  const double* in[3];
  double* out[3];
  if (dir == X3D2C_DIR_X) { out[0] = du; out[1] = dv; out[2] = dw; in[0] = u; in[1] = v; in[2] = w; }
  else if (dir == X3D2C_DIR_Y) { out[0] = dv; out[1] = du; out[2] = dw; in[0] = v; in[1] = u; in[2] = w; }
  else { out[0] = dw; out[1] = du; out[2] = dv; in[0] = w; in[1] = u; in[2] = v; }
  const int G = ctx->n_groups[dir], n_pad = ctx->n_pad(dir);
  for (int f = 0; f < 3; ++f) {
    if (xtin ? !make_map_xt(&p.in[f], in[f], lay_in, dir, L, nseg, ctx) : !make_line_map(&p.in[f], in[f], L, nseg, n_pad, G))
      return X3D2C_EUNSUPPORTED;
    if (!make_line_map(&p.out[f], out[f], L, nseg, n_pad, G)) return X3D2C_EUNSUPPORTED;
  }
  p.tiles = G * (SZ / L);
  p.nb = ctx->nx_pad / SZ;
  if (dry_run) return X3D2C_OK;  // the call qualifies (x3d2c_transeq_r_fused)
  if (xtin) {
    if (L == 32) return launch4<32, 128, false, true>(ctx, p);
    if (L == 16) return launch4<16, 128, false, true>(ctx, p);
    if (L == 8) return launch4<8, 128, false, true>(ctx, p);
    if (NT == 128) return launch4<4, 128, false, true>(ctx, p);
    return launch4<4, 256, false, true>(ctx, p);
  }
  if (!split) {
    if (L == 32) return launch4<32, 128, false>(ctx, p);
    if (L == 16) return launch4<16, 128, false>(ctx, p);
    if (L == 8) return launch4<8, 128, false>(ctx, p);
    if (NT == 128) return launch4<4, 128, false>(ctx, p);
    return launch4<4, 256, false>(ctx, p);
  }
  // rank-split direction: halos and boundary carries first (m3_edge.cu), then the main kernel
  DistBufs b = carve_dist(ctx);
  EdgeParams ep{};
  ep.n = n;
  ep.n_pad = n_pad;
  ep.nseg = nseg;
  ep.ns = NS;
  ep.transeq = 1;
  ep.ops[0] = p.o_du;
  ep.ops[1] = p.o_dud;
  ep.ops[2] = p.o_d2u;
  for (int f = 0; f < 3; ++f) ep.f[f] = in[f];
  // In-kernel carry exchange for thin slabs only: it costs the main kernel about 40% more instructions (measured:
  // 2.70 vs 3.62 ms at 1024 x 1024 x 128 per rank, but 2.10 vs 1.92 ms at 512^3 per rank, where the edge kernel's six
  // segments are a small share of the 32 of a line)
  static const bool force_inline = std::getenv("X3D2C_INLINE_TRANSEQ") != nullptr;
  int rc = exchange_edges(ctx, dir, in, 3, ep, b, (n <= 256 || force_inline) ? &p.inl : nullptr);
  if (rc) return rc;
  p.halo_s = b.halo_recv_s;
  p.halo_e = b.halo_recv_e;
  p.from_prev = b.carr_from_prev;
  p.from_next = b.carr_from_next;
  if (p.inl.from_prev) {
    if (NT == 64) return L == 8 ? launch4i<8, 64>(ctx, p) : launch4i<4, 64>(ctx, p);
    if (NT == 128) return launch4i<4, 128>(ctx, p);
    return launch4i<4, 256>(ctx, p);
  }
  if (NT == 64) return L == 8 ? launch4<8, 64, true>(ctx, p) : launch4<4, 64, true>(ctx, p);
  if (NT == 128) return launch4<4, 128, true>(ctx, p);
  return launch4<4, 256, true>(ctx, p);
}

}  // namespace x3d2c
