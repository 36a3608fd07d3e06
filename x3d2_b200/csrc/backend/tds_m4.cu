// tds_solve and its fused pairs with the tile copies done by the TMA engine ("m4"; same arithmetic as tds_m3.cu and
// tds_pair_m3.cu, shared-memory layout [row in segment][segment][lane] as in transeq_m4.cu).
//   SINGLE: out   = A(in)                  16 B/pt
//   SUM   : out   = A(in_a) + B(in_b)      24 B/pt
//   DUAL  : out_a = A(in), out_b = B(in)   24 B/pt
//   AXPY  : y     = y + a A(in)            24 B/pt
// Periodic, uniform directions with n = 64..1024 a power of two; other shapes use the cp.async kernels. Rank-split
// directions (DIST) stage the halo rows with cp.async as transeq_m4.cu does and exchange the boundary carries with the
// neighbouring ranks from inside the kernel (m3_common.cuh: InlineCarries; without peer-mapped buffers the carries come
// from the edge kernel and are staged like the halos); their inputs must be in the direction's own layout (the halo pack
// kernel reads them), the outputs may go through a tensor map.
// Measured at 512^3: tds_solve 0.395 -> 0.376 ms (87% of HBM peak) with 64-byte rows (L = 8, 256 threads); with
// 32-byte rows (L = 4, 128 threads) the TMA version is slower (0.52 ms), so narrow tiles stay on the cp.async path.
#include "m4_common.cuh"

using namespace m4;

namespace {

enum Mode { SINGLE = 0, SUM = 1, DUAL = 2, AXPY = 3 };

struct TdsParams4 {
  CUtensorMap in_a, in_b;    // SUM: two inputs; AXPY: in_b = y
  CUtensorMap out_a, out_b;  // DUAL: two outputs; AXPY: out_a = y
  int tiles, nb;  // tile coordinates: (lane0, 0, 0, group % nb, group / nb)
  int n4;         // > 0: tiles are numbered with c4 fastest (extent n4), see tile_coords()
  Op oa, ob;
  // rank-split direction only: received halos (SZ, 4, NF, G) and carries (SZ, 3, NR, G)
  const double *halo_s, *halo_e, *from_prev, *from_next;
  InlineCarries inl;  // from_prev != null: the carries are exchanged inside this kernel
};

// F: offset of the own rows' tile; om / op: base of the four rows before / after the own segment
template <int NT, unsigned M>
__device__ __forceinline__ void local_sweeps4(const int F, const Op& o, const int om, const int b0, const int op,
                                              double (&z)[S], double& ze) {
  auto at = [&](int t) {  // window element t: row j0 - 4 + t
    return smem4[t < 4 ? om + t * NT : (t < S + 4 ? F + b0 + (t - 4) * NT : op + (t - S - 4) * NT)];
  };
  double wf[9];
#pragma unroll
  for (int t = 0; t < 8; ++t) wf[t] = at(t);
  double pz = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    wf[8] = at(k + 8);
    pz = fma(o.a, pz, o.fs * sten_exact<M>(o.cfw, wf));
    z[k] = pz;
#pragma unroll
    for (int t = 0; t < 8; ++t) wf[t] = wf[t + 1];
  }
  ze = pz;
  double y = 0.0;
#pragma unroll
  for (int k = S - 1; k >= 0; --k) {
    y = fma(o.cb, y, z[k]);
    z[k] = y;
  }
}


// ---- x lines of fields kept in an x-fastest layout (XT: reading or writing through an X <-> Y reorder) ------------
// In DIR_Y (and DIR_Z, DIR_C) 16 consecutive x of one (y, z) are 128 contiguous bytes, so a 16-point segment of an x
// line is one row of a 128-byte-swizzled TMA box. The tensor map orders the box as (x in segment, lane y, segment), which
// lands in shared memory as [segment q][lane l][k]: row r = q L + l = threadIdx.x holds the thread's own 16 points, and
// the hardware swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7)) makes the eight threads of a quarter warp hit
// eight different chunks, so 128-bit accesses are conflict-free. This is the 32 x 32 transpose of reorder.cu done by the
// TMA engine on the way in or out: no separate pass, no padded staging tile.
__device__ __forceinline__ double2 lds128(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(unsigned addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
// window (rows j0 - 4 .. j0 + 19) of thread r from a swizzled tile at shared byte address tile
template <int L, int NT>
__device__ __forceinline__ void window_sw(const unsigned tile, const int r, const int q, double (&w)[S + 8]) {
  constexpr int nseg = NT / L;
  const int rp = q == 0 ? r + NT - L : r - L, rn = q == nseg - 1 ? r - (NT - L) : r + L;
  {
    const unsigned row = tile + rp * 128, x = rp & 7;
    const double2 a = lds128(row + ((6 ^ x) << 4)), b = lds128(row + ((7 ^ x) << 4));
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
  }
  {
    const unsigned row = tile + r * 128, x = r & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double2 v = lds128(row + ((c ^ x) << 4));
      w[4 + 2 * c] = v.x;
      w[5 + 2 * c] = v.y;
    }
  }
  {
    const unsigned row = tile + rn * 128, x = rn & 7;
    const double2 a = lds128(row + ((0 ^ x) << 4)), b = lds128(row + ((1 ^ x) << 4));
    w[S + 4] = a.x; w[S + 5] = a.y; w[S + 6] = b.x; w[S + 7] = b.y;
  }
}
// local_sweeps4 on a window held in registers
template <unsigned M>
__device__ __forceinline__ void local_sweeps_w(const Op& o, const double (&w)[S + 8], double (&z)[S], double& ze) {
  double pz = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    double wf[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wf[t] = w[k + t];
    pz = fma(o.a, pz, o.fs * sten_exact<M>(o.cfw, wf));
    z[k] = pz;
  }
  ze = pz;
  double y = 0.0;
#pragma unroll
  for (int k = S - 1; k >= 0; --k) {
    y = fma(o.cb, y, z[k]);
    z[k] = y;
  }
}

// tile -> (first lane, c3, c4). Default order: lane part, then c3, then c4 (tiles that run at the same time are
// neighbours in the direction's own layout). With n4 > 0, c4 runs fastest: for a solve that stores into a foreign
// layout (z lines -> y layout, y lines -> z layout) c4 is the row index of that layout, so the 64-byte pieces written
// at the same time by neighbouring CTAs fall next to each other instead of one plane (2 MB) apart.
template <int L>
__device__ __forceinline__ void tile_coords(const int tile, const int nb, const int n4, int& l0, int& c3, int& c4) {
  constexpr int tpg = SZ / L;
  if (n4 > 0) {
    c4 = tile % n4;
    const int r = tile / n4;
    c3 = r / tpg;
    l0 = (r - c3 * tpg) * L;
  } else {
    const int grp = tile / tpg;
    l0 = (tile - grp * tpg) * L;
    c4 = grp / nb;
    c3 = grp - c4 * nb;
  }
}

template <int L, int NT, int MODE>
struct Shape {
  static constexpr int NSLOT = MODE == SINGLE ? 1 : 2;
  static constexpr int NR = (MODE == SUM || MODE == DUAL) ? 2 : 1;     // recurrences
  static constexpr int NLOAD = (MODE == SUM || MODE == AXPY) ? 2 : 1;  // tiles loaded
  static constexpr int NF = MODE == SUM ? 2 : 1;                       // fields with halos
  static constexpr int nseg = NT / L;
  static constexpr int cz = 2 * NSLOT * S * NT;                        // carries
  static constexpr int hs0 = cz + 2 * NR * NT + 2;                     // halo staging (after the two mbarriers)
  static constexpr int stage = ((4 * NF + nseg - 1) / nseg) * 4 * NT;  // groups (buffer, field, side) of 4 rows x L
  static constexpr int xb0 = hs0 + stage;                              // neighbour carries
  static constexpr int xbuf = 2 * NR * EXP_ROWS * L;
  static constexpr size_t smem(bool dist) { return sizeof(double) * (dist ? xb0 + 2 * xbuf : hs0); }
  static __device__ __forceinline__ int stage_off(int j) { return hs0 + (j / nseg) * 4 * NT + (j % nseg) * L; }
};

// shared memory: [2 buffers][NSLOT tiles of 16 x NT] | NR x (ze, ys)[NT] | 2 mbarriers | DIST: halo staging, carries
// XT (x lines only, not rank-split): 1 = out_a is stored through a swizzled map into an x-fastest layout, 2 = in_a is loaded
// through one (see above); tile coordinates of such a map: (0, y of the first lane, 0, 0, z).
template <int L, int NT, unsigned M, int MODE, bool DIST, int XT = 0>
__global__ void __launch_bounds__(NT, 512 / NT) tds_m4_kernel(const __grid_constant__ TdsParams4 p) {
  using Sh = Shape<L, NT, MODE>;
  constexpr int nseg = Sh::nseg, fd = S * NT, tpg = SZ / L, cpr = L / 2;
  constexpr int NSLOT = Sh::NSLOT, NR = Sh::NR, NLOAD = Sh::NLOAD, NF = Sh::NF, cz = Sh::cz;
  constexpr unsigned tile_bytes = fd * sizeof(double);
  const int tid = threadIdx.x, l = tid & (L - 1), q = tid / L;
  const int b0 = tid, bm = tid - L + (q == 0 ? NT : 0), bp = tid + L - (q == nseg - 1 ? NT : 0);
  const unsigned bar0 = saddr(smem4 + cz + 2 * NR * NT), bar1 = bar0 + 8;
  auto issue_loads = [&](int buf, int tile) {  // thread 0
    int l0, c3, c4;
    tile_coords<L>(tile, p.nb, DIST ? 0 : p.n4, l0, c3, c4);
    const unsigned bar = buf ? bar1 : bar0;
    mbar_expect_tx(bar, NLOAD * tile_bytes);
    if (XT == 2) tma_load_5d(saddr(smem4 + buf * NSLOT * fd), &p.in_a, bar, 0, SZ * c3 + l0, 0, 0, c4);
    else tma_load_5d(saddr(smem4 + buf * NSLOT * fd), &p.in_a, bar, l0, 0, 0, c3, c4);
    if (NLOAD == 2) tma_load_5d(saddr(smem4 + (buf * NSLOT + 1) * fd), &p.in_b, bar, l0, 0, 0, c3, c4);
  };
  auto stage_neighbours = [&](int buf, int tile) {  // all threads
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    for (int idx = tid; idx < (8 * NF + 2 * NR * EXP_ROWS) * cpr; idx += NT) {
      const int row = idx / cpr, c = 2 * (idx - row * cpr);
      if (row < 8 * NF) {
        const int f = row >> 3, side = (row >> 2) & 1, r = row & 3;
        const double* src = (side ? p.halo_e : p.halo_s) + ((size_t)(grp * NF + f) * 4 + r) * SZ + l0 + c;
        cp_async16(smem4 + Sh::stage_off((buf * NF + f) * 2 + side) + r * NT + c, src);
      } else {
        const int e = row - 8 * NF, second = e >= NR * EXP_ROWS, rr = e - second * NR * EXP_ROWS;
        const double* src = (second ? p.from_next : p.from_prev) + ((size_t)grp * NR * EXP_ROWS + rr) * SZ + l0 + c;
        cp_async16(smem4 + Sh::xb0 + buf * Sh::xbuf + e * L + c, src);
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
    const int F0 = buf * NSLOT * fd, F1 = F0 + fd;
    // rows before / after the own segment: the neighbouring segment or the staged halo rows (slot = field index)
    auto before = [&](int F, int f) {
      return (DIST && q == 0) ? Sh::stage_off((buf * NF + f) * 2) + l : F + bm + (S - 4) * NT;
    };
    auto after = [&](int F, int f) {
      return (DIST && q == nseg - 1) ? Sh::stage_off((buf * NF + f) * 2 + 1) + l : F + bp;
    };
    double za[S], zb[S], ze;
    if (XT == 2) {
      double w[S + 8];
      window_sw<L, NT>(saddr(smem4 + F0), tid, q, w);
      local_sweeps_w<M>(p.oa, w, za, ze);
    } else {
      local_sweeps4<NT, M>(F0, p.oa, before(F0, 0), b0, after(F0, 0), za, ze);
    }
    smem4[cz + b0] = ze;
    smem4[cz + NT + b0] = za[0];
    if (NR == 2) {
      const int Fb = MODE == SUM ? F1 : F0, fb = MODE == SUM ? 1 : 0;
      local_sweeps4<NT, M>(Fb, p.ob, before(Fb, fb), b0, after(Fb, fb), zb, ze);
      smem4[cz + 2 * NT + b0] = ze;
      smem4[cz + 3 * NT + b0] = zb[0];
    }
    __syncthreads();
    const int xp = Sh::xb0 + buf * Sh::xbuf + l, xn = xp + NR * EXP_ROWS * L;
    double zin, yin;
    carries<L, DIST>(cz + l, cz + NT + l, L, xp, xn, p.oa, q, nseg, zin, yin);
#pragma unroll
    for (int k = 0; k < S; ++k) za[k] = fma(p.oa.Cp[k], yin, fma(p.oa.W[k], zin, za[k]));
    if (NR == 2) {
      carries<L, DIST>(cz + 2 * NT + l, cz + 3 * NT + l, L, xp + EXP_ROWS * L, xn + EXP_ROWS * L, p.ob, q, nseg, zin,
                       yin);
#pragma unroll
      for (int k = 0; k < S; ++k) zb[k] = fma(p.ob.Cp[k], yin, fma(p.ob.W[k], zin, zb[k]));
    }
    if (MODE == SINGLE && XT == 1) {
      const unsigned row = saddr(smem4 + F0) + tid * 128, x = tid & 7;
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(row + ((c ^ x) << 4), za[2 * c], za[2 * c + 1]);
    } else if (MODE == SINGLE) {
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[F0 + b0 + k * NT] = za[k];
    } else if (MODE == SUM) {
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[F0 + b0 + k * NT] = za[k] + zb[k];
    } else if (MODE == DUAL) {
#pragma unroll
      for (int k = 0; k < S; ++k) { smem4[F0 + b0 + k * NT] = za[k]; smem4[F1 + b0 + k * NT] = zb[k]; }
    } else {  // AXPY: the scale is folded into oa.cfw
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[F1 + b0 + k * NT] += za[k];
    }
    fence_async_smem();
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (DIST && nn < p.tiles) stage_neighbours(buf, nn);  // the staging areas of this buffer are free
    if (tid == 0) {
      int l0, c3, c4;
      tile_coords<L>(tile, p.nb, DIST ? 0 : p.n4, l0, c3, c4);
      if (XT == 1) tma_store_5d(&p.out_a, saddr(smem4 + F0), 0, SZ * c3 + l0, 0, 0, c4);
      else tma_store_5d(&p.out_a, saddr(smem4 + (MODE == AXPY ? F1 : F0)), l0, 0, 0, c3, c4);
      if (MODE == DUAL) tma_store_5d(&p.out_b, saddr(smem4 + F1), l0, 0, 0, c3, c4);
      tma_commit();
      if (nn < p.tiles) {
        tma_wait_read();  // the stores have read this buffer
        issue_loads(buf, nn);
      }
    }
  }
  if (tid == 0) tma_wait_all();
}


// ---- rank-split direction with the carry exchange inside the kernel (m3_common.cuh: InlineCarries) -----------------
// The wait for the neighbours' carries is taken out of the critical path by finishing every tile one iteration late:
//   iteration i:  local sweeps of tile i (results stay in registers) -> push its boundary carries
//                 poll the carries of tile i - 1 (pushed by the neighbours one tile ago), add the carry terms to the
//                 registers of tile i - 1, write the result into tile i's buffer (its input is dead after the sweeps),
//                 store it, reload that buffer with tile i + 2.
// SUM keeps the sum of its two local solutions, AXPY adds y at sweep time (both are linear), DUAL keeps two tiles.
template <int L, int NT, int MODE>
struct ShapeI {
  using Sh = Shape<L, NT, MODE>;
  static constexpr int NR = Sh::NR, NF = Sh::NF, nseg = Sh::nseg;
  static constexpr int cset = 2 * NR * NT;                       // (ze, ys) of NR recurrences; two sets
  static constexpr int cz = 2 * Sh::NSLOT * S * NT;
  static constexpr int hs0 = cz + 2 * cset + 2;                  // halo staging (after the two mbarriers)
  static constexpr int xb0 = hs0 + Sh::stage;                    // neighbour carries, two sets
  static constexpr int xset = 2 * NR * EXP_ROWS * L;
  static constexpr size_t smem() { return sizeof(double) * (xb0 + 2 * xset); }
  static __device__ __forceinline__ int stage_off(int j) { return hs0 + (j / nseg) * 4 * NT + (j % nseg) * L; }
};

template <int L, int NT, unsigned M, int MODE>
__global__ void __launch_bounds__(NT, MODE == DUAL ? 1 : 512 / NT) tds_m4i_kernel(const __grid_constant__ TdsParams4 p) {
  using Sh = ShapeI<L, NT, MODE>;
  constexpr int nseg = Sh::nseg, fd = S * NT, tpg = SZ / L, cpr = L / 2;
  constexpr int NSLOT = Shape<L, NT, MODE>::NSLOT, NR = Sh::NR, NLOAD = Shape<L, NT, MODE>::NLOAD, NF = Sh::NF, cz = Sh::cz;
  constexpr int NK = MODE == DUAL ? 2 : 1;  // register tiles kept per tile
  constexpr unsigned tile_bytes = fd * sizeof(double);
  const int tid = threadIdx.x, l = tid & (L - 1), q = tid / L;
  const int b0 = tid, bm = tid - L + (q == 0 ? NT : 0), bp = tid + L - (q == nseg - 1 ? NT : 0);
  const unsigned bar0 = saddr(smem4 + cz + 2 * Sh::cset), bar1 = bar0 + 8;
  auto issue_loads = [&](int buf, int tile) {  // thread 0
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    const unsigned bar = buf ? bar1 : bar0;
    mbar_expect_tx(bar, NLOAD * tile_bytes);
    const int c4 = grp / p.nb, c3 = grp - c4 * p.nb;
    tma_load_5d(saddr(smem4 + buf * NSLOT * fd), &p.in_a, bar, l0, 0, 0, c3, c4);
    if (NLOAD == 2) tma_load_5d(saddr(smem4 + (buf * NSLOT + 1) * fd), &p.in_b, bar, l0, 0, 0, c3, c4);
  };
  auto stage_halos = [&](int buf, int tile) {  // all threads
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    for (int idx = tid; idx < 8 * NF * cpr; idx += NT) {
      const int row = idx / cpr, c = 2 * (idx - row * cpr);
      const int f = row >> 3, side = (row >> 2) & 1, r = row & 3;
      const double* src = (side ? p.halo_e : p.halo_s) + ((size_t)(grp * NF + f) * 4 + r) * SZ + l0 + c;
      cp_async16(smem4 + Sh::stage_off((buf * NF + f) * 2 + side) + r * NT + c, src);
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(buf ? bar1 : bar0) : "memory");
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
  double zp[NK][S];  // tile i - 1
  int prev = -1;
  for (int tile = blockIdx.x, it = 0;; tile += gridDim.x, ++it) {
    const bool have = tile < p.tiles;
    if (!have && prev < 0) break;
    const int buf = it & 1, cs = cz + buf * Sh::cset, cp = cz + (buf ^ 1) * Sh::cset;
    const int F0 = buf * NSLOT * fd, F1 = F0 + fd;
    double zn[NK][S];
    if (have) {
      mbar_wait(buf ? bar1 : bar0, (it >> 1) & 1);
      auto before = [&](int F, int f) { return q == 0 ? Sh::stage_off((buf * NF + f) * 2) + l : F + bm + (S - 4) * NT; };
      auto after = [&](int F, int f) { return q == nseg - 1 ? Sh::stage_off((buf * NF + f) * 2 + 1) + l : F + bp; };
      double ze;
      local_sweeps4<NT, M>(F0, p.oa, before(F0, 0), b0, after(F0, 0), zn[0], ze);
      smem4[cs + b0] = ze;
      smem4[cs + NT + b0] = zn[0][0];
      if (NR == 2) {
        const int Fb = MODE == SUM ? F1 : F0, fb = MODE == SUM ? 1 : 0;
        double zb[S];
        local_sweeps4<NT, M>(Fb, p.ob, before(Fb, fb), b0, after(Fb, fb), zb, ze);
        smem4[cs + 2 * NT + b0] = ze;
        smem4[cs + 3 * NT + b0] = zb[0];
#pragma unroll
        for (int k = 0; k < S; ++k) {
          if (MODE == SUM) zn[0][k] += zb[k];
          else zn[NK - 1][k] = zb[k];
        }
      }
      if (MODE == AXPY) {
#pragma unroll
        for (int k = 0; k < S; ++k) zn[0][k] += smem4[F1 + b0 + k * NT];
      }
    }
    __syncthreads();  // the carries of tile i are complete; its input tile is dead
    if (have) {
      const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
      const size_t slot = (size_t)grp * NR * EXP_ROWS * SZ + l0 + l;
      push_carries_inline<L>(p.inl, slot, cs + l, cs + NT + l, p.oa, q, nseg);
      if (NR == 2) push_carries_inline<L>(p.inl, slot + (size_t)EXP_ROWS * SZ, cs + 2 * NT + l, cs + 3 * NT + l, p.ob, q, nseg);
    }
    if (prev >= 0) {
      const int grp = prev / tpg, l0 = (prev - grp * tpg) * L;
      const size_t slot = (size_t)grp * NR * EXP_ROWS * SZ + l0 + l;
      const int xp = Sh::xb0 + (buf ^ 1) * Sh::xset + l, xn = xp + NR * EXP_ROWS * L;
      poll_carries_inline<L, NR>(p.inl, slot, (size_t)EXP_ROWS * SZ, q, nseg, xp, xn, EXP_ROWS * L);
      __syncthreads();
      double zin, yin;
      carries<L, true>(cp + l, cp + NT + l, L, xp, xn, p.oa, q, nseg, zin, yin);
#pragma unroll
      for (int k = 0; k < S; ++k) zp[0][k] = fma(p.oa.Cp[k], yin, fma(p.oa.W[k], zin, zp[0][k]));
      if (NR == 2) {
        carries<L, true>(cp + 2 * NT + l, cp + 3 * NT + l, L, xp + EXP_ROWS * L, xn + EXP_ROWS * L, p.ob, q, nseg, zin, yin);
#pragma unroll
        for (int k = 0; k < S; ++k) zp[NK - 1][k] = fma(p.ob.Cp[k], yin, fma(p.ob.W[k], zin, zp[NK - 1][k]));
      }
      // AXPY stores from the y slot like the plain kernel, everything else from slot 0
      const int O0 = MODE == AXPY ? F1 : F0;
#pragma unroll
      for (int k = 0; k < S; ++k) smem4[O0 + b0 + k * NT] = zp[0][k];
      if (MODE == DUAL) {
#pragma unroll
        for (int k = 0; k < S; ++k) smem4[F1 + b0 + k * NT] = zp[1][k];
      }
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        const int c4 = grp / p.nb, c3 = grp - c4 * p.nb;
        tma_store_5d(&p.out_a, saddr(smem4 + O0), l0, 0, 0, c3, c4);
        if (MODE == DUAL) tma_store_5d(&p.out_b, saddr(smem4 + F1), l0, 0, 0, c3, c4);
        tma_commit();
      }
    }
    if (!have) break;
    const int nn = tile + 2 * gridDim.x;
    if (nn < p.tiles) {
      stage_halos(buf, nn);  // this buffer's staging area was read by the sweeps above
      if (tid == 0) {
        tma_wait_read();  // the store of tile i - 1 has read this buffer
        issue_loads(buf, nn);
      }
    }
#pragma unroll
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int k = 0; k < S; ++k) zp[j][k] = zn[j][k];
    prev = tile;
  }
  if (tid == 0) tma_wait_all();
}

template <int L, int NT, unsigned M, int MODE>
int launch_inline(x3d2c_ctx* ctx, const TdsParams4& p) {
  constexpr size_t smem = ShapeI<L, NT, MODE>::smem();
  static int per_sm_dev[x3d2c::kMaxDevices] = {};
  int& per_sm = per_sm_dev[ctx->device];
  if (!per_sm) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(tds_m4i_kernel<L, NT, M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    X3D2C_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tds_m4i_kernel<L, NT, M, MODE>, NT, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  tds_m4i_kernel<L, NT, M, MODE><<<grid, NT, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

// ---------------------------------------------------------------------------------------------- host side
template <int L, int NT, unsigned M, int MODE, bool DIST>
int launch(x3d2c_ctx* ctx, const TdsParams4& p) {
  if (DIST && p.inl.from_prev) return launch_inline<L, NT, M, MODE>(ctx, p);
  constexpr size_t smem = Shape<L, NT, MODE>::smem(DIST);
  static int per_sm_dev[x3d2c::kMaxDevices] = {};  // resident CTAs per SM (registers and shared memory), once per device
  int& per_sm = per_sm_dev[ctx->device];
  if (!per_sm) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(tds_m4_kernel<L, NT, M, MODE, DIST>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    X3D2C_CHECK_CUDA(
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tds_m4_kernel<L, NT, M, MODE, DIST>, NT, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  tds_m4_kernel<L, NT, M, MODE, DIST><<<grid, NT, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L, int NT, unsigned M, int MODE, int XT>
int launch_xt(x3d2c_ctx* ctx, const TdsParams4& p) {
  constexpr size_t smem = Shape<L, NT, MODE>::smem(false);
  static int per_sm_dev[x3d2c::kMaxDevices] = {};
  int& per_sm = per_sm_dev[ctx->device];
  if (!per_sm) {
    X3D2C_CHECK_CUDA(cudaFuncSetAttribute(tds_m4_kernel<L, NT, M, MODE, false, XT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    X3D2C_CHECK_CUDA(
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tds_m4_kernel<L, NT, M, MODE, false, XT>, NT, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = num_sms(ctx) * per_sm;
  if (grid > p.tiles) grid = p.tiles;
  tds_m4_kernel<L, NT, M, MODE, false, XT><<<grid, NT, smem, ctx->stream>>>(p);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

// the staggered operators of the divergence (v2p, 0x78) and of the gradient (p2v, 0x3C) only
template <int L, int NT, int MODE, int XT>
int dispatch_xt_mask(x3d2c_ctx* ctx, const TdsParams4& p, unsigned mask) {
  switch (mask) {
    case 0x78u: return launch_xt<L, NT, 0x78u, MODE, XT>(ctx, p);
    case 0x3Cu: return launch_xt<L, NT, 0x3Cu, MODE, XT>(ctx, p);
    default: return X3D2C_EUNSUPPORTED;
  }
}
template <int MODE, int XT>
int dispatch_xt(x3d2c_ctx* ctx, const TdsParams4& p, int L, int NT, unsigned mask) {
  if (NT == 128) return dispatch_xt_mask<32, 128, MODE, XT>(ctx, p, mask);
  if (NT == 512) return MODE == SINGLE ? dispatch_xt_mask<8, 512, SINGLE, XT>(ctx, p, mask) : X3D2C_EUNSUPPORTED;
  switch (L) {
    case 4: return dispatch_xt_mask<4, 256, MODE, XT>(ctx, p, mask);
    case 8: return dispatch_xt_mask<8, 256, MODE, XT>(ctx, p, mask);
    case 16: return dispatch_xt_mask<16, 256, MODE, XT>(ctx, p, mask);
    default: return dispatch_xt_mask<32, 256, MODE, XT>(ctx, p, mask);
  }
}

template <int L, int NT, int MODE, bool DIST>
int dispatch_mask(x3d2c_ctx* ctx, const TdsParams4& p, unsigned mask) {
  switch (mask) {
    case 0x78u: return launch<L, NT, 0x78u, MODE, DIST>(ctx, p);  // staggered derivative / interpolation v2p
    case 0x3Cu: return launch<L, NT, 0x3Cu, MODE, DIST>(ctx, p);  // p2v
    case 0x6Cu: if (MODE == SINGLE) return launch<L, NT, 0x6Cu, SINGLE, DIST>(ctx, p);  // first derivative
    case 0x7Cu: if (MODE == SINGLE) return launch<L, NT, 0x7Cu, SINGLE, DIST>(ctx, p);  // second derivative
    default: return launch<L, NT, 0x1FFu, MODE, DIST>(ctx, p);
  }
}

// Tiles of 256 threads: the copy-bound tds kernels want rows of at least 64 bytes (L >= 8) where the line allows it
bool tds_shape(int n, int mode, int* L, int* NT) {
  // 1024-point lines: a single solve has room for 64-byte rows (one CTA of 512 threads, two 64 KB tile buffers); the
  // two-tile modes keep 32-byte rows
  if (n == 1024 && mode == SINGLE) { *L = 8; *NT = 512; return true; }
  switch (n) {
    case 64: *L = 32; *NT = 128; return true;
    case 128: *L = 32; *NT = 256; return true;
    case 256: *L = 16; *NT = 256; return true;
    case 512: *L = 8; *NT = 256; return true;
    case 1024: *L = 4; *NT = 256; return true;  // 32-byte rows: no faster than the cp.async kernel (0.51 ms), but the
                                                // output reorders can go through the tensor map
    default: return false;
  }
}

template <int MODE, bool DIST>
int dispatch_shape(x3d2c_ctx* ctx, const TdsParams4& p, int L, int NT, unsigned mask) {
  if (NT == 128) return dispatch_mask<32, 128, MODE, DIST>(ctx, p, mask);
  if (NT == 512) {
    if (MODE == SINGLE) return dispatch_mask<8, 512, SINGLE, DIST>(ctx, p, mask);
    return X3D2C_EUNSUPPORTED;
  }
  switch (L) {
    case 4: return dispatch_mask<4, 256, MODE, DIST>(ctx, p, mask);
    case 8: return dispatch_mask<8, 256, MODE, DIST>(ctx, p, mask);
    case 16: return dispatch_mask<16, 256, MODE, DIST>(ctx, p, mask);
    default: return dispatch_mask<32, 256, MODE, DIST>(ctx, p, mask);
  }
}

template <bool DIST>
int dispatch_mode(x3d2c_ctx* ctx, const TdsParams4& p, int mode, int L, int NT, unsigned mask) {
  switch (mode) {
    case SINGLE: return dispatch_shape<SINGLE, DIST>(ctx, p, L, NT, mask);
    case SUM: return dispatch_shape<SUM, DIST>(ctx, p, L, NT, mask);
    case DUAL: return dispatch_shape<DUAL, DIST>(ctx, p, L, NT, mask);
    default: return dispatch_shape<AXPY, DIST>(ctx, p, L, NT, mask);
  }
}

}  // namespace

namespace x3d2c {

// mode: 0 single (out_a = A(in_a)), 1 sum, 2 dual, 3 axpy (out_a = y, scale_a folded into A).
// lay_in / lay_out: layout (DIR_*) in which the input / output fields are stored; different from `dir` when the
// kernel reads or writes through a reorder (Y and Z lines, layouts Y, Z, C).
int tds_m4(x3d2c_ctx* ctx, int dir, int mode, double* out_a, double* out_b, const double* in_a, const double* in_b,
           const x3d2c_tdsops* ta, const x3d2c_tdsops* tb, double scale_a, int lay_in, int lay_out) {
  static const bool disabled = std::getenv("X3D2C_NO_TMA") != nullptr;
  if (disabled) return X3D2C_EUNSUPPORTED;
  const int n = ta->n_tds;
  const bool split = ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist;
  if (split && (lay_in != dir || !dist_supported(ctx, dir, n))) return X3D2C_EUNSUPPORTED;
  int L = 0, NT = 0;
  if (!tds_shape(n, mode, &L, &NT)) return X3D2C_EUNSUPPORTED;
  const bool two_ops = mode == SUM || mode == DUAL;
  if (two_ops && (tb->n_tds != n || tb->n_rhs != ta->n_rhs)) return X3D2C_EUNSUPPORTED;
  TdsParams4 p{};
  if (!make_op(ta, mode == AXPY ? scale_a : 1.0, split, &p.oa)) return X3D2C_EUNSUPPORTED;
  if (two_ops && !make_op(tb, 1.0, split, &p.ob)) return X3D2C_EUNSUPPORTED;
  const int G = ctx->n_groups[dir], nseg = n / S;
  auto map = [&](CUtensorMap* m, const double* f, int layout) { return make_map5(m, f, layout, dir, L, nseg, ctx, &p.nb); };
  // x lines reading or writing a field kept in the y layout (an X <-> Y reorder folded into the solve)
  const int xt = dir != X3D2C_DIR_X ? 0 : (lay_out != dir ? 1 : (lay_in != dir ? 2 : 0));
  if (xt) {
    static const bool no_xt = std::getenv("X3D2C_NO_XT") != nullptr;
    if (no_xt || split || (lay_in != dir && lay_out != dir)) return X3D2C_EUNSUPPORTED;
    if (!(xt == 1 && mode == SINGLE) && !(xt == 2 && (mode == SINGLE || mode == AXPY))) return X3D2C_EUNSUPPORTED;
    const int lay_native = dir;
    if (xt == 1) {
      if (!map(&p.in_a, in_a, lay_native) || !make_map_xt(&p.out_a, out_a, lay_out, dir, L, nseg, ctx)) return X3D2C_EUNSUPPORTED;
    } else {
      if (!make_map_xt(&p.in_a, in_a, lay_in, dir, L, nseg, ctx) || !map(&p.out_a, out_a, lay_native)) return X3D2C_EUNSUPPORTED;
      if (mode == AXPY && !map(&p.in_b, in_b, lay_native)) return X3D2C_EUNSUPPORTED;
    }
    p.tiles = G * (SZ / L);
    if (xt == 1) return dispatch_xt<SINGLE, 1>(ctx, p, L, NT, ta->tap_mask);
    return mode == SINGLE ? dispatch_xt<SINGLE, 2>(ctx, p, L, NT, ta->tap_mask) : dispatch_xt<AXPY, 2>(ctx, p, L, NT, ta->tap_mask);
  }
  if (!map(&p.in_a, in_a, lay_in) || !map(&p.out_a, out_a, lay_out)) return X3D2C_EUNSUPPORTED;
  {
    // tile order with the foreign layout's row index fastest: measured SLOWER for the y <-> z stores of the pressure path
    // (pressure correction 10.9 against 10.1 ms at 512^3), so it stays an experiment
    static const bool swap = std::getenv("X3D2C_TILE_ORDER") != nullptr;
    const bool yz = (dir == X3D2C_DIR_Z && lay_out == X3D2C_DIR_Y) || (dir == X3D2C_DIR_Y && lay_out == X3D2C_DIR_Z);
    if (yz && !split && swap) p.n4 = G / p.nb;  // c4: y for z lines, z for y lines
  }
  if (mode == SUM && !map(&p.in_b, in_b, lay_in)) return X3D2C_EUNSUPPORTED;
  if (mode == AXPY && !map(&p.in_b, in_b, lay_out)) return X3D2C_EUNSUPPORTED;  // y
  if (mode == DUAL && !map(&p.out_b, out_b, lay_out)) return X3D2C_EUNSUPPORTED;
  p.tiles = G * (SZ / L);
  const unsigned mask = two_ops ? (ta->tap_mask | tb->tap_mask) : ta->tap_mask;
  if (!split) return dispatch_mode<false>(ctx, p, mode, L, NT, mask);
  // rank-split direction: halos and boundary carries first (m3_edge.cu), then the main kernel
  DistBufs b = carve_dist(ctx);
  EdgeParams ep{};
  ep.n = n;
  ep.n_pad = ctx->n_pad(dir);
  ep.nseg = nseg;
  ep.ns = two_ops ? 2 : 1;
  ep.ops[0] = p.oa;
  ep.ops[1] = p.ob;
  ep.f[0] = in_a;
  ep.f[1] = in_b;
  const double* fields[2] = {in_a, in_b};
  int rc = exchange_edges(ctx, dir, fields, mode == SUM ? 2 : 1, ep, b, &p.inl);
  if (rc) return rc;
  p.halo_s = b.halo_recv_s;
  p.halo_e = b.halo_recv_e;
  p.from_prev = b.carr_from_prev;
  p.from_next = b.carr_from_next;
  return dispatch_mode<true>(ctx, p, mode, L, NT, mask);
}

}  // namespace x3d2c
