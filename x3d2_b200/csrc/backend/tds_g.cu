// Generic segment-parallel DistD2-TDS kernels ("g"): any boundary condition (Dirichlet / Neumann walls, periodic),
// stretched meshes (stretch, stretch_correct), n_rhs = n_tds + 1 operators and line lengths that are not a multiple of
// 16 (257-point channel lines), for directions held by one rank. They replace, for these operators, the
// one-thread-per-line kernels of tds_m1.cu (which sweep through global memory three times, like the reference's
// der_univ_dist / der_univ_subs / der_univ_fused_subs: omp/kernels/distributed.f90:11-337) on the fast path.
//
// The algorithm is the reference's, reorganised exactly like the periodic fast path (tds_m3.cu) but with per-row
// coefficient tables instead of the Toeplitz constants:
//   forward   z_j = A_j z_{j-1} + B_j r_j        A_j = -fw_j af_j, B_j = fw_j (rows 1, 2: A = 0, B = af_j)   :37-96
//   backward  y_j = C_j y_{j+1} + E_j z_j        C_j = -bw_j (2 <= j <= n-2), row 1: C = -fw_1 bw_1, E = fw_1     :153-166
//   ends      s = (y_1 - sa_1 z_n) / (1 - sa_1^2),  e = (z_n - sc_n y_1) / (1 - sc_n^2)   (single rank: the exchanged
//             rows are the line's own, omp/sendrecv.f90:20-22)                                                   :196-206
//   result    x_j = (y_j - sa_j s - sc_j e) stretch_j,  x_1 = s stretch_1,  x_n = e stretch_n                   :208-228
// A line is cut into 16-row segments, one thread per (lane, segment). A thread evaluates its rows' 9-point stencils
// in the reference's summation order (first / last four rows with coeffs_s / coeffs_e), sweeps forwards and
// backwards locally in registers and exchanges one carry per sweep through shared memory; the carries of the three
// nearest segments on either side are combined with per-segment weights prepared on the host (products of A_j / C_j
// over whole segments; the truncation is checked there: < 1e-17, else the operator stays on tds_m1.cu).
// HBM sees one read of each input and one write of each output.
#include <algorithm>

#include "m3_common.cuh"

using namespace m3;

namespace {

enum { T_A, T_B, T_C, T_W, T_CP, T_SA, T_SC, T_ST, NTAB };  // per-row tables
constexpr int SEGW = 12;  // per-segment weights: ZW[3], YW[3], OM[5], pad

struct GenOp {
  const double* tab;  // NTAB tables of np doubles
  const double* seg;  // nseg x SEGW
  const double* rc;   // [8][9]: stencils of rows 1..4 and n_rhs-3..n_rhs (device memory)
  double c[9];        // bulk stencil
  double e1;          // E of row 1
  double sa1, s_rcp, scn, e_rcp;
  unsigned mask;
  int n_tds, n_rhs;
};

struct GGeom {
  int np, nseg, n_pad, tiles, fd;  // rows processed (nseg * 16), padded line length, tiles, doubles per field tile
};

__device__ __forceinline__ double sten_rows(const double* c, const double (&w)[9]) {
  double t = __dmul_rn(__ldg(c), w[0]);
#pragma unroll
  for (int k = 1; k < 9; ++k) t = __dadd_rn(t, __dmul_rn(__ldg(c + k), w[k]));
  return t;
}

// stencil of row j (1-based) of an operator: boundary rows take their own coefficient rows
template <unsigned M>
__device__ __forceinline__ double rhs_row(const GenOp& o, const int j, const bool interior, const double (&w)[9]) {
  if (interior) return sten_exact<M>(o.c, w);
  if (j <= 4) return sten_rows(o.rc + (j - 1) * 9, w);
  if (j > o.n_rhs) return 0.0;
  if (j >= o.n_rhs - 3) return sten_rows(o.rc + (4 + j - (o.n_rhs - 3)) * 9, w);
  return sten_exact<0x1FFu>(o.c, w);
}

// local sweeps of one recurrence over the 16 rows of segment q; w(k): window provider
template <unsigned M, class Win>
__device__ __forceinline__ void local_sweeps(const GenOp& o, const int np, const int q, const bool interior, Win&& win,
                                             double (&z)[S], double& fe, double& bs) {
  const double* tA = o.tab + T_A * np + q * S;
  const double* tB = o.tab + T_B * np + q * S;
  const double* tC = o.tab + T_C * np + q * S;
  double w[9];
#pragma unroll
  for (int t = 0; t < 8; ++t) w[t] = win(t);
  double pz = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    w[8] = win(k + 8);
    const double r = rhs_row<M>(o, q * S + k + 1, interior, w);
    pz = fma(__ldg(tA + k), pz, __ldg(tB + k) * r);
    z[k] = pz;
#pragma unroll
    for (int t = 0; t < 8; ++t) w[t] = w[t + 1];
  }
  fe = pz;
  double y = 0.0;
#pragma unroll
  for (int k = S - 1; k >= 0; --k) {
    const double ez = (q == 0 && k == 0) ? o.e1 * z[k] : z[k];
    y = fma(__ldg(tC + k), y, ez);
    z[k] = y;
  }
  bs = y;
}

// carries from the three nearest segments on either side (ze / ys: shared offsets of segment 0, lane applied)
template <int L>
__device__ __forceinline__ void gen_carries(const GenOp& o, const int ze, const int ys, const int q, const int nseg,
                                            double& zin, double& yin) {
  const double* sw = o.seg + q * SEGW;
  zin = 0.0;
  yin = 0.0;
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) {
    const int sm_ = q - d < 0 ? 0 : q - d, sp = q + d >= nseg ? nseg - 1 : q + d;  // weights are zero out of range
    zin = fma(__ldg(sw + d - 1), smem[ze + sm_ * L], zin);
    yin = fma(__ldg(sw + 3 + d - 1), smem[ys + sp * L], yin);
  }
#pragma unroll
  for (int m = -(DMAX - 1); m <= DMAX - 1; ++m) {
    int s = q + m;
    s = s < 0 ? 0 : (s >= nseg ? nseg - 1 : s);
    yin = fma(__ldg(sw + 6 + m + DMAX - 1), smem[ze + s * L], yin);
  }
}

// after the carries: y_j complete. Publishes y_1 and z_n (= y_n) of the line for the substitution.
template <int L>
__device__ __forceinline__ void finish_sweeps(const GenOp& o, const int np, const int q, const int l, const double zin,
                                              const double yin, double (&z)[S], const int ends) {
  const double* tW = o.tab + T_W * np + q * S;
  const double* tCp = o.tab + T_CP * np + q * S;
#pragma unroll
  for (int k = 0; k < S; ++k) z[k] = fma(__ldg(tCp + k), yin, fma(__ldg(tW + k), zin, z[k]));
  if (q == 0) smem[ends + l] = z[0];
  const int kn = o.n_tds - 1 - q * S;  // position of row n in this segment
  if (kn >= 0 && kn < S) {
#pragma unroll
    for (int k = 0; k < S; ++k)
      if (k == kn) smem[ends + L + l] = z[k];
  }
}

// substitution: value of row k of segment q given the line's y_1 and z_n (without the stretch factor)
__device__ __forceinline__ void line_ends(const GenOp& o, const double y1, const double zn, double& s, double& e) {
  s = o.s_rcp * (y1 - o.sa1 * zn);
  e = o.e_rcp * (zn - o.scn * y1);
}
__device__ __forceinline__ double subs_row(const GenOp& o, const int np, const int q, const int k, const double y,
                                           const double s, const double e) {
  const int j = q * S + k + 1;
  if (j == 1) return s;
  if (j == o.n_tds) return e;
  const int r = q * S + k;
  return y - __ldg(o.tab + T_SA * np + r) * s - __ldg(o.tab + T_SC * np + r) * e;
}

// tile copies: (32 lanes, n_pad rows, G groups) <-> shared [segment][SP rows][L lanes]; any thread count
template <int L>
struct GCopier {
  __device__ __forceinline__ static const double* base(const double* g, const GGeom& q, int tile) {
    constexpr int tpg = SZ / L;
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    return g + (size_t)grp * q.n_pad * SZ + l0;
  }
  __device__ __forceinline__ static void load(double* sm, const double* g, const GGeom& q, int tile) {
    constexpr int cpr = L / 2;
    const double* src = base(g, q, tile);
    for (int idx = threadIdx.x; idx < q.np * cpr; idx += blockDim.x) {
      const int r = idx / cpr, c2 = 2 * (idx - r * cpr);
      cp_async16(sm + (r + (r >> LOG2S)) * L + c2, src + (size_t)r * SZ + c2);
    }
  }
  __device__ __forceinline__ static void store(double* g, const double* sm, const GGeom& q, int tile, int rows) {
    constexpr int cpr = L / 2;
    double* dst = const_cast<double*>(base(g, q, tile));
    for (int idx = threadIdx.x; idx < rows * cpr; idx += blockDim.x) {
      const int r = idx / cpr, c2 = 2 * (idx - r * cpr);
      __stcs(reinterpret_cast<double2*>(dst + (size_t)r * SZ + c2),
             *reinterpret_cast<const double2*>(sm + (r + (r >> LOG2S)) * L + c2));
    }
  }
  // rows [from, np) of a tile are padding of the input line: make them finite (they only meet zero coefficients)
  __device__ __forceinline__ static void zero_tail(double* sm, const GGeom& q, int from) {
    for (int idx = threadIdx.x; idx < (q.np - from) * L; idx += blockDim.x) {
      const int r = from + idx / L, c = idx - (idx / L) * L;
      sm[(r + (r >> LOG2S)) * L + c] = 0.0;
    }
  }
};

// ------------------------------------------------------------------------------------------------ tds_solve
struct TdsG {
  const double* in;
  double* out;
  GGeom g;
  GenOp o;
};

// shared memory: [2 tiles][fd] | ze, ys [nseg * L] | y_1, z_n [2 L]
template <int L, unsigned M>
__global__ void __launch_bounds__(288) tds_g_kernel(const __grid_constant__ TdsG p) {
  const GGeom& g = p.g;
  const int fd = g.fd, nseg = g.nseg, np = g.np;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  const int qm = q == 0 ? nseg - 1 : q - 1, qp = q == nseg - 1 ? 0 : q + 1;
  const int bm = qm * SP * L + l, b0 = q * SP * L + l, bp = qp * SP * L + l;
  const int ze = 2 * fd + l, ys = ze + nseg * L, ends = 2 * fd + 2 * nseg * L;
  const bool interior = q > 0 && (q + 1) * S < p.o.n_rhs - 3;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
    if (it == 0) {
      GCopier<L>::load(smem, p.in, g, tile);
      cp_async_commit();
      const int nx = tile + gridDim.x;
      if (nx < g.tiles) GCopier<L>::load(smem + fd, p.in, g, nx);
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncthreads();
    double* F = smem + (it & 1) * fd;
    if (p.o.n_rhs < np) {
      GCopier<L>::zero_tail(F, g, p.o.n_rhs);
      __syncthreads();
    }
    double z[S], fe, bs;
    local_sweeps<M>(p.o, np, q, interior, [&](int t) { return F[woff<L>(t, bm, b0, bp)]; }, z, fe, bs);
    smem[ze + q * L] = fe;
    smem[ys + q * L] = bs;
    __syncthreads();
    double zin, yin;
    gen_carries<L>(p.o, ze, ys, q, nseg, zin, yin);
    finish_sweeps<L>(p.o, np, q, l, zin, yin, z, ends);
    __syncthreads();
    double s, e;
    line_ends(p.o, smem[ends + l], smem[ends + L + l], s, e);
    const double* tS = p.o.tab + T_ST * np + q * S;
#pragma unroll
    for (int k = 0; k < S; ++k) F[b0 + k * L] = subs_row(p.o, np, q, k, z[k], s, e) * __ldg(tS + k);
    __syncthreads();
    GCopier<L>::store(p.out, F, g, tile, np);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) GCopier<L>::load(smem + (it & 1) * fd, p.in, g, nn);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ transeq
struct TransG {
  const double* in[3];  // in[0]: the line-aligned velocity (conv)
  double* out[3];
  GGeom g;
  GenOp du[2], dud[2], d2u[2];  // [0]: aligned component (der1st, der1st_sym, der2nd), [1]: the other two
  const double* stc;            // stretch_correct of der2nd, np entries
  double nu;
};

// One velocity component of one tile (exec_dist_transeq_compact + der_univ_fused_subs, exec_dist.f90:67-186,
// distributed.f90:231-337). fF: field tile (in place), fC: conv tile; cz: carries of three recurrences; ends: y_1 / z_n.
template <int L, unsigned M1, unsigned M2, bool SELF>
__device__ __forceinline__ void component_g(const TransG& p, const int t, const int fF, const int fC, const int cz,
                                            const int ends, const int q, const int l, const int bm, const int b0,
                                            const int bp, const bool interior) {
  const int nseg = p.g.nseg, np = p.g.np;
  const GenOp &o1 = p.du[t], &o2 = p.dud[t], &o3 = p.d2u[t];
  double z1[S], z2[S], z3[S];
  const int zeo = cz + l, yso = cz + nseg * L + l, stride = 2 * nseg * L;
  {
    const double *a1 = o1.tab + T_A * np + q * S, *b1 = o1.tab + T_B * np + q * S;
    const double *a2 = o2.tab + T_A * np + q * S, *b2 = o2.tab + T_B * np + q * S;
    const double *a3 = o3.tab + T_A * np + q * S, *b3 = o3.tab + T_B * np + q * S;
    double wf[9], wp[9];
    auto load = [&](int t, double& f, double& pr) {
      const int o = woff<L>(t, bm, b0, bp);
      f = smem[fF + o];
      pr = f * (SELF ? f : smem[fC + o]);
    };
#pragma unroll
    for (int t = 0; t < 8; ++t) load(t, wf[t], wp[t]);
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      load(k + 8, wf[8], wp[8]);
      const int j = q * S + k + 1;
      p1 = fma(__ldg(a1 + k), p1, __ldg(b1 + k) * rhs_row<M1>(o1, j, interior, wf));
      p2 = fma(__ldg(a2 + k), p2, __ldg(b2 + k) * rhs_row<M1>(o2, j, interior, wp));
      p3 = fma(__ldg(a3 + k), p3, __ldg(b3 + k) * rhs_row<M2>(o3, j, interior, wf));
      z1[k] = p1; z2[k] = p2; z3[k] = p3;
#pragma unroll
      for (int t = 0; t < 8; ++t) { wf[t] = wf[t + 1]; wp[t] = wp[t + 1]; }
    }
    smem[zeo + q * L] = p1;
    smem[zeo + stride + q * L] = p2;
    smem[zeo + 2 * stride + q * L] = p3;
  }
  {
    const double *c1 = o1.tab + T_C * np + q * S, *c2 = o2.tab + T_C * np + q * S, *c3 = o3.tab + T_C * np + q * S;
    double y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
    for (int k = S - 1; k >= 0; --k) {
      const bool first = q == 0 && k == 0;
      y1 = fma(__ldg(c1 + k), y1, first ? o1.e1 * z1[k] : z1[k]);
      y2 = fma(__ldg(c2 + k), y2, first ? o2.e1 * z2[k] : z2[k]);
      y3 = fma(__ldg(c3 + k), y3, first ? o3.e1 * z3[k] : z3[k]);
      z1[k] = y1; z2[k] = y2; z3[k] = y3;
    }
    smem[yso + q * L] = y1;
    smem[yso + stride + q * L] = y2;
    smem[yso + 2 * stride + q * L] = y3;
  }
  __syncthreads();
  double zin, yin;
  gen_carries<L>(o1, zeo, yso, q, nseg, zin, yin);
  finish_sweeps<L>(o1, np, q, l, zin, yin, z1, ends);
  gen_carries<L>(o2, zeo + stride, yso + stride, q, nseg, zin, yin);
  finish_sweeps<L>(o2, np, q, l, zin, yin, z2, ends + 2 * L);
  gen_carries<L>(o3, zeo + 2 * stride, yso + 2 * stride, q, nseg, zin, yin);
  finish_sweeps<L>(o3, np, q, l, zin, yin, z3, ends + 4 * L);
  __syncthreads();
  double s1, e1, s2, e2, s3, e3;
  line_ends(o1, smem[ends + l], smem[ends + L + l], s1, e1);
  line_ends(o2, smem[ends + 2 * L + l], smem[ends + 3 * L + l], s2, e2);
  line_ends(o3, smem[ends + 4 * L + l], smem[ends + 5 * L + l], s3, e3);
  const double* st1 = o1.tab + T_ST * np + q * S;
  const double* st2 = o2.tab + T_ST * np + q * S;
  const double* st3 = o3.tab + T_ST * np + q * S;
  const double* stc = p.stc + q * S;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    const double t_du = __ldg(st1 + k) * subs_row(o1, np, q, k, z1[k], s1, e1);
    const double t_dud = __ldg(st2 + k) * subs_row(o2, np, q, k, z2[k], s2, e2);
    const double t_d2u = __ldg(st3 + k) * subs_row(o3, np, q, k, z3[k], s3, e3) + t_du * __ldg(stc + k);
    const double conv = SELF ? smem[fF + b0 + k * L] : smem[fC + b0 + k * L];
    smem[fF + b0 + k * L] = -0.5 * (conv * t_du + t_dud) + p.nu * t_d2u;
  }
  __syncthreads();
}

// shared memory: [2 buffers][3 fields][fd] | 3 x (ze, ys)[nseg * L] | 3 x (y_1, z_n)[L]
template <int L, unsigned M1, unsigned M2>
__global__ void __launch_bounds__(256, 1) transeq_g_kernel(const __grid_constant__ TransG p) {
  const GGeom& g = p.g;
  const int fd = g.fd, nseg = g.nseg;
  const int l = threadIdx.x & (L - 1), q = threadIdx.x / L;
  const int qm = q == 0 ? nseg - 1 : q - 1, qp = q == nseg - 1 ? 0 : q + 1;
  const int bm = qm * SP * L + l, b0 = q * SP * L + l, bp = qp * SP * L + l;
  const int cz = 6 * fd, ends = cz + 6 * nseg * L;
  const int n_rhs = p.du[0].n_rhs;
  const bool interior = q > 0 && (q + 1) * S < n_rhs - 3;
  auto load_tile = [&](int buf, int tile) {
#pragma unroll
    for (int f = 0; f < 3; ++f) GCopier<L>::load(smem + (3 * buf + f) * fd, p.in[f], g, tile);
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
    const int bo = (it & 1) * 3 * fd;
    if (n_rhs < g.np) {
#pragma unroll
      for (int f = 0; f < 3; ++f) GCopier<L>::zero_tail(smem + bo + f * fd, g, n_rhs);
      __syncthreads();
    }
    // components 1 and 2 first: they read the aligned velocity (field 0) as conv; field 0 is overwritten last
    component_g<L, M1, M2, false>(p, 1, bo + 1 * fd, bo, cz, ends, q, l, bm, b0, bp, interior);
    component_g<L, M1, M2, false>(p, 1, bo + 2 * fd, bo, cz, ends, q, l, bm, b0, bp, interior);
    component_g<L, M1, M2, true>(p, 0, bo, bo, cz, ends, q, l, bm, b0, bp, interior);
#pragma unroll
    for (int f = 0; f < 3; ++f) GCopier<L>::store(p.out[f], smem + bo + f * fd, g, tile, g.np);
    __syncthreads();
    const int nn = tile + 2 * gridDim.x;
    if (nn < g.tiles) load_tile(it & 1, nn);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ host side
struct GenHost {
  int nseg = 0, np = 0;
  double e1 = 1, sa1 = 0, s_rcp = 1, scn = 0, e_rcp = 1;
};

// Builds the per-row / per-segment tables of one operator from the tables of tdsops_init (as passed to
// x3d2c_tdsops_create). Returns false when the carries do not decay fast enough for DMAX segments.
bool build_tables(const x3d2c_tdsops* t, std::vector<double>& tab, std::vector<double>& seg, GenHost& h) {
  const int n = t->n_tds, n_rhs = t->n_rhs;
  if (n < 2 * S) return false;
  const int nseg = (std::max(n, n_rhs) + S - 1) / S, np = nseg * S;
  h.nseg = nseg; h.np = np;
  tab.assign((size_t)NTAB * np, 0.0);
  seg.assign((size_t)nseg * SEGW, 0.0);
  double *A = &tab[T_A * np], *B = &tab[T_B * np], *C = &tab[T_C * np], *W = &tab[T_W * np], *Cp = &tab[T_CP * np],
         *SA = &tab[T_SA * np], *SC = &tab[T_SC * np], *ST = &tab[T_ST * np];
  const auto &fw = t->h_fw, &bw = t->h_bw, &sa = t->h_sa, &sc = t->h_sc, &af = t->h_af;
  std::vector<double> E(np, 0.0);
  for (int j = 1; j <= n; ++j) {  // 1-based row j at index j - 1
    const int r = j - 1;
    if (j <= 2) { A[r] = 0.0; B[r] = af[r]; }
    else { A[r] = -fw[r] * af[r]; B[r] = fw[r]; }
    if (j == 1) { C[r] = -fw[0] * bw[0]; E[r] = fw[0]; }
    else { C[r] = (j <= n - 2) ? -bw[r] : 0.0; E[r] = 1.0; }
    if (j >= 2 && j <= n - 1) { SA[r] = sa[r]; SC[r] = sc[r]; }
    ST[r] = t->h_stretch[r];
  }
  h.e1 = E[0];
  h.sa1 = sa[0]; h.s_rcp = 1.0 / (1.0 - sa[0] * sa[0]);
  h.scn = sc[n - 1]; h.e_rcp = 1.0 / (1.0 - sc[n - 1] * sc[n - 1]);
  // per segment: alpha (forward propagation across the segment), W, Cp; g = W(first row), beta = Cp(first row)
  std::vector<double> alpha(nseg), g(nseg), beta(nseg);
  for (int q = 0; q < nseg; ++q) {
    double pf[S], pr = 1.0;
    for (int k = 0; k < S; ++k) { pr *= A[q * S + k]; pf[k] = pr; }
    alpha[q] = pr;
    double y = 0.0, cp = 1.0;
    for (int k = S - 1; k >= 0; --k) {
      const int r = q * S + k;
      y = C[r] * y + E[r] * pf[k];
      W[r] = y;
      cp *= C[r];
      Cp[r] = cp;
    }
    g[q] = W[q * S];
    beta[q] = Cp[q * S];
  }
  auto ZW = [&](int q, int d) -> double {  // weight of fe(q - d) in zin(q)
    if (q - d < 0) return 0.0;
    double w = 1.0;
    for (int s = q - d + 1; s <= q - 1; ++s) w *= alpha[s];
    return w;
  };
  for (int q = 0; q < nseg; ++q) {
    double* sw = &seg[(size_t)q * SEGW];
    for (int d = 1; d <= DMAX; ++d) sw[d - 1] = ZW(q, d);
    double yw[DMAX + 1] = {0, 0, 0, 0};
    for (int d = 1; d <= DMAX; ++d) {
      if (q + d >= nseg) continue;
      double w = 1.0;
      for (int s = q + 1; s <= q + d - 1; ++s) w *= beta[s];
      yw[d] = w;
      sw[3 + d - 1] = w;
    }
    for (int d = 1; d <= DMAX; ++d)
      for (int e = 1; e <= DMAX; ++e) {
        if (q + d >= nseg) continue;
        const int m = d - e;  // fe(q + d - e)
        if (m < -(DMAX - 1) || m > DMAX - 1) continue;
        sw[6 + m + DMAX - 1] += yw[d] * g[q + d] * ZW(q + d, e);
      }
    // truncation: what a fourth segment would contribute
    const double tz = q - DMAX - 1 >= 0 ? std::fabs(ZW(q, DMAX) * alpha[q - DMAX]) : 0.0;
    double ty = 0.0;
    if (q + DMAX + 1 < nseg) ty = std::fabs(yw[DMAX] * beta[q + DMAX]);
    if (tz > 1e-17 || ty > 1e-17) return false;
  }
  for (size_t i = 0; i < tab.size(); ++i)
    if (!std::isfinite(tab[i])) return false;
  return std::isfinite(h.s_rcp) && std::isfinite(h.e_rcp);
}

// device copy of an operator's tables, built on first use and kept with the handle (x3d2c_tdsops::d_m3)
bool gen_op(x3d2c_ctx* ctx, const x3d2c_tdsops* tc, GenOp* o, GenHost* hh) {
  auto* t = const_cast<x3d2c_tdsops*>(tc);
  if (t->gen_state < 0) return false;
  if (t->gen_state == 0) {
    std::vector<double> tab, seg;
    GenHost h;
    if (!build_tables(t, tab, seg, h)) { t->gen_state = -1; return false; }
    std::vector<double> blk(tab);
    blk.insert(blk.end(), seg.begin(), seg.end());
    blk.insert(blk.end(), &t->dev.coeffs_s[0][0], &t->dev.coeffs_s[0][0] + 36);
    blk.insert(blk.end(), &t->dev.coeffs_e[0][0], &t->dev.coeffs_e[0][0] + 36);
    const double sc[5] = {h.e1, h.sa1, h.s_rcp, h.scn, h.e_rcp};
    blk.insert(blk.end(), sc, sc + 5);
    if (cudaMalloc(&t->d_m3, sizeof(double) * blk.size()) != cudaSuccess) { t->gen_state = -1; return false; }
    cudaMemcpy(t->d_m3, blk.data(), sizeof(double) * blk.size(), cudaMemcpyHostToDevice);
    t->gen_nseg = h.nseg;
    t->gen_scalars.assign(sc, sc + 5);
    t->gen_state = 1;
  }
  const int np = t->gen_nseg * S;
  o->tab = t->d_m3;
  o->seg = t->d_m3 + (size_t)NTAB * np;
  o->rc = o->seg + (size_t)t->gen_nseg * SEGW;
  std::memcpy(o->c, t->dev.coeffs, sizeof(double) * 9);
  o->e1 = t->gen_scalars[0]; o->sa1 = t->gen_scalars[1]; o->s_rcp = t->gen_scalars[2];
  o->scn = t->gen_scalars[3]; o->e_rcp = t->gen_scalars[4];
  o->mask = t->tap_mask;
  o->n_tds = t->n_tds; o->n_rhs = t->n_rhs;
  hh->nseg = t->gen_nseg; hh->np = np;
  return true;
}

constexpr size_t kSmemCap = 227 * 1024 - 1024;

template <class K>
int grid_for(x3d2c_ctx* ctx, K kernel, int threads, size_t smem, int tiles, int* grid) {
  X3D2C_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  X3D2C_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) {
    x3d2c::set_error("generic DistD2 kernel: the tile does not fit one SM");
    return X3D2C_EUNSUPPORTED;
  }
  *grid = std::min(tiles, num_sms(ctx) * per_sm);
  return X3D2C_OK;
}

template <int L>
int launch_tds_g(x3d2c_ctx* ctx, const TdsG& p, size_t smem) {
  const int threads = L * p.g.nseg;
  int grid = 0, rc;
  const unsigned m = p.o.mask;
#define X3D2C_TDS_G(MASK)                                                             \
  {                                                                                   \
    if ((rc = grid_for(ctx, tds_g_kernel<L, MASK>, threads, smem, p.g.tiles, &grid))) return rc; \
    tds_g_kernel<L, MASK><<<grid, threads, smem, ctx->stream>>>(p);                   \
  }
  if (m == 0x78u) X3D2C_TDS_G(0x78u)
  else if (m == 0x3Cu) X3D2C_TDS_G(0x3Cu)
  else if (m == 0x6Cu) X3D2C_TDS_G(0x6Cu)
  else if (m == 0x7Cu) X3D2C_TDS_G(0x7Cu)
  else X3D2C_TDS_G(0x1FFu)
#undef X3D2C_TDS_G
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

template <int L>
int launch_transeq_g(x3d2c_ctx* ctx, const TransG& p, size_t smem, bool compact) {
  const int threads = L * p.g.nseg;
  int grid = 0, rc;
  if (compact) {
    if ((rc = grid_for(ctx, transeq_g_kernel<L, 0x6Cu, 0x7Cu>, threads, smem, p.g.tiles, &grid))) return rc;
    transeq_g_kernel<L, 0x6Cu, 0x7Cu><<<grid, threads, smem, ctx->stream>>>(p);
  } else {
    if ((rc = grid_for(ctx, transeq_g_kernel<L, 0x1FFu, 0x1FFu>, threads, smem, p.g.tiles, &grid))) return rc;
    transeq_g_kernel<L, 0x1FFu, 0x1FFu><<<grid, threads, smem, ctx->stream>>>(p);
  }
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

// lanes per tile: the power of two that keeps most threads resident per SM (shared memory and the register file both
// limit the number of CTAs), wider tiles (longer contiguous rows) on a tie
int pick_lanes_g(int nseg, int max_threads, size_t bytes_per_lane, int regs_per_thread) {
  int best = 0, best_threads = 0;
  for (int L = 2; L <= 32; L <<= 1) {
    const int threads = L * nseg;
    const size_t smem = L * bytes_per_lane;
    if (threads > max_threads || smem > kSmemCap) continue;
    const int by_smem = (int)((227 * 1024) / (smem + 1024));
    const int by_regs = 65536 / (regs_per_thread * ((threads + 31) / 32 * 32));
    const int resident = threads * std::min(by_smem, std::max(by_regs, 1));
    if (resident >= best_threads) { best_threads = resident; best = L; }
  }
  return best;
}

}  // namespace

namespace x3d2c {

// tds_solve for a direction held by one rank, any operator (non-strict mode). X3D2C_EUNSUPPORTED -> tds_m1.cu.
int tds_g(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops) {
  static const bool disabled = std::getenv("X3D2C_NO_GENERIC") != nullptr;
  if (disabled || ctx->strict || ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist) return X3D2C_EUNSUPPORTED;
  TdsG p{};
  GenHost h;
  if (!gen_op(ctx, ops, &p.o, &h)) return X3D2C_EUNSUPPORTED;
  if (h.np > ctx->n_pad(dir)) return X3D2C_EUNSUPPORTED;
  // shared memory per lane: 2 tiles + carries + ends
  const size_t per_lane = sizeof(double) * (2 * (size_t)h.nseg * SP + 2 * h.nseg + 2);
  const int L = pick_lanes_g(h.nseg, 288, per_lane, 128);
  if (!L) return X3D2C_EUNSUPPORTED;
  p.in = u; p.out = du;
  p.g.np = h.np; p.g.nseg = h.nseg; p.g.n_pad = ctx->n_pad(dir);
  p.g.tiles = ctx->n_groups[dir] * (SZ / L);
  p.g.fd = h.nseg * SP * L;
  const size_t smem = per_lane * L;
  switch (L) {
    case 2: return launch_tds_g<2>(ctx, p, smem);
    case 4: return launch_tds_g<4>(ctx, p, smem);
    case 8: return launch_tds_g<8>(ctx, p, smem);
    case 16: return launch_tds_g<16>(ctx, p, smem);
    default: return launch_tds_g<32>(ctx, p, smem);
  }
}

// transeq for a direction held by one rank; out / in already permuted (index 0 = the line-aligned velocity).
int transeq_g(x3d2c_ctx* ctx, int dir, double* const out[3], const double* const in[3], double nu,
              const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym, const x3d2c_tdsops* der2nd,
              const x3d2c_tdsops* der2nd_sym) {
  // Measured on a B200 (profiles/r02_opbench_channel_512x257x512.txt): with 17 segments per 257-point line, 255
  // registers per thread and 24 table look-ups per row this kernel keeps only 4-5 warps per SM busy and needs 14.8 ms per
  // call at 512 x 257 x 512, the one-thread-per-line kernels 7.5 ms. It is therefore opt-in (X3D2C_TRANSEQ_GENERIC=1;
  // tests run it for parity); tds_solve uses its generic kernel by default (0.59 ms against 0.66 ms).
  static const bool enabled = std::getenv("X3D2C_TRANSEQ_GENERIC") != nullptr;
  static const bool disabled = std::getenv("X3D2C_NO_GENERIC") != nullptr;
  if (!enabled || disabled || ctx->strict || ctx->cfg.nproc_dir[dir - 1] > 1 || ctx->force_dist) return X3D2C_EUNSUPPORTED;
  TransG p{};
  GenHost h, h2;
  // component 0: (der1st, der1st_sym, der2nd); components 1, 2: (der1st_sym, der1st, der2nd_sym)  omp/backend.f90:246-260
  if (!gen_op(ctx, der1st, &p.du[0], &h) || !gen_op(ctx, der1st_sym, &p.dud[0], &h2) || h2.np != h.np) return X3D2C_EUNSUPPORTED;
  if (!gen_op(ctx, der2nd, &p.d2u[0], &h2) || h2.np != h.np) return X3D2C_EUNSUPPORTED;
  if (!gen_op(ctx, der1st_sym, &p.du[1], &h2) || !gen_op(ctx, der1st, &p.dud[1], &h2)) return X3D2C_EUNSUPPORTED;
  if (!gen_op(ctx, der2nd_sym, &p.d2u[1], &h2) || h2.np != h.np) return X3D2C_EUNSUPPORTED;
  if (h.np > ctx->n_pad(dir)) return X3D2C_EUNSUPPORTED;
  if (der2nd->h_stretch_correct != der2nd_sym->h_stretch_correct) return X3D2C_EUNSUPPORTED;  // one table serves both
  // stretch_correct of the second derivative: one more device table (np entries), kept with the der2nd handle
  auto* d2 = const_cast<x3d2c_tdsops*>(der2nd);
  if (!d2->d_stc) {
    std::vector<double> stc(h.np, 0.0);
    std::copy(d2->h_stretch_correct.begin(), d2->h_stretch_correct.end(), stc.begin());
    X3D2C_CHECK_CUDA(cudaMalloc(&d2->d_stc, sizeof(double) * h.np));
    X3D2C_CHECK_CUDA(cudaMemcpy(d2->d_stc, stc.data(), sizeof(double) * h.np, cudaMemcpyHostToDevice));
  }
  p.stc = d2->d_stc;
  p.nu = nu;
  const size_t per_lane = sizeof(double) * (6 * (size_t)h.nseg * SP + 6 * h.nseg + 6);
  const int L = pick_lanes_g(h.nseg, 256, per_lane, 255);
  if (!L) return X3D2C_EUNSUPPORTED;
  for (int f = 0; f < 3; ++f) { p.in[f] = in[f]; p.out[f] = out[f]; }
  p.g.np = h.np; p.g.nseg = h.nseg; p.g.n_pad = ctx->n_pad(dir);
  p.g.tiles = ctx->n_groups[dir] * (SZ / L);
  p.g.fd = h.nseg * SP * L;
  const size_t smem = per_lane * L;
  const bool compact = der1st->tap_mask == 0x6Cu && der2nd->tap_mask == 0x7Cu && der1st_sym->tap_mask == 0x6Cu &&
                       der2nd_sym->tap_mask == 0x7Cu;
  switch (L) {
    case 2: return launch_transeq_g<2>(ctx, p, smem, compact);
    case 4: return launch_transeq_g<4>(ctx, p, smem, compact);
    case 8: return launch_transeq_g<8>(ctx, p, smem, compact);
    case 16: return launch_transeq_g<16>(ctx, p, smem, compact);
    default: return launch_transeq_g<32>(ctx, p, smem, compact);
  }
}

}  // namespace x3d2c
