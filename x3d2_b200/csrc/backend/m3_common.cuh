// Shared pieces of the "m3" fast-path kernels (tds_m3.cu, transeq_m3.cu, m3_edge.cu).
//
// Method (see tds_m3.cu for the derivation): the compact operators of a periodic, uniform direction are
// Toeplitz, A = L U holds with the converged factors of the tdsops tables, and a line is cut into 16-point
// segments that are swept independently; segments are coupled through carries (the end value ze of the local
// forward sweep, the start value ys of the local backward sweep) that decay like (alpha fw)^16 per segment, so
// only the DMAX = 3 nearest segments on either side contribute (< 1e-18 relative).
//
// Rank-split directions use the same decomposition ACROSS ranks: the carries of the DMAX segments next to a
// rank boundary only depend on local data and the 4-row halo, so they are computed by a small edge kernel
// (m3_edge.cu), exchanged once, and the main kernel then treats the neighbour's segments like its own. This
// replaces the reference's 2x2 reduced systems (der_univ_dist / der_univ_subs, omp/kernels/distributed.f90:11-200)
// for periodic uniform directions; the result is the same solution of the same global periodic system.
#pragma once
#include <cmath>

#include "common.cuh"

namespace m3 {

constexpr int S = 16;         // points per segment
constexpr int SP = S + 1;     // rows per segment in shared memory (the pad row removes bank conflicts)
constexpr int LOG2S = 4;
constexpr int DMAX = 3;       // neighbouring segments that contribute to a carry
constexpr int HALO_ROWS = 8;  // rank-split lines: 4 rows before the line + 4 rows after, kept behind the segments
constexpr int EXP_ROWS = DMAX;  // rows per recurrence in an exchange buffer: to next, ze of the last 3 segments;
                                // to prev, what this rank's first segments add to yin of prev's last 3 segments
constexpr int EXT_ROWS = 2 * EXP_ROWS;  // shared-memory rows per recurrence: EXP_ROWS from prev + EXP_ROWS from next

struct Op {
  double cfw[9];            // scale * fw * coeffs; exact operators: the plain coeffs of the tdsops table
  double fs;                // exact operators: scale * fw, applied to the finished stencil sum (1 otherwise)
  int exact;                // the stencil sum is evaluated as the reference writes it (sten_exact)
  double a, cb;             // forward / backward propagators: a = -fw*alpha, cb = -bw
  double zw[DMAX], yw[DMAX];
  double om[2 * DMAX - 1];  // index m + DMAX - 1, m = d - d'
  double W[S], Cp[S];
  unsigned mask;
};

struct Geom {
  int n, n_pad, nseg, tiles, field_doubles;  // field_doubles includes the halo rows of a rank-split line
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

// The same sum in the reference's order and rounding (omp/kernels/distributed.f90:88-93: products rounded one by one,
// added from tap -4 to tap +4, no FMA contraction; zero taps skipped, which does not change a bit). The second
// derivative of a smooth field cancels to O(dx^2) of its terms, so at n = 512 / 1024 any other evaluation order
// differs from the reference by more than the 1e-12 parity bar - with this one the right-hand
// side is bit-identical and only the (well conditioned) sweeps round differently.
template <unsigned M>
__device__ __forceinline__ double sten_exact(const double (&c)[9], const double (&w)[9]) {
  double t = 0.0;
  bool first = true;
#pragma unroll
  for (int k = 0; k < 9; ++k)
    if (M & (1u << k)) {
      const double pr = __dmul_rn(c[k], w[k]);
      t = first ? pr : __dadd_rn(t, pr);
      first = false;
    }
  return t;
}

// sten_exact for a symmetric stencil (c[8 - k] == c[k], checked by the host): both mirror taps use the SAME coefficient
// register, so in the fully unrolled sweeps the product c[k] * u_i needed by row i - (k - 4) and again by row
// i + (k - 4) is one instruction (common sub-expression): 3 multiplications + 4 additions per point for compact6.
template <unsigned M>
__device__ __forceinline__ double sten_exact_sym(const double (&c)[9], const double (&w)[9]) {
  double t = 0.0;
  bool first = true;
#pragma unroll
  for (int k = 0; k < 9; ++k)
    if (M & (1u << k)) {
      const double pr = __dmul_rn(c[k <= 4 ? k : 8 - k], w[k]);
      t = first ? pr : __dadd_rn(t, pr);
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
  // Neighbour data of a rank-split tile: nr rows of the tile's L lanes from each of two (SZ, nr, G) arrays into
  // shared [2 nr][L] (g0's rows first).
  __device__ __forceinline__ void load_rows2(double* sm, const double* g0, const double* g1, int nr, int tile) const {
    constexpr int tpg = SZ / L, cpr = L / 2;
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    const size_t go = (size_t)grp * nr * SZ + l0;
    for (int idx = threadIdx.x; idx < 2 * nr * cpr; idx += blockDim.x) {
      const int row = idx / cpr, c = 2 * (idx - row * cpr);
      const bool second = row >= nr;
      const double* src = (second ? g1 : g0) + go + (size_t)(second ? row - nr : row) * SZ + c;
      cp_async16(sm + row * L + c, src);
    }
  }
  // Halo rows of nf fields: (SZ, 4, nf, G) arrays hs (rows before the line) and he (rows after) into the 8 halo
  // rows behind the segments of each field tile (field stride fd, halo offset hoff).
  __device__ __forceinline__ void load_halos(double* sm, int nf, int fd, int hoff, const double* hs, const double* he,
                                             int tile) const {
    constexpr int tpg = SZ / L, cpr = L / 2;
    const int grp = tile / tpg, l0 = (tile - grp * tpg) * L;
    const size_t go = (size_t)grp * 4 * nf * SZ + l0;
    for (int idx = threadIdx.x; idx < 8 * nf * cpr; idx += blockDim.x) {
      const int row = idx / cpr, c = 2 * (idx - row * cpr);
      const int f = row >> 3, r = row & 3;
      const double* src = ((row & 4) ? he : hs) + go + (size_t)(f * 4 + r) * SZ + c;
      cp_async16(sm + f * fd + hoff + (row & 7) * L + c, src);
    }
  }
};

// Carries of one recurrence for segment q: zin (from the left), yin (from the right).
// ze / ys: shared-memory offsets (lane applied) of the carries of segment 0, `stride` doubles between segments.
// DIST: segments beyond the line ends belong to the neighbouring ranks: extp rows 0..2 hold ze of prev's last three
// segments, extn rows 0..2 hold the sum of all terms that next's first segments contribute to yin of this rank's
// last three segments (computed by next's edge kernel), lane applied. Otherwise the line is periodic and segment
// indices wrap.
// EXT = false (rank-split lines only): the neighbouring ranks' terms are left out; carries_ext() supplies them later.
template <int L, bool DIST, bool EXT = true>
__device__ __forceinline__ void carries(const int ze, const int ys, const int stride, const int extp, const int extn,
                                        const Op& o, const int q, const int nseg, double& zin, double& yin) {
  double zv[2 * DMAX];  // ze(q - DMAX .. q + DMAX - 1)
#pragma unroll
  for (int t = 0; t < 2 * DMAX; ++t) {
    int s = q - DMAX + t;
    if (DIST) {
      const bool beyond = s >= nseg || (!EXT && s < 0);
      const int a = (EXT && s < 0) ? extp + (s + DMAX) * L : ze + (beyond ? 0 : s) * stride;
      zv[t] = smem[a];
      if (beyond) zv[t] = 0.0;
    } else {
      if (s < 0) s += nseg;
      if (s >= nseg) s -= nseg;
      zv[t] = smem[ze + s * stride];
    }
  }
  zin = 0.0;
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) zin = fma(o.zw[d - 1], zv[DMAX - d], zin);
  yin = 0.0;
  if (DIST && EXT) {
    const int r = q - (nseg - DMAX);
    if (r >= 0) yin = smem[extn + r * L];
  }
#pragma unroll
  for (int d = 1; d <= DMAX; ++d) {
    int s = q + d;
    if (DIST) {
      const bool beyond = s >= nseg;
      double v = smem[ys + (beyond ? 0 : s) * stride];
      if (beyond) v = 0.0;
      yin = fma(o.yw[d - 1], v, yin);
    } else {
      if (s >= nseg) s -= nseg;
      yin = fma(o.yw[d - 1], smem[ys + s * stride], yin);
    }
  }
#pragma unroll
  for (int m = -(DMAX - 1); m <= DMAX - 1; ++m) yin = fma(o.om[m + DMAX - 1], zv[DMAX + m], yin);
}

// ---- in-kernel carry exchange of a rank-split line --------------------------------------------------------------
// Instead of a separate edge kernel that repeats the sweeps of the six segments next to the rank boundaries, the main
// kernel itself hands its boundary carries to the neighbours: right after the local sweeps of a tile the threads of the
// last three segments store their ze into the next rank's receive slots and the threads of the first three segments
// what they add to yin of the previous rank's last segments (peer stores over NVLink), then poll their own slots for the
// neighbours' values of the same tile. Every 8-byte slot validates itself: it holds kCarrySentinel until the neighbour
// has written it and is put back to the sentinel by its (only) reader, so no flags or fences are needed. Two sets of
// slots alternate between consecutive exchanges: a neighbour can be at most one operator ahead (it needs this rank's
// carries of that operator to finish it), so it never writes a slot that is still unread.
struct InlineCarries {
  double *to_prev, *to_next;      // the neighbours' from_next / from_prev (peer mapped; this rank's own with one rank)
  double *from_prev, *from_next;  // (SZ, EXP_ROWS, ns, G); null: carries come from the edge kernel
};

__device__ __forceinline__ void push_carry(double* dst, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
}
// N slots at once: all loads are in flight together, so that a poll whose data has arrived costs one trip to L2
template <int N>
__device__ __forceinline__ void take_carries(double* const (&src)[N], double (&out)[N]) {
  unsigned long long v[N];
  long long t0 = 0;
  for (unsigned spins = 0;; ++spins) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v[i]) : "l"(src[i]) : "memory");
    bool all = true;
#pragma unroll
    for (int i = 0; i < N; ++i) all = all && v[i] != x3d2c::kCarrySentinel;
    if (all) break;
    if (spins == 64) t0 = clock64();
    if (spins > 64) {
      __nanosleep(40);
      if (clock64() - t0 > 40000000000ll) {
        printf("x3d2c: carry exchange timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(src[i]), "l"(x3d2c::kCarrySentinel) : "memory");
    out[i] = __longlong_as_double((long long)v[i]);
  }
}

// N recurrences whose slots are `sstride` doubles apart and whose shared rows are `xstride` doubles apart
template <int L, int N>
__device__ __forceinline__ void poll_carries_inline(const InlineCarries& c, const size_t slot, const size_t sstride,
                                                    const int q, const int nseg, const int xp, const int xn,
                                                    const int xstride) {
  const int r = q - (nseg - DMAX);
  if (r >= 0 || q < DMAX) {
    const bool nx = r >= 0;
    double* const base = (nx ? c.from_next : c.from_prev) + slot + (size_t)(nx ? r : q) * SZ;
    double* src[N];
    double v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) src[i] = base + i * sstride;
    take_carries<N>(src, v);
    const int x = nx ? xn + r * L : xp + q * L;
#pragma unroll
    for (int i = 0; i < N; ++i) smem[x + i * xstride] = v[i];
  }
}
// One recurrence of one tile. ze / ys: shared offsets of the carries of segment 0 (lane applied, L doubles between
// segments); slot: offset of row 0 of this recurrence for this thread's lane in the (SZ, EXP_ROWS, ns, G) arrays.
// Pushing and polling are separate so that a kernel can put other work between them (the neighbour needs about as
// long as this rank to reach the same tile, and the stores take a few microseconds over NVLink).
template <int L>
__device__ __forceinline__ void push_carries_inline(const InlineCarries& c, const size_t slot, const int ze, const int ys,
                                                    const Op& o, const int q, const int nseg) {
  const int r = q - (nseg - DMAX);
  if (r >= 0) {
    push_carry(c.to_next + slot + (size_t)r * SZ, smem[ze + q * L]);
  } else if (q < DMAX) {
    // what this rank's first segments add to yin of the previous rank's segment nseg' - 3 + q (see carries())
    double acc = 0.0;
#pragma unroll
    for (int d = 1; d <= DMAX; ++d)
      if (q + d - DMAX >= 0) acc = fma(o.yw[d - 1], smem[ys + (q + d - DMAX) * L], acc);
#pragma unroll
    for (int m = 1; m <= DMAX - 1; ++m)
      if (q + m - DMAX >= 0) acc = fma(o.om[m + DMAX - 1], smem[ze + (q + m - DMAX) * L], acc);
    push_carry(c.to_prev + slot + (size_t)q * SZ, acc);
  }
}

// The terms of carries() that come from the neighbouring ranks (zero for all but the first / last DMAX segments)
template <int L>
__device__ __forceinline__ void carries_ext(const int extp, const int extn, const Op& o, const int q, const int nseg,
                                            double& zin, double& yin) {
  zin = 0.0;
  yin = 0.0;
  const int r = q - (nseg - DMAX);
  if (r >= 0) yin = smem[extn + r * L];
  if (q < DMAX) {
#pragma unroll
    for (int d = 1; d <= DMAX; ++d)
      if (q - d < 0) zin = fma(o.zw[d - 1], smem[extp + (q - d + DMAX) * L], zin);
#pragma unroll
    for (int m = -(DMAX - 1); m <= -1; ++m)
      if (q + m < 0) yin = fma(o.om[m + DMAX - 1], smem[extp + (q + m + DMAX) * L], yin);
  }
}

// window element t (row j0 - 4 + t, t = 0..S+7) given the bases of the previous, own and next segment
template <int L>
__device__ __forceinline__ int woff(int t, int bm, int b0, int bp) {
  return t < 4 ? bm + (S - 4 + t) * L : (t < S + 4 ? b0 + (t - 4) * L : bp + (t - S - 4) * L);
}

// segment bases of thread (lane l, segment q). Rank-split lines read the rows before / after the line from the
// halo rows stored behind the last segment (4 "before" rows, then 4 "after" rows).
template <int L, bool DIST>
__device__ __forceinline__ void segment_bases(int q, int l, int nseg, int& bm, int& b0, int& bp) {
  const int qm = q == 0 ? nseg - 1 : q - 1, qp = q == nseg - 1 ? 0 : q + 1;
  bm = qm * SP * L + l;
  b0 = q * SP * L + l;
  bp = qp * SP * L + l;
  if (DIST) {
    const int h = nseg * SP * L + l;
    if (q == 0) bm = h - (S - 4) * L;
    if (q == nseg - 1) bp = h + 4 * L;
  }
}

// ------------------------------------------------------------------------------------------------ host side
bool make_op(const x3d2c_tdsops* t, double scale, bool dist, Op* o, bool exact = true);
bool same_tables(const x3d2c_tdsops* a, const x3d2c_tdsops* b);
int num_sms(const x3d2c_ctx* ctx);

// Exchange buffers of a rank-split direction, carved from ctx->halo. Layouts:
//   halo_*:  (SZ, 4 rows, nf fields, G)      carr_*: (SZ, EXP_ROWS, ns recurrences, G)
struct DistBufs {
  double *halo_send_s, *halo_send_e, *halo_recv_s, *halo_recv_e;
  double *carr_to_prev, *carr_to_next, *carr_from_prev, *carr_from_next;
};
constexpr int kDistRows = 4 * (3 * 4) + 4 * (9 * EXP_ROWS);  // rows of SZ*G doubles that DistBufs needs
static_assert(kDistRows + x3d2c::kHaloRowsRecv2 + x3d2c::kHaloRowsInline <= x3d2c::kHaloRows, "ctx->halo (ctx.cu) is too small for the rank-split exchange buffers");
DistBufs carve_dist(x3d2c_ctx* ctx, int recv_set = 0);  // recv_set 1: the second set of receive buffers
bool dist_supported(const x3d2c_ctx* ctx, int dir, int n);

// Edge kernel input: nf fields with ns / nf recurrences each (m3_edge.cu: edge_kernel);
// recurrence index = field * (ns / nf) + k.
struct EdgeParams {
  int n, n_pad, nseg, G, ns, nf, transeq;
  const double* f[3];
  Op ops[3];
  const double *halo_s, *halo_e;  // received halos (SZ, 4, nf, G)
  double *to_prev, *to_next;      // (SZ, EXP_ROWS, ns, G)
};
// packs the halos of nf fields, exchanges them, computes and exchanges the boundary carries
// (with peer stores the receive buffers alternate between two sets: `b` is updated to the set this exchange filled).
// inl != null and the in-kernel carry exchange is possible (peer-mapped buffers, or a single rank that is its own
// neighbour): only the halos are exchanged and *inl describes the slots of this exchange; otherwise inl->from_prev = null.
int exchange_edges(x3d2c_ctx* ctx, int dir, const double* const* fields, int nf, EdgeParams& ep, DistBufs& b,
                   InlineCarries* inl = nullptr);

}  // namespace m3
