// Host layer of the cuda_c backend: tdsops_t coefficient setup (the role of /root/reference/src/tdsops.f90:63-931:
// tdsops_init, deriv_1st, deriv_2nd, interpl_mid, stagder_1st, preprocess_dist). The pentadiagonal scheme
// (:235-250,322-335,383-395,971-1103) is not reachable from tds_solve / transeq and is omitted.
//
// The reference writes every boundary row out by hand. This generator DERIVES them instead:
//   * an operator is its interior 9-tap stencil, stored symbolically as small integer multiples of the scheme
//     constants (a, b, c, d), the grids it reads from / writes to (vertices or cell midpoints) and the parity of the
//     field it acts on at a free-slip wall;
//   * Neumann rows follow by folding the taps that stick out of the domain back onto their mirror points (even or odd
//     extension about the wall; the staggered derivative towards the cells extends oddly about the wall VALUE) - the
//     multipliers are exact integers, and a weight is evaluated as sum_q m_q k_q from q = a to d, which is the order
//     the reference spells its sums in, so the tables are bit-identical to tdsops_init's;
//   * Dirichlet rows are Lele's one-sided closures (literal constants), written once and mirrored for the far wall;
//   * the left-hand side follows from the same parities (alpha f_0 folds onto f_2 or f_1).
// tests/test_host_logic.py compares the tables bit for bit with the oracle's line-by-line transcription of the
// Fortran, tests/test_independent_pin.py pins both to dense long-double solves of the published schemes.
// In a Fortran build this file is not needed: the reference's own tdsops_init computes the tables and the
// iso_c_binding shim passes them to x3d2c_tdsops_create. Arrays are 1-based (index 0 unused).
#pragma once
#include <array>

#include "common.hpp"

namespace x3d2h {

struct Tdsops {
  // tdsops.f90:27-41
  std::vector<double> dist_fw, dist_bw, dist_sa, dist_sc, dist_af;
  std::vector<double> stretch, stretch_correct;
  double coeffs[10];        // coeffs(1:9)
  double coeffs_s[5][10];   // coeffs_s(k, i) stored as [i][k], i = 1..4 rows, k = 1..9 taps
  double coeffs_e[5][10];
  double alpha = 0, a = 0, b = 0, c = 0, d = 0;
  bool periodic = false;
  int n_tds = 0, n_rhs = 0, move = 0, n_halo = 4;
};

namespace detail {

// one tap weight: m[0] a + m[1] b + m[2] c + m[3] d with integer m
struct Weight {
  int m[4] = {0, 0, 0, 0};
  void add(const Weight& o, int s) { for (int q = 0; q < 4; ++q) m[q] += s * o.m[q]; }
  bool zero() const { return !(m[0] | m[1] | m[2] | m[3]); }
  double eval(const double k[4]) const {
    double acc = 0.0;
    bool first = true;
    for (int q = 0; q < 4; ++q) {
      if (!m[q]) continue;
      const double term = m[q] * k[q];
      acc = first ? term : acc + term;
      first = false;
    }
    return acc;
  }
};
using Row = std::array<Weight, 9>;  // taps -4 .. +4 at index tap + 4

enum Grid { VERTEX_GRID = 0, CELL_GRID = 1 };

struct OpDef {
  double alpha = 0, k[4] = {0, 0, 0, 0};
  Row interior;
  Grid in = VERTEX_GRID, out = VERTEX_GRID;
  int parity_in = 1, parity_out = 1;  // +1 even, -1 odd about a free-slip wall
  bool about_wall_value = false;      // odd extension u(-x) = 2 u(0) - u(x) instead of u(-x) = -u(x)
  int order = 0;                      // derivative order (parity of the mirror image of a one-sided closure)
  void tap(int t, int q, int mult) { interior[t + 4].m[q] += mult; }
};

// one-sided closure row at a Dirichlet wall, written for the near (start) wall
struct Closure {
  double sa, sc;       // sub / super diagonal of the left-hand side
  double w[9];         // taps -4 .. +4, already divided by delta^order
};

inline void write_row(double* dst, const Row& r, const double k[4]) {
  for (int t = 0; t < 9; ++t) dst[t + 1] = r[t].eval(k);
}
inline void write_closure(double* dst, const Closure& c, bool mirrored, int order) {
  for (int t = 0; t < 9; ++t) {
    const double v = c.w[mirrored ? 8 - t : t];
    dst[t + 1] = (mirrored && (order & 1)) ? -v : v;
  }
}

// Row `row` of a line whose input points are 1..n_in (the wall sits on input vertex 1 / n_in for a vertex grid, half
// a cell outside input cell 1 / n_in for a cell grid): taps that leave the domain on the side `lo` are folded back.
inline Row fold_row(const OpDef& op, int row, int n_in, bool lo) {
  Row r;
  for (int t = -4; t <= 4; ++t) {
    const Weight& w = op.interior[t + 4];
    if (w.zero()) continue;
    const int j = row + t;
    const bool outside = lo ? j < 1 : j > n_in;
    if (!outside) {
      r[t + 4].add(w, 1);
      continue;
    }
    // mirror image of point j: vertices reflect about the wall vertex, cells about the wall face
    const int jm = lo ? (op.in == VERTEX_GRID ? 2 - j : 1 - j) : (op.in == VERTEX_GRID ? 2 * n_in - j : 2 * n_in + 1 - j);
    r[jm - row + 4].add(w, op.parity_in);
    if (op.about_wall_value && op.parity_in < 0) r[(lo ? 1 : n_in) - row + 4].add(w, 2);
  }
  return r;
}

// DistD2 factorisation (Algorithm 3 of doi:10.1109/MCSE.2021.3130544 as pre-computed by tdsops.f90:874-931):
// on entry sa / sc are the sub / super diagonals and `diag` the main diagonal of the n_tds x n_tds system.
inline void factorise_dist(Tdsops& t, const std::vector<double>& diag) {
  const int n = t.n_tds;
  double *sa = t.dist_sa.data(), *sc = t.dist_sc.data(), *fw = t.dist_fw.data(), *bw = t.dist_bw.data(),
         *af = t.dist_af.data();
  // rows 1, 2: normalised, kept out of the forward elimination (they couple to the previous rank)
  for (int i : {1, 2}) {
    sa[i] /= diag[i];
    sc[i] /= diag[i];
    bw[i] = sc[i];
    af[i] = 1.0 / diag[i];
  }
  // forward elimination: fw = pivot reciprocal, af = the original sub-diagonal, sa / sc = fill-in towards row 1 / n
  for (int i = 3; i <= n; ++i) {
    const double sub = sa[i];
    fw[i] = 1.0 / (diag[i] - sub * sc[i - 1]);
    af[i] = sub;
    sa[i] = -fw[i] * sub * sa[i - 1];
    sc[i] = fw[i] * sc[i];
  }
  // backward elimination
  for (int i = n - 2; i >= 2; --i) {
    sa[i] -= sc[i] * sa[i + 1];
    bw[i] = sc[i];
    sc[i] = -sc[i] * sc[i + 1];
  }
  // row 1 last: its pivot reciprocal is parked in fw(1)
  fw[1] = 1.0 / (1.0 - sc[1] * sa[2]);
  sa[1] = fw[1] * sa[1];
  sc[1] = -fw[1] * sc[1] * sc[2];
}

// Fills coeffs, coeffs_s / coeffs_e and the tridiagonal left-hand side for the two boundary conditions, then factorises.
inline void assemble(Tdsops& t, const OpDef& op, int bc_start, int bc_end, const Closure* closures, int n_closures) {
  const int n = t.n_tds, n_in = t.n_rhs;
  t.alpha = op.alpha; t.a = op.k[0]; t.b = op.k[1]; t.c = op.k[2]; t.d = op.k[3];
  write_row(t.coeffs, op.interior, op.k);
  for (int i = 1; i <= 4; ++i) { write_row(t.coeffs_s[i], op.interior, op.k); write_row(t.coeffs_e[i], op.interior, op.k); }
  std::fill(t.dist_sa.begin(), t.dist_sa.end(), op.alpha);
  std::fill(t.dist_sc.begin(), t.dist_sc.end(), op.alpha);
  std::vector<double> diag(t.n_rhs + 1, 1.0);
  const bool odd_on_wall = op.out == VERTEX_GRID && op.parity_out < 0;  // the result vanishes on a free-slip wall

  if (bc_start == BC_NEUMANN) {
    for (int i = 1; i <= 4; ++i) write_row(t.coeffs_s[i], fold_row(op, i, n_in, true), op.k);
    t.dist_sa[1] = 0.0;
    if (op.out == VERTEX_GRID) {  // alpha f_0 folds onto f_2
      t.dist_sc[1] = odd_on_wall ? 0.0 : op.alpha + op.alpha;
      if (odd_on_wall) for (int k = 1; k <= 9; ++k) t.coeffs_s[1][k] = 0.0;
    } else {                      // alpha f_0 folds onto f_1
      diag[1] = 1.0 + op.parity_out * op.alpha;
    }
  } else if (bc_start == BC_DIRICHLET) {
    if (!closures) fail(op.order ? "Dirichlet BC is not supported for midpoint derivatives!"
                                 : "Dirichlet BC is not supported for midpoint interpolations!");
    for (int i = 1; i <= n_closures; ++i) {
      write_closure(t.coeffs_s[i], closures[i - 1], false, op.order);
      t.dist_sa[i] = closures[i - 1].sa;
      t.dist_sc[i] = closures[i - 1].sc;
    }
  }

  if (bc_end == BC_NEUMANN) {
    // a walled line has one cell less than vertices: the last input cell of a cell-grid operator is n_in - 1
    const int n_last = op.in == CELL_GRID ? n_in - 1 : n_in;
    for (int i = 1; i <= 4; ++i) write_row(t.coeffs_e[i], fold_row(op, n_in - 4 + i, n_last, false), op.k);
    t.dist_sc[n] = 0.0;
    if (op.out == VERTEX_GRID) {
      t.dist_sa[n] = odd_on_wall ? 0.0 : op.alpha + op.alpha;
      if (odd_on_wall) for (int k = 1; k <= 9; ++k) t.coeffs_e[4][k] = 0.0;
    } else {
      diag[n] = 1.0 + op.parity_out * op.alpha;
      // one more input vertex than output cells: the row behind the last cell produces nothing
      if (n_in > n) for (int k = 1; k <= 9; ++k) t.coeffs_e[4][k] = 0.0;
    }
  } else if (bc_end == BC_DIRICHLET) {
    if (!closures) fail(op.order ? "Dirichlet BC is not supported for midpoint derivatives!"
                                 : "Dirichlet BC is not supported for midpoint interpolations!");
    for (int i = 1; i <= n_closures; ++i) {
      write_closure(t.coeffs_e[5 - i], closures[i - 1], true, op.order);
      t.dist_sa[n + 1 - i] = closures[i - 1].sc;  // mirrored: sub and super diagonal swap
      t.dist_sc[n + 1 - i] = closures[i - 1].sa;
    }
  }
  factorise_dist(t, diag);
}

inline OpDef first_derivative(double delta, const std::string& scheme, bool sym) {
  if (scheme != "compact6") fail("scheme is not defined");
  OpDef op;
  op.order = 1;
  op.alpha = 1.0 / 3.0;
  op.k[0] = 7.0 / 9.0 / delta;
  op.k[1] = 1.0 / 36.0 / delta;
  for (int q = 0; q < 3; ++q) { op.tap(q + 1, q, 1); op.tap(-(q + 1), q, -1); }  // a, b, c on +-1, +-2, +-3
  op.parity_in = sym ? 1 : -1;
  op.parity_out = -op.parity_in;
  return op;
}

inline OpDef second_derivative(double delta, const std::string& scheme, bool sym, bool has_hv, double c_nu, double nu0_nu) {
  OpDef op;
  op.order = 2;
  const double d2 = delta * delta;
  if (scheme == "compact6") {
    op.alpha = 2.0 / 11.0;
    op.k[0] = 12.0 / 11.0 / d2;
    op.k[1] = 3.0 / 44.0 / d2;
  } else if (scheme == "compact6-hyperviscous") {
    if (!has_hv) fail("compact6-hyperviscous requires c_nu and nu0_nu");
    // Lamballais et al. (2011) spectral-vanishing-viscosity-like second derivative, tdsops.f90:443-457
    const double dpis3 = 2.0 * pi / 3.0;
    const double xnpi2 = pi * pi * (1.0 + nu0_nu);
    const double xmpi2 = dpis3 * dpis3 * (1.0 + c_nu * nu0_nu);
    const double den = 405.0 * xnpi2 - 640.0 * xmpi2 + 144.0;
    op.alpha = 0.5 - (320.0 * xmpi2 - 1296.0) / den;
    op.k[0] = -(4329.0 * xnpi2 / 8.0 - 32.0 * xmpi2 - 140.0 * xnpi2 * xmpi2 + 286.0) / den / d2;
    op.k[1] = (2115.0 * xnpi2 - 1792.0 * xmpi2 - 280.0 * xnpi2 * xmpi2 + 1328.0) / den / (4.0 * d2);
    op.k[2] = -(7695.0 * xnpi2 / 8.0 + 288.0 * xmpi2 - 180.0 * xnpi2 * xmpi2 - 2574.0) / den / (9.0 * d2);
    op.k[3] = (198.0 * xnpi2 + 128.0 * xmpi2 - 40.0 * xnpi2 * xmpi2 - 736.0) / den / (16.0 * d2);
  } else {
    fail("scheme is not defined");
  }
  for (int q = 0; q < 4; ++q) { op.tap(q + 1, q, 1); op.tap(-(q + 1), q, 1); op.tap(0, q, -2); }
  op.parity_in = op.parity_out = sym ? 1 : -1;
  return op;
}

inline OpDef midpoint_interpolation(const std::string& scheme, const std::string& from_to) {
  OpDef op;
  op.order = 0;
  double& alpha = op.alpha;
  double &a = op.k[0], &b = op.k[1], &c = op.k[2], &d = op.k[3];
  if (scheme == "classic") {
    alpha = 0.3; a = 0.75; b = 0.05;
  } else if (scheme == "optimised") {
    alpha = 0.461658; d = 0.00146508;
    a = (75.0 + 70.0 * alpha - 640.0 * d) / 128.0;
    b = (-25.0 + 126.0 * alpha + 2304.0 * d) / 256.0;
    c = (3.0 - 10.0 * alpha - 1280.0 * d) / 256.0;
  } else if (scheme == "aggressive") {
    alpha = 0.49;
    a = (75.0 + 70.0 * alpha) / 128.0;
    b = (-25.0 + 126.0 * alpha) / 256.0;
    c = (3.0 - 10.0 * alpha) / 256.0;
  } else {
    fail("scheme is not defined");
  }
  // pairs of points at +-1/2, +-3/2, +-5/2, +-7/2 cells from the output point
  const int first = from_to == "v2p" ? 0 : -1;  // tap of the nearest input point on the low side
  for (int q = 0; q < 4; ++q) { op.tap(first - q, q, 1); op.tap(first + 1 + q, q, 1); }
  op.in = from_to == "v2p" ? VERTEX_GRID : CELL_GRID;
  op.out = from_to == "v2p" ? CELL_GRID : VERTEX_GRID;
  op.parity_in = op.parity_out = 1;  // the interpolated quantities are even about a free-slip wall
  return op;
}

inline OpDef staggered_derivative(double delta, const std::string& scheme, const std::string& from_to) {
  if (scheme != "compact6") fail("scheme is not defined");
  OpDef op;
  op.order = 1;
  op.alpha = 9.0 / 62.0;
  op.k[0] = 63.0 / 62.0 / delta;
  op.k[1] = 17.0 / 62.0 / 3.0 / delta;
  const int first = from_to == "v2p" ? 0 : -1;
  for (int q = 0; q < 2; ++q) { op.tap(first - q, q, -1); op.tap(first + 1 + q, q, 1); }
  op.in = from_to == "v2p" ? VERTEX_GRID : CELL_GRID;
  op.out = from_to == "v2p" ? CELL_GRID : VERTEX_GRID;
  if (from_to == "v2p") {  // wall-normal velocity towards the cells: odd about its wall value, the derivative is even
    op.parity_in = -1; op.parity_out = 1; op.about_wall_value = true;
  } else {                 // pressure-like quantity back to the vertices: even, the derivative vanishes on the wall
    op.parity_in = 1; op.parity_out = -1;
  }
  return op;
}

}  // namespace detail

// tdsops.f90:63-203.  stretch / stretch_correct may be null (=> 1 / 0).
inline Tdsops tdsops_init(int n_tds, double delta, const std::string& operation,
                          const std::string& scheme, int bc_start, int bc_end,
                          const double* stretch = nullptr, const double* stretch_correct = nullptr,
                          int n_halo = 4, const std::string& from_to = "", bool sym = false,
                          bool has_hv = false, double c_nu = 0, double nu0_nu = 0) {
  using namespace detail;
  Tdsops t;
  t.n_tds = n_tds;
  // towards the cells of a walled line there is one more input vertex than output cells (tdsops.f90:114-123)
  const bool walled_end = bc_end == BC_NEUMANN || bc_end == BC_DIRICHLET;
  t.n_rhs = (from_to == "v2p" && walled_end) ? n_tds + 1 : n_tds;
  t.n_halo = n_halo;
  const int n = t.n_rhs;
  // entries the factorisation never assigns (dist_fw(2), dist_bw(n-1:n), entry n_rhs > n_tds) stay zero
  for (auto* v : {&t.dist_fw, &t.dist_bw, &t.dist_sa, &t.dist_sc, &t.dist_af}) v->assign(n + 1, 0.0);
  t.stretch.assign(n_tds + 1, 1.0);
  t.stretch_correct.assign(n_tds + 1, 0.0);
  if (stretch) std::copy(stretch, stretch + n_tds, t.stretch.begin() + 1);
  if (stretch_correct) std::copy(stretch_correct, stretch_correct + n_tds, t.stretch_correct.begin() + 1);
  t.periodic = bc_start == BC_PERIODIC && bc_end == BC_PERIODIC;
  std::fill(&t.coeffs[0], &t.coeffs[0] + 10, 0.0);
  std::fill(&t.coeffs_s[0][0], &t.coeffs_s[0][0] + 50, 0.0);
  std::fill(&t.coeffs_e[0][0], &t.coeffs_e[0][0] + 50, 0.0);

  if (operation == "first-deriv") {
    if (n_halo < 2) fail("First derivative require n_halo >= 2");
    // Lele (1992): f'_1 + 2 f'_2 = (-5 u_1 + 4 u_2 + u_3) / (2 h);  f'_1 / 4 + f'_2 + f'_3 / 4 = 3 (u_3 - u_1) / (4 h)
    const Closure cl[2] = {{0.0, 2.0, {0, 0, 0, 0, -2.5 / delta, 2.0 / delta, 0.5 / delta, 0, 0}},
                           {0.25, 0.25, {0, 0, 0, -0.75 / delta, 0.0 / delta, 0.75 / delta, 0, 0, 0}}};
    assemble(t, first_derivative(delta, scheme, sym), bc_start, bc_end, cl, 2);
  } else if (operation == "second-deriv") {
    if (n_halo < 4) fail("Second derivative require n_halo >= 4");
    const double d2 = delta * delta;
    // f''_1 + 11 f''_2 = (13 u_1 - 27 u_2 + 15 u_3 - u_4) / h^2;  f''_1 / 10 + f''_2 + f''_3 / 10 = 6 (u_1 - 2 u_2 + u_3) / (5 h^2);
    // rows 3 and 4: the plain sixth-order scheme (alpha = 2/11), whatever the interior scheme is
    const double b6 = 3.0 / 44.0 / d2, a6 = 12.0 / 11.0 / d2, c6 = -2.0 * (b6 + a6);
    const Closure cl[4] = {{0.0, 11.0, {0, 0, 0, 0, 13.0 / d2, -27.0 / d2, 15.0 / d2, -1.0 / d2, 0}},
                           {0.1, 0.1, {0, 0, 0, 1.2 / d2, -2.4 / d2, 1.2 / d2, 0, 0, 0}},
                           {2.0 / 11.0, 2.0 / 11.0, {0, 0, b6, a6, c6, a6, b6, 0, 0}},
                           {2.0 / 11.0, 2.0 / 11.0, {0, 0, b6, a6, c6, a6, b6, 0, 0}}};
    assemble(t, second_derivative(delta, scheme, sym, has_hv, c_nu, nu0_nu), bc_start, bc_end, cl, 4);
  } else if (operation == "interpolate") {
    if (n_halo < 4) fail("Interpolation require n_halo >= 4");
    assemble(t, midpoint_interpolation(scheme, from_to), bc_start, bc_end, nullptr, 0);
  } else if (operation == "stag-deriv") {
    if (n_halo < 2) fail("Staggared deriv require n_halo >= 2");
    assemble(t, staggered_derivative(delta, scheme, from_to), bc_start, bc_end, nullptr, 0);
  } else {
    fail("operation is not defined");
  }
  t.move = from_to == "v2p" ? 1 : (from_to == "p2v" ? -1 : 0);
  return t;
}

// tdsops.f90:51-59
struct Dirps {
  Tdsops der1st, der1st_sym, der2nd, der2nd_sym, stagder_v2p, stagder_p2v, interpl_v2p, interpl_p2v;
  int dir = 0;
};

}  // namespace x3d2h
