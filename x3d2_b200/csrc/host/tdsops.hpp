// Host layer of the cuda_c backend: tdsops_t coefficient setup.
// Mirrors /root/reference/src/tdsops.f90:63-931 (tdsops_init, deriv_1st, deriv_2nd, interpl_mid,
// stagder_1st, preprocess_dist); the pentadiagonal scheme (:235-250,322-335,383-395,971-1103) is not
// reachable from tds_solve/transeq and is omitted.
// In a Fortran build this file is not needed: the reference's own tdsops_init computes the tables and
// the iso_c_binding shim passes them to x3d2c_tdsops_create.
// All arrays are 1-based (index 0 unused) so the formulas read like the Fortran source.
#pragma once
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
inline void set9(double* dst, std::initializer_list<double> v) {
  int k = 1;
  for (double x : v) dst[k++] = x;
}
inline void scale9(double* dst, double s) {
  for (int k = 1; k <= 9; ++k) dst[k] = dst[k] / s;
}
inline void copy9(double* dst, const double* src) {
  for (int k = 1; k <= 9; ++k) dst[k] = src[k];
}

// tdsops.f90:874-931
inline void preprocess_dist(Tdsops& t, const std::vector<double>& dist_b) {
  auto &sa = t.dist_sa, &sc = t.dist_sc, &fw = t.dist_fw, &bw = t.dist_bw, &af = t.dist_af;
  for (int i = 1; i <= 2; ++i) {
    sa[i] = sa[i] / dist_b[i];
    sc[i] = sc[i] / dist_b[i];
    bw[i] = sc[i];
    af[i] = 1.0 / dist_b[i];
  }
  for (int i = 3; i <= t.n_tds; ++i) {
    fw[i] = 1.0 / (dist_b[i] - sa[i] * sc[i - 1]);
    af[i] = sa[i];
    sa[i] = -fw[i] * sa[i] * sa[i - 1];
    sc[i] = fw[i] * sc[i];
  }
  for (int i = t.n_tds - 2; i >= 2; --i) {
    sa[i] = sa[i] - sc[i] * sa[i + 1];
    bw[i] = sc[i];
    sc[i] = -sc[i] * sc[i + 1];
  }
  fw[1] = 1.0 / (1.0 - sc[1] * sa[2]);
  sa[1] = fw[1] * sa[1];
  sc[1] = -fw[1] * sc[1] * sc[2];
}

// tdsops.f90:205-405 (tridiagonal compact6 only)
inline void deriv_1st(Tdsops& t, double delta, const std::string& scheme, int bc_start, int bc_end,
                      bool symmetry) {
  if (t.n_halo < 2) fail("First derivative require n_halo >= 2");
  double alpha, afi, bfi, cfi;
  if (scheme == "compact6") {
    alpha = 1.0 / 3.0;
    afi = 7.0 / 9.0 / delta;
    bfi = 1.0 / 36.0 / delta;
    cfi = 0.0;
  } else {
    fail("scheme is not defined");
  }
  t.alpha = alpha; t.a = afi; t.b = bfi; t.c = cfi;
  set9(t.coeffs, {0.0, -cfi, -bfi, -afi, 0.0, afi, bfi, cfi, 0.0});
  for (int i = 1; i <= t.n_halo; ++i) { copy9(t.coeffs_s[i], t.coeffs); copy9(t.coeffs_e[i], t.coeffs); }
  std::fill(t.dist_sa.begin(), t.dist_sa.end(), alpha);
  std::fill(t.dist_sc.begin(), t.dist_sc.end(), alpha);
  const int n = t.n_tds, n_halo = t.n_halo;
  std::vector<double> dist_b(t.n_rhs + 1, 1.0);

  if (bc_start == BC_NEUMANN) {
    if (symmetry) {
      t.dist_sa[1] = 0.0; t.dist_sc[1] = 0.0;
      set9(t.coeffs_s[1], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_s[2], {0, 0, 0, -afi, -bfi, afi, bfi, 0, 0});
    } else {
      t.dist_sa[1] = 0.0; t.dist_sc[1] = 2 * alpha;
      set9(t.coeffs_s[1], {0, 0, 0, 0, 0, 2 * afi, 2 * bfi, 0, 0});
      set9(t.coeffs_s[2], {0, 0, 0, -afi, bfi, afi, bfi, 0, 0});
    }
  } else if (bc_start == BC_DIRICHLET) {
    t.dist_sa[1] = 0.0; t.dist_sc[1] = 2.0;
    set9(t.coeffs_s[1], {0, 0, 0, 0, -2.5, 2.0, 0.5, 0, 0});
    scale9(t.coeffs_s[1], delta);
    t.dist_sa[2] = 0.25; t.dist_sc[2] = 0.25;
    set9(t.coeffs_s[2], {0, 0, 0, -0.75, 0, 0.75, 0, 0, 0});
    scale9(t.coeffs_s[2], delta);
  }

  if (bc_end == BC_NEUMANN) {
    if (symmetry) {
      t.dist_sa[n] = 0.0; t.dist_sc[n] = 0.0;
      set9(t.coeffs_e[n_halo], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[n_halo - 1], {0, 0, -bfi, -afi, bfi, afi, 0, 0, 0});
    } else {
      t.dist_sa[n] = 2 * alpha; t.dist_sc[n] = 0.0;
      set9(t.coeffs_e[n_halo], {0, 0, -2 * bfi, -2 * afi, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[n_halo - 1], {0, 0, -bfi, -afi, -bfi, afi, 0, 0, 0});
    }
  } else if (bc_end == BC_DIRICHLET) {
    t.dist_sa[n] = 2.0; t.dist_sc[n] = 0.0;
    set9(t.coeffs_e[n_halo], {0, 0, -0.5, -2.0, 2.5, 0, 0, 0, 0});
    scale9(t.coeffs_e[n_halo], delta);
    t.dist_sa[n - 1] = 0.25; t.dist_sc[n - 1] = 0.25;
    set9(t.coeffs_e[n_halo - 1], {0, 0, 0, -0.75, 0, 0.75, 0, 0, 0});
    scale9(t.coeffs_e[n_halo - 1], delta);
  }
  preprocess_dist(t, dist_b);
}

// tdsops.f90:407-618
inline void deriv_2nd(Tdsops& t, double delta, const std::string& scheme, int bc_start, int bc_end,
                      bool symmetry, bool has_hv, double c_nu, double nu0_nu) {
  if (t.n_halo < 4) fail("Second derivative require n_halo >= 4");
  const double d2 = delta * delta;
  double alpha, asi, bsi, csi, dsi;
  if (scheme == "compact6") {
    alpha = 2.0 / 11.0;
    asi = 12.0 / 11.0 / d2;
    bsi = 3.0 / 44.0 / d2;
    csi = 0.0;
    dsi = 0.0;
  } else if (scheme == "compact6-hyperviscous") {
    if (!has_hv) fail("compact6-hyperviscous requires c_nu and nu0_nu");
    double dpis3 = 2.0 * pi / 3.0;
    double xnpi2 = pi * pi * (1.0 + nu0_nu);
    double xmpi2 = dpis3 * dpis3 * (1.0 + c_nu * nu0_nu);
    double den = 405.0 * xnpi2 - 640.0 * xmpi2 + 144.0;
    alpha = 0.5 - (320.0 * xmpi2 - 1296.0) / den;
    asi = -(4329.0 * xnpi2 / 8.0 - 32.0 * xmpi2 - 140.0 * xnpi2 * xmpi2 + 286.0) / den / d2;
    bsi = (2115.0 * xnpi2 - 1792.0 * xmpi2 - 280.0 * xnpi2 * xmpi2 + 1328.0) / den / (4.0 * d2);
    csi = -(7695.0 * xnpi2 / 8.0 + 288.0 * xmpi2 - 180.0 * xnpi2 * xmpi2 - 2574.0) / den / (9.0 * d2);
    dsi = (198.0 * xnpi2 + 128.0 * xmpi2 - 40.0 * xnpi2 * xmpi2 - 736.0) / den / (16.0 * d2);
  } else {
    fail("scheme is not defined");
  }
  t.alpha = alpha; t.a = asi; t.b = bsi; t.c = csi; t.d = dsi;
  set9(t.coeffs, {dsi, csi, bsi, asi, -2.0 * (asi + bsi + csi + dsi), asi, bsi, csi, dsi});
  for (int i = 1; i <= t.n_halo; ++i) { copy9(t.coeffs_s[i], t.coeffs); copy9(t.coeffs_e[i], t.coeffs); }
  std::fill(t.dist_sa.begin(), t.dist_sa.end(), alpha);
  std::fill(t.dist_sc.begin(), t.dist_sc.end(), alpha);
  const int n = t.n_tds;
  std::vector<double> dist_b(t.n_rhs + 1, 1.0);
  double temp1, temp2;

  if (bc_start == BC_NEUMANN) {
    if (symmetry) {
      t.dist_sa[1] = 0.0; t.dist_sc[1] = 2 * alpha;
      set9(t.coeffs_s[1], {0, 0, 0, 0, -2 * asi - 2 * bsi - 2 * csi - 2 * dsi, 2 * asi, 2 * bsi, 2 * csi, 2 * dsi});
      set9(t.coeffs_s[2], {0, 0, 0, asi, -2 * asi - bsi - 2 * csi - 2 * dsi, asi + csi, bsi + dsi, csi, dsi});
      set9(t.coeffs_s[3], {0, 0, bsi, asi + csi, -2 * asi - 2 * bsi - 2 * csi - dsi, asi, bsi, csi, dsi});
      set9(t.coeffs_s[4], {0, csi, bsi + dsi, asi, -2 * asi - 2 * bsi - 2 * csi - 2 * dsi, asi, bsi, csi, dsi});
    } else {
      t.dist_sa[1] = 0.0; t.dist_sc[1] = 0.0;
      set9(t.coeffs_s[1], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_s[2], {0, 0, 0, asi, -2 * asi - 3 * bsi - 2 * csi - 2 * dsi, asi - csi, bsi - dsi, csi, dsi});
      set9(t.coeffs_s[3], {0, 0, bsi, asi - csi, -2 * asi - 2 * bsi - 2 * csi - 3 * dsi, asi, bsi, csi, dsi});
      set9(t.coeffs_s[4], {0, -csi, bsi - dsi, asi, -2 * asi - 2 * bsi - 2 * csi - 2 * dsi, asi, bsi, csi, dsi});
    }
  } else if (bc_start == BC_DIRICHLET) {
    t.dist_sa[1] = 0.0; t.dist_sc[1] = 11.0;
    set9(t.coeffs_s[1], {0, 0, 0, 0, 13.0 / d2, -27.0 / d2, 15.0 / d2, -1.0 / d2, 0});
    t.dist_sa[2] = 0.1; t.dist_sc[2] = 0.1;
    set9(t.coeffs_s[2], {0, 0, 0, 1.2 / d2, -2.4 / d2, 1.2 / d2, 0, 0, 0});
    t.dist_sa[3] = 2.0 / 11.0; t.dist_sc[3] = 2.0 / 11.0;
    temp1 = 3.0 / 44.0 / d2; temp2 = 12.0 / 11.0 / d2;
    set9(t.coeffs_s[3], {0, 0, temp1, temp2, -2.0 * (temp1 + temp2), temp2, temp1, 0, 0});
    t.dist_sa[4] = 2.0 / 11.0; t.dist_sc[4] = 2.0 / 11.0;
    copy9(t.coeffs_s[4], t.coeffs_s[3]);
  }

  if (bc_end == BC_NEUMANN) {
    if (symmetry) {
      t.dist_sa[n] = 2 * alpha; t.dist_sc[n] = 0.0;
      set9(t.coeffs_e[4], {2 * dsi, 2 * csi, 2 * bsi, 2 * asi, -2 * asi - 2 * bsi - 2 * csi - 2 * dsi, 0, 0, 0, 0});
      set9(t.coeffs_e[3], {dsi, csi, bsi + dsi, asi + csi, -2 * asi - bsi - 2 * csi - 2 * dsi, asi, 0, 0, 0});
      set9(t.coeffs_e[2], {dsi, csi, bsi, asi, -2 * asi - 2 * bsi - 2 * csi - dsi, asi + csi, bsi, 0, 0});
      set9(t.coeffs_e[1], {dsi, csi, bsi, asi, -2 * asi - 2 * bsi - 2 * csi - 2 * dsi, asi, bsi + dsi, csi, 0});
    } else {
      t.dist_sa[n] = 0.0; t.dist_sc[n] = 0.0;
      set9(t.coeffs_e[4], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[3], {dsi, csi, bsi - dsi, asi - csi, -2 * asi - 3 * bsi - 2 * csi - 2 * dsi, asi, 0, 0, 0});
      set9(t.coeffs_e[2], {dsi, csi, bsi, asi, -2 * asi - 2 * bsi - 2 * csi - 3 * dsi, asi - csi, bsi, 0, 0});
      set9(t.coeffs_e[1], {dsi, csi, bsi, asi, -2 * asi - 2 * bsi - 2 * csi - 2 * dsi, asi, bsi - dsi, -csi, 0});
    }
  } else if (bc_end == BC_DIRICHLET) {
    t.dist_sa[n] = 11.0; t.dist_sc[n] = 0.0;
    set9(t.coeffs_e[4], {0, -1.0 / d2, 15.0 / d2, -27.0 / d2, 13.0 / d2, 0, 0, 0, 0});
    t.dist_sa[n - 1] = 0.1; t.dist_sc[n - 1] = 0.1;
    set9(t.coeffs_e[3], {0, 0, 0, 1.2 / d2, -2.4 / d2, 1.2 / d2, 0, 0, 0});
    t.dist_sa[n - 2] = 2.0 / 11.0; t.dist_sc[n - 2] = 2.0 / 11.0;
    temp1 = 3.0 / 44.0 / d2; temp2 = 12.0 / 11.0 / d2;
    set9(t.coeffs_e[2], {0, 0, temp1, temp2, -2.0 * (temp1 + temp2), temp2, temp1, 0, 0});
    t.dist_sa[n - 3] = 2.0 / 11.0; t.dist_sc[n - 3] = 2.0 / 11.0;
    copy9(t.coeffs_e[1], t.coeffs_e[2]);
  }
  preprocess_dist(t, dist_b);
}

// tdsops.f90:620-764
inline void interpl_mid(Tdsops& t, const std::string& scheme, const std::string& from_to,
                        int bc_start, int bc_end) {
  if (t.n_halo < 4) fail("Interpolation require n_halo >= 4");
  double alpha, aici, bici, cici, dici;
  if (scheme == "classic") {
    alpha = 0.3; aici = 0.75; bici = 0.05; cici = 0.0; dici = 0.0;
  } else if (scheme == "optimised") {
    alpha = 0.461658;
    dici = 0.00146508;
    aici = (75.0 + 70.0 * alpha - 640.0 * dici) / 128.0;
    bici = (-25.0 + 126.0 * alpha + 2304.0 * dici) / 256.0;
    cici = (3.0 - 10.0 * alpha - 1280.0 * dici) / 256.0;
  } else if (scheme == "aggressive") {
    alpha = 0.49;
    aici = (75.0 + 70.0 * alpha) / 128.0;
    bici = (-25.0 + 126.0 * alpha) / 256.0;
    cici = (3.0 - 10.0 * alpha) / 256.0;
    dici = 0.0;
  } else {
    fail("scheme is not defined");
  }
  t.alpha = alpha; t.a = aici; t.b = bici; t.c = cici; t.d = dici;
  if (from_to == "v2p")
    set9(t.coeffs, {0.0, dici, cici, bici, aici, aici, bici, cici, dici});
  else if (from_to == "p2v")
    set9(t.coeffs, {dici, cici, bici, aici, aici, bici, cici, dici, 0.0});
  for (int i = 1; i <= t.n_halo; ++i) { copy9(t.coeffs_s[i], t.coeffs); copy9(t.coeffs_e[i], t.coeffs); }
  std::fill(t.dist_sa.begin(), t.dist_sa.end(), alpha);
  std::fill(t.dist_sc.begin(), t.dist_sc.end(), alpha);
  const int n = t.n_tds;
  std::vector<double> dist_b(t.n_rhs + 1, 1.0);

  if (bc_start == BC_NEUMANN) {
    t.dist_sa[1] = 0.0;
    if (from_to == "v2p") {
      dist_b[1] = 1.0 + alpha;
      set9(t.coeffs_s[1], {0, 0, 0, 0, aici, aici + bici, bici + cici, cici + dici, dici});
      set9(t.coeffs_s[2], {0, 0, 0, bici, aici + cici, aici + dici, bici, cici, dici});
      set9(t.coeffs_s[3], {0, 0, cici, bici + dici, aici, aici, bici, cici, dici});
    } else if (from_to == "p2v") {
      t.dist_sc[1] = 2 * alpha;
      set9(t.coeffs_s[1], {0, 0, 0, 0, 2 * aici, 2 * bici, 2 * cici, 2 * dici, 0});
      set9(t.coeffs_s[2], {0, 0, 0, aici + bici, aici + cici, bici + dici, cici, dici, 0});
      set9(t.coeffs_s[3], {0, 0, bici + cici, aici + dici, aici, bici, cici, dici, 0});
      set9(t.coeffs_s[4], {0, cici + dici, bici, aici, aici, bici, cici, dici, 0});
    }
  } else if (bc_start == BC_DIRICHLET) {
    fail("Dirichlet BC is not supported for midpoint interpolations!");
  }

  if (bc_end == BC_NEUMANN) {
    t.dist_sc[n] = 0.0;
    if (from_to == "v2p") {
      dist_b[n] = 1.0 + alpha;
      set9(t.coeffs_e[4], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[3], {0, dici, cici + dici, bici + cici, aici + bici, aici, 0, 0, 0});
      set9(t.coeffs_e[2], {0, dici, cici, bici, aici + dici, aici + cici, bici, 0, 0});
      set9(t.coeffs_e[1], {0, dici, cici, bici, aici, aici, bici + dici, cici, 0});
    } else if (from_to == "p2v") {
      t.dist_sa[n] = 2 * alpha;
      set9(t.coeffs_e[4], {2 * dici, 2 * cici, 2 * bici, 2 * aici, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[3], {dici, cici, bici + dici, aici + cici, aici + bici, 0, 0, 0, 0});
      set9(t.coeffs_e[2], {dici, cici, bici, aici, aici + dici, bici + cici, 0, 0, 0});
      set9(t.coeffs_e[1], {dici, cici, bici, aici, aici, bici, cici + dici, 0, 0});
    }
  } else if (bc_end == BC_DIRICHLET) {
    fail("Dirichlet BC is not supported for midpoint interpolations!");
  }
  preprocess_dist(t, dist_b);
}

// tdsops.f90:766-872
inline void stagder_1st(Tdsops& t, double delta, const std::string& scheme,
                        const std::string& from_to, int bc_start, int bc_end) {
  if (t.n_halo < 2) fail("Staggared deriv require n_halo >= 2");
  double alpha, aci, bci;
  if (scheme == "compact6") {
    alpha = 9.0 / 62.0;
    aci = 63.0 / 62.0 / delta;
    bci = 17.0 / 62.0 / 3.0 / delta;
  } else {
    fail("scheme is not defined");
  }
  t.alpha = alpha; t.a = aci; t.b = bci;
  if (from_to == "v2p")
    set9(t.coeffs, {0, 0, 0, -bci, -aci, aci, bci, 0, 0});
  else if (from_to == "p2v")
    set9(t.coeffs, {0, 0, -bci, -aci, aci, bci, 0, 0, 0});
  for (int i = 1; i <= t.n_halo; ++i) { copy9(t.coeffs_s[i], t.coeffs); copy9(t.coeffs_e[i], t.coeffs); }
  std::fill(t.dist_sa.begin(), t.dist_sa.end(), alpha);
  std::fill(t.dist_sc.begin(), t.dist_sc.end(), alpha);
  const int n = t.n_tds, n_halo = t.n_halo;
  std::vector<double> dist_b(t.n_rhs + 1, 1.0);

  if (bc_start == BC_NEUMANN) {
    t.dist_sa[1] = 0.0;
    if (from_to == "v2p") {
      dist_b[1] = 1.0 + alpha;
      set9(t.coeffs_s[1], {0, 0, 0, 0, -aci - 2 * bci, aci + bci, bci, 0, 0});
      set9(t.coeffs_s[2], {0, 0, 0, -bci, -aci, aci, bci, 0, 0});
    } else if (from_to == "p2v") {
      t.dist_sc[1] = 0.0;
      set9(t.coeffs_s[1], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_s[2], {0, 0, 0, -aci - bci, aci, bci, 0, 0, 0});
    }
  } else if (bc_start == BC_DIRICHLET) {
    fail("Dirichlet BC is not supported for midpoint derivatives!");
  }

  if (bc_end == BC_NEUMANN) {
    t.dist_sc[n] = 0.0;
    if (from_to == "v2p") {
      dist_b[n] = 1.0 + alpha;
      set9(t.coeffs_e[n_halo], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[n_halo - 1], {0, 0, 0, -bci, -aci - bci, aci + 2 * bci, 0, 0, 0});
    } else if (from_to == "p2v") {
      t.dist_sa[n] = 0.0;
      set9(t.coeffs_e[n_halo], {0, 0, 0, 0, 0, 0, 0, 0, 0});
      set9(t.coeffs_e[n_halo - 1], {0, 0, -bci, -aci, aci + bci, 0, 0, 0, 0});
    }
  } else if (bc_end == BC_DIRICHLET) {
    fail("Dirichlet BC is not supported for midpoint derivatives!");
  }
  preprocess_dist(t, dist_b);
}
}  // namespace detail

// tdsops.f90:63-203.  stretch / stretch_correct may be null (=> 1 / 0).
inline Tdsops tdsops_init(int n_tds, double delta, const std::string& operation,
                          const std::string& scheme, int bc_start, int bc_end,
                          const double* stretch = nullptr, const double* stretch_correct = nullptr,
                          int n_halo = 4, const std::string& from_to = "", bool sym = false,
                          bool has_hv = false, double c_nu = 0, double nu0_nu = 0) {
  Tdsops t;
  t.n_tds = n_tds;
  if (!from_to.empty() && (bc_end == BC_NEUMANN || bc_end == BC_DIRICHLET) && from_to == "v2p")
    t.n_rhs = n_tds + 1;
  else
    t.n_rhs = n_tds;
  t.n_halo = n_halo;
  const int n = t.n_rhs;
  // Fortran leaves unassigned entries (dist_fw(2), dist_bw(n-1:n), entry n_rhs>n_tds) undefined;
  // the oracle zero-fills them. No valid output depends on them.
  t.dist_fw.assign(n + 1, 0.0); t.dist_bw.assign(n + 1, 0.0);
  t.dist_sa.assign(n + 1, 0.0); t.dist_sc.assign(n + 1, 0.0); t.dist_af.assign(n + 1, 0.0);
  t.stretch.assign(n_tds + 1, 1.0);
  t.stretch_correct.assign(n_tds + 1, 0.0);
  if (stretch) for (int i = 1; i <= n_tds; ++i) t.stretch[i] = stretch[i - 1];
  if (stretch_correct) for (int i = 1; i <= n_tds; ++i) t.stretch_correct[i] = stretch_correct[i - 1];
  t.periodic = bc_start == BC_PERIODIC && bc_end == BC_PERIODIC;
  for (int k = 0; k < 10; ++k) t.coeffs[k] = 0;
  for (int i = 0; i < 5; ++i) for (int k = 0; k < 10; ++k) { t.coeffs_s[i][k] = 0; t.coeffs_e[i][k] = 0; }

  if (operation == "first-deriv")
    detail::deriv_1st(t, delta, scheme, bc_start, bc_end, sym);
  else if (operation == "second-deriv")
    detail::deriv_2nd(t, delta, scheme, bc_start, bc_end, sym, has_hv, c_nu, nu0_nu);
  else if (operation == "interpolate")
    detail::interpl_mid(t, scheme, from_to, bc_start, bc_end);
  else if (operation == "stag-deriv")
    detail::stagder_1st(t, delta, scheme, from_to, bc_start, bc_end);
  else
    fail("operation is not defined");

  if (from_to == "v2p") t.move = 1;
  else if (from_to == "p2v") t.move = -1;
  else t.move = 0;
  return t;
}

// tdsops.f90:51-59
struct Dirps {
  Tdsops der1st, der1st_sym, der2nd, der2nd_sym, stagder_v2p, stagder_p2v, interpl_v2p, interpl_p2v;
  int dir = 0;
};

}  // namespace x3d2h
