// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.hpp).
// Restatement of the OMP DistD2 kernels, /root/reference/src/backend/omp/kernels/distributed.f90:11-337,
// and their orchestration, /root/reference/src/backend/omp/exec_dist.f90:16-186.
// Strict IEEE: sums are evaluated left to right exactly as written in the Fortran source; this file
// must be compiled with -ffp-contract=off and without -ffast-math.
//
// A group slice is a Fortran (SZ, n) array: element (i, j) lives at p[(j-1)*SZ + i], i = 0..SZ-1, j = 1..n.
#pragma once
#include "orc_tdsops.hpp"

namespace orc {

#define ORC_AT(p, i, j) (p)[((j)-1) * SZ + (i)]

// distributed.f90:11-168
inline void der_univ_dist(double* du, double* send_u_s, double* send_u_e, const double* u,
                          const double* u_s, const double* u_e, int n_tds, int n_rhs,
                          const double (*cs)[10], const double (*ce)[10], const double* coeffs,
                          const double* ffr, const double* fbc, const double* faf) {
  const double c_m4 = coeffs[1], c_m3 = coeffs[2], c_m2 = coeffs[3], c_m1 = coeffs[4], c_j = coeffs[5],
               c_p1 = coeffs[6], c_p2 = coeffs[7], c_p3 = coeffs[8], c_p4 = coeffs[9];
  const double last_r = ffr[1];

#pragma omp simd
  for (int i = 0; i < SZ; ++i) {
    double t;
    t = cs[1][1] * ORC_AT(u_s, i, 1) + cs[1][2] * ORC_AT(u_s, i, 2) + cs[1][3] * ORC_AT(u_s, i, 3) +
        cs[1][4] * ORC_AT(u_s, i, 4) + cs[1][5] * ORC_AT(u, i, 1) + cs[1][6] * ORC_AT(u, i, 2) +
        cs[1][7] * ORC_AT(u, i, 3) + cs[1][8] * ORC_AT(u, i, 4) + cs[1][9] * ORC_AT(u, i, 5);
    ORC_AT(du, i, 1) = t * faf[1];
    t = cs[2][1] * ORC_AT(u_s, i, 2) + cs[2][2] * ORC_AT(u_s, i, 3) + cs[2][3] * ORC_AT(u_s, i, 4) +
        cs[2][4] * ORC_AT(u, i, 1) + cs[2][5] * ORC_AT(u, i, 2) + cs[2][6] * ORC_AT(u, i, 3) +
        cs[2][7] * ORC_AT(u, i, 4) + cs[2][8] * ORC_AT(u, i, 5) + cs[2][9] * ORC_AT(u, i, 6);
    ORC_AT(du, i, 2) = t * faf[2];
    t = cs[3][1] * ORC_AT(u_s, i, 3) + cs[3][2] * ORC_AT(u_s, i, 4) + cs[3][3] * ORC_AT(u, i, 1) +
        cs[3][4] * ORC_AT(u, i, 2) + cs[3][5] * ORC_AT(u, i, 3) + cs[3][6] * ORC_AT(u, i, 4) +
        cs[3][7] * ORC_AT(u, i, 5) + cs[3][8] * ORC_AT(u, i, 6) + cs[3][9] * ORC_AT(u, i, 7);
    ORC_AT(du, i, 3) = ffr[3] * (t - faf[3] * ORC_AT(du, i, 2));
    t = cs[4][1] * ORC_AT(u_s, i, 4) + cs[4][2] * ORC_AT(u, i, 1) + cs[4][3] * ORC_AT(u, i, 2) +
        cs[4][4] * ORC_AT(u, i, 3) + cs[4][5] * ORC_AT(u, i, 4) + cs[4][6] * ORC_AT(u, i, 5) +
        cs[4][7] * ORC_AT(u, i, 6) + cs[4][8] * ORC_AT(u, i, 7) + cs[4][9] * ORC_AT(u, i, 8);
    ORC_AT(du, i, 4) = ffr[4] * (t - faf[4] * ORC_AT(du, i, 3));
  }

  const double alpha = faf[5];
  for (int j = 5; j <= n_rhs - 4; ++j) {
#pragma omp simd
    for (int i = 0; i < SZ; ++i) {
      double t = c_m4 * ORC_AT(u, i, j - 4) + c_m3 * ORC_AT(u, i, j - 3) + c_m2 * ORC_AT(u, i, j - 2) +
                 c_m1 * ORC_AT(u, i, j - 1) + c_j * ORC_AT(u, i, j) + c_p1 * ORC_AT(u, i, j + 1) +
                 c_p2 * ORC_AT(u, i, j + 2) + c_p3 * ORC_AT(u, i, j + 3) + c_p4 * ORC_AT(u, i, j + 4);
      ORC_AT(du, i, j) = ffr[j] * (t - alpha * ORC_AT(du, i, j - 1));
    }
  }

#pragma omp simd
  for (int i = 0; i < SZ; ++i) {
    double t;
    int j = n_rhs - 3;
    t = ce[1][1] * ORC_AT(u, i, j - 4) + ce[1][2] * ORC_AT(u, i, j - 3) + ce[1][3] * ORC_AT(u, i, j - 2) +
        ce[1][4] * ORC_AT(u, i, j - 1) + ce[1][5] * ORC_AT(u, i, j) + ce[1][6] * ORC_AT(u, i, j + 1) +
        ce[1][7] * ORC_AT(u, i, j + 2) + ce[1][8] * ORC_AT(u, i, j + 3) + ce[1][9] * ORC_AT(u_e, i, 1);
    ORC_AT(du, i, j) = ffr[j] * (t - faf[j] * ORC_AT(du, i, j - 1));
    j = n_rhs - 2;
    t = ce[2][1] * ORC_AT(u, i, j - 4) + ce[2][2] * ORC_AT(u, i, j - 3) + ce[2][3] * ORC_AT(u, i, j - 2) +
        ce[2][4] * ORC_AT(u, i, j - 1) + ce[2][5] * ORC_AT(u, i, j) + ce[2][6] * ORC_AT(u, i, j + 1) +
        ce[2][7] * ORC_AT(u, i, j + 2) + ce[2][8] * ORC_AT(u_e, i, 1) + ce[2][9] * ORC_AT(u_e, i, 2);
    ORC_AT(du, i, j) = ffr[j] * (t - faf[j] * ORC_AT(du, i, j - 1));
    j = n_rhs - 1;
    t = ce[3][1] * ORC_AT(u, i, j - 4) + ce[3][2] * ORC_AT(u, i, j - 3) + ce[3][3] * ORC_AT(u, i, j - 2) +
        ce[3][4] * ORC_AT(u, i, j - 1) + ce[3][5] * ORC_AT(u, i, j) + ce[3][6] * ORC_AT(u, i, j + 1) +
        ce[3][7] * ORC_AT(u_e, i, 1) + ce[3][8] * ORC_AT(u_e, i, 2) + ce[3][9] * ORC_AT(u_e, i, 3);
    ORC_AT(du, i, j) = ffr[j] * (t - faf[j] * ORC_AT(du, i, j - 1));
    j = n_rhs;
    t = ce[4][1] * ORC_AT(u, i, j - 4) + ce[4][2] * ORC_AT(u, i, j - 3) + ce[4][3] * ORC_AT(u, i, j - 2) +
        ce[4][4] * ORC_AT(u, i, j - 1) + ce[4][5] * ORC_AT(u, i, j) + ce[4][6] * ORC_AT(u_e, i, 1) +
        ce[4][7] * ORC_AT(u_e, i, 2) + ce[4][8] * ORC_AT(u_e, i, 3) + ce[4][9] * ORC_AT(u_e, i, 4);
    ORC_AT(du, i, j) = ffr[j] * (t - faf[j] * ORC_AT(du, i, j - 1));
  }

#pragma omp simd
  for (int i = 0; i < SZ; ++i) ORC_AT(send_u_e, i, 1) = ORC_AT(du, i, n_tds);

  for (int j = n_tds - 2; j >= 2; --j) {
#pragma omp simd
    for (int i = 0; i < SZ; ++i) ORC_AT(du, i, j) = ORC_AT(du, i, j) - fbc[j] * ORC_AT(du, i, j + 1);
  }
#pragma omp simd
  for (int i = 0; i < SZ; ++i) {
    ORC_AT(du, i, 1) = last_r * (ORC_AT(du, i, 1) - fbc[1] * ORC_AT(du, i, 2));
    ORC_AT(send_u_s, i, 1) = ORC_AT(du, i, 1);
  }
}

// distributed.f90:170-229
inline void der_univ_subs(double* du, const double* recv_u_s, const double* recv_u_e, int n,
                          const double* dist_sa, const double* dist_sc, const double* strch) {
  double du_s[SZ], du_e[SZ];
#pragma omp simd
  for (int i = 0; i < SZ; ++i) {
    double bl = dist_sa[1], ur = dist_sa[1];
    double recp = 1.0 / (1.0 - ur * bl);
    du_s[i] = recp * (ORC_AT(du, i, 1) - bl * ORC_AT(recv_u_s, i, 1));
    bl = dist_sc[n]; ur = dist_sc[n];
    recp = 1.0 / (1.0 - ur * bl);
    du_e[i] = recp * (ORC_AT(du, i, n) - ur * ORC_AT(recv_u_e, i, 1));
  }
#pragma omp simd
  for (int i = 0; i < SZ; ++i) ORC_AT(du, i, 1) = du_s[i] * strch[1];
  for (int j = 2; j <= n - 1; ++j) {
#pragma omp simd
    for (int i = 0; i < SZ; ++i)
      ORC_AT(du, i, j) = (ORC_AT(du, i, j) - dist_sa[j] * du_s[i] - dist_sc[j] * du_e[i]) * strch[j];
  }
#pragma omp simd
  for (int i = 0; i < SZ; ++i) ORC_AT(du, i, n) = du_e[i] * strch[n];
}

// distributed.f90:231-337
inline void der_univ_fused_subs(double* rhs_du, const double* dud, const double* d2u, const double* v,
                                const double* du_recv_s, const double* du_recv_e,
                                const double* dud_recv_s, const double* dud_recv_e,
                                const double* d2u_recv_s, const double* d2u_recv_e, double nu, int n,
                                const double* du_sa, const double* du_sc, const double* du_strch,
                                const double* dud_sa, const double* dud_sc, const double* dud_strch,
                                const double* d2u_sa, const double* d2u_sc, const double* d2u_strch,
                                const double* d2u_strch_cor) {
  double du_s[SZ], du_e[SZ], dud_s[SZ], dud_e[SZ], d2u_s[SZ], d2u_e[SZ];
#pragma omp simd
  for (int i = 0; i < SZ; ++i) {
    double bl, ur, recp;
    bl = du_sa[1]; ur = du_sa[1]; recp = 1.0 / (1.0 - ur * bl);
    du_s[i] = recp * (ORC_AT(rhs_du, i, 1) - bl * ORC_AT(du_recv_s, i, 1));
    bl = dud_sa[1]; ur = dud_sa[1]; recp = 1.0 / (1.0 - ur * bl);
    dud_s[i] = recp * (ORC_AT(dud, i, 1) - bl * ORC_AT(dud_recv_s, i, 1));
    bl = d2u_sa[1]; ur = d2u_sa[1]; recp = 1.0 / (1.0 - ur * bl);
    d2u_s[i] = recp * (ORC_AT(d2u, i, 1) - bl * ORC_AT(d2u_recv_s, i, 1));
    bl = du_sc[n]; ur = du_sc[n]; recp = 1.0 / (1.0 - ur * bl);
    du_e[i] = recp * (ORC_AT(rhs_du, i, n) - ur * ORC_AT(du_recv_e, i, 1));
    bl = dud_sc[n]; ur = dud_sc[n]; recp = 1.0 / (1.0 - ur * bl);
    dud_e[i] = recp * (ORC_AT(dud, i, n) - ur * ORC_AT(dud_recv_e, i, 1));
    bl = d2u_sc[n]; ur = d2u_sc[n]; recp = 1.0 / (1.0 - ur * bl);
    d2u_e[i] = recp * (ORC_AT(d2u, i, n) - ur * ORC_AT(d2u_recv_e, i, 1));
  }
#pragma omp simd
  for (int i = 0; i < SZ; ++i)
    ORC_AT(rhs_du, i, 1) = -0.5 * (ORC_AT(v, i, 1) * du_s[i] * du_strch[1] + dud_s[i] * dud_strch[1]) +
                           nu * (d2u_s[i] * d2u_strch[1] + du_s[i] * du_strch[1] * d2u_strch_cor[1]);
  for (int j = 2; j <= n - 1; ++j) {
#pragma omp simd
    for (int i = 0; i < SZ; ++i) {
      double temp_du = du_strch[j] * (ORC_AT(rhs_du, i, j) - du_sa[j] * du_s[i] - du_sc[j] * du_e[i]);
      double temp_dud = dud_strch[j] * (ORC_AT(dud, i, j) - dud_sa[j] * dud_s[i] - dud_sc[j] * dud_e[i]);
      double temp_d2u = d2u_strch[j] * (ORC_AT(d2u, i, j) - d2u_sa[j] * d2u_s[i] - d2u_sc[j] * d2u_e[i]) +
                        temp_du * d2u_strch_cor[j];
      ORC_AT(rhs_du, i, j) = -0.5 * (ORC_AT(v, i, j) * temp_du + temp_dud) + nu * temp_d2u;
    }
  }
#pragma omp simd
  for (int i = 0; i < SZ; ++i)
    ORC_AT(rhs_du, i, n) = -0.5 * (ORC_AT(v, i, n) * du_e[i] * du_strch[n] + dud_e[i] * dud_strch[n]) +
                           nu * (d2u_e[i] * d2u_strch[n] + du_e[i] * du_strch[n] * d2u_strch_cor[n]);
}

// ---------------------------------------------------------------------------------------------
// Per-rank directional data used by the exec_* orchestration. A "block" is the Fortran array
// (SZ, n_pad, n_groups) held in one std::vector<double>; element (i, j, k) is at
// i + SZ*(j-1) + SZ*n_pad*(k-1)  (i 0-based, j, k 1-based).
// ---------------------------------------------------------------------------------------------
struct Halo {  // (SZ, rows, n_groups)
  std::vector<double> d;
  int rows = 0;
  void resize(int r, int n_groups) { rows = r; d.assign((size_t)SZ * r * n_groups, 0.0); }
  double* grp(int k) { return d.data() + (size_t)SZ * rows * (k - 1); }
  const double* grp(int k) const { return d.data() + (size_t)SZ * rows * (k - 1); }
};

// omp/backend.f90:714-737
inline void copy_into_buffers(Halo& send_s, Halo& send_e, const double* u, int n_pad, int n, int n_groups) {
#pragma omp parallel for
  for (int k = 1; k <= n_groups; ++k) {
    const double* ug = u + (size_t)SZ * n_pad * (k - 1);
    double* ss = send_s.grp(k);
    double* se = send_e.grp(k);
    for (int j = 1; j <= 4; ++j)
      for (int i = 0; i < SZ; ++i) {
        ORC_AT(ss, i, j) = ORC_AT(ug, i, j);
        ORC_AT(se, i, j) = ORC_AT(ug, i, n - 4 + j);
      }
  }
}

// omp/sendrecv.f90:10-36 for an in-process set of P ranks along one direction:
// recv_s(r) <- send_e(prev(r)); recv_e(r) <- send_s(next(r)). nproc == 1 is the self copy of :20-22.
template <class GetPrev, class GetNext>
inline void sendrecv_fields(std::vector<Halo*>& recv_s, std::vector<Halo*>& recv_e,
                            std::vector<Halo*>& send_s, std::vector<Halo*>& send_e, GetPrev prev,
                            GetNext next) {
  const int P = (int)recv_s.size();
  for (int r = 0; r < P; ++r) {
    recv_s[r]->d = send_e[prev(r)]->d;
    recv_e[r]->d = send_s[next(r)]->d;
  }
}

}  // namespace orc
