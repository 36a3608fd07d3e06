// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// C entry points (ctypes-friendly) over the CPU restatement of the xcompact3d/x3d2 OpenMP backend.
// The reference (Fortran 2018 + OpenMP + MPI + 2DECOMP&FFT) cannot be compiled in this image
// (no Fortran compiler, no MPI), so this restatement is the parity oracle and the CPU
// baseline ("port"). It is pinned against the reference's own known-answer tests (analytic solutions
// at the reference's tolerances, tests/test_oracle_*.py); the reference stores no golden vectors, so
// bit-level parity with the reference *binary* is unpinned.
//
// Build: make -C oracle   (g++ -O3 -march=native -fopenmp -ffp-contract=off, no -ffast-math)
#include "orc_world.hpp"

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret)                    \
  }                                       \
  catch (const std::exception& e) {       \
    g_err = e.what();                     \
    return ret;                           \
  }

#include <omp.h>

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_sz() { return SZ; }
// OpenMP thread control: torchrun exports OMP_NUM_THREADS=1, so the timed legs set the count explicitly and report
// the number of threads a parallel region really gets.
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_num_threads() {
  int n = 1;
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
  return n;
}

// ------------------------------------------------------------------------------------ tdsops
void* orc_tdsops_create(int n_tds, double delta, const char* operation, const char* scheme, int bc_start,
                        int bc_end, const double* stretch, const double* stretch_correct, int n_halo,
                        const char* from_to, int sym, int has_hv, double c_nu, double nu0_nu) {
  ORC_TRY
  auto* t = new Tdsops(tdsops_init(n_tds, delta, operation, scheme, bc_start, bc_end, stretch, stretch_correct,
                                   n_halo, from_to ? from_to : "", sym != 0, has_hv != 0, c_nu, nu0_nu));
  return (void*)t;
  ORC_CATCH(nullptr)
}
void orc_tdsops_destroy(void* h) { delete (Tdsops*)h; }

// info[0..3] = n_tds, n_rhs, move, periodic ; sc[0..4] = alpha, a, b, c, d
void orc_tdsops_info(void* h, int* info, double* sc) {
  auto* t = (Tdsops*)h;
  info[0] = t->n_tds; info[1] = t->n_rhs; info[2] = t->move; info[3] = t->periodic;
  sc[0] = t->alpha; sc[1] = t->a; sc[2] = t->b; sc[3] = t->c; sc[4] = t->d;
}
// 0-based flat copies. coeffs_s/e are [row 0..3][tap 0..8]. fw/bw/sa/sc/af have n_rhs entries,
// stretch/stretch_correct n_tds entries.
void orc_tdsops_arrays(void* h, double* coeffs, double* coeffs_s, double* coeffs_e, double* fw, double* bw,
                       double* sa, double* sc, double* af, double* stretch, double* stretch_correct) {
  auto* t = (Tdsops*)h;
  for (int k = 1; k <= 9; ++k) coeffs[k - 1] = t->coeffs[k];
  for (int i = 1; i <= 4; ++i)
    for (int k = 1; k <= 9; ++k) {
      coeffs_s[(i - 1) * 9 + (k - 1)] = t->coeffs_s[i][k];
      coeffs_e[(i - 1) * 9 + (k - 1)] = t->coeffs_e[i][k];
    }
  for (int i = 1; i <= t->n_rhs; ++i) {
    fw[i - 1] = t->dist_fw[i]; bw[i - 1] = t->dist_bw[i]; sa[i - 1] = t->dist_sa[i];
    sc[i - 1] = t->dist_sc[i]; af[i - 1] = t->dist_af[i];
  }
  for (int i = 1; i <= t->n_tds; ++i) { stretch[i - 1] = t->stretch[i]; stretch_correct[i - 1] = t->stretch_correct[i]; }
}

// ------------------------------------------------------------------------------------ line-level exec
// P emulated ranks along the line direction. u is [P][n_lines][n_pad] (line-major, j fastest);
// n_lines must be a multiple of SZ. Output du has the same shape (entries 0..n_tds-1 are valid).
static void pack_lines(std::vector<double>& blk, const double* lines, int n_lines, int n_pad) {
  int G = n_lines / SZ;
  blk.assign((size_t)SZ * n_pad * G, 0.0);
  for (int g = 0; g < G; ++g)
    for (int j = 0; j < n_pad; ++j)
      for (int i = 0; i < SZ; ++i) blk[i + (size_t)SZ * (j + (size_t)n_pad * g)] = lines[(size_t)(g * SZ + i) * n_pad + j];
}
static void unpack_lines(double* lines, const std::vector<double>& blk, int n_lines, int n_pad) {
  int G = n_lines / SZ;
  for (int g = 0; g < G; ++g)
    for (int j = 0; j < n_pad; ++j)
      for (int i = 0; i < SZ; ++i) lines[(size_t)(g * SZ + i) * n_pad + j] = blk[i + (size_t)SZ * (j + (size_t)n_pad * g)];
}

struct LineRank {
  std::vector<double> u, v, du, dud, d2u;
  Halo u_send_s, u_send_e, u_recv_s, u_recv_e, v_send_s, v_send_e, v_recv_s, v_recv_e;
  Halo s_s[3], s_e[3], r_s[3], r_e[3];
};

// tests/verification/test_omp_tridiag.f90:365-404 (run_kernel): halo copy, sendrecv, exec_dist_tds_compact
int orc_lines_tds_solve(int P, void** ops, int n_lines, int n_pad, const double* u, double* du) {
  ORC_TRY
  if (n_lines % SZ) fail("n_lines must be a multiple of SZ");
  const int G = n_lines / SZ;
  std::vector<LineRank> R(P);
  for (int r = 0; r < P; ++r) {
    auto* t = (Tdsops*)ops[r];
    if (n_pad < t->n_rhs) fail("n_pad < n_rhs");
    pack_lines(R[r].u, u + (size_t)r * n_lines * n_pad, n_lines, n_pad);
    R[r].du.assign(R[r].u.size(), 0.0);
    for (Halo* h : {&R[r].u_send_s, &R[r].u_send_e, &R[r].u_recv_s, &R[r].u_recv_e}) h->resize(4, G);
    for (Halo* h : {&R[r].s_s[0], &R[r].s_e[0], &R[r].r_s[0], &R[r].r_e[0]}) h->resize(1, G);
    copy_into_buffers(R[r].u_send_s, R[r].u_send_e, R[r].u.data(), n_pad, t->n_tds, G);
  }
  for (int r = 0; r < P; ++r) {
    R[r].u_recv_s.d = R[(r - 1 + P) % P].u_send_e.d;
    R[r].u_recv_e.d = R[(r + 1) % P].u_send_s.d;
  }
  for (int r = 0; r < P; ++r) {
    auto* t = (Tdsops*)ops[r];
#pragma omp parallel for
    for (int k = 1; k <= G; ++k) {
      size_t off = (size_t)SZ * n_pad * (k - 1);
      der_univ_dist(R[r].du.data() + off, R[r].s_s[0].grp(k), R[r].s_e[0].grp(k), R[r].u.data() + off,
                    R[r].u_recv_s.grp(k), R[r].u_recv_e.grp(k), t->n_tds, t->n_rhs, t->coeffs_s, t->coeffs_e,
                    t->coeffs, t->dist_fw.data(), t->dist_bw.data(), t->dist_af.data());
    }
  }
  for (int r = 0; r < P; ++r) {
    R[r].r_s[0].d = R[(r - 1 + P) % P].s_e[0].d;
    R[r].r_e[0].d = R[(r + 1) % P].s_s[0].d;
  }
  for (int r = 0; r < P; ++r) {
    auto* t = (Tdsops*)ops[r];
#pragma omp parallel for
    for (int k = 1; k <= G; ++k) {
      size_t off = (size_t)SZ * n_pad * (k - 1);
      der_univ_subs(R[r].du.data() + off, R[r].r_s[0].grp(k), R[r].r_e[0].grp(k), t->n_tds, t->dist_sa.data(),
                    t->dist_sc.data(), t->stretch.data());
    }
    unpack_lines(du + (size_t)r * n_lines * n_pad, R[r].du, n_lines, n_pad);
  }
  return 0;
  ORC_CATCH(1)
}

// tests/verification/test_omp_dist_transeq.f90 shape: exec_dist_transeq_compact on (u, v=conv)
int orc_lines_transeq(int P, void** ops_du, void** ops_dud, void** ops_d2u, double nu, int n_lines, int n_pad,
                      const double* u, const double* v, double* rhs) {
  ORC_TRY
  if (n_lines % SZ) fail("n_lines must be a multiple of SZ");
  const int G = n_lines / SZ;
  std::vector<LineRank> R(P);
  for (int r = 0; r < P; ++r) {
    auto* a = (Tdsops*)ops_du[r];
    pack_lines(R[r].u, u + (size_t)r * n_lines * n_pad, n_lines, n_pad);
    pack_lines(R[r].v, v + (size_t)r * n_lines * n_pad, n_lines, n_pad);
    R[r].du.assign(R[r].u.size(), 0.0); R[r].dud = R[r].du; R[r].d2u = R[r].du;
    for (Halo* h : {&R[r].u_send_s, &R[r].u_send_e, &R[r].u_recv_s, &R[r].u_recv_e, &R[r].v_send_s,
                    &R[r].v_send_e, &R[r].v_recv_s, &R[r].v_recv_e}) h->resize(4, G);
    for (int q = 0; q < 3; ++q)
      for (Halo* h : {&R[r].s_s[q], &R[r].s_e[q], &R[r].r_s[q], &R[r].r_e[q]}) h->resize(1, G);
    copy_into_buffers(R[r].u_send_s, R[r].u_send_e, R[r].u.data(), n_pad, a->n_tds, G);
    copy_into_buffers(R[r].v_send_s, R[r].v_send_e, R[r].v.data(), n_pad, a->n_tds, G);
  }
  for (int r = 0; r < P; ++r) {
    int pv = (r - 1 + P) % P, nx = (r + 1) % P;
    R[r].u_recv_s.d = R[pv].u_send_e.d; R[r].u_recv_e.d = R[nx].u_send_s.d;
    R[r].v_recv_s.d = R[pv].v_send_e.d; R[r].v_recv_e.d = R[nx].v_send_s.d;
  }
  for (int r = 0; r < P; ++r) {
    auto *a = (Tdsops*)ops_du[r], *b = (Tdsops*)ops_dud[r], *c = (Tdsops*)ops_d2u[r];
#pragma omp parallel
    {
      std::vector<double> ud((size_t)SZ * n_pad), ud_s(SZ * 4), ud_e(SZ * 4);
#pragma omp for
      for (int k = 1; k <= G; ++k) {
        size_t off = (size_t)SZ * n_pad * (k - 1);
        const double *ug = R[r].u.data() + off, *vg = R[r].v.data() + off;
        der_univ_dist(R[r].du.data() + off, R[r].s_s[0].grp(k), R[r].s_e[0].grp(k), ug, R[r].u_recv_s.grp(k),
                      R[r].u_recv_e.grp(k), a->n_tds, a->n_rhs, a->coeffs_s, a->coeffs_e, a->coeffs,
                      a->dist_fw.data(), a->dist_bw.data(), a->dist_af.data());
        der_univ_dist(R[r].d2u.data() + off, R[r].s_s[2].grp(k), R[r].s_e[2].grp(k), ug, R[r].u_recv_s.grp(k),
                      R[r].u_recv_e.grp(k), c->n_tds, c->n_rhs, c->coeffs_s, c->coeffs_e, c->coeffs,
                      c->dist_fw.data(), c->dist_bw.data(), c->dist_af.data());
        for (int j = 1; j <= b->n_tds; ++j)
          for (int i = 0; i < SZ; ++i) ORC_AT(ud.data(), i, j) = ORC_AT(ug, i, j) * ORC_AT(vg, i, j);
        for (int j = 1; j <= 4; ++j)
          for (int i = 0; i < SZ; ++i) {
            ORC_AT(ud_s.data(), i, j) = ORC_AT(R[r].u_recv_s.grp(k), i, j) * ORC_AT(R[r].v_recv_s.grp(k), i, j);
            ORC_AT(ud_e.data(), i, j) = ORC_AT(R[r].u_recv_e.grp(k), i, j) * ORC_AT(R[r].v_recv_e.grp(k), i, j);
          }
        der_univ_dist(R[r].dud.data() + off, R[r].s_s[1].grp(k), R[r].s_e[1].grp(k), ud.data(), ud_s.data(),
                      ud_e.data(), b->n_tds, b->n_rhs, b->coeffs_s, b->coeffs_e, b->coeffs, b->dist_fw.data(),
                      b->dist_bw.data(), b->dist_af.data());
      }
    }
  }
  for (int r = 0; r < P; ++r)
    for (int q = 0; q < 3; ++q) {
      R[r].r_s[q].d = R[(r - 1 + P) % P].s_e[q].d;
      R[r].r_e[q].d = R[(r + 1) % P].s_s[q].d;
    }
  for (int r = 0; r < P; ++r) {
    auto *a = (Tdsops*)ops_du[r], *b = (Tdsops*)ops_dud[r], *c = (Tdsops*)ops_d2u[r];
#pragma omp parallel for
    for (int k = 1; k <= G; ++k) {
      size_t off = (size_t)SZ * n_pad * (k - 1);
      der_univ_fused_subs(R[r].du.data() + off, R[r].dud.data() + off, R[r].d2u.data() + off, R[r].v.data() + off,
                          R[r].r_s[0].grp(k), R[r].r_e[0].grp(k), R[r].r_s[1].grp(k), R[r].r_e[1].grp(k),
                          R[r].r_s[2].grp(k), R[r].r_e[2].grp(k), nu, a->n_tds, a->dist_sa.data(),
                          a->dist_sc.data(), a->stretch.data(), b->dist_sa.data(), b->dist_sc.data(),
                          b->stretch.data(), c->dist_sa.data(), c->dist_sc.data(), c->stretch.data(),
                          c->stretch_correct.data());
    }
    unpack_lines(rhs + (size_t)r * n_lines * n_pad, R[r].du, n_lines, n_pad);
  }
  return 0;
  ORC_CATCH(1)
}

// ------------------------------------------------------------------------------------ world
// bcs: [dir][side] flattened (6 ints); stretching may be null => uniform
void* orc_world_create(const int* dims_global, const int* nproc_dir, const double* L, const int* bcs, double Re,
                       double dt, const char* time_intg, const char* der1st, const char* der2nd,
                       const char* interpl, const char* stagder, const char* const* stretching,
                       const double* beta) {
  ORC_TRY
  SolverConfig cfg;
  cfg.Re = Re; cfg.dt = dt; cfg.time_intg = time_intg; cfg.der1st = der1st; cfg.der2nd = der2nd;
  cfg.interpl = interpl; cfg.stagder = stagder;
  int b[3][2] = {{bcs[0], bcs[1]}, {bcs[2], bcs[3]}, {bcs[4], bcs[5]}};
  auto* w = new World(dims_global, nproc_dir, L, b, cfg);
  if (stretching) {
    std::string st[3] = {stretching[0], stretching[1], stretching[2]};
    double be[3] = {beta ? beta[0] : 1.0, beta ? beta[1] : 1.0, beta ? beta[2] : 1.0};
    w->set_stretching(st, be);
  }
  w->init_solver();
  return (void*)w;
  ORC_CATCH(nullptr)
}
void orc_world_destroy(void* h) { delete (World*)h; }

int orc_world_init_tgv(void* h) {
  ORC_TRY((World*)h)->init_tgv(); return 0; ORC_CATCH(1)
}
// velocity in/out as global vertex-located Cartesian arrays (nx, ny, nz), x fastest
int orc_world_set_uvw(void* h, const double* u, const double* v, const double* w) {
  ORC_TRY
  auto* W = (World*)h;
  W->set_field_from_global(*W->u, u, VERT);
  W->set_field_from_global(*W->v, v, VERT);
  W->set_field_from_global(*W->w, w, VERT);
  return 0;
  ORC_CATCH(1)
}
int orc_world_get_uvw(void* h, double* u, double* v, double* w) {
  ORC_TRY
  auto* W = (World*)h;
  W->get_field_to_global(u, *W->u, VERT);
  W->get_field_to_global(v, *W->v, VERT);
  W->get_field_to_global(w, *W->w, VERT);
  return 0;
  ORC_CATCH(1)
}
int orc_world_set_case_channel(void* h, double omega_rot, int n_rotate) {
  ORC_TRY
  ((World*)h)->set_case_channel(omega_rot, n_rotate);
  return 0;
  ORC_CATCH(1)
}
int orc_world_step(void* h, int nsteps) {
  ORC_TRY
  for (int i = 0; i < nsteps; ++i) ((World*)h)->step();
  return 0;
  ORC_CATCH(1)
}
// out[0..3] = enstrophy, kinetic energy, div_u_max, div_u_mean
int orc_world_monitor(void* h, double* out) {
  ORC_TRY
  auto* W = (World*)h;
  out[0] = W->enstrophy();
  out[1] = W->kinetic_energy();
  W->divergence_max_mean(out[2], out[3]);
  return 0;
  ORC_CATCH(1)
}

// transeq_default of the given velocity (does not touch the world's own u, v, w)
int orc_world_transeq(void* h, const double* u, const double* v, const double* w, double* du, double* dv, double* dw) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(DIR_X), *fv = W->get_block(DIR_X), *fw = W->get_block(DIR_X);
  WField *a = W->get_block(DIR_X), *b = W->get_block(DIR_X), *c = W->get_block(DIR_X);
  W->set_field_from_global(*fu, u, VERT); W->set_field_from_global(*fv, v, VERT); W->set_field_from_global(*fw, w, VERT);
  W->transeq_default(*a, *b, *c, *fu, *fv, *fw);
  W->get_field_to_global(du, *a, VERT); W->get_field_to_global(dv, *b, VERT); W->get_field_to_global(dw, *c, VERT);
  for (WField* f : {fu, fv, fw, a, b, c}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}

// transeq_lowmem (solver.f90:391-505): same right-hand side as transeq_default, velocities handed back in new blocks
int orc_world_transeq_lowmem(void* h, const double* u, const double* v, const double* w, double* du, double* dv,
                             double* dw, double* u_back) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(DIR_X), *fv = W->get_block(DIR_X), *fw = W->get_block(DIR_X);
  WField *a = W->get_block(DIR_X), *b = W->get_block(DIR_X), *c = W->get_block(DIR_X);
  W->set_field_from_global(*fu, u, VERT); W->set_field_from_global(*fv, v, VERT); W->set_field_from_global(*fw, w, VERT);
  W->transeq_lowmem(*a, *b, *c, fu, fv, fw);
  W->get_field_to_global(du, *a, VERT); W->get_field_to_global(dv, *b, VERT); W->get_field_to_global(dw, *c, VERT);
  fu->data_loc = VERT;
  if (u_back) W->get_field_to_global(u_back, *fu, VERT);
  for (WField* f : {fu, fv, fw, a, b, c}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}

// transeq_species (solver.f90:507-600 with omp/backend.f90:186-233) for one scalar
int orc_world_transeq_species(void* h, const double* u, const double* v, const double* w, const double* spec,
                              double nu_s, double* dspec) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(DIR_X), *fv = W->get_block(DIR_X), *fw = W->get_block(DIR_X), *fs = W->get_block(DIR_X);
  WField* a = W->get_block(DIR_X);
  W->set_field_from_global(*fu, u, VERT); W->set_field_from_global(*fv, v, VERT); W->set_field_from_global(*fw, w, VERT);
  W->set_field_from_global(*fs, spec, VERT);
  W->transeq_species(*a, *fu, *fv, *fw, *fs, nu_s);
  W->get_field_to_global(dspec, *a, VERT);
  for (WField* f : {fu, fv, fw, fs, a}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}

// one directional transeq (backend%transeq_x/y/z) on vertex data given in Cartesian order
int orc_world_transeq_dir(void* h, int dir, const double* u, const double* v, const double* w, double* du,
                          double* dv, double* dw) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(dir), *fv = W->get_block(dir), *fw = W->get_block(dir);
  WField *a = W->get_block(dir), *b = W->get_block(dir), *c = W->get_block(dir);
  W->set_field_from_global(*fu, u, VERT); W->set_field_from_global(*fv, v, VERT); W->set_field_from_global(*fw, w, VERT);
  if (dir == DIR_X) W->transeq_x(*a, *b, *c, *fu, *fv, *fw, W->nu);
  else if (dir == DIR_Y) W->transeq_y(*a, *b, *c, *fu, *fv, *fw, W->nu);
  else W->transeq_z(*a, *b, *c, *fu, *fv, *fw, W->nu);
  W->get_field_to_global(du, *a, VERT); W->get_field_to_global(dv, *b, VERT); W->get_field_to_global(dw, *c, VERT);
  for (WField* f : {fu, fv, fw, a, b, c}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}

static std::vector<const Tdsops*> pick(World* W, int dir, const std::string& name) {
  const std::vector<Dirps>& d = dir == DIR_X ? W->xdirps : (dir == DIR_Y ? W->ydirps : W->zdirps);
  if (name == "der1st") return W->op(d, &Dirps::der1st);
  if (name == "der1st_sym") return W->op(d, &Dirps::der1st_sym);
  if (name == "der2nd") return W->op(d, &Dirps::der2nd);
  if (name == "der2nd_sym") return W->op(d, &Dirps::der2nd_sym);
  if (name == "stagder_v2p") return W->op(d, &Dirps::stagder_v2p);
  if (name == "stagder_p2v") return W->op(d, &Dirps::stagder_p2v);
  if (name == "interpl_v2p") return W->op(d, &Dirps::interpl_v2p);
  if (name == "interpl_p2v") return W->op(d, &Dirps::interpl_p2v);
  fail("unknown operator name " + name);
}

// backend%tds_solve with one of the solver's operators. `in` has extents of in_loc; `out` of the moved loc.
int orc_world_tds_solve(void* h, int dir, const char* opname, int in_loc, const double* in, double* out, int* out_loc) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fi = W->get_block(dir), *fo = W->get_block(dir);
  W->set_field_from_global(*fi, in, in_loc);
  W->tds_solve(*fo, *fi, pick(W, dir, opname));
  *out_loc = fo->data_loc;
  W->get_field_to_global(out, *fo, fo->data_loc);
  W->release_block(fi); W->release_block(fo);
  return 0;
  ORC_CATCH(1)
}

int orc_world_divergence(void* h, const double* u, const double* v, const double* w, double* div) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(DIR_X), *fv = W->get_block(DIR_X), *fw = W->get_block(DIR_X), *d = W->get_block(DIR_Z);
  W->set_field_from_global(*fu, u, VERT); W->set_field_from_global(*fv, v, VERT); W->set_field_from_global(*fw, w, VERT);
  W->divergence_v2c(*d, *fu, *fv, *fw);
  W->get_field_to_global(div, *d, CELL);
  for (WField* f : {fu, fv, fw, d}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}
int orc_world_gradient(void* h, const double* p, double* gx, double* gy, double* gz) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fp = W->get_block(DIR_Z), *a = W->get_block(DIR_X), *b = W->get_block(DIR_X), *c = W->get_block(DIR_X);
  W->set_field_from_global(*fp, p, CELL);
  W->gradient_c2v(*a, *b, *c, *fp);
  W->get_field_to_global(gx, *a, VERT); W->get_field_to_global(gy, *b, VERT); W->get_field_to_global(gz, *c, VERT);
  for (WField* f : {fp, a, b, c}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}
int orc_world_interpl_c2v(void* h, const double* p, double* out) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fp = W->get_block(DIR_Z), *a = W->get_block(DIR_X);
  W->set_field_from_global(*fp, p, CELL);
  W->interpl_c2v(*a, *fp);
  W->get_field_to_global(out, *a, VERT);
  for (WField* f : {fp, a}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}
int orc_world_laplacian(void* h, const double* u, double* out) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(DIR_X), *a = W->get_block(DIR_X);
  W->set_field_from_global(*fu, u, VERT);
  W->laplacian(*a, *fu);
  W->get_field_to_global(out, *a, VERT);
  for (WField* f : {fu, a}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}
int orc_world_curl(void* h, const double* u, const double* v, const double* w, double* ox, double* oy, double* oz) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fu = W->get_block(DIR_X), *fv = W->get_block(DIR_X), *fw = W->get_block(DIR_X);
  WField *a = W->get_block(DIR_X), *b = W->get_block(DIR_X), *c = W->get_block(DIR_X);
  W->set_field_from_global(*fu, u, VERT); W->set_field_from_global(*fv, v, VERT); W->set_field_from_global(*fw, w, VERT);
  W->curl(*a, *b, *c, *fu, *fv, *fw);
  W->get_field_to_global(ox, *a, VERT); W->get_field_to_global(oy, *b, VERT); W->get_field_to_global(oz, *c, VERT);
  for (WField* f : {fu, fv, fw, a, b, c}) W->release_block(f);
  return 0;
  ORC_CATCH(1)
}
// solver%poisson: f and p are CELL-located Cartesian arrays
int orc_world_poisson(void* h, const double* f, double* p) {
  ORC_TRY
  auto* W = (World*)h;
  WField *ff = W->get_block(DIR_Z), *fp = W->get_block(DIR_Z);
  W->set_field_from_global(*ff, f, CELL);
  W->poisson_fft(*fp, *ff);
  W->get_field_to_global(p, *fp, CELL);
  W->release_block(ff); W->release_block(fp);
  return 0;
  ORC_CATCH(1)
}
// tests/verification/test_fft.f90:156-169: forward then backward, no normalisation. spec may be null.
int orc_world_fft_roundtrip(void* h, const double* f, double* out, double* spec_re_im) {
  ORC_TRY
  auto* W = (World*)h;
  WField* c = W->get_block(DIR_C, CELL);
  W->set_field_from_global(*c, f, CELL);
  W->fft_forward(*c);
  if (spec_re_im)
    for (size_t i = 0; i < W->c_x.size(); ++i) { spec_re_im[2 * i] = W->c_x[i].real(); spec_re_im[2 * i + 1] = W->c_x[i].imag(); }
  W->fft_backward(*c);
  W->get_field_to_global(out, *c, CELL);
  W->release_block(c);
  return 0;
  ORC_CATCH(1)
}
int orc_world_spec_dims(void* h, int* d) {
  auto* W = (World*)h;
  d[0] = W->nx_spec; d[1] = W->ny_spec; d[2] = W->nz_spec;
  return 0;
}
int orc_world_waves(void* h, double* re_im) {
  auto* W = (World*)h;
  for (size_t i = 0; i < W->waves.size(); ++i) { re_im[2 * i] = W->waves[i].real(); re_im[2 * i + 1] = W->waves[i].imag(); }
  return 0;
}
// stretching_matrix (poisson_fft.f90:275-652): info = {stretched (0 | 1 sym | 2 bottom), rows}; tensors may be null
int orc_world_stretching_matrix(void* h, int* info, double* a_odd, double* a_even) {
  ORC_TRY
  auto* W = (World*)h;
  info[0] = !W->stretched_y ? 0 : (W->stretched_y_sym ? 1 : 2);
  info[1] = W->stretched_y_sym ? W->ny_spec / 2 : W->ny_spec;
  const std::vector<double>& o = W->stretched_y_sym ? W->a_odd : W->a_full;
  if (a_odd && !o.empty()) std::memcpy(a_odd, o.data(), sizeof(double) * o.size());
  if (a_even && !W->a_even.empty()) std::memcpy(a_even, W->a_even.data(), sizeof(double) * W->a_even.size());
  return 0;
  ORC_CATCH(1)
}

int orc_world_pressure_correction(void* h) {
  ORC_TRY
  auto* W = (World*)h;
  W->pressure_correction(*W->u, *W->v, *W->w);
  return 0;
  ORC_CATCH(1)
}

// reorder chain test helper (tests/unit/test_reordering.f90): put `in` (VERT, Cartesian) into a DIR_C block,
// apply the reorders in `rdrs` one after another, then bring the result back to Cartesian.
int orc_world_reorder_chain(void* h, const double* in, const int* rdrs, int n_rdr, double* out) {
  ORC_TRY
  auto* W = (World*)h;
  WField* cur = W->get_block(DIR_C, VERT);
  W->set_field_from_global(*cur, in, VERT);
  for (int q = 0; q < n_rdr; ++q) {
    int from, to;
    get_dirs_from_rdr(from, to, rdrs[q]);
    if (from != cur->dir) fail("reorder chain: direction mismatch");
    WField* nxt = W->get_block(to);
    W->reorder(*nxt, *cur, rdrs[q]);
    W->release_block(cur);
    cur = nxt;
  }
  W->get_field_to_global(out, *cur, VERT);
  W->release_block(cur);
  return 0;
  ORC_CATCH(1)
}
// tests/unit/test_sum_intox.f90: a (DIR_X) += b given in dir_from (DIR_Y or DIR_Z); all Cartesian in/out
int orc_world_sum_intox(void* h, int dir_from, const double* a, const double* b, double* out) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fa = W->get_block(DIR_X), *fb = W->get_block(dir_from);
  W->set_field_from_global(*fa, a, VERT); W->set_field_from_global(*fb, b, VERT);
  if (dir_from == DIR_Y) W->sum_yintox(*fa, *fb); else W->sum_zintox(*fa, *fb);
  W->get_field_to_global(out, *fa, VERT);
  W->release_block(fa); W->release_block(fb);
  return 0;
  ORC_CATCH(1)
}
int orc_world_vecadd(void* h, int dir, double a, const double* x, double b, const double* y, double* out) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fx = W->get_block(dir), *fy = W->get_block(dir);
  W->set_field_from_global(*fx, x, VERT); W->set_field_from_global(*fy, y, VERT);
  W->vecadd(a, *fx, b, *fy);
  W->get_field_to_global(out, *fy, VERT);
  W->release_block(fx); W->release_block(fy);
  return 0;
  ORC_CATCH(1)
}
int orc_world_scalar_product(void* h, int dir, int data_loc, const double* x, const double* y, double* s) {
  ORC_TRY
  auto* W = (World*)h;
  WField *fx = W->get_block(dir), *fy = W->get_block(dir);
  W->set_field_from_global(*fx, x, data_loc); W->set_field_from_global(*fy, y, data_loc);
  *s = W->scalar_product(*fx, *fy);
  W->release_block(fx); W->release_block(fy);
  return 0;
  ORC_CATCH(1)
}
int orc_world_field_max_mean(void* h, int dir, int data_loc, const double* x, double* mx, double* mean) {
  ORC_TRY
  auto* W = (World*)h;
  WField* fx = W->get_block(dir);
  W->set_field_from_global(*fx, x, data_loc);
  W->field_max_mean(*mx, *mean, *fx);
  W->release_block(fx);
  return 0;
  ORC_CATCH(1)
}
// mesh / allocator facts for unit tests: out = local vert dims(3), cell dims(3), padded cart dims(3),
// n_groups(3), BCs of rank `r` (6)
int orc_world_mesh_info(void* h, int r, int* out) {
  auto* W = (World*)h;
  const RankMesh& m = W->rm[r];
  for (int d = 0; d < 3; ++d) {
    out[d] = m.vert_dims[d]; out[3 + d] = m.cell_dims[d]; out[6 + d] = W->alloc.padded(DIR_C)[d];
    out[9 + d] = W->alloc.n_groups(d + 1); out[12 + 2 * d] = m.BCs[d][0]; out[13 + 2 * d] = m.BCs[d][1];
  }
  return 0;
}
// coordinates of rank 0 along `dir` (0-based): vert_coords, vert_ds, vert_ds2, vert_d2s (n_vert each),
// midp_coords, midp_ds (n_cell each)
int orc_world_geo(void* h, int dir, double* vc, double* vds, double* vds2, double* vd2s, double* mc, double* mds) {
  auto* W = (World*)h;
  const Geo& g = W->rm[0].geo;
  for (size_t i = 0; i < g.vert_coords[dir].size(); ++i) {
    vc[i] = g.vert_coords[dir][i]; vds[i] = g.vert_ds[dir][i]; vds2[i] = g.vert_ds2[dir][i]; vd2s[i] = g.vert_d2s[dir][i];
  }
  for (size_t i = 0; i < g.midp_coords[dir].size(); ++i) { mc[i] = g.midp_coords[dir][i]; mds[i] = g.midp_ds[dir][i]; }
  return 0;
}

}  // extern "C"
