// Host layer of the cuda_c backend: the reference's solver-side modules, restated in C++ on top of the
// C ABI of include/x3d2c.h (the Fortran toolchain is absent in this image; with one, these classes are
// the reference's own Fortran modules and only cuda_c_backend_t below becomes an iso_c_binding shim).
//
//   mesh_t / par_t / geo_t       src/mesh.f90:37-306, src/mesh_content.f90:6-253
//   allocator_t / field_t        src/allocator.f90:9-162, src/field.f90:5-83
//   cuda_c_backend_t             a new `extends(base_backend_t)` (src/backend/backend.f90:13-62),
//                                modelled on cuda_backend_t (src/backend/cuda/backend.f90:42-81)
//   poisson_fft_t (000)          src/poisson_fft.f90:120-226,654-882
//   vector_calculus_t            src/vector_calculus.f90:40-332
//   time_intg_t                  src/time_integrator.f90:70-300
//   solver_t                     src/solver.f90:111-389,603-739
//   base_case_t%run / case_tgv_t src/case/base_case.f90:139-289, src/case/tgv.f90:41-72
//   monitoring_t                 src/postprocess/monitoring.f90:46-90
#pragma once
#include <array>
#include <complex>

#include "../../../include/x3d2h.h"
#include "tdsops.hpp"

namespace x3d2h {

using cplx = std::complex<double>;

struct Config {
  int dims_global[3];
  int nproc_dir[3];
  double L_global[3];
  int bc[3][2];
  double Re = 1600, dt = 1e-3;
  std::string time_intg = "RK3", der1st = "compact6", der2nd = "compact6", interpl = "classic",
              stagder = "compact6";
  int rank = 0, nproc = 1, device = -1, flags = 0;
  const void* nccl_unique_id = nullptr;
  std::string stretching[3] = {"uniform", "uniform", "uniform"};
  double beta[3] = {1.0, 1.0, 1.0};
};

// ------------------------------------------------------------------------------------ mesh
struct Mesh {
  // grid_t
  int global_vert_dims[3], global_cell_dims[3], vert_dims[3], cell_dims[3];
  bool periodic_BC[3];
  int BCs_global[3][2], BCs[3][2];
  // par_t
  int nrank = 0, nproc = 1, nrank_dir[3], nproc_dir[3], n_offset[3], pprev[3], pnext[3];
  // geo_t (mesh_content.f90:6-27)
  double d[3], L[3], alpha[3] = {0, 0, 0}, beta[3] = {1, 1, 1};
  std::string stretching[3];
  bool stretched[3] = {false, false, false};
  std::vector<double> vert_coords[3], midp_coords[3], vert_ds[3], vert_ds2[3], vert_d2s[3], midp_ds[3], midp_ds2[3],
      midp_d2s[3];

  // mesh.f90:37-158 with the generic decomposition of :160-194
  void init(const Config& c) {
    nrank = c.rank;
    nproc = c.nproc;
    for (int dir = 0; dir < 3; ++dir) {
      BCs_global[dir][0] = c.bc[dir][0]; BCs_global[dir][1] = c.bc[dir][1];
      bool p0 = c.bc[dir][0] == BC_PERIODIC, p1 = c.bc[dir][1] == BC_PERIODIC;
      if (p0 != p1) fail("BCs are incompatible: in a direction make sure to have either both sides periodic or none.");
      periodic_BC[dir] = p0 && p1;
      global_vert_dims[dir] = c.dims_global[dir];
      global_cell_dims[dir] = periodic_BC[dir] ? c.dims_global[dir] : c.dims_global[dir] - 1;
      nproc_dir[dir] = c.nproc_dir[dir];
      L[dir] = c.L_global[dir];
      d[dir] = L[dir] / global_cell_dims[dir];
    }
    if (nproc_dir[0] * nproc_dir[1] * nproc_dir[2] != nproc) {  // xcompact.f90:69-74
      nproc_dir[0] = 1; nproc_dir[1] = 1; nproc_dir[2] = nproc;
    }
    // global_ranks = reshape([0..nproc-1], [px, py, pz]) (x fastest), mesh.f90:186-187
    auto rank_of = [&](int px, int py, int pz) { return px + nproc_dir[0] * (py + nproc_dir[1] * pz); };
    int pos[3] = {nrank % nproc_dir[0], (nrank / nproc_dir[0]) % nproc_dir[1], nrank / (nproc_dir[0] * nproc_dir[1])};
    for (int dir = 0; dir < 3; ++dir) {
      nrank_dir[dir] = pos[dir];
      if (global_vert_dims[dir] % nproc_dir[dir]) fail("dims_global must be divisible by nproc_dir");
      vert_dims[dir] = global_vert_dims[dir] / nproc_dir[dir];
      const int np = nproc_dir[dir];
      int prev[3] = {pos[0], pos[1], pos[2]}, next[3] = {pos[0], pos[1], pos[2]};
      prev[dir] = ((pos[dir] - 1) % np + np) % np;  // mesh_content.f90:87-100
      next[dir] = (pos[dir] + 1) % np;
      pprev[dir] = rank_of(prev[0], prev[1], prev[2]);
      pnext[dir] = rank_of(next[0], next[1], next[2]);
      const bool first = pos[dir] == 0, last = pos[dir] + 1 == np;
      cell_dims[dir] = (last && !periodic_BC[dir]) ? vert_dims[dir] - 1 : vert_dims[dir];  // :104-121
      n_offset[dir] = vert_dims[dir] * pos[dir];
      if (first && last) { BCs[dir][0] = BCs_global[dir][0]; BCs[dir][1] = BCs_global[dir][1]; }  // mesh.f90:119-136
      else if (first) { BCs[dir][0] = BCs_global[dir][0]; BCs[dir][1] = BC_HALO; }
      else if (last) { BCs[dir][0] = BC_HALO; BCs[dir][1] = BCs_global[dir][1]; }
      else { BCs[dir][0] = BC_HALO; BCs[dir][1] = BC_HALO; }
      stretching[dir] = c.stretching[dir];
      beta[dir] = c.beta[dir];
      obtain_coordinates(dir);
    }
  }

  // geo_t%obtain_coordinates, mesh_content.f90:142-253
  void obtain_coordinates(int dir) {
    const int nv = vert_dims[dir], nc = cell_dims[dir];
    vert_coords[dir].assign(nv, 0); vert_ds[dir].assign(nv, 1); vert_ds2[dir].assign(nv, 1); vert_d2s[dir].assign(nv, 0);
    midp_coords[dir].assign(nc, 0); midp_ds[dir].assign(nc, 1); midp_ds2[dir].assign(nc, 1); midp_d2s[dir].assign(nc, 0);
    if (stretching[dir] == "uniform") {
      stretched[dir] = false;
      alpha[dir] = 0;
      for (int i = 1; i <= nv; ++i) vert_coords[dir][i - 1] = (n_offset[dir] + i - 1) * d[dir];
      for (int i = 1; i <= nc; ++i) midp_coords[dir][i - 1] = (n_offset[dir] + i - 0.5) * d[dir];
      return;
    }
    stretched[dir] = true;
    const std::string& st = stretching[dir];
    if (st != "centred" && st != "top-bottom" && st != "bottom") fail("Invalid stretching type");
    const double L_inf = L[dir] / 2, be = beta[dir];
    if (be <= 2.220446049250313e-16) fail("Invalid beta in domain_settings");
    const double al = std::fabs((L_inf - std::sqrt((pi * be) * (pi * be) + L_inf * L_inf)) / (2 * be * L_inf));
    alpha[dir] = al;
    const double r = std::sqrt((al * be + 1) / (al * be));
    const double cst = std::sqrt(be) / (2 * std::sqrt(al) * std::sqrt(al * be + 1));
    const double s = d[dir] / L[dir];
    auto eta = [&](double idx) { return st == "centred" ? idx * s : (st == "top-bottom" ? idx * s - 0.5 : idx * s / 2 - 0.5); };
    auto fill = [&](double y, double& coord, double& ds, double& ds2, double& d2s) {
      const double sp = std::sin(pi * y), cp = std::cos(pi * y);
      coord = cst * std::atan2(r * sp, cp) * (2 * al * be - std::cos(2 * pi * y) + 1) / (sp * sp + al * be) + pi * cst;
      ds = L[dir] * (al / pi + sp * sp / (pi * be));
      ds2 = ds * ds;
      d2s = 2 * cp * sp / be;
    };
    for (int i = 1; i <= nv; ++i)
      fill(eta((double)(i + n_offset[dir] - 1)), vert_coords[dir][i - 1], vert_ds[dir][i - 1], vert_ds2[dir][i - 1],
           vert_d2s[dir][i - 1]);
    for (int i = 1; i <= nc; ++i)
      fill(eta(i + n_offset[dir] - 0.5), midp_coords[dir][i - 1], midp_ds[dir][i - 1], midp_ds2[dir][i - 1],
           midp_d2s[dir][i - 1]);
    if (st == "centred") {
      for (auto& x : vert_coords[dir]) x -= L_inf;
      for (auto& x : midp_coords[dir]) x -= L_inf;
    } else if (st == "bottom") {
      for (auto& x : vert_coords[dir]) x = 2 * x;
      for (auto& x : vert_d2s[dir]) x = x / 2;
      for (auto& x : midp_coords[dir]) x = 2 * x;
      for (auto& x : midp_d2s[dir]) x = x / 2;
    }
  }

  void get_dims(int dims[3], int data_loc, bool global = false) const {  // mesh.f90:196-261
    const int* v = global ? global_vert_dims : vert_dims;
    const int* c = global ? global_cell_dims : cell_dims;
    switch (data_loc) {
      case VERT: dims[0] = v[0]; dims[1] = v[1]; dims[2] = v[2]; break;
      case CELL: dims[0] = c[0]; dims[1] = c[1]; dims[2] = c[2]; break;
      case X_FACE: dims[0] = v[0]; dims[1] = c[1]; dims[2] = c[2]; break;
      case Y_FACE: dims[0] = c[0]; dims[1] = v[1]; dims[2] = c[2]; break;
      case Z_FACE: dims[0] = c[0]; dims[1] = c[1]; dims[2] = v[2]; break;
      case X_EDGE: dims[0] = c[0]; dims[1] = v[1]; dims[2] = v[2]; break;
      case Y_EDGE: dims[0] = v[0]; dims[1] = c[1]; dims[2] = v[2]; break;
      case Z_EDGE: dims[0] = v[0]; dims[1] = v[1]; dims[2] = c[2]; break;
      default: fail("Unknown location in get_dims_dataloc");
    }
  }
  int get_n(int dir, int data_loc) const {  // mesh.f90:263-306
    int n_cell = cell_dims[dir - 1], n_vert = vert_dims[dir - 1], n = n_vert;
    switch (data_loc) {
      case CELL: n = n_cell; break;
      case VERT: n = n_vert; break;
      case X_FACE: if (dir != DIR_X) n = n_cell; break;
      case Y_FACE: if (dir != DIR_Y) n = n_cell; break;
      case Z_FACE: if (dir != DIR_Z) n = n_cell; break;
      case X_EDGE: if (dir == DIR_X) n = n_cell; break;
      case Y_EDGE: if (dir == DIR_Y) n = n_cell; break;
      case Z_EDGE: if (dir == DIR_Z) n = n_cell; break;
      default: fail("Unknown direction in get_n_dir");
    }
    return n;
  }
};

// ------------------------------------------------------------------------------------ field / allocator
struct Field {  // field.f90:5-23 with the device pointer of cuda_field_t (cuda/allocator.f90:9-29)
  double* dev = nullptr;
  int dir = DIR_X, data_loc = NULL_LOC, id = 0;
  Field* next = nullptr;
};

[[noreturn]] inline void fail_c(const char* what) { fail(std::string(what) + ": " + x3d2c_last_error()); }
#define X3D2H_CALL(expr)            \
  do {                              \
    if ((expr) != X3D2C_OK) fail_c(#expr); \
  } while (0)

class Allocator {  // allocator.f90:9-162; storage comes from x3d2c_field_alloc (create_block override)
 public:
  x3d2c_ctx* ctx = nullptr;
  int next_id = 0;
  Field* first = nullptr;
  std::vector<std::unique_ptr<Field>> all;
  int dims_padded[3] = {0, 0, 0}, n_groups_dir[3] = {0, 0, 0};
  long long ngrid = 0;

  void init(x3d2c_ctx* c) {
    ctx = c;
    X3D2H_CALL(x3d2c_get_padded_dims(ctx, dims_padded, n_groups_dir, &ngrid));
  }
  Field* get_block(int dir, int data_loc = NULL_LOC) {
    if (!first) {
      all.emplace_back(new Field);
      Field* f = all.back().get();
      f->id = ++next_id;
      X3D2H_CALL(x3d2c_field_alloc(ctx, &f->dev));
      f->next = first;
      first = f;
    }
    Field* h = first;
    first = first->next;
    h->next = nullptr;
    h->dir = dir;
    h->data_loc = data_loc;
    return h;
  }
  void release_block(Field* h) {
    h->next = first;
    first = h;
  }
  void destroy() {
    for (auto& f : all)
      if (f->dev) x3d2c_field_free(ctx, f->dev);
    all.clear();
    first = nullptr;
  }
};

struct DevTdsops {  // cuda_tdsops_t role (cuda/tdsops.f90:9-23): host tables + device handle
  Tdsops t;
  x3d2c_tdsops* h = nullptr;
};
struct DevDirps {
  DevTdsops der1st, der1st_sym, der2nd, der2nd_sym, stagder_v2p, stagder_p2v, interpl_v2p, interpl_p2v;
  int dir = 0;
};

// ------------------------------------------------------------------------------------ backend
class Backend {  // cuda_c_backend_t: every method is one deferred procedure of base_backend_t
 public:
  x3d2c_ctx* ctx = nullptr;
  x3d2c_poisson* poisson = nullptr;
  Mesh* mesh = nullptr;
  Allocator* allocator = nullptr;

  void alloc_tdsops(DevTdsops& o, int n_tds, double delta, const std::string& operation, const std::string& scheme,
                    int bc_start, int bc_end, const double* stretch = nullptr, const double* stretch_correct = nullptr,
                    int n_halo = 4, const std::string& from_to = "", bool sym = false) {
    o.t = tdsops_init(n_tds, delta, operation, scheme, bc_start, bc_end, stretch, stretch_correct, n_halo, from_to, sym);
    const Tdsops& t = o.t;
    double cs[36], ce[36];
    for (int i = 1; i <= 4; ++i)
      for (int k = 1; k <= 9; ++k) { cs[(i - 1) * 9 + k - 1] = t.coeffs_s[i][k]; ce[(i - 1) * 9 + k - 1] = t.coeffs_e[i][k]; }
    X3D2H_CALL(x3d2c_tdsops_create(ctx, t.n_tds, t.n_rhs, t.move, t.periodic, &t.coeffs[1], cs, ce, &t.dist_fw[1],
                                   &t.dist_bw[1], &t.dist_sa[1], &t.dist_sc[1], &t.dist_af[1], &t.stretch[1],
                                   &t.stretch_correct[1], &o.h));
  }
  void transeq(int dir, Field& du, Field& dv, Field& dw, const Field& u, const Field& v, const Field& w, double nu,
               const DevDirps& dp) {
    X3D2H_CALL(x3d2c_transeq(ctx, dir, du.dev, dv.dev, dw.dev, u.dev, v.dev, w.dev, nu, dp.der1st.h, dp.der1st_sym.h,
                             dp.der2nd.h, dp.der2nd_sym.h));
    du.data_loc = u.data_loc; dv.data_loc = v.data_loc; dw.data_loc = w.data_loc;  // omp/backend.f90:333
  }
  // reorder(u, v, w; rdr_in) + transeq in `dir`; transeq_r_fused: the backend does it without a reorder pass
  bool transeq_r_fused(int dir, const DevDirps& dp, int rdr_in) {
    return x3d2c_transeq_r_fused(ctx, dir, dp.der1st.h, dp.der1st_sym.h, dp.der2nd.h, dp.der2nd_sym.h, rdr_in) == 1;
  }
  void transeq_r(int dir, Field& du, Field& dv, Field& dw, const Field& u, const Field& v, const Field& w, double nu,
                 const DevDirps& dp, int rdr_in) {
    X3D2H_CALL(x3d2c_transeq_r(ctx, dir, du.dev, dv.dev, dw.dev, u.dev, v.dev, w.dev, nu, dp.der1st.h, dp.der1st_sym.h,
                               dp.der2nd.h, dp.der2nd_sym.h, rdr_in));
    du.data_loc = u.data_loc; dv.data_loc = v.data_loc; dw.data_loc = w.data_loc;
  }
  void transeq_x(Field& du, Field& dv, Field& dw, const Field& u, const Field& v, const Field& w, double nu, const DevDirps& dp) { transeq(DIR_X, du, dv, dw, u, v, w, nu, dp); }
  void transeq_y(Field& du, Field& dv, Field& dw, const Field& u, const Field& v, const Field& w, double nu, const DevDirps& dp) { transeq(DIR_Y, du, dv, dw, u, v, w, nu, dp); }
  void transeq_z(Field& du, Field& dv, Field& dw, const Field& u, const Field& v, const Field& w, double nu, const DevDirps& dp) { transeq(DIR_Z, du, dv, dw, u, v, w, nu, dp); }
  // omp/backend.f90:186-233
  void transeq_species(Field& dspec, const Field& uvw, const Field& spec, double nu_s, const DevDirps& dp, bool sync) {
    if (dspec.dir != dp.dir || uvw.dir != dp.dir || spec.dir != dp.dir) fail("DIR mismatch between fields in transeq_species.");
    X3D2H_CALL(x3d2c_transeq_species(ctx, dp.dir, dspec.dev, uvw.dev, spec.dev, nu_s, dp.der1st.h, dp.der1st_sym.h,
                                     dp.der2nd.h, sync ? 1 : 0));
    dspec.data_loc = spec.data_loc;
  }
  // omp/backend.f90:816-881: rank-local signed max and sum of one plane
  void slice_max_sum(double& mx, double& sum, const Field& f, int i_slice, int enforced_data_loc = -999) {
    if (f.data_loc == NULL_LOC && enforced_data_loc == -999) fail("slice_max_sum: the field has no valid data_loc");
    if (f.dir == DIR_C) fail("slice_max_sum does not support DIR_C fields!");
    X3D2H_CALL(x3d2c_slice_max_sum(ctx, f.dir, enforced_data_loc != -999 ? enforced_data_loc : f.data_loc, f.dev, i_slice, &mx, &sum));
  }
  // omp/backend.f90:616-649
  void compute_vorticity(Field& out, const Field* const g[9]) {
    X3D2H_CALL(x3d2c_compute_vorticity(ctx, out.dev, g[0]->dev, g[1]->dev, g[2]->dev, g[3]->dev, g[4]->dev, g[5]->dev,
                                       g[6]->dev, g[7]->dev, g[8]->dev));
  }
  void compute_qcriterion(Field& out, const Field* const g[9]) {
    X3D2H_CALL(x3d2c_compute_qcriterion(ctx, out.dev, g[0]->dev, g[1]->dev, g[2]->dev, g[3]->dev, g[4]->dev, g[5]->dev,
                                        g[6]->dev, g[7]->dev, g[8]->dev));
  }
  void tds_solve(Field& du, const Field& u, const DevTdsops& ops) {  // cuda/backend.f90:449-470
    if (u.dir != du.dir) fail("DIR mismatch between fields in tds_solve.");
    if (u.data_loc != NULL_LOC) du.data_loc = move_data_loc(u.data_loc, u.dir, ops.t.move);
    X3D2H_CALL(x3d2c_tds_solve(ctx, u.dir, du.dev, u.dev, ops.h));
  }
  // fused combinations (x3d2c.h): identical to the reference's call sequences, one pass over memory on the fast path
  void tds_solve_sum(Field& out, const Field& a, const DevTdsops& opa, const Field& b, const DevTdsops& opb) {
    if (a.dir != out.dir || b.dir != out.dir) fail("DIR mismatch between fields in tds_solve.");
    if (a.data_loc != NULL_LOC) out.data_loc = move_data_loc(a.data_loc, a.dir, opa.t.move);
    X3D2H_CALL(x3d2c_tds_solve_sum(ctx, a.dir, out.dev, a.dev, opa.h, b.dev, opb.h));
  }
  void tds_solve_dual(Field& oa, Field& ob, const Field& u, const DevTdsops& opa, const DevTdsops& opb) {
    if (u.dir != oa.dir || u.dir != ob.dir) fail("DIR mismatch between fields in tds_solve.");
    if (u.data_loc != NULL_LOC) {
      oa.data_loc = move_data_loc(u.data_loc, u.dir, opa.t.move);
      ob.data_loc = move_data_loc(u.data_loc, u.dir, opb.t.move);
    }
    X3D2H_CALL(x3d2c_tds_solve_dual(ctx, u.dir, oa.dev, ob.dev, u.dev, opa.h, opb.h));
  }
  // ... through reorders (x3d2c.h): rdr_in brings the input(s) into DIR_`dir`, rdr_out takes the output(s) out of it
  void check_r(int dir, const Field& in, const Field& out, int rdr_in, int rdr_out) {
    if (in.dir != (rdr_in ? rdr_in / 10 : dir) || out.dir != (rdr_out ? rdr_out % 10 : dir))
      fail("DIR mismatch between fields in tds_solve.");
  }
  void tds_solve_r(int dir, Field& out, const Field& u, const DevTdsops& op, int rdr_in, int rdr_out) {
    check_r(dir, u, out, rdr_in, rdr_out);
    if (u.data_loc != NULL_LOC) out.data_loc = move_data_loc(u.data_loc, dir, op.t.move);
    X3D2H_CALL(x3d2c_tds_solve_r(ctx, dir, out.dev, u.dev, op.h, rdr_in, rdr_out));
  }
  void tds_solve_sum_r(int dir, Field& out, const Field& a, const DevTdsops& opa, const Field& b, const DevTdsops& opb,
                       int rdr_in, int rdr_out) {
    check_r(dir, a, out, rdr_in, rdr_out);
    check_r(dir, b, out, rdr_in, rdr_out);
    if (a.data_loc != NULL_LOC) out.data_loc = move_data_loc(a.data_loc, dir, opa.t.move);
    X3D2H_CALL(x3d2c_tds_solve_sum_r(ctx, dir, out.dev, a.dev, opa.h, b.dev, opb.h, rdr_in, rdr_out));
  }
  void tds_solve_dual_r(int dir, Field& oa, Field& ob, const Field& u, const DevTdsops& opa, const DevTdsops& opb,
                        int rdr_in, int rdr_out) {
    check_r(dir, u, oa, rdr_in, rdr_out);
    check_r(dir, u, ob, rdr_in, rdr_out);
    if (u.data_loc != NULL_LOC) {
      oa.data_loc = move_data_loc(u.data_loc, dir, opa.t.move);
      ob.data_loc = move_data_loc(u.data_loc, dir, opb.t.move);
    }
    X3D2H_CALL(x3d2c_tds_solve_dual_r(ctx, dir, oa.dev, ob.dev, u.dev, opa.h, opb.h, rdr_in, rdr_out));
  }
  void tds_solve_axpy(Field& y, double a, const Field& u, const DevTdsops& op) {
    if (u.dir != y.dir) fail("DIR mismatch between fields in tds_solve.");
    X3D2H_CALL(x3d2c_tds_solve_axpy(ctx, u.dir, y.dev, a, u.dev, op.h));
  }
  void tds_solve_axpy_r(int dir, Field& y, double a, const Field& u, const DevTdsops& op, int rdr_in) {
    check_r(dir, u, y, rdr_in, 0);
    X3D2H_CALL(x3d2c_tds_solve_axpy_r(ctx, dir, y.dev, a, u.dev, op.h, rdr_in));
  }
  void reorder(Field& u_, const Field& u, int direction) {
    X3D2H_CALL(x3d2c_reorder(ctx, direction, u_.dev, u.dev));
    u_.data_loc = u.data_loc;
  }
  void reorder_x2yz(Field& u_y, Field& u_z, const Field& u) {
    if (u.dir != DIR_X || u_y.dir != DIR_Y || u_z.dir != DIR_Z) fail("reorder_x2yz: fields must be DIR_X -> DIR_Y, DIR_Z");
    X3D2H_CALL(x3d2c_reorder_x2yz(ctx, u_y.dev, u_z.dev, u.dev));
    u_y.data_loc = u.data_loc; u_z.data_loc = u.data_loc;
  }
  void sum_yintox(Field& u, const Field& u_) { X3D2H_CALL(x3d2c_sum_yintox(ctx, u.dev, u_.dev)); }
  void sum_zintox(Field& u, const Field& u_) { X3D2H_CALL(x3d2c_sum_zintox(ctx, u.dev, u_.dev)); }
  void sum_yzintox(Field& u, const Field& u_y, const Field& u_z) { X3D2H_CALL(x3d2c_sum_yzintox(ctx, u.dev, u_y.dev, u_z.dev)); }
  // backend.f90:255-308 (channel case): volume integral of a DIR_X field, shift, wall rows from a field
  double field_volume_integral(const Field& f) {
    if (f.data_loc == NULL_LOC) fail("You must set the data_loc before calling volume integral.");
    if (f.dir != DIR_X) fail("Volume integral can only be called on DIR_X fields.");
    double s = 0.0;
    X3D2H_CALL(x3d2c_field_volume_integral(ctx, f.data_loc, f.dev, &s));
    return s;
  }
  void field_shift(Field& f, double a) { X3D2H_CALL(x3d2c_field_shift(ctx, f.dev, a)); }
  void field_set_face_from_field(Field& f, const Field& f_start, double c_end, int face, double flow_rate_diff = 0.0) {
    if (f.dir != DIR_X || f_start.dir != DIR_X) fail("field_set_face_from_field: only supported for DIR_X fields.");
    if (f.data_loc == NULL_LOC) fail("field_set_face_from_field: requires a valid data_loc.");
    X3D2H_CALL(x3d2c_field_set_face_from_field(ctx, f.dev, f_start.dev, f.data_loc, c_end, face, flow_rate_diff));
  }
  // sum_yzintox(u, u_y, u_z) + veclincomb(out, base, terms + (c_u, u)) in one pass
  void sum_yzintox_lincomb(Field& u, const Field& u_y, const Field& u_z, bool store_u, Field& out, const Field& base,
                           const std::vector<std::pair<double, const Field*>>& terms, double c_u) {
    if (terms.size() > 3) fail("sum_yzintox_lincomb takes at most 3 terms");
    double c[3];
    const double* x[3];
    for (size_t k = 0; k < terms.size(); ++k) { c[k] = terms[k].first; x[k] = terms[k].second->dev; }
    X3D2H_CALL(x3d2c_sum_yzintox_lincomb(ctx, u.dev, u_y.dev, u_z.dev, store_u ? 1 : 0, out.dev, base.dev, (int)terms.size(),
                                         c, x, c_u));
  }
  void veccopy(Field& dst, const Field& src) {
    if (src.dir != dst.dir) fail("Called vector copy with incompatible fields");
    if (dst.dir == DIR_C) fail("veccopy does not support DIR_C fields");
    X3D2H_CALL(x3d2c_veccopy(ctx, dst.dev, src.dev));
  }
  void vecadd(double a, const Field& x, double b, Field& y) {
    if (x.dir != y.dir) fail("Called vector add with incompatible fields");
    if (y.dir == DIR_C) fail("vecadd does not support DIR_C fields");
    X3D2H_CALL(x3d2c_vecadd(ctx, a, x.dev, b, y.dev));
  }
  // out = base, then out = c[k] x[k] + out for every term: the vecadd chains of the time integrators in one pass
  void veclincomb(Field& out, const Field& base, const std::vector<std::pair<double, const Field*>>& terms) {
    if (terms.empty() || terms.size() > 4) fail("veclincomb takes 1 to 4 terms");
    double c[4];
    const double* x[4];
    for (size_t k = 0; k < terms.size(); ++k) {
      if (terms[k].second->dir != out.dir) fail("Called vector add with incompatible fields");
      c[k] = terms[k].first;
      x[k] = terms[k].second->dev;
    }
    X3D2H_CALL(x3d2c_veclincomb(ctx, out.dev, base.dev, (int)terms.size(), c, x));
  }
  double scalar_product(const Field& x, const Field& y) {
    if (x.data_loc == NULL_LOC || y.data_loc == NULL_LOC) fail("You must set the data_loc before calling scalar product");
    if (x.data_loc != y.data_loc) fail("Called scalar product with incompatible fields");
    if (x.dir != y.dir) fail("Called scalar product with incompatible fields");
    double s = 0;
    X3D2H_CALL(x3d2c_scalar_product(ctx, x.dir, x.data_loc, x.dev, y.dev, &s));
    return s;
  }
  void field_max_mean(double& mx, double& mean, const Field& f, int enforced_data_loc = -999) {
    if (f.data_loc == NULL_LOC && enforced_data_loc == -999) fail("field_max_mean: the field has no valid data_loc");
    X3D2H_CALL(x3d2c_field_max_mean(ctx, f.dir, enforced_data_loc != -999 ? enforced_data_loc : f.data_loc, f.dev, &mx, &mean));
  }
  // backend.f90:402-466 (get_field_data / set_field_data through a DIR_C block); `data` is the full padded block
  void set_field_data(Field& f, const double* data_padded_c) {
    int rdr = get_rdr_from_dirs(DIR_C, f.dir);
    if (rdr) {
      Field* tmp = allocator->get_block(DIR_C, f.data_loc);
      X3D2H_CALL(x3d2c_copy_data_to_f(ctx, tmp->dev, data_padded_c));
      int loc = f.data_loc;
      reorder(f, *tmp, rdr);
      f.data_loc = loc;
      allocator->release_block(tmp);
    } else {
      X3D2H_CALL(x3d2c_copy_data_to_f(ctx, f.dev, data_padded_c));
    }
  }
  void get_field_data(double* data_padded_c, const Field& f) {
    int rdr = get_rdr_from_dirs(f.dir, DIR_C);
    if (rdr) {
      Field* tmp = allocator->get_block(DIR_C);
      reorder(*tmp, f, rdr);
      X3D2H_CALL(x3d2c_copy_f_to_data(ctx, data_padded_c, tmp->dev));
      allocator->release_block(tmp);
    } else {
      X3D2H_CALL(x3d2c_copy_f_to_data(ctx, data_padded_c, f.dev));
    }
  }
};

// ------------------------------------------------------------------------------------ poisson_fft_t (000)
struct PoissonFFT {
  int nx_glob, ny_glob, nz_glob, nx_spec, ny_spec, nz_spec, sp_st[3];
  std::vector<double> ax, bx, ay, by, az, bz;  // 1-based
  std::vector<cplx> kx, ky, kz, exs, eys, ezs, k2x, k2y, k2z;
  std::vector<cplx> waves;
  // non-periodic y (poisson_fft.f90:29-37,176-190): bc_case 10; stretched y adds the pentadiagonal spectral operator
  int bc_case = 0;       // 0: 000, 10: 010
  int stretched = 0;     // 0 uniform, 1 'centred' / 'top-bottom' (odd and even modes decouple), 2 'bottom'
  std::vector<double> trans[3];                    // trans_x/y/z_re (== _im), 1-based
  std::vector<double> a_odd, a_even;               // (nx_spec, rows, nz_spec, 5) Fortran order; 'bottom': a_odd = a_re
  int penta_rows = 0;                              // ny_spec / 2 (decoupled families) or ny_spec ('bottom')

  // poisson_fft.f90:833-882
  static void wave_numbers(std::vector<double>& a, std::vector<double>& b, std::vector<cplx>& k, std::vector<cplx>& e,
                           std::vector<cplx>& k2, int n, double L, double d, bool periodic, double c_a, double c_b,
                           double c_alpha) {
    a.assign(n + 1, 0); b.assign(n + 1, 0);
    k.assign(n + 1, 0); e.assign(n + 1, 0); k2.assign(n + 1, 0);
    for (int i = 1; i <= n; ++i) {
      if (periodic) { a[i] = std::sin((i - 1) * pi / n); b[i] = std::cos((i - 1) * pi / n); }
      else { a[i] = std::sin((i - 1) * pi / 2 / n); b[i] = std::cos((i - 1) * pi / 2 / n); }
    }
    auto one = [&](int i, double w) {
      double wp = c_a * 2 * d * std::sin(0.5 * w) + c_b * 2 * d * std::sin(1.5 * w);
      wp = wp / (1.0 + 2 * c_alpha * std::cos(w));
      k[i] = cplx(1.0, 1.0) * (n * wp / L);
      e[i] = cplx(1.0, 1.0) * (n * w / L);
      const double q = n * wp / L;
      k2[i] = cplx(1.0, 1.0) * (q * q);
    };
    if (periodic) {
      for (int i = 1; i <= n / 2 + 1; ++i) one(i, 2 * pi * (i - 1) / n);
      for (int i = n / 2 + 2; i <= n; ++i) { k[i] = k[n - i + 2]; e[i] = e[n - i + 2]; k2[i] = k2[n - i + 2]; }
    } else {
      for (int i = 1; i <= n; ++i) one(i, pi * (i - 1) / n);
    }
  }

  // base_init (:120-204) + waves_set, periodic-z branch (:777-819)
  void base_init(const Mesh& mesh, const DevDirps& xd, const DevDirps& yd, const DevDirps& zd, const int n_spec[3],
                 const int n_sp_st[3]) {
    if (mesh.nproc_dir[0] != 1) fail("nproc_dir in x-dir must be 1");
    nx_glob = mesh.global_cell_dims[0]; ny_glob = mesh.global_cell_dims[1]; nz_glob = mesh.global_cell_dims[2];
    nx_spec = n_spec[0]; ny_spec = n_spec[1]; nz_spec = n_spec[2];
    for (int q = 0; q < 3; ++q) sp_st[q] = n_sp_st[q];
    const bool p000 = mesh.periodic_BC[0] && mesh.periodic_BC[1] && mesh.periodic_BC[2];
    const bool p010 = mesh.periodic_BC[0] && !mesh.periodic_BC[1] && mesh.periodic_BC[2];
    if (!p000 && !p010) fail("Requested BCs are not supported in the cuda_c FFT-based Poisson solver (000 and 010 only)");
    if (p010 && mesh.nproc > 1) fail("Multiple ranks are not yet supported for non-periodic BCs!");  // poisson_fft.f90:178-180
    if (mesh.stretched[0] || mesh.stretched[2])
      fail("FFT based Poisson solver does not support stretching in x- or z-directions!");
    bc_case = p010 ? 10 : 0;
    const Tdsops &sx = xd.stagder_v2p.t, &sy = yd.stagder_v2p.t, &sz_ = zd.stagder_v2p.t;
    const Tdsops &ix = xd.interpl_v2p.t, &iy = yd.interpl_v2p.t, &iz = zd.interpl_v2p.t;
    wave_numbers(ax, bx, kx, exs, k2x, nx_glob, mesh.L[0], mesh.d[0], mesh.periodic_BC[0], sx.a, sx.b, sx.alpha);
    wave_numbers(ay, by, ky, eys, k2y, ny_glob, mesh.L[1], mesh.d[1], mesh.periodic_BC[1], sy.a, sy.b, sy.alpha);
    wave_numbers(az, bz, kz, ezs, k2z, nz_glob, mesh.L[2], mesh.d[2], mesh.periodic_BC[2], sz_.a, sz_.b, sz_.alpha);
    waves.assign((size_t)nx_spec * ny_spec * nz_spec, 0);
    for (int k = 1; k <= nz_spec; ++k)
      for (int j = 1; j <= ny_spec; ++j)
        for (int i = 1; i <= nx_spec; ++i) {
          const int ixx = i + sp_st[0], iyy = j + sp_st[1], izz = k + sp_st[2];
          const double rlexs = exs[ixx].real() * mesh.d[0], rleys = eys[iyy].real() * mesh.d[1], rlezs = ezs[izz].real() * mesh.d[2];
          const double xtt = 2 * (ix.a * std::cos(rlexs * 0.5) + ix.b * std::cos(rlexs * 1.5) + ix.c * std::cos(rlexs * 2.5) + ix.d * std::cos(rlexs * 3.5));
          const double ytt = 2 * (iy.a * std::cos(rleys * 0.5) + iy.b * std::cos(rleys * 1.5) + iy.c * std::cos(rleys * 2.5) + iy.d * std::cos(rleys * 3.5));
          const double ztt = 2 * (iz.a * std::cos(rlezs * 0.5) + iz.b * std::cos(rlezs * 1.5) + iz.c * std::cos(rlezs * 2.5) + iz.d * std::cos(rlezs * 3.5));
          const double xt1 = 1.0 + 2 * ix.alpha * std::cos(rlexs);
          const double yt1 = 1.0 + 2 * iy.alpha * std::cos(rleys);
          const double zt1 = 1.0 + 2 * iz.alpha * std::cos(rlezs);
          const double fx = (ytt / yt1) * (ztt / zt1), fy = (xtt / xt1) * (ztt / zt1), fz = (xtt / xt1) * (ytt / yt1);
          const cplx xt2 = k2x[ixx] * (fx * fx), yt2 = k2y[iyy] * (fy * fy), zt2 = k2z[izz] * (fz * fz);
          waves[(i - 1) + (size_t)nx_spec * ((j - 1) + (size_t)ny_spec * (k - 1))] = xt2 + yt2 + zt2;
        }
    if (p010 && mesh.stretched[1]) stretching_matrix(mesh, xd, yd, zd);
  }

  // stretching_matrix (poisson_fft.f90:275-652; JCP 228 (2009) 5989, sec. 5). The metric of the stretched mesh is
  // a0 + 2 a1 cos(2 pi eta), so multiplying by it couples cosine mode m with m +- 2 ('bottom': m +- 1): one
  // pentadiagonal system per (kx, kz) and mode family. Every complex number of the reference is (1 + i) x, so one
  // real tensor stands for a_*_re and a_*_im. A row is assembled from the band description below instead of the
  // reference's five separate triple loops; each entry is the reference's expression, evaluated in its order.
  double km(int i, int j, int k) const { return trans[0][i] * ky[j].real() * trans[2][k]; }  // get_km, :893-903
  static double sq(double x) { return x * x; }
  void stretching_matrix(const Mesh& mesh, const DevDirps& xd, const DevDirps& yd, const DevDirps& zd) {
    const Tdsops* ip[3] = {&xd.interpl_v2p.t, &yd.interpl_v2p.t, &zd.interpl_v2p.t};
    const std::vector<cplx>* es[3] = {&exs, &eys, &ezs};
    const int ns[3] = {nx_spec, ny_spec, nz_spec};
    for (int q = 0; q < 3; ++q) {
      trans[q].assign(ns[q] + 1, 0.0);
      const Tdsops& t = *ip[q];
      for (int m = 1; m <= ns[q]; ++m) {
        const double temp = (*es[q])[m].real() * mesh.d[q];
        trans[q][m] = 2 * (t.a * std::cos(temp * 0.5) + t.b * std::cos(temp * 1.5) + t.c * std::cos(temp * 2.5) +
                           t.d * std::cos(temp * 3.5)) / (1.0 + 2 * t.alpha * std::cos(temp));
      }
    }
    const std::string& st = mesh.stretching[1];
    const double a0 = (mesh.alpha[1] / pi + 1.0 / (2 * pi * mesh.beta[1])) * mesh.L[1];
    const double a1 = (st == "centred" ? 1.0 : (st == "top-bottom" || st == "bottom" ? -1.0 : 0.0)) / (4 * pi * mesh.beta[1]) * mesh.L[1];
    const bool bottom = st == "bottom";
    stretched = bottom ? 2 : 1;
    const int n = penta_rows = bottom ? ny_spec : ny_spec / 2;
    const int hop = bottom ? 1 : 2;  // distance in y between neighbouring rows of one family
    a_odd.assign((size_t)nx_spec * n * nz_spec * 5, 0.0);
    if (!bottom) a_even.assign(a_odd.size(), 0.0);
    auto at = [&](std::vector<double>& a, int i, int j, int k, int d) -> double& {
      return a[(size_t)(i - 1) + (size_t)nx_spec * ((j - 1) + (size_t)n * ((k - 1) + (size_t)nz_spec * (d - 1)))];
    };
    for (int fam = 0; fam < (bottom ? 1 : 2); ++fam) {
      std::vector<double>& a = fam == 0 ? a_odd : a_even;
      const bool even = fam == 1;
      for (int k = 1; k <= nz_spec; ++k)
        for (int j = 1; j <= n; ++j)
          for (int i = 1; i <= nx_spec; ++i) {
            const int y = bottom ? j : 2 * j - 1 + fam;  // the y mode of row j
            // main diagonal: the coefficients of km(y)^2 and of km(y) (km(y - hop) + km(y + hop)); the first and last
            // rows of a family see their neighbour once (mirror) and the even family has modified end rows
            double c1 = a0 * a0, c2 = a1 * a1, nb;
            if (j == 1) { nb = km(i, y + hop, k); if (even) c1 = a0 * a0 - a1 * a1; }
            else if (j == n) { nb = km(i, y - hop, k); if (even) c1 = (a0 + a1) * (a0 + a1); }
            else nb = km(i, y - hop, k) + km(i, y + hop, k);
            at(a, i, j, k, 3) = -sq(kx[i].real() * trans[1][y] * trans[2][k]) - sq(kz[k].real() * trans[1][y] * trans[0][i]) -
                                c1 * sq(km(i, y, k)) - c2 * km(i, y, k) * nb;
            if (bottom) {  // :368-420
              if (j + 1 <= n) at(a, i, j, k, 4) = a0 * a1 * km(i, y + 1, k) * (km(i, y, k) + km(i, y + 1, k));
              if (j + 2 <= n) at(a, i, j, k, 5) = -a1 * a1 * km(i, y + 1, k) * km(i, y + 2, k);
              if (j >= 2) at(a, i, j, k, 2) = a0 * a1 * km(i, y - 1, k) * (km(i, y, k) + km(i, y - 1, k));
              if (j >= 3) at(a, i, j, k, 1) = -a1 * a1 * km(i, y - 1, k) * km(i, y - 2, k);
              continue;
            }
            if (j + 1 <= n) {  // first super-diagonal (:500-538); entries of the last row are never used
              double u1 = a0 * a1, u2 = a0 * a1;
              if (j == 1 && !even) { u1 = 2 * a0 * a1; u2 = 2 * a0 * a1; }
              if (j == 1 && even) { u1 = a0 * a1 - a1 * a1; u2 = a0 * a1; }
              if (j == n - 1 && j != 1 && even) { u1 = a0 * a1; u2 = (a0 + a1) * a1; }
              at(a, i, j, k, 4) = u1 * (km(i, y, k) * km(i, y + 2, k)) + u2 * sq(km(i, y + 2, k));
            }
            if (j + 2 <= n) {  // second super-diagonal (:541-567)
              const double w = (j == 1 && !even) ? 2 * a1 * a1 : a1 * a1;
              at(a, i, j, k, 5) = -(w * km(i, y + 2, k) * km(i, y + 4, k));
            }
            if (j >= 2) {      // first sub-diagonal (:570-608)
              double l1 = a0 * a1, l2 = a0 * a1;
              if (j == 2 && even) { l1 = a0 * a1; l2 = (a0 + a1) * a1; }
              else if (j == n && even) { l1 = (a0 + a1) * a1; l2 = a0 * a1; }
              at(a, i, j, k, 2) = l1 * (km(i, y, k) * km(i, y - 2, k)) + l2 * sq(km(i, y - 2, k));
            }
            if (j >= 3)        // second sub-diagonal (:611-631)
              at(a, i, j, k, 1) = -(a1 * a1 * km(i, y - 2, k) * km(i, y - 4, k));
          }
    }
    // the mean mode: the operator is singular there, the first row becomes the identity
    if (bottom) {  // :418-422
      at(a_odd, 1, 1, 1, 3) = 1.0; at(a_odd, 1, 1, 1, 4) = 0.0; at(a_odd, 1, 1, 1, 5) = 0.0;
    } else {       // :633-648
      for (int k = 1; k <= nz_spec; ++k)
        for (int i = 1; i <= nx_spec; ++i)
          if (k2x[i + sp_st[0]].real() < 1e-15 && k2z[k + sp_st[2]].real() < 1e-15) {
            at(a_odd, i, 1, k, 3) = 1.0; at(a_odd, i, 1, k, 4) = 0.0; at(a_odd, i, 1, k, 5) = 0.0;
          }
    }
  }
};

// ------------------------------------------------------------------------------------ solver + time integrator + case
class Sim {
 public:
  Config cfg;
  Mesh mesh;
  x3d2c_ctx* ctx = nullptr;
  Allocator allocator;
  Backend backend;
  DevDirps xdirps, ydirps, zdirps;
  PoissonFFT pfft;
  double nu = 0, dt = 0;
  long long ngrid = 0;
  Field *u = nullptr, *v = nullptr, *w = nullptr;
  // time_intg_t (time_integrator.f90:11-26)
  int ti_istep = 1, ti_istage = 1, ti_order = 3, ti_nstep = 1, ti_nstage = 3, ti_nolds = 3;
  bool ti_is_ab = false;
  double ti_coeffs[5][5], ti_rk_b[5][5], ti_rk_a[4][4][5];
  std::vector<std::vector<Field*>> olds;

  explicit Sim(const Config& c) : cfg(c) {
    mesh.init(cfg);
    x3d2c_config bc;
    std::memset(&bc, 0, sizeof bc);
    for (int q = 0; q < 3; ++q) {
      bc.dims_vert[q] = mesh.vert_dims[q]; bc.dims_cell[q] = mesh.cell_dims[q];
      bc.dims_vert_global[q] = mesh.global_vert_dims[q]; bc.dims_cell_global[q] = mesh.global_cell_dims[q];
      bc.nproc_dir[q] = mesh.nproc_dir[q]; bc.nrank_dir[q] = mesh.nrank_dir[q]; bc.n_offset[q] = mesh.n_offset[q];
      bc.pprev[q] = mesh.pprev[q]; bc.pnext[q] = mesh.pnext[q]; bc.periodic[q] = mesh.periodic_BC[q];
    }
    bc.sz = SZ; bc.rank = mesh.nrank; bc.nproc = mesh.nproc; bc.device = cfg.device; bc.flags = cfg.flags & 0xff;
    bc.nccl_unique_id = cfg.nccl_unique_id;
    X3D2H_CALL(x3d2c_create(&bc, &ctx));
    try {
      construct();
    } catch (...) {  // a constructor that throws does not run the destructor: release the context, fields and handles
      release();
      throw;
    }
  }
  void construct() {
    allocator.init(ctx);
    backend.ctx = ctx; backend.mesh = &mesh; backend.allocator = &allocator;
    // solver init (solver.f90:111-212)
    u = allocator.get_block(DIR_X); v = allocator.get_block(DIR_X); w = allocator.get_block(DIR_X);
    init_time_integrator();
    dt = cfg.dt;
    nu = 1.0 / cfg.Re;
    ngrid = (long long)mesh.global_vert_dims[0] * mesh.global_vert_dims[1] * mesh.global_vert_dims[2];
    xdirps.dir = DIR_X; ydirps.dir = DIR_Y; zdirps.dir = DIR_Z;
    allocate_tdsops(xdirps); allocate_tdsops(ydirps); allocate_tdsops(zdirps);
    init_poisson_fft();
  }
  ~Sim() { release(); }
  void release() {
    if (!ctx) return;
    x3d2c_sync(ctx);
    if (backend.poisson) x3d2c_poisson_destroy(ctx, backend.poisson);
    for (DevDirps* d : {&xdirps, &ydirps, &zdirps})
      for (DevTdsops* o : {&d->der1st, &d->der1st_sym, &d->der2nd, &d->der2nd_sym, &d->stagder_v2p, &d->stagder_p2v,
                           &d->interpl_v2p, &d->interpl_p2v})
        if (o->h) x3d2c_tdsops_destroy(ctx, o->h);
    allocator.destroy();
    x3d2c_destroy(ctx);
    ctx = nullptr;
  }

  // solver.f90:214-289
  void allocate_tdsops(DevDirps& dp) {
    const int dir = dp.dir;
    const double d = mesh.d[dir - 1];
    const int bc_start = mesh.BCs[dir - 1][0], bc_end = mesh.BCs[dir - 1][1];
    const int bc_mp_start = bc_start == BC_DIRICHLET ? BC_NEUMANN : bc_start;
    const int bc_mp_end = bc_end == BC_DIRICHLET ? BC_NEUMANN : bc_end;
    const int n_vert = mesh.get_n(dir, VERT), n_cell = mesh.get_n(dir, CELL);
    const double* vds = mesh.vert_ds[dir - 1].data();
    const double* vds2 = mesh.vert_ds2[dir - 1].data();
    const double* vd2s = mesh.vert_d2s[dir - 1].data();
    const double* mds = mesh.midp_ds[dir - 1].data();
    backend.alloc_tdsops(dp.der1st, n_vert, d, "first-deriv", cfg.der1st, bc_start, bc_end, vds);
    backend.alloc_tdsops(dp.der1st_sym, n_vert, d, "first-deriv", cfg.der1st, bc_start, bc_end, vds, nullptr, 4, "", true);
    backend.alloc_tdsops(dp.der2nd, n_vert, d, "second-deriv", cfg.der2nd, bc_start, bc_end, vds2, vd2s);
    backend.alloc_tdsops(dp.der2nd_sym, n_vert, d, "second-deriv", cfg.der2nd, bc_start, bc_end, vds2, vd2s, 4, "", true);
    backend.alloc_tdsops(dp.stagder_v2p, n_cell, d, "stag-deriv", cfg.stagder, bc_mp_start, bc_mp_end, mds, nullptr, 4, "v2p");
    backend.alloc_tdsops(dp.stagder_p2v, n_vert, d, "stag-deriv", cfg.stagder, bc_mp_start, bc_mp_end, vds, nullptr, 4, "p2v");
    backend.alloc_tdsops(dp.interpl_v2p, n_cell, d, "interpolate", cfg.interpl, bc_mp_start, bc_mp_end, nullptr, nullptr, 4, "v2p");
    backend.alloc_tdsops(dp.interpl_p2v, n_vert, d, "interpolate", cfg.interpl, bc_mp_start, bc_mp_end, nullptr, nullptr, 4, "p2v");
  }

  // init_poisson_fft (cuda/poisson_fft.f90:97-260 pattern): layout from the backend, base_init on the host, upload
  void init_poisson_fft() {
    const bool p000 = mesh.periodic_BC[0] && mesh.periodic_BC[1] && mesh.periodic_BC[2];
    const bool p010 = mesh.periodic_BC[0] && !mesh.periodic_BC[1] && mesh.periodic_BC[2];
    // other cases (and stretching in x or z, poisson_fft.f90:166-169): the solver has no FFT Poisson, poisson_fft() fails
    if ((!p000 && !(p010 && mesh.nproc == 1)) || mesh.stretched[0] || mesh.stretched[2]) return;
    if (p010 && (mesh.global_cell_dims[1] % 2 || mesh.global_cell_dims[1] < 8)) return;  // the paired-mode step needs an even count
    int n_spec[3], n_sp_st[3];
    X3D2H_CALL(x3d2c_poisson_spec_layout(ctx, n_spec, n_sp_st));
    pfft.base_init(mesh, xdirps, ydirps, zdirps, n_spec, n_sp_st);
    const double* w = reinterpret_cast<const double*>(pfft.waves.data());
    if (p000) {
      X3D2H_CALL(x3d2c_poisson_create(ctx, w, &pfft.ax[1], &pfft.bx[1], &pfft.ay[1], &pfft.by[1], &pfft.az[1], &pfft.bz[1],
                                      &backend.poisson));
    } else {  // a_*_re and a_*_im hold the same numbers (see stretching_matrix)
      const double* ao = pfft.stretched ? pfft.a_odd.data() : nullptr;
      const double* ae = pfft.stretched == 1 ? pfft.a_even.data() : nullptr;
      X3D2H_CALL(x3d2c_poisson_create_010(ctx, w, &pfft.ax[1], &pfft.bx[1], &pfft.ay[1], &pfft.by[1], &pfft.az[1],
                                          &pfft.bz[1], pfft.stretched, ao, ao, ae, ae, &backend.poisson));
    }
  }

  // ---- base-ops mode (X3D2H_FLAG_BASE_OPS): the operator graph of the UNCHANGED reference solver, issued call by
  // call through the deferred procedures of base_backend_t only (no fused extension entry point). This is what a
  // Fortran cuda_c_backend_t sees from solver.f90 / vector_calculus.f90 / time_integrator.f90 as they are.
  bool base_ops() const { return (cfg.flags & X3D2H_FLAG_BASE_OPS) != 0; }

  // solver.f90:291-389, statement by statement
  void transeq_default_base(Field& du, Field& dv, Field& dw, const Field& uu, const Field& vv, const Field& ww) {
    Allocator& A = allocator;
    backend.transeq_x(du, dv, dw, uu, vv, ww, nu, xdirps);
    Field* in[3] = {A.get_block(DIR_Y), A.get_block(DIR_Y), A.get_block(DIR_Y)};
    Field* out[3] = {A.get_block(DIR_Y), A.get_block(DIR_Y), A.get_block(DIR_Y)};
    const Field* vel[3] = {&uu, &vv, &ww};
    Field* rhs[3] = {&du, &dv, &dw};
    for (int i = 0; i < 3; ++i) backend.reorder(*in[i], *vel[i], RDR_X2Y);
    backend.transeq_y(*out[0], *out[1], *out[2], *in[0], *in[1], *in[2], nu, ydirps);
    for (int i = 0; i < 3; ++i) A.release_block(in[i]);
    for (int i = 0; i < 3; ++i) backend.sum_yintox(*rhs[i], *out[i]);
    for (int i = 0; i < 3; ++i) A.release_block(out[i]);
    for (int i = 0; i < 3; ++i) in[i] = A.get_block(DIR_Z);
    for (int i = 0; i < 3; ++i) out[i] = A.get_block(DIR_Z);
    for (int i = 0; i < 3; ++i) backend.reorder(*in[i], *vel[i], RDR_X2Z);
    backend.transeq_z(*out[0], *out[1], *out[2], *in[0], *in[1], *in[2], nu, zdirps);
    for (int i = 0; i < 3; ++i) A.release_block(in[i]);
    for (int i = 0; i < 3; ++i) backend.sum_zintox(*rhs[i], *out[i]);
    for (int i = 0; i < 3; ++i) A.release_block(out[i]);
  }

  // vector_calculus.f90:142-246: 8 tds_solve, 5 reorder, 2 vecadd
  void divergence_v2c_base(Field& div_u, const Field& uu, const Field& vv, const Field& ww) {
    if (div_u.dir != DIR_Z || uu.dir != DIR_X || vv.dir != DIR_X || ww.dir != DIR_X)
      fail("Error in divergence_v2c input/output field dirs: output must be in DIR_Z, inputs must be in DIR_X layout.");
    Allocator& A = allocator;
    Field* x[3] = {A.get_block(DIR_X), A.get_block(DIR_X), A.get_block(DIR_X)};
    backend.tds_solve(*x[0], uu, xdirps.stagder_v2p);
    backend.tds_solve(*x[1], vv, xdirps.interpl_v2p);
    backend.tds_solve(*x[2], ww, xdirps.interpl_v2p);
    Field* y[3] = {A.get_block(DIR_Y), A.get_block(DIR_Y), A.get_block(DIR_Y)};
    for (int i = 0; i < 3; ++i) backend.reorder(*y[i], *x[i], RDR_X2Y);
    for (int i = 0; i < 3; ++i) A.release_block(x[i]);
    Field* dy[3] = {A.get_block(DIR_Y), A.get_block(DIR_Y), A.get_block(DIR_Y)};
    backend.tds_solve(*dy[0], *y[0], ydirps.interpl_v2p);
    backend.tds_solve(*dy[1], *y[1], ydirps.stagder_v2p);
    backend.tds_solve(*dy[2], *y[2], ydirps.interpl_v2p);
    for (int i = 0; i < 3; ++i) A.release_block(y[i]);
    Field *u_z = A.get_block(DIR_Z), *w_z = A.get_block(DIR_Z);
    backend.vecadd(1.0, *dy[1], 1.0, *dy[0]);
    backend.reorder(*u_z, *dy[0], RDR_Y2Z);
    backend.reorder(*w_z, *dy[2], RDR_Y2Z);
    for (int i = 0; i < 3; ++i) A.release_block(dy[i]);
    Field* dw_z = A.get_block(DIR_Z);
    backend.tds_solve(div_u, *u_z, zdirps.interpl_v2p);
    backend.tds_solve(*dw_z, *w_z, zdirps.stagder_v2p);
    backend.vecadd(1.0, *dw_z, 1.0, div_u);
    A.release_block(u_z); A.release_block(w_z); A.release_block(dw_z);
  }

  // vector_calculus.f90:248-332: 8 tds_solve, 5 reorder
  void gradient_c2v_base(Field& dpdx, Field& dpdy, Field& dpdz, const Field& p) {
    if (dpdx.dir != DIR_X || dpdy.dir != DIR_X || dpdz.dir != DIR_X || p.dir != DIR_Z)
      fail("Error in gradient_c2v input/output field dirs: outputs must be in DIR_X, input must be in DIR_Z layout.");
    Allocator& A = allocator;
    Field *p_z = A.get_block(DIR_Z), *dz_z = A.get_block(DIR_Z);
    backend.tds_solve(*p_z, p, zdirps.interpl_p2v);
    backend.tds_solve(*dz_z, p, zdirps.stagder_p2v);
    Field *p_y = A.get_block(DIR_Y), *dz_y = A.get_block(DIR_Y);
    backend.reorder(*p_y, *p_z, RDR_Z2Y);
    backend.reorder(*dz_y, *dz_z, RDR_Z2Y);
    A.release_block(p_z); A.release_block(dz_z);
    Field *q_y = A.get_block(DIR_Y), *dy_y = A.get_block(DIR_Y);
    backend.tds_solve(*q_y, *p_y, ydirps.interpl_p2v);
    backend.tds_solve(*dy_y, *p_y, ydirps.stagder_p2v);
    A.release_block(p_y);
    Field* dz2_y = A.get_block(DIR_Y);
    backend.tds_solve(*dz2_y, *dz_y, ydirps.interpl_p2v);
    A.release_block(dz_y);
    Field* q_x = A.get_block(DIR_X);
    backend.reorder(*q_x, *q_y, RDR_Y2X); A.release_block(q_y);
    Field* dy_x = A.get_block(DIR_X);
    backend.reorder(*dy_x, *dy_y, RDR_Y2X); A.release_block(dy_y);
    Field* dz_x = A.get_block(DIR_X);
    backend.reorder(*dz_x, *dz2_y, RDR_Y2X); A.release_block(dz2_y);
    backend.tds_solve(dpdx, *q_x, xdirps.stagder_p2v);
    backend.tds_solve(dpdy, *dy_x, xdirps.interpl_p2v);
    backend.tds_solve(dpdz, *dz_x, xdirps.interpl_p2v);
    A.release_block(q_x); A.release_block(dy_x); A.release_block(dz_x);
  }

  // solver.f90:693-739
  void pressure_correction_base(Field& uu, Field& vv, Field& ww) {
    Allocator& A = allocator;
    Field* div_u = A.get_block(DIR_Z);
    divergence_v2c_base(*div_u, uu, vv, ww);
    Field* p = A.get_block(DIR_Z);
    poisson_fft(*p, *div_u);  // reorder Z2C, solve_poisson, reorder C2Z (solver.f90:653-678)
    A.release_block(div_u);
    Field *dpdx = A.get_block(DIR_X), *dpdy = A.get_block(DIR_X), *dpdz = A.get_block(DIR_X);
    gradient_c2v_base(*dpdx, *dpdy, *dpdz, *p);
    A.release_block(p);
    backend.vecadd(-1.0, *dpdx, 1.0, uu);
    backend.vecadd(-1.0, *dpdy, 1.0, vv);
    backend.vecadd(-1.0, *dpdz, 1.0, ww);
    A.release_block(dpdx); A.release_block(dpdy); A.release_block(dpdz);
  }

  // time_integrator.f90:166-231 with veccopy / vecadd only
  void runge_kutta_base(Field* curr[3], Field* deriv[3], double dt_) {
    const int S = ti_nstage, s = ti_istage;
    for (int i = 0; i < 3; ++i) {
      if (s == S) {
        if (S > 1) backend.veccopy(*curr[i], *olds[i][1]);
        for (int j = 1; j <= S - 1; ++j) backend.vecadd(ti_rk_b[j][S] * dt_, *olds[i][j + 1], 1.0, *curr[i]);
        backend.vecadd(ti_rk_b[S][S] * dt_, *deriv[i], 1.0, *curr[i]);
      } else {
        if (s == 1) backend.veccopy(*olds[i][1], *curr[i]);
        backend.veccopy(*olds[i][s + 1], *deriv[i]);
        if (s > 1) backend.veccopy(*curr[i], *olds[i][1]);
        for (int j = 1; j <= s; ++j) backend.vecadd(ti_rk_a[j][s][S] * dt_, *olds[i][j + 1], 1.0, *curr[i]);
      }
    }
    ti_istage = (s == S) ? 1 : s + 1;
  }
  // time_integrator.f90:233-300 with veccopy / vecadd only
  void adams_bashforth_base(Field* curr[3], Field* deriv[3], double dt_) {
    const int nstep = std::min(ti_istep, ti_nstep);
    for (int i = 0; i < 3; ++i) {
      backend.vecadd(ti_coeffs[1][nstep] * dt_, *deriv[i], 1.0, *curr[i]);
      for (int j = 2; j <= nstep; ++j) backend.vecadd(ti_coeffs[j][nstep] * dt_, *olds[i][j - 1], 1.0, *curr[i]);
      const int nrot = nstep < ti_nstep ? (ti_istep > 1 ? nstep : 0) : (ti_nstep > 2 ? nstep - 1 : 0);
      if (nrot) {
        Field* last = olds[i][nrot];
        for (int q = nrot; q >= 2; --q) olds[i][q] = olds[i][q - 1];
        olds[i][1] = last;
      }
      if (ti_nstep > 1) backend.veccopy(*olds[i][1], *deriv[i]);
    }
    ti_istep = ti_istep + 1;
  }

  // solver.f90:391-505: the low-memory variant. The velocity blocks are given back to the pool once they exist in the
  // y layout and return from the z layout (reorder Z2X) in fresh blocks: uu, vv, ww are re-pointed.
  void transeq_lowmem(Field& du, Field& dv, Field& dw, Field*& uu, Field*& vv, Field*& ww) {
    Allocator& A = allocator;
    backend.transeq_x(du, dv, dw, *uu, *vv, *ww, nu, xdirps);
    Field* vel[3] = {uu, vv, ww};
    Field* rhs[3] = {&du, &dv, &dw};
    Field* y[3];
    for (int i = 0; i < 3; ++i) y[i] = A.get_block(DIR_Y);
    for (int i = 0; i < 3; ++i) backend.reorder(*y[i], *vel[i], RDR_X2Y);
    for (int i = 0; i < 3; ++i) A.release_block(vel[i]);
    Field* d[3];
    for (int i = 0; i < 3; ++i) d[i] = A.get_block(DIR_Y);
    backend.transeq_y(*d[0], *d[1], *d[2], *y[0], *y[1], *y[2], nu, ydirps);
    for (int i = 0; i < 3; ++i) backend.sum_yintox(*rhs[i], *d[i]);
    for (int i = 0; i < 3; ++i) A.release_block(d[i]);
    Field* z[3];
    for (int i = 0; i < 3; ++i) z[i] = A.get_block(DIR_Z);
    for (int i = 0; i < 3; ++i) backend.reorder(*z[i], *y[i], RDR_Y2Z);
    for (int i = 0; i < 3; ++i) A.release_block(y[i]);
    for (int i = 0; i < 3; ++i) d[i] = A.get_block(DIR_Z);
    backend.transeq_z(*d[0], *d[1], *d[2], *z[0], *z[1], *z[2], nu, zdirps);
    for (int i = 0; i < 3; ++i) backend.sum_zintox(*rhs[i], *d[i]);
    for (int i = 0; i < 3; ++i) A.release_block(d[i]);
    for (int i = 0; i < 3; ++i) vel[i] = A.get_block(DIR_X);
    for (int i = 0; i < 3; ++i) backend.reorder(*vel[i], *z[i], RDR_Z2X);
    for (int i = 0; i < 3; ++i) A.release_block(z[i]);
    uu = vel[0]; vv = vel[1]; ww = vel[2];
  }

  // solver.f90:507-600: convection-diffusion of the species fields spec[0..n) with diffusivities nu_s[i]
  void transeq_species(Field* const* rhs, int n, const Field& uu, const Field& vv, const Field& ww, Field* const* spec,
                       const double* nu_s) {
    Allocator& A = allocator;
    for (int i = 0; i < n; ++i) backend.transeq_species(*rhs[i], uu, *spec[i], nu_s[i], xdirps, i <= 0);
    {
      Field *v_y = A.get_block(DIR_Y), *spec_y = A.get_block(DIR_Y), *dspec_y = A.get_block(DIR_Y);
      backend.reorder(*v_y, vv, RDR_X2Y);
      for (int i = 0; i < n; ++i) {
        backend.reorder(*spec_y, *spec[i], RDR_X2Y);
        backend.transeq_species(*dspec_y, *v_y, *spec_y, nu_s[i], ydirps, i <= 0);
        backend.sum_yintox(*rhs[i], *dspec_y);
      }
      A.release_block(v_y); A.release_block(spec_y); A.release_block(dspec_y);
    }
    {
      Field *w_z = A.get_block(DIR_Z), *spec_z = A.get_block(DIR_Z), *dspec_z = A.get_block(DIR_Z);
      backend.reorder(*w_z, ww, RDR_X2Z);
      for (int i = 0; i < n; ++i) {
        backend.reorder(*spec_z, *spec[i], RDR_X2Z);
        backend.transeq_species(*dspec_z, *w_z, *spec_z, nu_s[i], zdirps, i <= 0);
        backend.sum_zintox(*rhs[i], *dspec_z);
      }
      A.release_block(w_z); A.release_block(spec_z); A.release_block(dspec_z);
    }
  }

  // solver.f90:291-389
  void transeq_default(Field& du, Field& dv, Field& dw, const Field& uu, const Field& vv, const Field& ww) {
    if (base_ops()) return transeq_default_base(du, dv, dw, uu, vv, ww);
    Field *dy[3], *dz[3];
    transeq_parts(du, dv, dw, dy, dz, uu, vv, ww);
    // one pass adds the y and the z contributions (sum_yintox then sum_zintox, in the reference's order)
    backend.sum_yzintox(du, *dy[0], *dz[0]); backend.sum_yzintox(dv, *dy[1], *dz[1]); backend.sum_yzintox(dw, *dy[2], *dz[2]);
    for (int i = 0; i < 3; ++i) { allocator.release_block(dy[i]); allocator.release_block(dz[i]); }
  }
  // transeq_default without its final sums: (du, dv, dw) hold the x contribution, dy / dz (blocks of the allocator, to be
  // released by the caller) the y and z contributions in their own layouts
  void transeq_parts(Field& du, Field& dv, Field& dw, Field* dy[3], Field* dz[3], const Field& uu, const Field& vv,
                     const Field& ww) {
    Allocator& A = allocator;
    backend.transeq_x(du, dv, dw, uu, vv, ww, nu, xdirps);
    Field *u_z = A.get_block(DIR_Z), *v_z = A.get_block(DIR_Z), *w_z = A.get_block(DIR_Z);
    for (int i = 0; i < 3; ++i) dy[i] = A.get_block(DIR_Y);
    if (backend.transeq_r_fused(DIR_Y, ydirps, RDR_X2Y)) {
      // the y sweep reads the x-layout velocity itself (swizzled tiles): only the z copies are made
      backend.transeq_r(DIR_Y, *dy[0], *dy[1], *dy[2], uu, vv, ww, nu, ydirps, RDR_X2Y);
      backend.reorder(*u_z, uu, RDR_X2Z); backend.reorder(*v_z, vv, RDR_X2Z); backend.reorder(*w_z, ww, RDR_X2Z);
    } else {
      // u, v, w into both pencil layouts with one read each (reference: 3 x reorder X2Y here, 3 x reorder X2Z below)
      Field *u_y = A.get_block(DIR_Y), *v_y = A.get_block(DIR_Y), *w_y = A.get_block(DIR_Y);
      backend.reorder_x2yz(*u_y, *u_z, uu); backend.reorder_x2yz(*v_y, *v_z, vv); backend.reorder_x2yz(*w_y, *w_z, ww);
      backend.transeq_y(*dy[0], *dy[1], *dy[2], *u_y, *v_y, *w_y, nu, ydirps);
      A.release_block(u_y); A.release_block(v_y); A.release_block(w_y);
    }
    for (int i = 0; i < 3; ++i) dz[i] = A.get_block(DIR_Z);
    backend.transeq_z(*dz[0], *dz[1], *dz[2], *u_z, *v_z, *w_z, nu, zdirps);
    A.release_block(u_z); A.release_block(v_z); A.release_block(w_z);
  }

  // vector_calculus.f90:142-246
  void divergence_v2c(Field& div_u, const Field& uu, const Field& vv, const Field& ww) {
    if (base_ops()) return divergence_v2c_base(div_u, uu, vv, ww);
    // div_u in DIR_Z as in the reference, or in DIR_C (then the z2c reorder of poisson_fft is part of the last solve)
    if ((div_u.dir != DIR_Z && div_u.dir != DIR_C) || uu.dir != DIR_X || vv.dir != DIR_X || ww.dir != DIR_X)
      fail("Error in divergence_v2c input/output field dirs: output must be in DIR_Z, inputs must be in DIR_X layout.");
    Allocator& A = allocator;
    // (u_y, v_y, w_y) = x2y(stagder(u), interpl(v), interpl(w))   (:160-183: tds_solve x3, reorder x3; the reorder is
    // part of the solve's store)
    Field *u_y = A.get_block(DIR_Y), *v_y = A.get_block(DIR_Y), *w_y = A.get_block(DIR_Y);
    backend.tds_solve_r(DIR_X, *u_y, uu, xdirps.stagder_v2p, 0, RDR_X2Y);
    backend.tds_solve_r(DIR_X, *v_y, vv, xdirps.interpl_v2p, 0, RDR_X2Y);
    backend.tds_solve_r(DIR_X, *w_y, ww, xdirps.interpl_v2p, 0, RDR_X2Y);
    // u_z = y2z(interpl(u_y) + stagder(v_y)); w_z = y2z(interpl(w_y))   (:185-203: tds_solve x3, vecadd, reorder x2)
    Field *u_z = A.get_block(DIR_Z), *w_z = A.get_block(DIR_Z);
    backend.tds_solve_sum_r(DIR_Y, *u_z, *u_y, ydirps.interpl_v2p, *v_y, ydirps.stagder_v2p, 0, RDR_Y2Z);
    backend.tds_solve_r(DIR_Y, *w_z, *w_y, ydirps.interpl_v2p, 0, RDR_Y2Z);
    A.release_block(u_y); A.release_block(v_y); A.release_block(w_y);
    // div = interpl(u_z) + stagder(w_z)   (:205-214)
    backend.tds_solve_sum_r(DIR_Z, div_u, *u_z, zdirps.interpl_v2p, *w_z, zdirps.stagder_v2p, 0,
                            div_u.dir == DIR_C ? RDR_Z2C : 0);
    A.release_block(u_z); A.release_block(w_z);
  }

  // vector_calculus.f90:248-332. With sub != nullptr the x stage subtracts the gradient from (sub[0], sub[1],
  // sub[2]) instead of writing it (the vecadd(-1, dpdx, 1, u) calls of solver.f90:296-298 fused into the last solve).
  void gradient_c2v(Field& dpdx, Field& dpdy, Field& dpdz, const Field& p, Field* const* sub = nullptr) {
    if (base_ops() && !sub) return gradient_c2v_base(dpdx, dpdy, dpdz, p);
    // p in DIR_Z as in the reference, or in DIR_C (then the c2z reorder of poisson_fft is part of the first solve)
    if (dpdx.dir != DIR_X || dpdy.dir != DIR_X || dpdz.dir != DIR_X || (p.dir != DIR_Z && p.dir != DIR_C))
      fail("Error in gradient_c2v input/output field dirs: outputs must be in DIR_X, input must be in DIR_Z layout.");
    Allocator& A = allocator;
    // (p_sxy_y, dpdz_sxy_y) = z2y(interpl(p), stagder(p))   (:275-285: tds_solve x2, reorder x2)
    Field *p_sxy_y = A.get_block(DIR_Y), *dpdz_sxy_y = A.get_block(DIR_Y);
    backend.tds_solve_dual_r(DIR_Z, *p_sxy_y, *dpdz_sxy_y, p, zdirps.interpl_p2v, zdirps.stagder_p2v,
                             p.dir == DIR_C ? RDR_C2Z : 0, RDR_Z2Y);
    Field *p_sx_y = A.get_block(DIR_Y), *dpdy_sx_y = A.get_block(DIR_Y);
    backend.tds_solve_dual(*p_sx_y, *dpdy_sx_y, *p_sxy_y, ydirps.interpl_p2v, ydirps.stagder_p2v);
    A.release_block(p_sxy_y);
    Field* dpdz_sx_y = A.get_block(DIR_Y);
    backend.tds_solve(*dpdz_sx_y, *dpdz_sxy_y, ydirps.interpl_p2v);
    A.release_block(dpdz_sxy_y);
    // y2x reorders (:302-318) are part of the x solves' loads
    if (sub) {
      backend.tds_solve_axpy_r(DIR_X, *sub[0], -1.0, *p_sx_y, xdirps.stagder_p2v, RDR_Y2X);
      backend.tds_solve_axpy_r(DIR_X, *sub[1], -1.0, *dpdy_sx_y, xdirps.interpl_p2v, RDR_Y2X);
      backend.tds_solve_axpy_r(DIR_X, *sub[2], -1.0, *dpdz_sx_y, xdirps.interpl_p2v, RDR_Y2X);
    } else {
      backend.tds_solve_r(DIR_X, dpdx, *p_sx_y, xdirps.stagder_p2v, RDR_Y2X, 0);
      backend.tds_solve_r(DIR_X, dpdy, *dpdy_sx_y, xdirps.interpl_p2v, RDR_Y2X, 0);
      backend.tds_solve_r(DIR_X, dpdz, *dpdz_sx_y, xdirps.interpl_p2v, RDR_Y2X, 0);
    }
    A.release_block(p_sx_y); A.release_block(dpdy_sx_y); A.release_block(dpdz_sx_y);
  }

  // vector_calculus.f90:334-378 with the operators of postprocess.f90:184-189: cell centres (DIR_Z, CELL) -> vertices
  // (DIR_X, VERT). The z2y and y2x reorders are part of the neighbouring solves (tds_solve_r).
  void interpl_c2v(Field& p_out, const Field& p) {
    if (p_out.dir != DIR_X || p.dir != DIR_Z)
      fail("Error in interpl_c2v input/output field dirs: output must be in DIR_X, input must be in DIR_Z layout.");
    Allocator& A = allocator;
    Field* p_sy_y = A.get_block(DIR_Y);
    backend.tds_solve_r(DIR_Z, *p_sy_y, p, zdirps.interpl_p2v, 0, RDR_Z2Y);
    Field* p_out_y = A.get_block(DIR_Y);
    backend.tds_solve(*p_out_y, *p_sy_y, ydirps.interpl_p2v);
    A.release_block(p_sy_y);
    backend.tds_solve_r(DIR_X, p_out, *p_out_y, xdirps.interpl_p2v, RDR_Y2X, 0);
    A.release_block(p_out_y);
  }
  // postprocess.f90:166-197: pressure at the vertices, rescaled from the pseudo-pressure p'/dt... (vecadd(1/dt, p, 0, p))
  void pressure_vert(Field& p_out, const Field& p) {
    interpl_c2v(p_out, p);
    backend.vecadd(1.0 / dt, p_out, 0.0, p_out);
  }
  // vector_calculus.f90:380-437 with der2nd in x, y, z: DIR_X in, DIR_X out, evaluated at the input's data_loc
  void laplacian(Field& lapl_u, const Field& uu) {
    if (uu.dir != DIR_X || lapl_u.dir != DIR_X)
      fail("Error in laplacian input/output field dirs: outputs and inputs must be in DIR_X layout.");
    Allocator& A = allocator;
    backend.tds_solve(lapl_u, uu, xdirps.der2nd);
    Field *u_y = A.get_block(DIR_Y), *u_z = A.get_block(DIR_Z);
    backend.reorder_x2yz(*u_y, *u_z, uu);
    Field *d2u_y = A.get_block(DIR_Y), *d2u_z = A.get_block(DIR_Z);
    backend.tds_solve(*d2u_y, *u_y, ydirps.der2nd);
    backend.tds_solve(*d2u_z, *u_z, zdirps.der2nd);
    backend.sum_yzintox(lapl_u, *d2u_y, *d2u_z);
    A.release_block(u_y); A.release_block(u_z); A.release_block(d2u_y); A.release_block(d2u_z);
  }

  // vector_calculus.f90:40-140
  void curl(Field& o_i, Field& o_j, Field& o_k, const Field& uu, const Field& vv, const Field& ww) {
    Allocator& A = allocator;
    Field *w_y = A.get_block(DIR_Y), *dwdy_y = A.get_block(DIR_Y);
    backend.reorder(*w_y, ww, RDR_X2Y); backend.tds_solve(*dwdy_y, *w_y, ydirps.der1st);
    backend.reorder(o_i, *dwdy_y, RDR_Y2X);
    A.release_block(w_y); A.release_block(dwdy_y);
    Field *v_z = A.get_block(DIR_Z), *dvdz_z = A.get_block(DIR_Z);
    backend.reorder(*v_z, vv, RDR_X2Z); backend.tds_solve(*dvdz_z, *v_z, zdirps.der1st);
    Field* dvdz_x = A.get_block(DIR_X);
    backend.reorder(*dvdz_x, *dvdz_z, RDR_Z2X);
    A.release_block(v_z); A.release_block(dvdz_z);
    backend.vecadd(-1.0, *dvdz_x, 1.0, o_i);
    A.release_block(dvdz_x);
    Field *u_z = A.get_block(DIR_Z), *dudz_z = A.get_block(DIR_Z);
    backend.reorder(*u_z, uu, RDR_X2Z); backend.tds_solve(*dudz_z, *u_z, zdirps.der1st);
    Field* dudz_x = A.get_block(DIR_X);
    backend.reorder(*dudz_x, *dudz_z, RDR_Z2X);
    A.release_block(u_z); A.release_block(dudz_z);
    backend.tds_solve(o_j, ww, xdirps.der1st);
    backend.vecadd(1.0, *dudz_x, -1.0, o_j);
    A.release_block(dudz_x);
    backend.tds_solve(o_k, vv, xdirps.der1st);
    Field *u_y = A.get_block(DIR_Y), *dudy_y = A.get_block(DIR_Y);
    backend.reorder(*u_y, uu, RDR_X2Y); backend.tds_solve(*dudy_y, *u_y, ydirps.der1st);
    Field* dudy_x = A.get_block(DIR_X);
    backend.reorder(*dudy_x, *dudy_y, RDR_Y2X);
    A.release_block(u_y); A.release_block(dudy_y);
    backend.vecadd(-1.0, *dudy_x, 1.0, o_k);
    A.release_block(dudy_x);
  }

  // solver.f90:653-678 + poisson_fft.f90:206-226
  // both fields in DIR_Z as in the reference, or the same DIR_C field (solved in place, no reorders)
  void poisson_fft(Field& pressure, const Field& div_u) {
    if (!backend.poisson) fail("FFT Poisson solver is not initialised for these BCs");
    if (pfft.bc_case == 10) {  // poisson_010 (poisson_fft.f90:228-242) between the reorders of solver.f90:653-678
      if (pressure.dir != DIR_Z || div_u.dir != DIR_Z) fail("poisson_fft: fields must be in DIR_Z layout");
      Field* p_temp = allocator.get_block(DIR_C);
      backend.reorder(*p_temp, div_u, RDR_Z2C);
      Field* temp = allocator.get_block(DIR_C);
      X3D2H_CALL(x3d2c_enforce_periodicity_y(ctx, backend.poisson, temp->dev, p_temp->dev));
      X3D2H_CALL(x3d2c_fft_forward(ctx, backend.poisson, temp->dev));
      X3D2H_CALL(x3d2c_fft_postprocess_010(ctx, backend.poisson));
      X3D2H_CALL(x3d2c_fft_backward(ctx, backend.poisson, temp->dev));
      X3D2H_CALL(x3d2c_undo_periodicity_y(ctx, backend.poisson, p_temp->dev, temp->dev));
      allocator.release_block(temp);
      backend.reorder(pressure, *p_temp, RDR_C2Z);
      allocator.release_block(p_temp);
      return;
    }
    if (pressure.dir == DIR_C && &pressure == &div_u) {
      X3D2H_CALL(x3d2c_fft_forward(ctx, backend.poisson, pressure.dev));
      X3D2H_CALL(x3d2c_fft_postprocess_000(ctx, backend.poisson));
      X3D2H_CALL(x3d2c_fft_backward(ctx, backend.poisson, pressure.dev));
      return;
    }
    Field* p_temp = allocator.get_block(DIR_C);
    backend.reorder(*p_temp, div_u, RDR_Z2C);
    Field* temp = allocator.get_block(DIR_C);
    X3D2H_CALL(x3d2c_fft_forward(ctx, backend.poisson, p_temp->dev));
    X3D2H_CALL(x3d2c_fft_postprocess_000(ctx, backend.poisson));
    X3D2H_CALL(x3d2c_fft_backward(ctx, backend.poisson, p_temp->dev));
    allocator.release_block(temp);
    backend.reorder(pressure, *p_temp, RDR_C2Z);
    allocator.release_block(p_temp);
  }

  // solver.f90:693-739
  void pressure_correction(Field& uu, Field& vv, Field& ww) {
    if (base_ops() || pfft.bc_case == 10) return pressure_correction_base(uu, vv, ww);
    Allocator& A = allocator;
    // the Cartesian field of the FFT is written by the last solve of the divergence and read by the first solve of
    // the gradient: no z2c / c2z passes (same values as the reference's sequence)
    Field* p = A.get_block(DIR_C);
    divergence_v2c(*p, uu, vv, ww);
    poisson_fft(*p, *p);
    Field* vel[3] = {&uu, &vv, &ww};
    gradient_c2v(uu, vv, ww, *p, vel);  // u -= dpdx, v -= dpdy, w -= dpdz
    A.release_block(p);
  }

  // time_integrator.f90:70-164
  void init_time_integrator() {
    std::memset(ti_coeffs, 0, sizeof ti_coeffs);
    std::memset(ti_rk_b, 0, sizeof ti_rk_b);
    std::memset(ti_rk_a, 0, sizeof ti_rk_a);
    ti_rk_b[1][1] = 1.0;
    ti_rk_a[1][1][2] = 0.5; ti_rk_b[2][2] = 1.0;
    ti_rk_a[1][1][3] = 0.5; ti_rk_a[2][2][3] = 3.0 / 4.0;
    ti_rk_b[1][3] = 2.0 / 9.0; ti_rk_b[2][3] = 1.0 / 3.0; ti_rk_b[3][3] = 4.0 / 9.0;
    ti_rk_a[1][1][4] = 0.5; ti_rk_a[2][2][4] = 0.5; ti_rk_a[3][3][4] = 1.0;
    ti_rk_b[1][4] = 1.0 / 6.0; ti_rk_b[2][4] = 1.0 / 3.0; ti_rk_b[3][4] = 1.0 / 3.0; ti_rk_b[4][4] = 1.0 / 6.0;
    ti_coeffs[1][1] = 1.0;
    ti_coeffs[1][2] = 1.5; ti_coeffs[2][2] = -0.5;
    ti_coeffs[1][3] = 23.0 / 12.0; ti_coeffs[2][3] = -4.0 / 3.0; ti_coeffs[3][3] = 5.0 / 12.0;
    ti_coeffs[1][4] = 55.0 / 24.0; ti_coeffs[2][4] = -59.0 / 24.0; ti_coeffs[3][4] = 37.0 / 24.0; ti_coeffs[4][4] = -3.0 / 8.0;
    const std::string& m = cfg.time_intg;
    if (m.size() != 3) fail("Integration method " + m + " is not defined");
    ti_order = m[2] - '0';
    if (ti_order < 1 || ti_order > 4) fail("Integration order >4 is not supported");
    if (m.substr(0, 2) == "AB") { ti_is_ab = true; ti_nstep = ti_order; ti_nstage = 1; ti_nolds = ti_nstep - 1; }
    else if (m.substr(0, 2) == "RK") { ti_is_ab = false; ti_nstep = 1; ti_nstage = ti_order; ti_nolds = ti_nstage; }
    else fail("Integration method " + m + " is not defined");
    ti_istep = 1; ti_istage = 1;
    olds.assign(3, std::vector<Field*>(ti_nolds + 1, nullptr));
    for (int i = 0; i < 3; ++i)
      for (int j = 1; j <= ti_nolds; ++j) olds[i][j] = allocator.get_block(DIR_X);
  }
  // time_integrator.f90:166-231
  // The vecadd / veccopy chains of the reference are executed as one veclincomb per field (same values, same
  // rounding sequence; terms with a zero coefficient are dropped on the fast path, where c*x + y == y), and the
  // copies "olds <- deriv", "olds(1) <- curr" become pointer swaps: deriv[i] / curr[i] are exchanged with the
  // block they would have been copied into (the caller releases whatever deriv[i] points to afterwards).
  using Terms = std::vector<std::pair<double, const Field*>>;
  void lincomb_or_copy(Field& out, const Field& base, Terms terms) {
    if (!(cfg.flags & X3D2C_FLAG_STRICT)) {
      Terms kept;
      for (auto& t : terms) if (t.first != 0.0) kept.push_back(t);
      terms.swap(kept);
    }
    if (terms.empty()) { if (&out != &base) backend.veccopy(out, base); return; }
    backend.veclincomb(out, base, terms);
  }
  // dy / dz != nullptr: deriv[i] still lacks its y and z contributions (transeq_parts); the sums are taken in the same
  // pass as the update (x3d2c_sum_yzintox_lincomb). The derivative is always the last term of the update.
  void update_with_deriv(Field& out, const Field& base, Terms terms, Field& d, Field* dy, Field* dz, bool keep_d) {
    if (!dy) return lincomb_or_copy(out, base, terms);
    const double c_d = terms.back().first;
    terms.pop_back();
    Terms kept;
    for (auto& t : terms) if (t.first != 0.0) kept.push_back(t);
    if (c_d == 0.0 || kept.size() > 3 || &out == &d) {  // not expressible in the fused form
      backend.sum_yzintox(d, *dy, *dz);
      kept.push_back({c_d, &d});
      return lincomb_or_copy(out, base, kept);
    }
    backend.sum_yzintox_lincomb(d, *dy, *dz, keep_d, out, base, kept, c_d);
  }
  void runge_kutta(Field* curr[3], Field* deriv[3], double dt_, Field* const* dy = nullptr, Field* const* dz = nullptr) {
    if (ti_istage == ti_nstage) {
      for (int i = 0; i < 3; ++i) {
        Terms terms;
        for (int j = 1; j <= ti_nstage - 1; ++j) terms.push_back({ti_rk_b[j][ti_nstage] * dt_, olds[i][j + 1]});
        terms.push_back({ti_rk_b[ti_nstage][ti_nstage] * dt_, deriv[i]});
        update_with_deriv(*curr[i], ti_nstage > 1 ? *olds[i][1] : *curr[i], terms, *deriv[i], dy ? dy[i] : nullptr,
                          dz ? dz[i] : nullptr, false);
      }
      ti_istage = 1;
    } else {
      for (int i = 0; i < 3; ++i) {
        if (ti_istage == 1) {  // olds(1) <- curr: keep the block as olds(1), continue in the former olds(1) block
          std::swap(olds[i][1], curr[i]);
          curr[i]->data_loc = olds[i][1]->data_loc;
        }
        std::swap(olds[i][ti_istage + 1], deriv[i]);  // olds(istage + 1) <- deriv
        Terms terms;
        for (int j = 1; j <= ti_istage; ++j) terms.push_back({ti_rk_a[j][ti_istage][ti_nstage] * dt_, olds[i][j + 1]});
        update_with_deriv(*curr[i], *olds[i][1], terms, *olds[i][ti_istage + 1], dy ? dy[i] : nullptr, dz ? dz[i] : nullptr,
                          true);
      }
      ti_istage = ti_istage + 1;
    }
  }
  void adams_bashforth(Field* curr[3], Field* deriv[3], double dt_) {
    const int nstep = std::min(ti_istep, ti_nstep);
    for (int i = 0; i < 3; ++i) {
      Terms terms;
      terms.push_back({ti_coeffs[1][nstep] * dt_, deriv[i]});
      for (int j = 2; j <= nstep; ++j) terms.push_back({ti_coeffs[j][nstep] * dt_, olds[i][j - 1]});
      lincomb_or_copy(*curr[i], *curr[i], terms);
      auto rotate = [&](int n) {
        Field* ptr = olds[i][n];
        for (int q = n; q >= 2; --q) olds[i][q] = olds[i][q - 1];
        olds[i][1] = ptr;
      };
      if (nstep < ti_nstep) { if (ti_istep > 1) rotate(nstep); }
      else { if (ti_nstep > 2) rotate(nstep - 1); }
      if (ti_nstep > 1) std::swap(olds[i][1], deriv[i]);  // olds(1) <- deriv
    }
    ti_istep = ti_istep + 1;
  }

  // ---- the channel case's hooks (case/channel.f90:59-228) around the generic loop. Wall values: zero (inlet_noise = 0,
  // the parity configuration). The bulk velocity is reduced once (see the oracle's note on channel.f90:76-78).
  int case_kind = 0;  // 0: none (TGV, generic), 1: channel
  double omega_rot = 0.0;
  int n_rotate = 0, iter = 1;
  Field *bc_u = nullptr, *bc_v = nullptr, *bc_w = nullptr;
  void set_case_channel(double omega, int n_rot) {
    case_kind = 1;
    omega_rot = omega;
    n_rotate = n_rot;
    if (!bc_u) {
      bc_u = allocator.get_block(DIR_X, VERT); bc_v = allocator.get_block(DIR_X, VERT); bc_w = allocator.get_block(DIR_X, VERT);
      for (Field* b : {bc_u, bc_v, bc_w}) { X3D2H_CALL(x3d2c_field_fill(ctx, b->dev, 0.0)); b->data_loc = VERT; }
    }
  }
  void define_BC() {  // channel.f90:59-80
    if (case_kind != 1) return;
    double ub = backend.field_volume_integral(*u);
    ub = ub / ((double)mesh.global_cell_dims[0] * mesh.global_cell_dims[1] * mesh.global_cell_dims[2]);
    backend.field_shift(*u, 2.0 / 3.0 - ub);
  }
  void forcings(Field& du, Field& dv) {  // channel.f90:189-204
    if (case_kind != 1 || omega_rot == 0.0 || iter >= n_rotate) return;
    backend.vecadd(-omega_rot, *v, 1.0, du);
    backend.vecadd(omega_rot, *u, 1.0, dv);
  }
  void apply_BC() {  // channel.f90:211-228
    if (case_kind != 1) return;
    backend.field_set_face_from_field(*u, *bc_u, 0.0, Y_FACE);
    backend.field_set_face_from_field(*v, *bc_v, 0.0, Y_FACE);
    backend.field_set_face_from_field(*w, *bc_w, 0.0, Y_FACE);
  }

  // base_case.f90:246-289: one time step = nstage x (define_BC, transeq, forcings, time integration, apply_BC, pressure
  // correction)
  void step() {
    Field* curr[3] = {u, v, w};
    for (int sub = 1; sub <= ti_nstage; ++sub) {
      define_BC();
      Field* deriv[3] = {allocator.get_block(DIR_X), allocator.get_block(DIR_X), allocator.get_block(DIR_X)};
      if (!base_ops() && !ti_is_ab && !(cfg.flags & X3D2C_FLAG_STRICT)) {  // Runge-Kutta, fast mode: the y / z sums of
        // transeq are taken inside the update pass
        Field *dy[3], *dz[3];
        transeq_parts(*deriv[0], *deriv[1], *deriv[2], dy, dz, *u, *v, *w);
        forcings(*deriv[0], *deriv[1]);  // linear source terms may join the x contribution before the sums
        runge_kutta(curr, deriv, dt, dy, dz);
        for (int i = 0; i < 3; ++i) { allocator.release_block(dy[i]); allocator.release_block(dz[i]); }
      } else {
        transeq_default(*deriv[0], *deriv[1], *deriv[2], *u, *v, *w);
        forcings(*deriv[0], *deriv[1]);
        if (base_ops()) { if (ti_is_ab) adams_bashforth_base(curr, deriv, dt); else runge_kutta_base(curr, deriv, dt); }
        else if (ti_is_ab) adams_bashforth(curr, deriv, dt);
        else runge_kutta(curr, deriv, dt);
      }
      u = curr[0]; v = curr[1]; w = curr[2];  // the integrators may continue in another block
      for (int i = 0; i < 3; ++i) allocator.release_block(deriv[i]);
      apply_BC();
      pressure_correction(*u, *v, *w);
    }
    iter = iter + 1;
  }

  // Independent batches streamed through one time step each (ensemble members, parameter sweeps): batch b's velocity is
  // uploaded from host_in, advanced by one step, and downloaded to host_out. The upload of batch b + 1 and the download
  // of batch b - 1 run on the backend's copy lanes while the kernels of batch b run on its stream; staging blocks in the
  // Cartesian layout are double-buffered. Host arrays: un-padded DIR_C blocks (the grid must need no padding), page-locked.
  void step_batches(int n_batches, const double* const host_in[3], double* const host_out[3]) {
    if (!is_unpadded(VERT)) fail("step_batches: the grid needs padding; use set_velocity / step / get_velocity");
    if (ti_is_ab) fail("step_batches: Adams-Bashforth carries history from step to step; batches are independent");
    Allocator& A = allocator;
    Field* in[2][3];
    Field* out[2][3];
    for (int b = 0; b < 2; ++b)
      for (int i = 0; i < 3; ++i) { in[b][i] = A.get_block(DIR_C); out[b][i] = A.get_block(DIR_C); }
    // events: 0,1 upload of set b done; 2,3 staging-in of set b consumed; 4,5 results staged in set b; 6,7 download of set b done
    auto upload = [&](int b) {
      const int sset = b & 1;
      X3D2H_CALL(x3d2c_lane_wait(ctx, 1, 2 + sset));  // the batch that used this set before has been reordered out of it
      for (int i = 0; i < 3; ++i) X3D2H_CALL(x3d2c_copy_data_to_f_async(ctx, in[sset][i]->dev, host_in[i], 1));
      X3D2H_CALL(x3d2c_lane_record(ctx, 1, 0 + sset));
    };
    if (n_batches > 0) upload(0);
    Field* state[3];
    for (int b = 0; b < n_batches; ++b) {
      const int sset = b & 1;
      if (b + 1 < n_batches) upload(b + 1);
      X3D2H_CALL(x3d2c_lane_wait(ctx, 0, 0 + sset));
      state[0] = u; state[1] = v; state[2] = w;
      for (int i = 0; i < 3; ++i) { backend.reorder(*state[i], *in[sset][i], RDR_C2X); state[i]->data_loc = VERT; }
      X3D2H_CALL(x3d2c_lane_record(ctx, 0, 2 + sset));
      step();
      X3D2H_CALL(x3d2c_lane_wait(ctx, 0, 6 + sset));  // the previous download from this set has finished
      state[0] = u; state[1] = v; state[2] = w;
      for (int i = 0; i < 3; ++i) backend.reorder(*out[sset][i], *state[i], RDR_X2C);
      X3D2H_CALL(x3d2c_lane_record(ctx, 0, 4 + sset));
      X3D2H_CALL(x3d2c_lane_wait(ctx, 2, 4 + sset));
      for (int i = 0; i < 3; ++i) X3D2H_CALL(x3d2c_copy_f_to_data_async(ctx, host_out[i], out[sset][i]->dev, 2));
      X3D2H_CALL(x3d2c_lane_record(ctx, 2, 6 + sset));
    }
    X3D2H_CALL(x3d2c_lane_sync(ctx, 2));
    X3D2H_CALL(x3d2c_lane_sync(ctx, 1));
    X3D2H_CALL(x3d2c_sync(ctx));
    for (int b = 0; b < 2; ++b)
      for (int i = 0; i < 3; ++i) { A.release_block(in[b][i]); A.release_block(out[b][i]); }
  }

  // ---- host <-> field helpers: local Cartesian un-padded arrays of the data_loc extents
  void pad(std::vector<double>& padded, const double* compact, int data_loc) const {
    int dims[3];
    mesh.get_dims(dims, data_loc);
    const int* p = allocator.dims_padded;
    padded.assign((size_t)allocator.ngrid, 0.0);
    for (int k = 0; k < dims[2]; ++k)
      for (int j = 0; j < dims[1]; ++j)
        std::memcpy(&padded[(size_t)p[0] * (j + (size_t)p[1] * k)], compact + (size_t)dims[0] * (j + (size_t)dims[1] * k),
                    sizeof(double) * dims[0]);
  }
  void unpad(double* compact, const std::vector<double>& padded, int data_loc) const {
    int dims[3];
    mesh.get_dims(dims, data_loc);
    const int* p = allocator.dims_padded;
    for (int k = 0; k < dims[2]; ++k)
      for (int j = 0; j < dims[1]; ++j)
        std::memcpy(compact + (size_t)dims[0] * (j + (size_t)dims[1] * k), &padded[(size_t)p[0] * (j + (size_t)p[1] * k)],
                    sizeof(double) * dims[0]);
  }
  bool is_unpadded(int data_loc) const {
    int dims[3];
    mesh.get_dims(dims, data_loc);
    const int* p = allocator.dims_padded;
    return dims[0] == p[0] && dims[1] == p[1] && dims[2] == p[2];
  }
  void set_field(Field& f, const double* compact, int data_loc) {
    f.data_loc = data_loc;
    if (is_unpadded(data_loc)) {  // the caller's array already is the DIR_C block: copy straight from it
      backend.set_field_data(f, compact);
    } else {
      std::vector<double> padded;
      pad(padded, compact, data_loc);
      backend.set_field_data(f, padded.data());
    }
    f.data_loc = data_loc;
  }
  void get_field(double* compact, const Field& f, int data_loc) {
    if (is_unpadded(data_loc)) {
      backend.get_field_data(compact, f);
      return;
    }
    std::vector<double> padded((size_t)allocator.ngrid);
    backend.get_field_data(padded.data(), f);
    unpad(compact, padded, data_loc);
  }

  // case/tgv.f90:41-72 through base_case.f90:139-179 (set_init)
  void init_tgv() {
    const int nx = mesh.vert_dims[0], ny = mesh.vert_dims[1], nz = mesh.vert_dims[2];
    std::vector<double> hu((size_t)nx * ny * nz), hv(hu.size());
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          const double x = mesh.vert_coords[0][i], y = mesh.vert_coords[1][j], z = mesh.vert_coords[2][k];
          hu[i + (size_t)nx * (j + (size_t)ny * k)] = std::sin(x) * std::cos(y) * std::cos(z);
          hv[i + (size_t)nx * (j + (size_t)ny * k)] = -std::cos(x) * std::sin(y) * std::cos(z);
        }
    set_field(*u, hu.data(), VERT);
    set_field(*v, hv.data(), VERT);
    X3D2H_CALL(x3d2c_field_fill(ctx, w->dev, 0.0));
    u->data_loc = VERT; v->data_loc = VERT; w->data_loc = VERT;
  }

  // monitoring.f90:46-90 (+ kinetic energy with the same normalisation)
  double enstrophy() {
    Field *du = allocator.get_block(DIR_X, VERT), *dv = allocator.get_block(DIR_X, VERT), *dw = allocator.get_block(DIR_X, VERT);
    curl(*du, *dv, *dw, *u, *v, *w);
    const double e = 0.5 * (backend.scalar_product(*du, *du) + backend.scalar_product(*dv, *dv) + backend.scalar_product(*dw, *dw)) / ngrid;
    allocator.release_block(du); allocator.release_block(dv); allocator.release_block(dw);
    return e;
  }
  double kinetic_energy() {
    return 0.5 * (backend.scalar_product(*u, *u) + backend.scalar_product(*v, *v) + backend.scalar_product(*w, *w)) / ngrid;
  }
  void divergence_max_mean(double& mx, double& mean) {
    Field* div_u = allocator.get_block(DIR_Z);
    divergence_v2c(*div_u, *u, *v, *w);
    backend.field_max_mean(mx, mean, *div_u);
    allocator.release_block(div_u);
  }

  const DevTdsops& pick(int dir, const std::string& name) const {
    const DevDirps& d = dir == DIR_X ? xdirps : (dir == DIR_Y ? ydirps : zdirps);
    if (name == "der1st") return d.der1st;
    if (name == "der1st_sym") return d.der1st_sym;
    if (name == "der2nd") return d.der2nd;
    if (name == "der2nd_sym") return d.der2nd_sym;
    if (name == "stagder_v2p") return d.stagder_v2p;
    if (name == "stagder_p2v") return d.stagder_p2v;
    if (name == "interpl_v2p") return d.interpl_v2p;
    if (name == "interpl_p2v") return d.interpl_p2v;
    fail("unknown operator name " + name);
  }
};

}  // namespace x3d2h
