// C entry points of the host layer (include/x3d2h.h).
#include "../../../include/x3d2h.h"

#include "sim.hpp"

using namespace x3d2h;

struct x3d2h_sim {
  std::unique_ptr<Sim> s;
};

static thread_local std::string g_err;
#define H_TRY try {
#define H_CATCH                                  \
  }                                              \
  catch (const std::exception& e) {              \
    g_err = e.what();                            \
    return X3D2C_EINVAL;                         \
  }                                              \
  return X3D2C_OK;

static Config to_config(const x3d2h_config* c) {
  Config k;
  for (int q = 0; q < 3; ++q) {
    k.dims_global[q] = c->dims_global[q];
    k.nproc_dir[q] = c->nproc_dir[q];
    k.L_global[q] = c->L_global[q];
    k.bc[q][0] = c->bc[2 * q];
    k.bc[q][1] = c->bc[2 * q + 1];
  }
  k.Re = c->Re; k.dt = c->dt;
  if (c->time_intg) k.time_intg = c->time_intg;
  if (c->der1st_scheme) k.der1st = c->der1st_scheme;
  if (c->der2nd_scheme) k.der2nd = c->der2nd_scheme;
  if (c->interpl_scheme) k.interpl = c->interpl_scheme;
  if (c->stagder_scheme) k.stagder = c->stagder_scheme;
  k.rank = c->rank; k.nproc = c->nproc; k.device = c->device; k.flags = c->flags;
  k.nccl_unique_id = c->nccl_unique_id;
  for (int q = 0; q < 3; ++q) {
    if (c->stretching[q]) k.stretching[q] = c->stretching[q];
    k.beta[q] = c->beta[q] > 0 ? c->beta[q] : 1.0;
  }
  return k;
}

namespace {
struct Tmp {  // RAII set of pool blocks
  Allocator& a;
  std::vector<Field*> f;
  explicit Tmp(Allocator& al) : a(al) {}
  Field* get(int dir, int loc = NULL_LOC) { f.push_back(a.get_block(dir, loc)); return f.back(); }
  ~Tmp() { for (auto it = f.rbegin(); it != f.rend(); ++it) a.release_block(*it); }
};
}  // namespace

extern "C" {

const char* x3d2h_last_error(void) { return g_err.c_str(); }

int x3d2h_decompose(const x3d2h_config* cfg, int* out) {
  H_TRY
  Mesh m;
  m.init(to_config(cfg));
  for (int q = 0; q < 3; ++q) {
    out[q] = m.vert_dims[q]; out[3 + q] = m.cell_dims[q]; out[6 + q] = m.n_offset[q]; out[9 + q] = m.nrank_dir[q];
    out[12 + q] = m.pprev[q]; out[15 + q] = m.pnext[q]; out[18 + 2 * q] = m.BCs[q][0]; out[19 + 2 * q] = m.BCs[q][1];
  }
  H_CATCH
}

int x3d2h_geo(const x3d2h_config* cfg, int dir, double* vc, double* vds, double* vds2, double* vd2s, double* mc, double* mds) {
  H_TRY
  Mesh m;
  m.init(to_config(cfg));
  for (size_t i = 0; i < m.vert_coords[dir].size(); ++i) {
    vc[i] = m.vert_coords[dir][i]; vds[i] = m.vert_ds[dir][i]; vds2[i] = m.vert_ds2[dir][i]; vd2s[i] = m.vert_d2s[dir][i];
  }
  for (size_t i = 0; i < m.midp_coords[dir].size(); ++i) { mc[i] = m.midp_coords[dir][i]; mds[i] = m.midp_ds[dir][i]; }
  H_CATCH
}

int x3d2h_tdsops_tables(int n_tds, double delta, const char* operation, const char* scheme, int bc_start, int bc_end,
                        const double* stretch, const double* stretch_correct, int n_halo, const char* from_to, int sym,
                        int* info, double* sc, double* coeffs, double* coeffs_s, double* coeffs_e, double* dist_fw,
                        double* dist_bw, double* dist_sa, double* dist_sc, double* dist_af, double* stretch_out,
                        double* stretch_correct_out) {
  H_TRY
  Tdsops t = tdsops_init(n_tds, delta, operation, scheme, bc_start, bc_end, stretch, stretch_correct, n_halo,
                         from_to ? from_to : "", sym != 0);
  info[0] = t.n_tds; info[1] = t.n_rhs; info[2] = t.move; info[3] = t.periodic;
  sc[0] = t.alpha; sc[1] = t.a; sc[2] = t.b; sc[3] = t.c; sc[4] = t.d;
  for (int k = 1; k <= 9; ++k) coeffs[k - 1] = t.coeffs[k];
  for (int i = 1; i <= 4; ++i)
    for (int k = 1; k <= 9; ++k) {
      coeffs_s[(i - 1) * 9 + k - 1] = t.coeffs_s[i][k];
      coeffs_e[(i - 1) * 9 + k - 1] = t.coeffs_e[i][k];
    }
  for (int i = 1; i <= t.n_rhs; ++i) {
    dist_fw[i - 1] = t.dist_fw[i]; dist_bw[i - 1] = t.dist_bw[i]; dist_sa[i - 1] = t.dist_sa[i];
    dist_sc[i - 1] = t.dist_sc[i]; dist_af[i - 1] = t.dist_af[i];
  }
  for (int i = 1; i <= t.n_tds; ++i) { stretch_out[i - 1] = t.stretch[i]; stretch_correct_out[i - 1] = t.stretch_correct[i]; }
  H_CATCH
}

int x3d2h_waves_000(const x3d2h_config* cfg, double* waves) {
  H_TRY
  Config k = to_config(cfg);
  k.rank = 0; k.nproc = 1; k.nproc_dir[0] = k.nproc_dir[1] = k.nproc_dir[2] = 1;
  Mesh m;
  m.init(k);
  DevDirps d[3];
  for (int q = 0; q < 3; ++q) {
    const int n_cell = m.get_n(q + 1, CELL);
    d[q].stagder_v2p.t = tdsops_init(n_cell, m.d[q], "stag-deriv", k.stagder, m.BCs[q][0], m.BCs[q][1], nullptr, nullptr, 4, "v2p");
    d[q].interpl_v2p.t = tdsops_init(n_cell, m.d[q], "interpolate", k.interpl, m.BCs[q][0], m.BCs[q][1], nullptr, nullptr, 4, "v2p");
  }
  PoissonFFT p;
  int n_spec[3] = {m.global_cell_dims[0] / 2 + 1, m.global_cell_dims[1], m.global_cell_dims[2]}, st[3] = {0, 0, 0};
  p.base_init(m, d[0], d[1], d[2], n_spec, st);
  std::memcpy(waves, p.waves.data(), sizeof(cplx) * p.waves.size());
  H_CATCH
}

// base_init of the Poisson solver for a whole (single-rank) domain with walls in y: the wave-number table and, on a
// stretched mesh, the pentadiagonal spectral operators (src/poisson_fft.f90:275-652). info = {stretched, rows}
int x3d2h_poisson_tables_010(const x3d2h_config* cfg, int* info, double* waves, double* a_odd, double* a_even) {
  H_TRY
  Config k = to_config(cfg);
  k.rank = 0; k.nproc = 1; k.nproc_dir[0] = k.nproc_dir[1] = k.nproc_dir[2] = 1;
  Mesh m;
  m.init(k);
  DevDirps d[3];
  for (int q = 0; q < 3; ++q) {
    const int n_cell = m.get_n(q + 1, CELL);
    const int b0 = m.BCs[q][0] == BC_DIRICHLET ? BC_NEUMANN : m.BCs[q][0], b1 = m.BCs[q][1] == BC_DIRICHLET ? BC_NEUMANN : m.BCs[q][1];
    d[q].stagder_v2p.t = tdsops_init(n_cell, m.d[q], "stag-deriv", k.stagder, b0, b1, m.midp_ds[q].data(), nullptr, 4, "v2p");
    d[q].interpl_v2p.t = tdsops_init(n_cell, m.d[q], "interpolate", k.interpl, b0, b1, nullptr, nullptr, 4, "v2p");
  }
  PoissonFFT p;
  int n_spec[3] = {m.global_cell_dims[0] / 2 + 1, m.global_cell_dims[1], m.global_cell_dims[2]}, st[3] = {0, 0, 0};
  p.base_init(m, d[0], d[1], d[2], n_spec, st);
  info[0] = p.stretched; info[1] = p.penta_rows;
  if (waves) std::memcpy(waves, p.waves.data(), sizeof(cplx) * p.waves.size());
  if (a_odd && !p.a_odd.empty()) std::memcpy(a_odd, p.a_odd.data(), sizeof(double) * p.a_odd.size());
  if (a_even && !p.a_even.empty()) std::memcpy(a_even, p.a_even.data(), sizeof(double) * p.a_even.size());
  H_CATCH
}

int x3d2h_create(const x3d2h_config* cfg, x3d2h_sim** out) {
  H_TRY
  auto* h = new x3d2h_sim;
  try {
    h->s.reset(new Sim(to_config(cfg)));
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  H_CATCH
}
int x3d2h_destroy(x3d2h_sim* sim) {
  delete sim;
  return X3D2C_OK;
}
x3d2c_ctx* x3d2h_backend(x3d2h_sim* sim) { return sim->s->ctx; }
int x3d2h_local_dims(x3d2h_sim* sim, int data_loc, int dims[3]) {
  H_TRY sim->s->mesh.get_dims(dims, data_loc); H_CATCH
}

int x3d2h_init_tgv(x3d2h_sim* sim) { H_TRY sim->s->init_tgv(); H_CATCH }
int x3d2h_set_velocity(x3d2h_sim* sim, const double* u, const double* v, const double* w) {
  H_TRY
  Sim& S = *sim->s;
  S.set_field(*S.u, u, VERT); S.set_field(*S.v, v, VERT); S.set_field(*S.w, w, VERT);
  H_CATCH
}
int x3d2h_get_velocity(x3d2h_sim* sim, double* u, double* v, double* w) {
  H_TRY
  Sim& S = *sim->s;
  S.get_field(u, *S.u, VERT); S.get_field(v, *S.v, VERT); S.get_field(w, *S.w, VERT);
  H_CATCH
}
int x3d2h_set_case_channel(x3d2h_sim* sim, double omega_rot, int n_rotate) {
  H_TRY
  sim->s->set_case_channel(omega_rot, n_rotate);
  H_CATCH
}
int x3d2h_step(x3d2h_sim* sim, int nsteps) {
  H_TRY
  for (int i = 0; i < nsteps; ++i) sim->s->step();
  H_CATCH
}
int x3d2h_step_batches(x3d2h_sim* sim, int n_batches, const double* u_in, const double* v_in, const double* w_in,
                       double* u_out, double* v_out, double* w_out) {
  H_TRY
  const double* in[3] = {u_in, v_in, w_in};
  double* out[3] = {u_out, v_out, w_out};
  if (!u_in || !v_in || !w_in || !u_out || !v_out || !w_out) fail("x3d2h_step_batches: null argument");
  sim->s->step_batches(n_batches, in, out);
  H_CATCH
}
int x3d2h_sync(x3d2h_sim* sim) { H_TRY X3D2H_CALL(x3d2c_sync(sim->s->ctx)); H_CATCH }
int x3d2h_monitor(x3d2h_sim* sim, double out[4]) {
  H_TRY
  Sim& S = *sim->s;
  out[0] = S.enstrophy();
  out[1] = S.kinetic_energy();
  S.divergence_max_mean(out[2], out[3]);
  H_CATCH
}

int x3d2h_transeq(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* du, double* dv, double* dw) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fu = t.get(DIR_X), *fv = t.get(DIR_X), *fw = t.get(DIR_X), *a = t.get(DIR_X), *b = t.get(DIR_X), *c = t.get(DIR_X);
  S.set_field(*fu, u, VERT); S.set_field(*fv, v, VERT); S.set_field(*fw, w, VERT);
  S.transeq_default(*a, *b, *c, *fu, *fv, *fw);
  S.get_field(du, *a, VERT); S.get_field(dv, *b, VERT); S.get_field(dw, *c, VERT);
  H_CATCH
}
int x3d2h_transeq_lowmem(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* du, double* dv,
                         double* dw, double* u_back) {
  H_TRY
  Sim& S = *sim->s;
  Allocator& A = S.allocator;
  Field *fu = A.get_block(DIR_X), *fv = A.get_block(DIR_X), *fw = A.get_block(DIR_X);
  Tmp t(A);
  Field *a = t.get(DIR_X), *b = t.get(DIR_X), *c = t.get(DIR_X);
  S.set_field(*fu, u, VERT); S.set_field(*fv, v, VERT); S.set_field(*fw, w, VERT);
  S.transeq_lowmem(*a, *b, *c, fu, fv, fw);  // fu, fv, fw now point at the blocks that came back from the z layout
  S.get_field(du, *a, VERT); S.get_field(dv, *b, VERT); S.get_field(dw, *c, VERT);
  if (u_back) S.get_field(u_back, *fu, VERT);
  A.release_block(fu); A.release_block(fv); A.release_block(fw);
  H_CATCH
}
int x3d2h_transeq_species(x3d2h_sim* sim, const double* u, const double* v, const double* w, const double* spec,
                          double nu_s, double* dspec) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fu = t.get(DIR_X), *fv = t.get(DIR_X), *fw = t.get(DIR_X), *fs = t.get(DIR_X), *a = t.get(DIR_X);
  S.set_field(*fu, u, VERT); S.set_field(*fv, v, VERT); S.set_field(*fw, w, VERT); S.set_field(*fs, spec, VERT);
  Field* rhs[1] = {a};
  Field* sp[1] = {fs};
  S.transeq_species(rhs, 1, *fu, *fv, *fw, sp, &nu_s);
  S.get_field(dspec, *a, VERT);
  H_CATCH
}
int x3d2h_derived(x3d2h_sim* sim, const char* what, const double* const* grads, double* out) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field* g[9];
  for (int k = 0; k < 9; ++k) { g[k] = t.get(DIR_X); S.set_field(*g[k], grads[k], VERT); }
  Field* o = t.get(DIR_X, VERT);
  if (std::string(what) == "vorticity") S.backend.compute_vorticity(*o, g);
  else if (std::string(what) == "qcriterion") S.backend.compute_qcriterion(*o, g);
  else fail("x3d2h_derived: unknown quantity");
  S.get_field(out, *o, VERT);
  H_CATCH
}
int x3d2h_slice_max_sum(x3d2h_sim* sim, int dir, int data_loc, const double* x, int i_slice, double* mx, double* sum) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field* f = t.get(dir, data_loc);
  S.set_field(*f, x, data_loc);
  S.backend.slice_max_sum(*mx, *sum, *f, i_slice);
  H_CATCH
}
int x3d2h_transeq_dir(x3d2h_sim* sim, int dir, const double* u, const double* v, const double* w, double* du,
                      double* dv, double* dw) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fu = t.get(dir), *fv = t.get(dir), *fw = t.get(dir), *a = t.get(dir), *b = t.get(dir), *c = t.get(dir);
  S.set_field(*fu, u, VERT); S.set_field(*fv, v, VERT); S.set_field(*fw, w, VERT);
  const DevDirps& dp = dir == DIR_X ? S.xdirps : (dir == DIR_Y ? S.ydirps : S.zdirps);
  S.backend.transeq(dir, *a, *b, *c, *fu, *fv, *fw, S.nu, dp);
  S.get_field(du, *a, VERT); S.get_field(dv, *b, VERT); S.get_field(dw, *c, VERT);
  H_CATCH
}
int x3d2h_tds_solve(x3d2h_sim* sim, int dir, const char* opname, int in_loc, const double* in, double* out, int* out_loc) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fi = t.get(dir), *fo = t.get(dir);
  S.set_field(*fi, in, in_loc);
  S.backend.tds_solve(*fo, *fi, S.pick(dir, opname));
  *out_loc = fo->data_loc;
  S.get_field(out, *fo, fo->data_loc);
  H_CATCH
}
int x3d2h_tds_fused(x3d2h_sim* sim, const char* mode, int dir, const char* op_a, const char* op_b, int in_loc,
                    int out_loc, const double* in_a, const double* in_b, double a, double* out_a, double* out_b) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  const std::string m = mode;
  Field *fa = t.get(dir), *fb = t.get(dir), *oa = t.get(dir), *ob = t.get(dir);
  S.set_field(*fa, in_a, in_loc);
  if (m == "sum") {
    S.set_field(*fb, in_b, in_loc);
    S.backend.tds_solve_sum(*oa, *fa, S.pick(dir, op_a), *fb, S.pick(dir, op_b));
    S.get_field(out_a, *oa, out_loc);
  } else if (m == "dual") {
    S.backend.tds_solve_dual(*oa, *ob, *fa, S.pick(dir, op_a), S.pick(dir, op_b));
    S.get_field(out_a, *oa, out_loc);
    S.get_field(out_b, *ob, out_loc);
  } else if (m == "axpy") {
    S.set_field(*oa, in_b, out_loc);  // y
    S.backend.tds_solve_axpy(*oa, a, *fa, S.pick(dir, op_a));
    S.get_field(out_a, *oa, out_loc);
  } else {
    fail("x3d2h_tds_fused: mode must be sum, dual or axpy");
  }
  H_CATCH
}
int x3d2h_tds_fused_r(x3d2h_sim* sim, const char* mode, int dir, const char* op_a, const char* op_b, int in_loc,
                      int out_loc, int rdr_in, int rdr_out, const double* in_a, const double* in_b, double* out_a,
                      double* out_b) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  const std::string m = mode;
  const int din = rdr_in ? rdr_in / 10 : dir, dout = rdr_out ? rdr_out % 10 : dir;
  Field *fa = t.get(din), *fb = t.get(din), *oa = t.get(dout), *ob = t.get(dout);
  S.set_field(*fa, in_a, in_loc);
  if (m == "single") {
    S.backend.tds_solve_r(dir, *oa, *fa, S.pick(dir, op_a), rdr_in, rdr_out);
  } else if (m == "sum") {
    S.set_field(*fb, in_b, in_loc);
    S.backend.tds_solve_sum_r(dir, *oa, *fa, S.pick(dir, op_a), *fb, S.pick(dir, op_b), rdr_in, rdr_out);
  } else if (m == "dual") {
    S.backend.tds_solve_dual_r(dir, *oa, *ob, *fa, S.pick(dir, op_a), S.pick(dir, op_b), rdr_in, rdr_out);
    S.get_field(out_b, *ob, out_loc);
  } else if (m == "axpy") {  // out_a = in_b - A(reorder(in_a)); in_b has the output's extents
    if (rdr_out) fail("x3d2h_tds_fused_r: axpy takes no output reorder");
    S.set_field(*oa, in_b, out_loc);
    S.backend.tds_solve_axpy_r(dir, *oa, -1.0, *fa, S.pick(dir, op_a), rdr_in);
  } else {
    fail("x3d2h_tds_fused_r: mode must be single, sum, dual or axpy");
  }
  S.get_field(out_a, *oa, out_loc);
  H_CATCH
}
int x3d2h_divergence(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* div) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fu = t.get(DIR_X), *fv = t.get(DIR_X), *fw = t.get(DIR_X), *d = t.get(DIR_Z);
  S.set_field(*fu, u, VERT); S.set_field(*fv, v, VERT); S.set_field(*fw, w, VERT);
  S.divergence_v2c(*d, *fu, *fv, *fw);
  S.get_field(div, *d, CELL);
  H_CATCH
}
int x3d2h_gradient(x3d2h_sim* sim, const double* p, double* gx, double* gy, double* gz) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fp = t.get(DIR_Z), *a = t.get(DIR_X), *b = t.get(DIR_X), *c = t.get(DIR_X);
  S.set_field(*fp, p, CELL);
  S.gradient_c2v(*a, *b, *c, *fp);
  S.get_field(gx, *a, VERT); S.get_field(gy, *b, VERT); S.get_field(gz, *c, VERT);
  H_CATCH
}
int x3d2h_interpl_c2v(x3d2h_sim* sim, const double* p, double* out) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fp = t.get(DIR_Z), *a = t.get(DIR_X);
  S.set_field(*fp, p, CELL);
  S.interpl_c2v(*a, *fp);
  S.get_field(out, *a, VERT);
  H_CATCH
}
int x3d2h_laplacian(x3d2h_sim* sim, const double* u, double* out) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fu = t.get(DIR_X), *a = t.get(DIR_X);
  S.set_field(*fu, u, VERT);
  S.laplacian(*a, *fu);
  S.get_field(out, *a, VERT);
  H_CATCH
}
int x3d2h_curl(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* ox, double* oy, double* oz) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fu = t.get(DIR_X), *fv = t.get(DIR_X), *fw = t.get(DIR_X), *a = t.get(DIR_X), *b = t.get(DIR_X), *c = t.get(DIR_X);
  S.set_field(*fu, u, VERT); S.set_field(*fv, v, VERT); S.set_field(*fw, w, VERT);
  S.curl(*a, *b, *c, *fu, *fv, *fw);
  S.get_field(ox, *a, VERT); S.get_field(oy, *b, VERT); S.get_field(oz, *c, VERT);
  H_CATCH
}
int x3d2h_poisson(x3d2h_sim* sim, const double* f, double* p) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *ff = t.get(DIR_Z), *fp = t.get(DIR_Z);
  S.set_field(*ff, f, CELL);
  S.poisson_fft(*fp, *ff);
  S.get_field(p, *fp, CELL);
  H_CATCH
}
int x3d2h_pressure_correction(x3d2h_sim* sim) {
  H_TRY
  Sim& S = *sim->s;
  S.pressure_correction(*S.u, *S.v, *S.w);
  H_CATCH
}
int x3d2h_fft_roundtrip(x3d2h_sim* sim, const double* f, double* out, double* spec_re_im) {
  H_TRY
  Sim& S = *sim->s;
  if (!S.backend.poisson) fail("FFT Poisson solver is not initialised for these BCs");
  Tmp t(S.allocator);
  Field* c = t.get(DIR_C, CELL);
  S.set_field(*c, f, CELL);
  X3D2H_CALL(x3d2c_fft_forward(S.ctx, S.backend.poisson, c->dev));
  if (spec_re_im) X3D2H_CALL(x3d2c_poisson_get_spectrum(S.ctx, S.backend.poisson, spec_re_im));
  X3D2H_CALL(x3d2c_fft_backward(S.ctx, S.backend.poisson, c->dev));
  S.get_field(out, *c, CELL);
  H_CATCH
}
int x3d2h_reorder_chain(x3d2h_sim* sim, const double* in, const int* rdrs, int n_rdr, double* out) {
  H_TRY
  Sim& S = *sim->s;
  Field* cur = S.allocator.get_block(DIR_C, VERT);
  S.set_field(*cur, in, VERT);
  for (int q = 0; q < n_rdr; ++q) {
    int from, to;
    get_dirs_from_rdr(from, to, rdrs[q]);
    if (from != cur->dir) { S.allocator.release_block(cur); fail("reorder chain: direction mismatch"); }
    Field* nxt = S.allocator.get_block(to);
    S.backend.reorder(*nxt, *cur, rdrs[q]);
    S.allocator.release_block(cur);
    cur = nxt;
  }
  S.get_field(out, *cur, VERT);
  S.allocator.release_block(cur);
  H_CATCH
}
int x3d2h_sum_intox(x3d2h_sim* sim, int dir_from, const double* a, const double* b, double* out) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fa = t.get(DIR_X), *fb = t.get(dir_from);
  S.set_field(*fa, a, VERT); S.set_field(*fb, b, VERT);
  if (dir_from == DIR_Y) S.backend.sum_yintox(*fa, *fb); else S.backend.sum_zintox(*fa, *fb);
  S.get_field(out, *fa, VERT);
  H_CATCH
}
int x3d2h_vecadd(x3d2h_sim* sim, int dir, double a, const double* x, double b, const double* y, double* out) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fx = t.get(dir), *fy = t.get(dir);
  S.set_field(*fx, x, VERT); S.set_field(*fy, y, VERT);
  S.backend.vecadd(a, *fx, b, *fy);
  S.get_field(out, *fy, VERT);
  H_CATCH
}
int x3d2h_scalar_product(x3d2h_sim* sim, int dir, int data_loc, const double* x, const double* y, double* s) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field *fx = t.get(dir), *fy = t.get(dir);
  S.set_field(*fx, x, data_loc); S.set_field(*fy, y, data_loc);
  *s = S.backend.scalar_product(*fx, *fy);
  H_CATCH
}
int x3d2h_field_max_mean(x3d2h_sim* sim, int dir, int data_loc, const double* x, double* mx, double* mean) {
  H_TRY
  Sim& S = *sim->s;
  Tmp t(S.allocator);
  Field* fx = t.get(dir);
  S.set_field(*fx, x, data_loc);
  S.backend.field_max_mean(*mx, *mean, *fx);
  H_CATCH
}

int x3d2h_bench_op(x3d2h_sim* sim, const char* op_c, int reps) {
  H_TRY
  Sim& S = *sim->s;
  const std::string op = op_c;
  Tmp t(S.allocator);
  auto dir_of = [](char c) { return c == 'x' ? DIR_X : (c == 'y' ? DIR_Y : (c == 'z' ? DIR_Z : DIR_C)); };
  if (op.rfind("transeq_", 0) == 0) {
    const int dir = dir_of(op[8]);
    const DevDirps& dp = dir == DIR_X ? S.xdirps : (dir == DIR_Y ? S.ydirps : S.zdirps);
    Field *a = t.get(dir), *b = t.get(dir), *c = t.get(dir);
    // the solver's velocity blocks are plain storage; any pencil layout is valid input for timing
    for (int r = 0; r < reps; ++r) S.backend.transeq(dir, *a, *b, *c, *S.u, *S.v, *S.w, S.nu, dp);
  } else if (op.rfind("tds_solve_", 0) == 0) {
    const int dir = dir_of(op[10]);
    std::string name = op.size() > 12 ? op.substr(12) : "der1st";
    Field *a = t.get(dir), *b = t.get(dir);
    b->dir = dir;
    Field src = *S.u;
    src.dir = dir;
    for (int r = 0; r < reps; ++r) S.backend.tds_solve(*a, src, S.pick(dir, name));
  } else if (op.rfind("reorder_", 0) == 0) {
    const int from = dir_of(op[8]), to = dir_of(op[10]);
    Field* a = t.get(to);
    for (int r = 0; r < reps; ++r) X3D2H_CALL(x3d2c_reorder(S.ctx, 10 * from + to, a->dev, S.u->dev));
  } else if (op == "sum_yintox" || op == "sum_zintox") {
    Field* a = t.get(DIR_X);
    for (int r = 0; r < reps; ++r) {
      if (op == "sum_yintox") X3D2H_CALL(x3d2c_sum_yintox(S.ctx, a->dev, S.u->dev));
      else X3D2H_CALL(x3d2c_sum_zintox(S.ctx, a->dev, S.u->dev));
    }
  } else if (op == "vecadd") {
    Field* a = t.get(DIR_X);
    for (int r = 0; r < reps; ++r) S.backend.vecadd(0.5, *S.u, 1.0, *a);
  } else if (op == "veccopy") {
    Field* a = t.get(DIR_X);
    for (int r = 0; r < reps; ++r) S.backend.veccopy(*a, *S.u);
  } else if (op == "poisson") {
    Field *a = t.get(DIR_Z), *b = t.get(DIR_Z);
    for (int r = 0; r < reps; ++r) S.poisson_fft(*a, *b);
  } else if (op == "pressure_correction") {
    for (int r = 0; r < reps; ++r) S.pressure_correction(*S.u, *S.v, *S.w);
  } else if (op == "transeq") {
    Field *a = t.get(DIR_X), *b = t.get(DIR_X), *c = t.get(DIR_X);
    for (int r = 0; r < reps; ++r) S.transeq_default(*a, *b, *c, *S.u, *S.v, *S.w);
  } else {
    fail("x3d2h_bench_op: unknown op " + op);
  }
  H_CATCH
}

}  // extern "C"

// remaining elementwise ops / reductions of base_backend_t on host data (parity tests):
// op = "scale" | "shift" | "vecmult" | "veccopy" | "fill" | "volume_integral"
extern "C" int x3d2h_fieldop(x3d2h_sim* sim, const char* op_c, int dir, int data_loc, double a, const double* x,
                             const double* y, double* out, double* s) {
  H_TRY
  Sim& S = *sim->s;
  const std::string op = op_c;
  Tmp t(S.allocator);
  Field *fx = t.get(dir, data_loc), *fy = t.get(dir, data_loc);
  S.set_field(*fx, x, data_loc);
  if (y) S.set_field(*fy, y, data_loc);
  if (op == "scale") X3D2H_CALL(x3d2c_field_scale(S.ctx, fx->dev, a));
  else if (op == "shift") X3D2H_CALL(x3d2c_field_shift(S.ctx, fx->dev, a));
  else if (op == "vecmult") { X3D2H_CALL(x3d2c_vecmult(S.ctx, fy->dev, fx->dev)); fx = fy; }
  else if (op == "veccopy") { S.backend.veccopy(*fy, *fx); fx = fy; }
  else if (op == "fill") X3D2H_CALL(x3d2c_field_fill(S.ctx, fx->dev, a));
  else if (op == "set_face") {  // a = c_start; s[0] = c_end, s[1] = face
    X3D2H_CALL(x3d2c_field_set_face(S.ctx, fx->dev, data_loc, a, s[0], (int)s[1]));
  } else if (op == "set_face_from_field") {  // y = f_start; a = c_end; s[0] = flow_rate_diff, s[1] = face
    X3D2H_CALL(x3d2c_field_set_face_from_field(S.ctx, fx->dev, fy->dev, data_loc, a, (int)s[1], s[0]));
  }
  else if (op == "lincomb") {  // out = x; out = a y + out; out = (-a / 2) y + out, written over x (out aliases base)
    S.backend.veclincomb(*fx, *fx, {{a, fy}, {-a / 2, fy}});
  }
  else if (op == "volume_integral") X3D2H_CALL(x3d2c_field_volume_integral(S.ctx, data_loc, fx->dev, s));
  else fail("x3d2h_fieldop: unknown op " + op);
  if (out) S.get_field(out, *fx, data_loc);
  H_CATCH
}
