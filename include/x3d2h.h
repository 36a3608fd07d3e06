/* x3d2h.h — C entry points of the HOST layer (libx3d2h.so) that drives the cuda_c backend through the
 * C ABI of x3d2c.h. The host layer mirrors the reference's solver-side Fortran modules
 * (src/solver.f90, src/vector_calculus.f90, src/time_integrator.f90, src/tdsops.f90, src/mesh.f90,
 * src/allocator.f90, src/poisson_fft.f90, src/case/base_case.f90, src/case/tgv.f90,
 * src/postprocess/monitoring.f90); it exists because this image has no Fortran compiler.
 * Python (x3d2_b200/, tests/, bench.py) binds these with ctypes.
 *
 * Host arrays are rank-local, un-padded Cartesian arrays (x fastest) with the extents of the stated
 * data location (mesh%get_dims(data_loc)).
 */
#ifndef X3D2H_H
#define X3D2H_H

#include "x3d2c.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct x3d2h_sim x3d2h_sim;

/* host-layer flag (same `flags` word as X3D2C_FLAG_*, bits >= 8 never reach the backend):
 * issue the operator graph of the UNCHANGED reference solver (src/solver.f90:291-389,693-739,
 * src/vector_calculus.f90:142-332, src/time_integrator.f90:166-300) call by call through the deferred procedures of
 * base_backend_t only - no fused extension entry point. This is the drop-in path a Fortran cuda_c_backend_t gets. */
#define X3D2H_FLAG_BASE_OPS 0x100

/* keys of the reference's namelists domain_settings / solver_params (src/config.f90:104-205) */
typedef struct {
  int dims_global[3];
  int nproc_dir[3];
  double L_global[3];
  int bc[6];            /* BC_x(1:2), BC_y(1:2), BC_z(1:2): 0 periodic, 1 neumann, 2 dirichlet */
  double Re, dt;
  const char* time_intg;      /* 'AB[1-4]' | 'RK[1-4]' */
  const char* der1st_scheme;  /* 'compact6' */
  const char* der2nd_scheme;  /* 'compact6' | 'compact6-hyperviscous' (the latter: tdsops tables only) */
  const char* interpl_scheme; /* 'classic' | 'optimised' | 'aggressive' */
  const char* stagder_scheme; /* 'compact6' */
  int rank, nproc;            /* position in the job; one process per GPU */
  int device;                 /* CUDA device ordinal, -1 = current */
  int flags;                  /* X3D2C_FLAG_* | X3D2H_FLAG_* */
  const void* nccl_unique_id; /* 128 bytes when nproc > 1 */
  const char* stretching[3];  /* 'uniform' | 'centred' | 'top-bottom' | 'bottom' per direction; NULL = uniform */
  double beta[3];             /* stretching parameter (src/config.f90 domain_settings) */
} x3d2h_config;

const char* x3d2h_last_error(void);

/* ---- pure host logic (no GPU needed) */
/* mesh_t decomposition (src/mesh.f90:160-194, src/mesh_content.f90:72-121). out: vert_dims[3], cell_dims[3],
 * n_offset[3], nrank_dir[3], pprev[3], pnext[3], BCs[6] = 24 ints */
int x3d2h_decompose(const x3d2h_config* cfg, int* out24);
/* geo_t of rank cfg->rank along `dir` (0..2): vert_coords, vert_ds, vert_ds2, vert_d2s (n_vert), midp_coords, midp_ds (n_cell) */
int x3d2h_geo(const x3d2h_config* cfg, int dir, double* vc, double* vds, double* vds2, double* vd2s, double* mc, double* mds);
/* tdsops_init (src/tdsops.f90:63-203): fills the tables exactly as they are passed to x3d2c_tdsops_create.
 * info[4] = n_tds, n_rhs, move, periodic; sc[5] = alpha, a, b, c, d; arrays as in x3d2c_tdsops_create. */
int x3d2h_tdsops_tables(int n_tds, double delta, const char* operation, const char* scheme, int bc_start, int bc_end,
                        const double* stretch, const double* stretch_correct, int n_halo, const char* from_to, int sym,
                        int* info, double* sc, double* coeffs, double* coeffs_s, double* coeffs_e, double* dist_fw,
                        double* dist_bw, double* dist_sa, double* dist_sc, double* dist_af, double* stretch_out,
                        double* stretch_correct_out);
/* wave numbers + modified-wavenumber table of the periodic Poisson solver for a whole (single-rank) domain
 * (src/poisson_fft.f90:654-882); waves: (nx/2+1, ny, nz) interleaved re/im */
int x3d2h_waves_000(const x3d2h_config* cfg, double* waves);

/* same for walls in y (010): waves (nx/2+1, ny_cell, nz) and, on a stretched mesh, the pentadiagonal spectral operators
 * a_odd / a_even (nx/2+1, rows, nz, 5) of src/poisson_fft.f90:275-652; info = {stretched (0 | 1 | 2), rows} */
int x3d2h_poisson_tables_010(const x3d2h_config* cfg, int* info, double* waves, double* a_odd, double* a_even);

/* ---- simulation object: xcompact.f90:48-131 + solver init (src/solver.f90:111-212) */
int x3d2h_create(const x3d2h_config* cfg, x3d2h_sim** out);
int x3d2h_destroy(x3d2h_sim* sim);
x3d2c_ctx* x3d2h_backend(x3d2h_sim* sim);
int x3d2h_local_dims(x3d2h_sim* sim, int data_loc, int dims[3]);

/* case_tgv_t%initial_conditions (src/case/tgv.f90:41-72) */
int x3d2h_init_tgv(x3d2h_sim* sim);
int x3d2h_set_velocity(x3d2h_sim* sim, const double* u, const double* v, const double* w);
int x3d2h_get_velocity(x3d2h_sim* sim, double* u, double* v, double* w);
/* base_case_t%run loop body (src/case/base_case.f90:246-289), nsteps full time steps, asynchronous */
/* the channel case's per-sub-stage hooks (src/case/channel.f90:59-228) around every following step: bulk-velocity
 * correction (field_volume_integral + field_shift towards 2/3), rotation forcing (omega_rot while the step counter is
 * below n_rotate), wall rows of u, v, w reset from zero boundary fields (inlet_noise = 0) before the pressure correction */
int x3d2h_set_case_channel(x3d2h_sim* sim, double omega_rot, int n_rotate);
int x3d2h_step(x3d2h_sim* sim, int nsteps);
/* n independent batches, one time step each: batch b = upload (u_in, v_in, w_in) -> step -> download into (u_out, v_out,
 * w_out); uploads and downloads run on the backend's copy lanes and overlap the kernels of the neighbouring batches.
 * Page-locked host arrays of the local vertex extents; grids that need padding are refused. Returns after everything
 * has completed. (The same host arrays serve every batch: this is the streaming form of set_velocity / step /
 * get_velocity for benchmarks and ensembles.) */
int x3d2h_step_batches(x3d2h_sim* sim, int n_batches, const double* u_in, const double* v_in, const double* w_in,
                       double* u_out, double* v_out, double* w_out);
int x3d2h_sync(x3d2h_sim* sim);
/* monitoring_t%write_step (src/postprocess/monitoring.f90:46-90): enstrophy, kinetic energy, div_u max, mean */
int x3d2h_monitor(x3d2h_sim* sim, double out[4]);

/* ---- single operators on host data, for parity tests. Inputs/outputs do not touch the solver state. */
int x3d2h_transeq(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* du, double* dv, double* dw);
int x3d2h_transeq_dir(x3d2h_sim* sim, int dir, const double* u, const double* v, const double* w, double* du,
                      double* dv, double* dw);
/* transeq_lowmem (src/solver.f90:391-505); u_back (may be NULL): u after its x -> y -> z -> x round trip */
int x3d2h_transeq_lowmem(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* du, double* dv,
                         double* dw, double* u_back);
/* transeq_species for one scalar (src/solver.f90:507-600) */
int x3d2h_transeq_species(x3d2h_sim* sim, const double* u, const double* v, const double* w, const double* spec,
                          double nu_s, double* dspec);
/* compute_vorticity / compute_qcriterion: what = "vorticity" | "qcriterion"; grads = dudx dudy dudz dvdx ... dwdz */
int x3d2h_derived(x3d2h_sim* sim, const char* what, const double* const* grads, double* out);
int x3d2h_slice_max_sum(x3d2h_sim* sim, int dir, int data_loc, const double* x, int i_slice, double* mx, double* sum);
int x3d2h_tds_solve(x3d2h_sim* sim, int dir, const char* opname, int in_loc, const double* in, double* out, int* out_loc);
/* x3d2c_tds_solve_sum / _dual / _axpy on host data. mode "sum": out_a = A(in_a) + B(in_b); "dual": out_a = A(in_a),
 * out_b = B(in_a); "axpy": out_a = in_b + a A(in_a) (in_b has the output's extents) */
int x3d2h_tds_fused(x3d2h_sim* sim, const char* mode, int dir, const char* op_a, const char* op_b, int in_loc,
                    int out_loc, const double* in_a, const double* in_b, double a, double* out_a, double* out_b);
/* x3d2c_tds_solve_r / _sum_r / _dual_r on host data (mode "single" | "sum" | "dual"); rdr_in / rdr_out: 0 or RDR codes */
int x3d2h_tds_fused_r(x3d2h_sim* sim, const char* mode, int dir, const char* op_a, const char* op_b, int in_loc,
                      int out_loc, int rdr_in, int rdr_out, const double* in_a, const double* in_b, double* out_a,
                      double* out_b);
int x3d2h_divergence(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* div);
int x3d2h_gradient(x3d2h_sim* sim, const double* p, double* gx, double* gy, double* gz);
/* vector_calculus_t%interpl_c2v with the interpl_p2v operators (src/vector_calculus.f90:334-378, src/postprocess/
 * postprocess.f90:184-189): p at the cell centres -> out at the vertices; vector_calculus_t%laplacian with the der2nd
 * operators (src/vector_calculus.f90:380-437): u and out at the vertices */
int x3d2h_interpl_c2v(x3d2h_sim* sim, const double* p, double* out);
int x3d2h_laplacian(x3d2h_sim* sim, const double* u, double* out);
int x3d2h_curl(x3d2h_sim* sim, const double* u, const double* v, const double* w, double* ox, double* oy, double* oz);
int x3d2h_poisson(x3d2h_sim* sim, const double* f, double* p);
int x3d2h_pressure_correction(x3d2h_sim* sim);
int x3d2h_fft_roundtrip(x3d2h_sim* sim, const double* f, double* out, double* spec_re_im);
int x3d2h_reorder_chain(x3d2h_sim* sim, const double* in, const int* rdrs, int n_rdr, double* out);
int x3d2h_sum_intox(x3d2h_sim* sim, int dir_from, const double* a, const double* b, double* out);
int x3d2h_vecadd(x3d2h_sim* sim, int dir, double a, const double* x, double b, const double* y, double* out);
int x3d2h_scalar_product(x3d2h_sim* sim, int dir, int data_loc, const double* x, const double* y, double* s);
int x3d2h_field_max_mean(x3d2h_sim* sim, int dir, int data_loc, const double* x, double* mx, double* mean);

/* remaining elementwise ops / reductions: op = "scale" | "shift" | "vecmult" | "veccopy" | "fill" | "volume_integral" */
int x3d2h_fieldop(x3d2h_sim* sim, const char* op, int dir, int data_loc, double a, const double* x, const double* y,
                  double* out, double* s);

/* ---- device-resident benchmark helpers: fields stay in HBM, only the op is enqueued (bench.py) */
/* op: "transeq_x|y|z", "tds_solve_x|y|z" (der1st), "reorder_x2y|x2z|y2z|z2c|...", "sum_yintox", "vecadd", ... */
int x3d2h_bench_op(x3d2h_sim* sim, const char* op, int reps);

#ifdef __cplusplus
}
#endif
#endif /* X3D2H_H */
