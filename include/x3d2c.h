/* x3d2c.h — C ABI of the `cuda_c` backend: a B200 (sm_100a) implementation of x3d2's per-timestep
 * right-hand-side + pressure hot path, behind the reference's own operator API.
 *
 * Every entry point replaces one deferred procedure of the reference's abstract backend
 *   type, abstract :: base_backend_t      /root/reference/src/backend/backend.f90:13-62
 *   type, abstract :: poisson_fft_t       /root/reference/src/poisson_fft.f90:9-70
 * and is what a Fortran `iso_c_binding` shim (`cuda_c_backend_t`) binds to.
 *
 * Conventions
 *  - All functions return an int status (X3D2C_OK == 0). The reference aborts (`error stop`) on misuse;
 *    the shim turns a non-zero status into `error stop x3d2c_last_error()`.
 *  - Plain pointers and sizes only. `double*` field arguments are DEVICE pointers obtained from
 *    x3d2c_field_alloc; host arrays are named `host_*`.
 *  - One CUDA stream per context; every op is enqueued in call order and is asynchronous unless it
 *    returns a scalar or copies to the host (those synchronise the stream).
 *  - Layout contract: only DIR_C (Cartesian, Fortran order (nx_pad, ny_pad, nz), x fastest) is defined
 *    externally (src/allocator.f90:91). The DIR_X/Y/Z pencil-group layouts are private to the backend
 *    (SURVEY.md F2): nothing outside a backend indexes directional data.
 *  - Padding follows src/allocator.f90:72-76 with sz = X3D2C_SZ.
 */
#ifndef X3D2C_H
#define X3D2C_H

#ifdef __cplusplus
extern "C" {
#endif

#define X3D2C_SZ 32 /* pencil-group width, role of src/backend/cuda/common.f90:4 */

/* status codes */
#define X3D2C_OK 0
#define X3D2C_EINVAL 1
#define X3D2C_ECUDA 2
#define X3D2C_ENCCL 3
#define X3D2C_EUNSUPPORTED 4
#define X3D2C_ENOMEM 5

/* src/common.f90:23-39 */
#define X3D2C_DIR_X 1
#define X3D2C_DIR_Y 2
#define X3D2C_DIR_Z 3
#define X3D2C_DIR_C 4
#define X3D2C_RDR_X2Y 12
#define X3D2C_RDR_X2Z 13
#define X3D2C_RDR_Y2X 21
#define X3D2C_RDR_Y2Z 23
#define X3D2C_RDR_Z2X 31
#define X3D2C_RDR_Z2Y 32
#define X3D2C_RDR_C2X 41
#define X3D2C_RDR_C2Y 42
#define X3D2C_RDR_C2Z 43
#define X3D2C_RDR_X2C 14
#define X3D2C_RDR_Y2C 24
#define X3D2C_RDR_Z2C 34
#define X3D2C_VERT 0
#define X3D2C_CELL 1110
#define X3D2C_X_FACE 1100
#define X3D2C_Y_FACE 1010
#define X3D2C_Z_FACE 110
#define X3D2C_X_EDGE 10
#define X3D2C_Y_EDGE 100
#define X3D2C_Z_EDGE 1000

/* context flags */
#define X3D2C_FLAG_STRICT 1 /* reference-order arithmetic, no FMA contraction: bit-exact vs the OMP backend */

typedef struct x3d2c_ctx x3d2c_ctx;
typedef struct x3d2c_tdsops x3d2c_tdsops;
typedef struct x3d2c_poisson x3d2c_poisson;

/* What cuda_backend_t%init receives through mesh_t + allocator_t
 * (src/backend/cuda/backend.f90:95-152, src/mesh.f90:37-158, src/mesh_content.f90:28-59). */
typedef struct {
  int dims_vert[3];        /* mesh%get_dims(VERT): local vertex counts            */
  int dims_cell[3];        /* mesh%get_dims(CELL)                                  */
  int dims_vert_global[3]; /* mesh%get_global_dims(VERT)                           */
  int dims_cell_global[3]; /* mesh%get_global_dims(CELL)                           */
  int nproc_dir[3];        /* mesh%par%nproc_dir                                   */
  int nrank_dir[3];        /* mesh%par%nrank_dir                                   */
  int n_offset[3];         /* mesh%par%n_offset                                    */
  int pprev[3], pnext[3];  /* mesh%par%pprev / pnext (ranks of the communicator)   */
  int periodic[3];         /* mesh%grid%periodic_BC                                */
  int sz;                  /* must be X3D2C_SZ; the Fortran allocator pads with it */
  int rank, nproc;         /* mesh%par%nrank / nproc                               */
  int device;              /* CUDA device ordinal; -1 keeps the current device     */
  int flags;               /* X3D2C_FLAG_*                                         */
  const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks; NULL when nproc == 1 */
} x3d2c_config;

const char* x3d2c_last_error(void);
int x3d2c_version(void);
/* fills 128 bytes with a fresh ncclUniqueId (rank 0 calls this and broadcasts the bytes to all ranks) */
int x3d2c_nccl_unique_id(void* out128);

/* ---- backend object: cuda_backend_t%init / finaliser (src/backend/cuda/backend.f90:95-152) */
int x3d2c_create(const x3d2c_config* cfg, x3d2c_ctx** out);
int x3d2c_destroy(x3d2c_ctx* ctx);
int x3d2c_sync(x3d2c_ctx* ctx);
/* padded Cartesian dims, ngrid = product (src/allocator.f90:64-93); n_groups per DIR_X/Y/Z */
int x3d2c_get_padded_dims(const x3d2c_ctx* ctx, int dims_padded[3], int n_groups[3], long long* ngrid);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long x3d2c_launch_count(const x3d2c_ctx* ctx);
/* the context's cudaStream_t as an opaque pointer (for event timing on the launching stream) */
void* x3d2c_stream(const x3d2c_ctx* ctx);

/* ---- field storage: cuda_allocator_t%create_block / cuda_field_t%fill
 *      (src/backend/cuda/allocator.f90:44-90). Pool semantics (get_block/release_block) stay with the caller. */
int x3d2c_field_alloc(x3d2c_ctx* ctx, double** dev);
int x3d2c_field_free(x3d2c_ctx* ctx, double* dev);
int x3d2c_field_fill(x3d2c_ctx* ctx, double* dev, double c);
/* copy_data_to_f / copy_f_to_data (src/backend/backend.f90:327-349): whole padded block, ngrid doubles */
int x3d2c_copy_data_to_f(x3d2c_ctx* ctx, double* dev, const double* host_data);
int x3d2c_copy_f_to_data(x3d2c_ctx* ctx, double* host_data, const double* dev);
/* extension - I/O lanes: the same copies, asynchronous, on dedicated copy streams that are ordered against the
 * context's stream with events, so that a caller can stream independent batches through the device (the upload of
 * batch b + 1 and the download of batch b - 1 overlap the kernels of batch b). lane: 0 = the context's stream,
 * 1 = upload stream, 2 = download stream; ev: 0..15. The host arrays must be page-locked and stay valid until the lane
 * has been synchronised. x3d2c_lane_wait on an event that was never recorded is a no-op. */
int x3d2c_copy_data_to_f_async(x3d2c_ctx* ctx, double* dev, const double* host_pinned, int lane);
int x3d2c_copy_f_to_data_async(x3d2c_ctx* ctx, double* host_pinned, const double* dev, int lane);
int x3d2c_lane_record(x3d2c_ctx* ctx, int lane, int ev);
int x3d2c_lane_wait(x3d2c_ctx* ctx, int lane, int ev);
int x3d2c_lane_sync(x3d2c_ctx* ctx, int lane);

/* ---- alloc_tdsops (src/backend/backend.f90:351-372; upload pattern of src/backend/cuda/tdsops.f90:31-90).
 * The host computes the tables with tdsops_init (src/tdsops.f90:63-203) and passes them in:
 *   coeffs[9]; coeffs_s / coeffs_e as [row 0..3][tap 0..8] (= Fortran coeffs_s(tap, row));
 *   dist_fw/bw/sa/sc/af with n_rhs entries; stretch / stretch_correct with n_tds entries. */
int x3d2c_tdsops_create(x3d2c_ctx* ctx, int n_tds, int n_rhs, int move, int periodic, const double* coeffs,
                        const double* coeffs_s, const double* coeffs_e, const double* dist_fw,
                        const double* dist_bw, const double* dist_sa, const double* dist_sc,
                        const double* dist_af, const double* stretch, const double* stretch_correct,
                        x3d2c_tdsops** out);
int x3d2c_tdsops_destroy(x3d2c_ctx* ctx, x3d2c_tdsops* ops);

/* ---- transeq_x / transeq_y / transeq_z (src/backend/backend.f90:64-86, cuda/backend.f90:244-319).
 * du, dv, dw, u, v, w are the caller's fields in natural order; `dir` selects which of them is the
 * line-aligned velocity (the permutation of src/backend/omp/backend.f90:154,168,182 is done here). */
int x3d2c_transeq(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
                  const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
                  const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym);
/* extension: reorder(u, v, w with rdr_in) followed by transeq in `dir` (transeq_default reorders the velocity before the
 * y and z sweeps, src/solver.f90:325-327,355-357); rdr_in = 0 is x3d2c_transeq. The fast path lets the y sweep read the
 * x-layout velocity directly. u, v, w are in the layout rdr_in starts from, du, dv, dw in DIR_`dir`. */
int x3d2c_transeq_r(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
                    const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
                    const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym, int rdr_in);
/* 1 when x3d2c_transeq_r with these arguments runs without a reorder pass (the caller may then skip producing the
 * reordered copies another way), 0 otherwise */
int x3d2c_transeq_r_fused(x3d2c_ctx* ctx, int dir, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
                          const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym, int rdr_in);

/* ---- tds_solve (src/backend/backend.f90:108-126, cuda/backend.f90:449-521). data_loc bookkeeping
 * (move_data_loc) stays on the caller's field_t. */
int x3d2c_tds_solve(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops);

/* ---- fused tds_solve combinations (extensions: no counterpart in base_backend_t). Each equals the sequence of
 * reference calls given below - and is executed as exactly that sequence in strict mode or when the operators
 * do not qualify for the fast path - but moves 24 B per point instead of 56 / 32 / 40:
 *   sum : tds_solve(out, in_a, op_a); tds_solve(tmp, in_b, op_b); vecadd(1, tmp, 1, out)
 *         (divergence_v2c, src/vector_calculus.f90:185-214)
 *   dual: tds_solve(out_a, in, op_a); tds_solve(out_b, in, op_b)   (gradient_c2v, src/vector_calculus.f90:275-300)
 *   axpy: tds_solve(tmp, in, op); vecadd(a, tmp, 1, y)             (pressure correction, src/solver.f90:279-301) */
int x3d2c_tds_solve_sum(x3d2c_ctx* ctx, int dir, double* out, const double* in_a, const x3d2c_tdsops* op_a,
                        const double* in_b, const x3d2c_tdsops* op_b);
int x3d2c_tds_solve_dual(x3d2c_ctx* ctx, int dir, double* out_a, double* out_b, const double* in,
                         const x3d2c_tdsops* op_a, const x3d2c_tdsops* op_b);
int x3d2c_tds_solve_axpy(x3d2c_ctx* ctx, int dir, double* y, double a, const double* in, const x3d2c_tdsops* op);
/* ... and through a reorder: reorder(in, rdr_in) -> operator(s) in `dir` -> reorder(out, rdr_out); rdr = 0 means the
 * field already is / stays in the DIR_`dir` layout, otherwise rdr_in must end in `dir` (e.g. X3D2C_RDR_C2Z for dir = Z)
 * and rdr_out must start from it (e.g. X3D2C_RDR_Y2Z for dir = Y). The fast path addresses the foreign layout from
 * inside the kernel (Y / Z lines, layouts Y, Z, C); otherwise the three steps run one after the other. Removes the
 * y2z, z2c, c2z and z2y reorder passes of divergence_v2c / poisson_fft / gradient_c2v
 * (src/vector_calculus.f90:200-203,275-285, src/solver.f90:755-772). */
int x3d2c_tds_solve_r(x3d2c_ctx* ctx, int dir, double* out, const double* in, const x3d2c_tdsops* op, int rdr_in,
                      int rdr_out);
int x3d2c_tds_solve_sum_r(x3d2c_ctx* ctx, int dir, double* out, const double* in_a, const x3d2c_tdsops* op_a,
                          const double* in_b, const x3d2c_tdsops* op_b, int rdr_in, int rdr_out);
int x3d2c_tds_solve_dual_r(x3d2c_ctx* ctx, int dir, double* out_a, double* out_b, const double* in,
                           const x3d2c_tdsops* op_a, const x3d2c_tdsops* op_b, int rdr_in, int rdr_out);
/* axpy through an input reorder: reorder(tmp, in, rdr_in); tds_solve(tmp2, tmp, op); vecadd(a, tmp2, 1, y)
 * (gradient_c2v's y2x reorders followed by the pressure correction, src/vector_calculus.f90:302-330 + src/solver.f90:296-298) */
int x3d2c_tds_solve_axpy_r(x3d2c_ctx* ctx, int dir, double* y, double a, const double* in, const x3d2c_tdsops* op,
                           int rdr_in);

/* ---- reorder (src/backend/backend.f90:128-144), rdr is one of X3D2C_RDR_* */
int x3d2c_reorder(x3d2c_ctx* ctx, int rdr, double* dst, const double* src);
/* extension: reorder(dst_y, src, RDR_X2Y) and reorder(dst_z, src, RDR_X2Z) with one read of src */
int x3d2c_reorder_x2yz(x3d2c_ctx* ctx, double* dst_y, double* dst_z, const double* src);
/* ---- sum_yintox / sum_zintox (src/backend/backend.f90:146-159): u (DIR_X) += reorder(u_) */
int x3d2c_sum_yintox(x3d2c_ctx* ctx, double* u, const double* u_y);
int x3d2c_sum_zintox(x3d2c_ctx* ctx, double* u, const double* u_z);
/* extension: sum_yintox(u, u_y) followed by sum_zintox(u, u_z) in one pass over u (same order of additions) */
int x3d2c_sum_yzintox(x3d2c_ctx* ctx, double* u, const double* u_y, const double* u_z);
/* extension: x3d2c_sum_yzintox(u, u_y, u_z) followed by x3d2c_veclincomb(out, base, [x..., u], [coef..., c_u]) in one
 * pass (the end of transeq + the Runge-Kutta update, src/solver.f90:340-372 + src/time_integrator.f90:166-231); n <= 3.
 * store_u == 0: u need not hold the sum on return (the last stage does not keep the derivative). Strict mode runs the
 * two calls. */
int x3d2c_sum_yzintox_lincomb(x3d2c_ctx* ctx, double* u, const double* u_y, const double* u_z, int store_u, double* out,
                              const double* base, int n, const double* coef, const double* const* x, double c_u);

/* ---- veccopy / vecadd / vecmult (src/backend/backend.f90:161-201): whole padded block */
int x3d2c_veccopy(x3d2c_ctx* ctx, double* dst, const double* src);
int x3d2c_vecadd(x3d2c_ctx* ctx, double a, const double* x, double b, double* y);
int x3d2c_vecmult(x3d2c_ctx* ctx, double* y, const double* x);
/* extension: out = base, then out = coef[k] * x[k] + 1.0 * out for k = 0..n-1 (n <= 4) in one pass; the same values as
 * veccopy(out, base) followed by vecadd(coef[k], x[k], 1.0, out) (time integrators, src/time_integrator.f90:166-231).
 * out may alias base, not a term. */
int x3d2c_veclincomb(x3d2c_ctx* ctx, double* out, const double* base, int n, const double* coef,
                     const double* const* x);
/* ---- field_scale / field_shift (src/backend/backend.f90:255-266) */
int x3d2c_field_scale(x3d2c_ctx* ctx, double* f, double a);
int x3d2c_field_shift(x3d2c_ctx* ctx, double* f, double a);

/* ---- scalar_product (src/backend/backend.f90:203-216): global (all-reduced) sum over the un-padded
 * entries of mesh%get_dims(data_loc); x and y share `dir`. Synchronous. */
int x3d2c_scalar_product(x3d2c_ctx* ctx, int dir, int data_loc, const double* x, const double* y, double* s);
/* ---- field_max_mean (src/backend/backend.f90:218-236): max|f| and sum|f| / N_global. Synchronous. */
int x3d2c_field_max_mean(x3d2c_ctx* ctx, int dir, int data_loc, const double* f, double* max_val, double* mean_val);
/* ---- field_volume_integral (src/backend/backend.f90:268-279): DIR_X only. Synchronous. */
int x3d2c_field_volume_integral(x3d2c_ctx* ctx, int data_loc, const double* f, double* s);
/* ---- field_set_face / field_set_face_from_field (src/backend/backend.f90:268-308, omp/backend.f90:903-1021): DIR_X
 * fields only, `data_loc` gives the extents. set_face: Y_FACE sets the rows y = 1 and y = ny to c_start / c_end (X_FACE,
 * Z_FACE: "not yet supported", as the reference; the top row is the one the reference's CUDA backend and the TODO at
 * omp/backend.f90:939 intend). from_field: Y_FACE copies both rows from f_start; X_FACE copies the inlet column x = 1
 * from f_start and applies the convective outflow f(nx) = f(nx) - c_end (f(nx) - f(nx-1)) + flow_rate_diff. */
int x3d2c_field_set_face(x3d2c_ctx* ctx, double* f, int data_loc, double c_start, double c_end, int face);
int x3d2c_field_set_face_from_field(x3d2c_ctx* ctx, double* f, const double* f_start, int data_loc, double c_end,
                                    int face, double flow_rate_diff);

/* ---- init_poisson_fft (src/backend/backend.f90:374-391) and the poisson_fft_t hooks
 * (src/poisson_fft.f90:45-62,72-116). The spectral buffer is hidden state of the handle between
 * fft_forward / fft_postprocess / fft_backward; both transforms are unnormalised
 * (tests/verification/test_fft.f90:162-166). */
/* spectral pencil owned by this rank: extents and offsets passed to poisson_fft_t%base_init */
int x3d2c_poisson_spec_layout(const x3d2c_ctx* ctx, int n_spec[3], int n_sp_st[3]);
/* waves: complex(dp) waves(nx_spec, ny_spec, nz_spec) as interleaved re/im (src/poisson_fft.f90:23,163);
 * ax..bz: the global wave-number tables of src/poisson_fft.f90:148-150 (nx_glob / ny_glob / nz_glob entries) */
int x3d2c_poisson_create(x3d2c_ctx* ctx, const double* waves, const double* ax, const double* bx,
                         const double* ay, const double* by, const double* az, const double* bz,
                         x3d2c_poisson** out);
int x3d2c_poisson_destroy(x3d2c_ctx* ctx, x3d2c_poisson* p);
int x3d2c_fft_forward(x3d2c_ctx* ctx, x3d2c_poisson* p, const double* f_c);  /* f_c: DIR_C block */
int x3d2c_fft_postprocess_000(x3d2c_ctx* ctx, x3d2c_poisson* p);
int x3d2c_fft_backward(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_c);
/* ---- non-periodic y (poisson_010, src/poisson_fft.f90:228-242; single rank, as the reference requires :178-180).
 * fft_forward_010 / fft_backward_010 are x3d2c_fft_forward / x3d2c_fft_backward (as in cuda/poisson_fft.f90:78-83).
 * stretched: 0 uniform mesh (waves table), 1 'centred' / 'top-bottom' (a_odd_*, a_even_*: (nx_spec, ny_spec / 2, nz_spec, 5)
 * Fortran order, src/poisson_fft.f90:423-650), 2 'bottom' (a_odd_* = a_re / a_im: (nx_spec, ny_spec, nz_spec, 5)).
 * The tensors are factorised ONCE here; fft_postprocess_010 never modifies or re-copies them (the reference's kernels
 * eliminate in place and cuda/poisson_fft.f90:870-895 restores the tensors before every solve). */
int x3d2c_poisson_create_010(x3d2c_ctx* ctx, const double* waves, const double* ax, const double* bx, const double* ay,
                             const double* by, const double* az, const double* bz, int stretched,
                             const double* a_odd_re, const double* a_odd_im, const double* a_even_re,
                             const double* a_even_im, x3d2c_poisson** out);
/* fft_postprocess_010 (cuda/poisson_fft.f90:822-924; omp/kernels/spectral_processing.f90:108-283 on a uniform mesh) */
int x3d2c_fft_postprocess_010(x3d2c_ctx* ctx, x3d2c_poisson* p);
/* enforce_periodicity_y / undo_periodicity_y (src/poisson_fft.f90:57-58; omp/poisson_fft.f90:237-285): the even / odd
 * reshuffle in y that turns the cosine transform into an FFT of the same length. DIR_C blocks, f_out != f_in. */
int x3d2c_enforce_periodicity_y(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in);
int x3d2c_undo_periodicity_y(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in);
/* ---- non-periodic x (100 / 110; src/poisson_fft.f90:47-62): no BASELINE.json configuration has walls in x; the
 * reference's own OMP backend stops with 'does not support ...' for these hooks (omp/poisson_fft.f90:99-127,183-235).
 * Exported so that a Fortran extends(poisson_fft_t) can bind every deferred procedure; they return X3D2C_EUNSUPPORTED. */
int x3d2c_fft_forward_100(x3d2c_ctx* ctx, x3d2c_poisson* p, const double* f_c);
int x3d2c_fft_forward_110(x3d2c_ctx* ctx, x3d2c_poisson* p, const double* f_c);
int x3d2c_fft_backward_100(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_c);
int x3d2c_fft_backward_110(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_c);
int x3d2c_fft_postprocess_100(x3d2c_ctx* ctx, x3d2c_poisson* p);
int x3d2c_fft_postprocess_110(x3d2c_ctx* ctx, x3d2c_poisson* p);
int x3d2c_enforce_periodicity_x(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in);
int x3d2c_undo_periodicity_x(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in);
int x3d2c_enforce_periodicity_xy(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in);
int x3d2c_undo_periodicity_xy(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in);
/* debugging / tests: copy the spectral buffer (reference index order (i, j, k), interleaved re/im) to the host */
int x3d2c_poisson_get_spectrum(x3d2c_ctx* ctx, x3d2c_poisson* p, double* host_spec);

/* ---- transeq_species (src/backend/backend.f90:88-106, omp/backend.f90:186-233): one scalar `spec` advected by the
 * line-aligned velocity `uvw`; der1st / der1st_sym / der2nd of dirps; sync != 0 also exchanges the velocity halos */
int x3d2c_transeq_species(x3d2c_ctx* ctx, int dir, double* dspec, const double* uvw, const double* spec, double nu,
                          const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym, const x3d2c_tdsops* der2nd,
                          int sync);

/* ---- slice_max_sum (src/backend/backend.f90:238-253, omp/backend.f90:816-881): signed max and signed sum of the
 * plane i_slice (1-based) along the field's own direction; rank-local, the caller reduces across ranks. Synchronous. */
int x3d2c_slice_max_sum(x3d2c_ctx* ctx, int dir, int data_loc, const double* f, int i_slice, double* max_val,
                        double* sum_val);

/* ---- compute_vorticity / compute_qcriterion (src/backend/backend.f90:310-325, omp/backend.f90:616-649): pointwise
 * |curl u| and Q = -1/2 (dudx^2 + dvdy^2 + dwdz^2) - dudy dvdx - dudz dwdx - dvdz dwdy over whole padded blocks */
int x3d2c_compute_vorticity(x3d2c_ctx* ctx, double* field_out, const double* dudx, const double* dudy,
                            const double* dudz, const double* dvdx, const double* dvdy, const double* dvdz,
                            const double* dwdx, const double* dwdy, const double* dwdz);
int x3d2c_compute_qcriterion(x3d2c_ctx* ctx, double* field_out, const double* dudx, const double* dudy,
                             const double* dudz, const double* dvdx, const double* dvdy, const double* dvdz,
                             const double* dwdx, const double* dwdy, const double* dwdz);

#ifdef __cplusplus
}
#endif
#endif /* X3D2C_H */
