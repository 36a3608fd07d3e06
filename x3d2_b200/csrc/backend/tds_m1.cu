// DistD2-TDS "reference-order" kernels (thread per line, lane = line inside a 32-wide pencil group).
//
// These follow the arithmetic of the reference OMP kernels statement by statement
//   der_univ_dist        src/backend/omp/kernels/distributed.f90:11-168
//   der_univ_subs        src/backend/omp/kernels/distributed.f90:170-229
//   der_univ_fused_subs  src/backend/omp/kernels/distributed.f90:231-337
//   exec_dist_tds_compact / exec_dist_transeq_compact   src/backend/omp/exec_dist.f90:16-186
// In STRICT mode every product and sum is a separately rounded __dmul_rn/__dadd_rn in source order, so
// the result is bit-identical to the strict-IEEE oracle. In the default mode the same statements are
// evaluated with FMA contraction and exactly-zero stencil taps are skipped.
// The intermediate sweeps live in the output arrays (as in the reference). These kernels serve strict mode and
// every operator the fast path does not cover (non-periodic boundaries, stretched meshes, short lines); periodic
// uniform directions use the segment-parallel kernels of *_m4.cu (TMA tiles) and *_m3.cu (cp.async tiles).
// This file also holds the C entry points x3d2c_tds_solve / _sum / _dual / _axpy / x3d2c_transeq and their dispatch.
#include "common.cuh"

// The reference-order kernels always use the reference's rounding (no FMA contraction), also outside strict mode:
// they serve the operators the segment kernels do not take (walls and stretched meshes in transeq), and on smooth
// wall-bounded data a contracted second-derivative stencil differs from the reference by up to 7e-12 (SURVEY.md F4;
// tests/test_gpu_nonperiodic.py::test_generic_kernels_walls_and_stretching). They are latency-bound, not FP64-bound.
constexpr bool kM1AlwaysExact = true;

namespace {

template <bool S>
struct Ar {
  static __device__ __forceinline__ double mul(double a, double b) { return S ? __dmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ double add(double a, double b) { return S ? __dadd_rn(a, b) : a + b; }
  // c + a*b
  static __device__ __forceinline__ double mad(double a, double b, double c) {
    return S ? __dadd_rn(c, __dmul_rn(a, b)) : fma(a, b, c);
  }
  // c - a*b
  static __device__ __forceinline__ double nmad(double a, double b, double c) {
    return S ? __dadd_rn(c, -__dmul_rn(a, b)) : fma(-a, b, c);
  }
};

// sum_{k} c[k] * w[k], left to right as written in the Fortran source
template <bool S>
__device__ __forceinline__ double stencil9(const double* __restrict__ c, const double (&w)[9], unsigned mask) {
  if (S) {
    double t = __dmul_rn(c[0], w[0]);
#pragma unroll
    for (int k = 1; k < 9; ++k) t = __dadd_rn(t, __dmul_rn(c[k], w[k]));
    return t;
  } else {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (mask & (1u << k)) t = fma(c[k], w[k], t);
    return t;
  }
}

struct LineIO {
  const double* u;       // line base (lane offset applied), row stride SZ
  const double* halo_s;  // (4 rows) or nullptr => periodic wrap onto the line itself
  const double* halo_e;
  int n_tds, n_rhs;
  // extended line e(k), k in [-3, n_rhs + 4]  (k is the 1-based Fortran row index)
  __device__ __forceinline__ double at(int k) const {
    if (k >= 1 && k <= n_rhs) return u[(size_t)(k - 1) * SZ];
    if (k < 1) {  // u_s(k + 4)
      if (halo_s) return halo_s[(k + 3) * SZ];
      return u[(size_t)(n_tds + k - 1) * SZ];  // copy_into_buffers: u_send_e(j) = u(n - 4 + j)
    }
    const int r = k - n_rhs;  // u_e(r)
    if (halo_e) return halo_e[(r - 1) * SZ];
    return u[(size_t)(r - 1) * SZ];
  }
};

enum { PH_DIST = 1, PH_SUBS = 2, PH_ALL = 3 };

// ------------------------------------------------------------------------------------------ tds_solve
template <bool S>
__global__ void __launch_bounds__(128)
tds_m1_kernel(double* __restrict__ du, const double* __restrict__ u, const double* __restrict__ halo_s,
              const double* __restrict__ halo_e, double* __restrict__ send_s, double* __restrict__ send_e,
              const double* __restrict__ recv_s, const double* __restrict__ recv_e,
              const __grid_constant__ TdsDev ops,
              const unsigned mask, const int n_pad, const int n_groups, const int phase) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n_groups) return;
  const size_t base = (size_t)SZ * n_pad * g + lane;
  double* d = du + base;
  const int n = ops.n_tds, n_rhs = ops.n_rhs;
  const size_t hrow = (size_t)g * SZ + lane;  // 1-row buffers (SZ, 1, G)
  double z_n = 0.0, y_1 = 0.0;

  if (phase & PH_DIST) {
    LineIO io{u + base, halo_s ? halo_s + (size_t)g * 4 * SZ + lane : nullptr,
              halo_e ? halo_e + (size_t)g * 4 * SZ + lane : nullptr, n, n_rhs};
    double w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = io.at(k - 3);  // window of row j = 1: e(-3 .. 5)
    double prev;
    // rows 1..4 (distributed.f90:37-80)
    {
      double t = stencil9<S>(ops.coeffs_s[0], w, 0x1ff);
      prev = Ar<S>::mul(t, ops.af[0]);
      d[0] = prev;
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = w[k + 1];
      w[8] = io.at(6);
      t = stencil9<S>(ops.coeffs_s[1], w, 0x1ff);
      prev = Ar<S>::mul(t, ops.af[1]);
      d[SZ] = prev;
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = w[k + 1];
      w[8] = io.at(7);
      t = stencil9<S>(ops.coeffs_s[2], w, 0x1ff);
      prev = Ar<S>::mul(ops.fw[2], Ar<S>::nmad(ops.af[2], prev, t));
      d[2 * SZ] = prev;
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = w[k + 1];
      w[8] = io.at(8);
      t = stencil9<S>(ops.coeffs_s[3], w, 0x1ff);
      prev = Ar<S>::mul(ops.fw[3], Ar<S>::nmad(ops.af[3], prev, t));
      d[3 * SZ] = prev;
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = w[k + 1];
      w[8] = io.at(9);
    }
    // bulk (distributed.f90:82-96); all taps of rows 5 .. n_rhs-4 are interior points
    const double alpha = ops.af[4];
    const double* up = u + base;
    for (int j = 5; j <= n_rhs - 4; ++j) {
      double t = stencil9<S>(ops.coeffs, w, mask);
      prev = Ar<S>::mul(ops.fw[j - 1], Ar<S>::nmad(alpha, prev, t));
      d[(size_t)(j - 1) * SZ] = prev;
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = w[k + 1];
      w[8] = (j + 5 <= n_rhs) ? up[(size_t)(j + 4) * SZ] : io.at(j + 5);
    }
    // last four rows (distributed.f90:98-145)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = n_rhs - 3 + r;
      double t = stencil9<S>(ops.coeffs_e[r], w, 0x1ff);
      prev = Ar<S>::mul(ops.fw[j - 1], Ar<S>::nmad(ops.af[j - 1], prev, t));
      d[(size_t)(j - 1) * SZ] = prev;
      if (r < 3) {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = w[k + 1];
        w[8] = io.at(j + 5);
      }
    }
    z_n = d[(size_t)(n - 1) * SZ];  // send_u_e (distributed.f90:147-151)
    // backward pass (distributed.f90:153-166)
    double nxt = d[(size_t)(n - 2) * SZ];
    for (int j = n - 2; j >= 2; --j) {
      nxt = Ar<S>::nmad(ops.bw[j - 1], nxt, d[(size_t)(j - 1) * SZ]);
      d[(size_t)(j - 1) * SZ] = nxt;
    }
    y_1 = Ar<S>::mul(ops.fw[0], Ar<S>::nmad(ops.bw[0], nxt, d[0]));
    d[0] = y_1;
    if (phase != PH_ALL) {
      send_e[hrow] = z_n;
      send_s[hrow] = y_1;
    }
  }

  if (phase & PH_SUBS) {
    double rs, re;
    if (phase == PH_ALL) {  // nproc == 1: recv_s = send_e, recv_e = send_s (omp/sendrecv.f90:20-22)
      rs = z_n;
      re = y_1;
    } else {
      rs = recv_s[hrow];
      re = recv_e[hrow];
      y_1 = d[0];
    }
    // distributed.f90:184-208
    const double sa1 = ops.sa[0], scn = ops.sc[n - 1];
    const double recp_s = S ? __ddiv_rn(1.0, __dadd_rn(1.0, -__dmul_rn(sa1, sa1))) : 1.0 / (1.0 - sa1 * sa1);
    const double recp_e = S ? __ddiv_rn(1.0, __dadd_rn(1.0, -__dmul_rn(scn, scn))) : 1.0 / (1.0 - scn * scn);
    const double y_n = d[(size_t)(n - 1) * SZ];
    const double s = Ar<S>::mul(recp_s, Ar<S>::nmad(sa1, rs, y_1));
    const double e = Ar<S>::mul(recp_e, Ar<S>::nmad(scn, re, y_n));
    // distributed.f90:210-227
    d[0] = Ar<S>::mul(s, ops.stretch[0]);
    for (int j = 2; j <= n - 1; ++j) {
      double v = d[(size_t)(j - 1) * SZ];
      v = Ar<S>::nmad(ops.sa[j - 1], s, v);
      v = Ar<S>::nmad(ops.sc[j - 1], e, v);
      d[(size_t)(j - 1) * SZ] = Ar<S>::mul(v, ops.stretch[j - 1]);
    }
    d[(size_t)(n - 1) * SZ] = Ar<S>::mul(e, ops.stretch[n - 1]);
  }
}

// ------------------------------------------------------------------------------------------ transeq, one component
// rhs = -1/2 (conv * du + d(u*conv)) + nu (d2u + du * stretch_correct)      (omp/backend.f90:299-338)
template <bool S>
__global__ void __launch_bounds__(128)
transeq_m1_kernel(double* __restrict__ rhs, double* __restrict__ dud, double* __restrict__ d2u,
                  const double* __restrict__ u, const double* __restrict__ conv,
                  const double* __restrict__ u_halo_s, const double* __restrict__ u_halo_e,
                  const double* __restrict__ c_halo_s, const double* __restrict__ c_halo_e,
                  double* __restrict__ send, const double* __restrict__ recv, const size_t row_stride,
                  const __grid_constant__ TdsDev o_du, const __grid_constant__ TdsDev o_dud,
                  const __grid_constant__ TdsDev o_d2u, const unsigned m_du,
                  const unsigned m_dud, const unsigned m_d2u, const double nu, const int n_pad,
                  const int n_groups, const int phase) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n_groups) return;
  const size_t base = (size_t)SZ * n_pad * g + lane;
  double* a = rhs + base;
  double* b = dud + base;
  double* c = d2u + base;
  const int n = o_du.n_tds, n_rhs = o_du.n_rhs;
  const size_t hrow = (size_t)g * SZ + lane;
  // send/recv hold six 1-row buffers each: [du_s, du_e, dud_s, dud_e, d2u_s, d2u_e] x row_stride
  double zn[3] = {0, 0, 0}, y1[3] = {0, 0, 0};

  if (phase & PH_DIST) {
    LineIO iu{u + base, u_halo_s ? u_halo_s + (size_t)g * 4 * SZ + lane : nullptr,
              u_halo_e ? u_halo_e + (size_t)g * 4 * SZ + lane : nullptr, n, n_rhs};
    LineIO ic{conv + base, c_halo_s ? c_halo_s + (size_t)g * 4 * SZ + lane : nullptr,
              c_halo_e ? c_halo_e + (size_t)g * 4 * SZ + lane : nullptr, n, n_rhs};
    double wu[9], wp[9];  // u window and (u * conv) window (exec_dist.f90:133-149)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      wu[k] = iu.at(k - 3);
      wp[k] = Ar<S>::mul(wu[k], ic.at(k - 3));
    }
    double p1 = 0, p2 = 0, p3 = 0;
    for (int j = 1; j <= n_rhs; ++j) {
      double t1, t2, t3;
      if (j <= 4) {
        t1 = stencil9<S>(o_du.coeffs_s[j - 1], wu, 0x1ff);
        t2 = stencil9<S>(o_dud.coeffs_s[j - 1], wp, 0x1ff);
        t3 = stencil9<S>(o_d2u.coeffs_s[j - 1], wu, 0x1ff);
      } else if (j <= n_rhs - 4) {
        t1 = stencil9<S>(o_du.coeffs, wu, m_du);
        t2 = stencil9<S>(o_dud.coeffs, wp, m_dud);
        t3 = stencil9<S>(o_d2u.coeffs, wu, m_d2u);
      } else {
        const int r = j - (n_rhs - 3);
        t1 = stencil9<S>(o_du.coeffs_e[r], wu, 0x1ff);
        t2 = stencil9<S>(o_dud.coeffs_e[r], wp, 0x1ff);
        t3 = stencil9<S>(o_d2u.coeffs_e[r], wu, 0x1ff);
      }
      if (j <= 2) {
        p1 = Ar<S>::mul(t1, o_du.af[j - 1]);
        p2 = Ar<S>::mul(t2, o_dud.af[j - 1]);
        p3 = Ar<S>::mul(t3, o_d2u.af[j - 1]);
      } else {
        // rows 5 .. n_rhs-4 use alpha = faf(5) in the source; faf is constant over the bulk
        const int ja = (j >= 5 && j <= n_rhs - 4) ? 4 : j - 1;
        p1 = Ar<S>::mul(o_du.fw[j - 1], Ar<S>::nmad(o_du.af[ja], p1, t1));
        p2 = Ar<S>::mul(o_dud.fw[j - 1], Ar<S>::nmad(o_dud.af[ja], p2, t2));
        p3 = Ar<S>::mul(o_d2u.fw[j - 1], Ar<S>::nmad(o_d2u.af[ja], p3, t3));
      }
      const size_t off = (size_t)(j - 1) * SZ;
      a[off] = p1; b[off] = p2; c[off] = p3;
      if (j < n_rhs) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { wu[k] = wu[k + 1]; wp[k] = wp[k + 1]; }
        wu[8] = iu.at(j + 5);
        wp[8] = Ar<S>::mul(wu[8], ic.at(j + 5));
      }
    }
    zn[0] = a[(size_t)(n - 1) * SZ]; zn[1] = b[(size_t)(n - 1) * SZ]; zn[2] = c[(size_t)(n - 1) * SZ];
    double q1 = a[(size_t)(n - 2) * SZ], q2 = b[(size_t)(n - 2) * SZ], q3 = c[(size_t)(n - 2) * SZ];
    for (int j = n - 2; j >= 2; --j) {
      const size_t off = (size_t)(j - 1) * SZ;
      q1 = Ar<S>::nmad(o_du.bw[j - 1], q1, a[off]);
      q2 = Ar<S>::nmad(o_dud.bw[j - 1], q2, b[off]);
      q3 = Ar<S>::nmad(o_d2u.bw[j - 1], q3, c[off]);
      a[off] = q1; b[off] = q2; c[off] = q3;
    }
    y1[0] = Ar<S>::mul(o_du.fw[0], Ar<S>::nmad(o_du.bw[0], q1, a[0]));
    y1[1] = Ar<S>::mul(o_dud.fw[0], Ar<S>::nmad(o_dud.bw[0], q2, b[0]));
    y1[2] = Ar<S>::mul(o_d2u.fw[0], Ar<S>::nmad(o_d2u.bw[0], q3, c[0]));
    a[0] = y1[0]; b[0] = y1[1]; c[0] = y1[2];
    if (phase != PH_ALL) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        send[(2 * q) * row_stride + hrow] = y1[q];      // *_send_s
        send[(2 * q + 1) * row_stride + hrow] = zn[q];  // *_send_e
      }
    }
  }

  if (phase & PH_SUBS) {
    double rs[3], re[3];
    if (phase == PH_ALL) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { rs[q] = zn[q]; re[q] = y1[q]; }
    } else {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        rs[q] = recv[(2 * q) * row_stride + hrow];      // *_recv_s
        re[q] = recv[(2 * q + 1) * row_stride + hrow];  // *_recv_e
      }
      y1[0] = a[0]; y1[1] = b[0]; y1[2] = c[0];
    }
    const TdsDev* o[3] = {&o_du, &o_dud, &o_d2u};
    double s[3], e[3];
    const double yn[3] = {a[(size_t)(n - 1) * SZ], b[(size_t)(n - 1) * SZ], c[(size_t)(n - 1) * SZ]};
#pragma unroll
    for (int q = 0; q < 3; ++q) {  // distributed.f90:259-303
      const double sa1 = o[q]->sa[0], scn = o[q]->sc[n - 1];
      const double recp_s = S ? __ddiv_rn(1.0, __dadd_rn(1.0, -__dmul_rn(sa1, sa1))) : 1.0 / (1.0 - sa1 * sa1);
      const double recp_e = S ? __ddiv_rn(1.0, __dadd_rn(1.0, -__dmul_rn(scn, scn))) : 1.0 / (1.0 - scn * scn);
      s[q] = Ar<S>::mul(recp_s, Ar<S>::nmad(sa1, rs[q], y1[q]));
      e[q] = Ar<S>::mul(recp_e, Ar<S>::nmad(scn, re[q], yn[q]));
    }
    const double* cv = conv + base;
    {  // row 1 (distributed.f90:305-312)
      const double dus = Ar<S>::mul(s[0], o_du.stretch[0]);
      double t = Ar<S>::mul(Ar<S>::mul(cv[0], s[0]), o_du.stretch[0]);
      t = Ar<S>::mad(s[1], o_dud.stretch[0], t);
      double t2 = Ar<S>::mul(s[2], o_d2u.stretch[0]);
      t2 = Ar<S>::mad(dus, o_d2u.stretch_correct[0], t2);
      a[0] = Ar<S>::mad(nu, t2, Ar<S>::mul(-0.5, t));
    }
    for (int j = 2; j <= n - 1; ++j) {  // distributed.f90:313-327
      const size_t off = (size_t)(j - 1) * SZ;
      double x1 = Ar<S>::nmad(o_du.sc[j - 1], e[0], Ar<S>::nmad(o_du.sa[j - 1], s[0], a[off]));
      const double temp_du = Ar<S>::mul(o_du.stretch[j - 1], x1);
      double x2 = Ar<S>::nmad(o_dud.sc[j - 1], e[1], Ar<S>::nmad(o_dud.sa[j - 1], s[1], b[off]));
      const double temp_dud = Ar<S>::mul(o_dud.stretch[j - 1], x2);
      double x3 = Ar<S>::nmad(o_d2u.sc[j - 1], e[2], Ar<S>::nmad(o_d2u.sa[j - 1], s[2], c[off]));
      const double temp_d2u = Ar<S>::mad(temp_du, o_d2u.stretch_correct[j - 1], Ar<S>::mul(o_d2u.stretch[j - 1], x3));
      const double conv_term = Ar<S>::mad(cv[off], temp_du, temp_dud);  // v*temp_du + temp_dud
      a[off] = Ar<S>::mad(nu, temp_d2u, Ar<S>::mul(-0.5, conv_term));
    }
    {  // row n (distributed.f90:328-335)
      const size_t off = (size_t)(n - 1) * SZ;
      const double due = Ar<S>::mul(e[0], o_du.stretch[n - 1]);
      double t = Ar<S>::mul(Ar<S>::mul(cv[off], e[0]), o_du.stretch[n - 1]);
      t = Ar<S>::mad(e[1], o_dud.stretch[n - 1], t);
      double t2 = Ar<S>::mul(e[2], o_d2u.stretch[n - 1]);
      t2 = Ar<S>::mad(due, o_d2u.stretch_correct[n - 1], t2);
      a[off] = Ar<S>::mad(nu, t2, Ar<S>::mul(-0.5, t));
    }
  }
}

// pack first / last four rows of a directional field (omp/backend.f90:714-737)
__global__ void halo_pack_kernel(double* __restrict__ send_s, double* __restrict__ send_e,
                                 const double* __restrict__ u, const int n, const int n_pad, const int n_groups) {
  const int lane = threadIdx.x & 31;
  const int row = threadIdx.x >> 5;  // 0..3
  const int g = blockIdx.x;
  if (g >= n_groups) return;
  const double* ug = u + (size_t)SZ * n_pad * g + lane;
  send_s[((size_t)g * 4 + row) * SZ + lane] = ug[(size_t)row * SZ];
  send_e[((size_t)g * 4 + row) * SZ + lane] = ug[(size_t)(n - 4 + row) * SZ];
}

}  // namespace

namespace x3d2c {
// nccl.cu: exchange along `dir`: recv_s <- prev.send_e, recv_e <- next.send_s, `count` doubles each
int sendrecv_fields(x3d2c_ctx* ctx, int dir, double* recv_s, double* recv_e, const double* send_s,
                    const double* send_e, size_t count);

// m3 fast path (tds_m3.cu); returns X3D2C_EUNSUPPORTED when the shape is not covered
int tds_solve_m3(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops);
// pair kernels (tds_pair_m3.cu); mode 0 sum, 1 dual, 2 axpy
// TMA kernels (tds_m4.cu); mode 0 single, 1 sum, 2 dual, 3 axpy
int tds_m4(x3d2c_ctx* ctx, int dir, int mode, double* out_a, double* out_b, const double* in_a, const double* in_b,
           const x3d2c_tdsops* ta, const x3d2c_tdsops* tb, double scale_a, int lay_in, int lay_out);
int tds_pair_m3(x3d2c_ctx* ctx, int dir, int mode, double* out_a, double* out_b, const double* in_a,
                const double* in_b, const x3d2c_tdsops* ta, const x3d2c_tdsops* tb, double scale_a);
int transeq_m4(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym, int lay_in, bool dry_run = false);
int transeq_m3(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
               const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
               const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym);
}  // namespace x3d2c

using namespace x3d2c;

// halo buffer carving inside ctx->halo (all sized for the largest cross-section)
namespace {
struct HaloBufs {
  double *send_s[3], *send_e[3], *recv_s[3], *recv_e[3];  // 4-row buffers for u, v, w
  double *rsend, *rrecv;                                  // 18 one-row buffers each? (3 comps x 6)
  size_t row;                                             // doubles in a 1-row buffer
};
HaloBufs carve(x3d2c_ctx* ctx) {
  HaloBufs h;
  int ng = ctx->n_groups[1] > ctx->n_groups[2] ? ctx->n_groups[1] : ctx->n_groups[2];
  if (ctx->n_groups[3] > ng) ng = ctx->n_groups[3];
  const size_t row = (size_t)SZ * ng;
  double* p = ctx->halo;
  for (int f = 0; f < 3; ++f) {
    h.send_s[f] = p; p += 4 * row;
    h.send_e[f] = p; p += 4 * row;
    h.recv_s[f] = p; p += 4 * row;
    h.recv_e[f] = p; p += 4 * row;
  }
  h.rsend = p; p += 18 * row;
  h.rrecv = p; p += 18 * row;
  static_assert(3 * 4 * 4 + 2 * 18 == x3d2c::kHaloRowsRef, "carve() and kHaloRowsRef disagree");
  if ((size_t)(p - ctx->halo) > ctx->halo_doubles) { std::fprintf(stderr, "x3d2c: halo buffer overflow in carve()\n"); std::abort(); }
  h.row = row;
  return h;
}

// X3D2C_TRACE=1: report on stderr which kernel family serves each call (used by tools/mgpu_check.py)
// family: 4 = TMA tiles (*_m4.cu), 3 = cp.async tiles (*_m3.cu), 1 = reference-order kernels (this file)
void trace_path(const char* op, int dir, int P, int family) {
  static int on = -1;
  if (on < 0) on = std::getenv("X3D2C_TRACE") ? 1 : 0;
  if (on)
    std::fprintf(stderr, "[x3d2c] %s dir=%d ranks=%d -> %s\n", op, dir, P,
                 family == 4 ? "m4 (tma tiles)"
                             : (family == 3 ? "m3 (cp.async tiles)"
                                            : (family == 5 ? "g (generic segment-parallel)" : "m1 (reference order)")));
}
}  // namespace

extern "C" {

int x3d2c_tds_solve(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && du && u && ops, "x3d2c_tds_solve: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(du != u, "x3d2c_tds_solve: du and u must be different fields");
  const int n_pad = ctx->n_pad(dir), G = ctx->n_groups[dir];
  X3D2C_REQUIRE(ops->n_rhs <= n_pad, "x3d2c_tds_solve: operator longer than the padded line");
  const int P = ctx->cfg.nproc_dir[dir - 1];
  if (!ctx->strict) {  // fast path: periodic uniform directions, single-rank or rank-split
    int rc = tds_m4(ctx, dir, 0, du, nullptr, u, nullptr, ops, ops, 1.0, dir, dir);
    const bool tma = rc != X3D2C_EUNSUPPORTED;
    if (!tma) rc = tds_solve_m3(ctx, dir, du, u, ops);
    bool gen = false;
    if (rc == X3D2C_EUNSUPPORTED) { rc = tds_g(ctx, dir, du, u, ops); gen = rc != X3D2C_EUNSUPPORTED; }
    trace_path("tds_solve", dir, P, tma ? 4 : (gen ? 5 : (rc != X3D2C_EUNSUPPORTED ? 3 : 1)));
    if (rc != X3D2C_EUNSUPPORTED) return rc;
  }
  const dim3 block(128), grid((G + 3) / 4);
  if (P == 1) {
    if (ctx->strict || kM1AlwaysExact)
      tds_m1_kernel<true><<<grid, block, 0, ctx->stream>>>(du, u, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                           nullptr, ops->dev, ops->tap_mask, n_pad, G, PH_ALL);
    else
      tds_m1_kernel<false><<<grid, block, 0, ctx->stream>>>(du, u, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                            nullptr, ops->dev, ops->tap_mask, n_pad, G, PH_ALL);
    X3D2C_CHECK_LAUNCH(ctx);
    return X3D2C_OK;
  }
  // multi-rank direction: pack + halo exchange, dist, reduced-row exchange, subs (omp/backend.f90:361-391)
  HaloBufs h = carve(ctx);
  halo_pack_kernel<<<G, 128, 0, ctx->stream>>>(h.send_s[0], h.send_e[0], u, ops->n_tds, n_pad, G);
  X3D2C_CHECK_LAUNCH(ctx);
  int rc = sendrecv_fields(ctx, dir, h.recv_s[0], h.recv_e[0], h.send_s[0], h.send_e[0], (size_t)SZ * 4 * G);
  if (rc) return rc;
  double *s_s = h.rsend, *s_e = h.rsend + h.row, *r_s = h.rrecv, *r_e = h.rrecv + h.row;
  if (ctx->strict || kM1AlwaysExact)
    tds_m1_kernel<true><<<grid, block, 0, ctx->stream>>>(du, u, h.recv_s[0], h.recv_e[0], s_s, s_e, nullptr, nullptr,
                                                         ops->dev, ops->tap_mask, n_pad, G, PH_DIST);
  else
    tds_m1_kernel<false><<<grid, block, 0, ctx->stream>>>(du, u, h.recv_s[0], h.recv_e[0], s_s, s_e, nullptr, nullptr,
                                                          ops->dev, ops->tap_mask, n_pad, G, PH_DIST);
  X3D2C_CHECK_LAUNCH(ctx);
  rc = sendrecv_fields(ctx, dir, r_s, r_e, s_s, s_e, (size_t)SZ * G);
  if (rc) return rc;
  if (ctx->strict || kM1AlwaysExact)
    tds_m1_kernel<true><<<grid, block, 0, ctx->stream>>>(du, u, nullptr, nullptr, nullptr, nullptr, r_s, r_e,
                                                         ops->dev, ops->tap_mask, n_pad, G, PH_SUBS);
  else
    tds_m1_kernel<false><<<grid, block, 0, ctx->stream>>>(du, u, nullptr, nullptr, nullptr, nullptr, r_s, r_e,
                                                          ops->dev, ops->tap_mask, n_pad, G, PH_SUBS);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_tds_solve_sum(x3d2c_ctx* ctx, int dir, double* out, const double* in_a, const x3d2c_tdsops* op_a,
                        const double* in_b, const x3d2c_tdsops* op_b) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out && in_a && in_b && op_a && op_b, "x3d2c_tds_solve_sum: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_sum: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(out != in_a && out != in_b, "x3d2c_tds_solve_sum: out must differ from the inputs");
  if (!ctx->strict) {
    int rc = tds_m4(ctx, dir, 1, out, nullptr, in_a, in_b, op_a, op_b, 1.0, dir, dir);
    const bool tma = rc != X3D2C_EUNSUPPORTED;
    if (!tma) rc = tds_pair_m3(ctx, dir, 0, out, nullptr, in_a, in_b, op_a, op_b, 1.0);
    trace_path("tds_solve_sum", dir, ctx->cfg.nproc_dir[dir - 1], tma ? 4 : (rc != X3D2C_EUNSUPPORTED ? 3 : 1));
    if (rc != X3D2C_EUNSUPPORTED) return rc;
  }
  int rc = ensure_scratch(ctx);
  if (rc) return rc;
  if ((rc = x3d2c_tds_solve(ctx, dir, out, in_a, op_a))) return rc;
  if ((rc = x3d2c_tds_solve(ctx, dir, ctx->scratch[0], in_b, op_b))) return rc;
  return x3d2c_vecadd(ctx, 1.0, ctx->scratch[0], 1.0, out);
}

int x3d2c_tds_solve_dual(x3d2c_ctx* ctx, int dir, double* out_a, double* out_b, const double* in,
                         const x3d2c_tdsops* op_a, const x3d2c_tdsops* op_b) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out_a && out_b && in && op_a && op_b, "x3d2c_tds_solve_dual: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_dual: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(out_a != in && out_b != in && out_a != out_b, "x3d2c_tds_solve_dual: fields must be distinct");
  if (!ctx->strict) {
    int rc = tds_m4(ctx, dir, 2, out_a, out_b, in, nullptr, op_a, op_b, 1.0, dir, dir);
    const bool tma = rc != X3D2C_EUNSUPPORTED;
    if (!tma) rc = tds_pair_m3(ctx, dir, 1, out_a, out_b, in, nullptr, op_a, op_b, 1.0);
    trace_path("tds_solve_dual", dir, ctx->cfg.nproc_dir[dir - 1], tma ? 4 : (rc != X3D2C_EUNSUPPORTED ? 3 : 1));
    if (rc != X3D2C_EUNSUPPORTED) return rc;
  }
  int rc = x3d2c_tds_solve(ctx, dir, out_a, in, op_a);
  if (rc) return rc;
  return x3d2c_tds_solve(ctx, dir, out_b, in, op_b);
}

int x3d2c_tds_solve_axpy(x3d2c_ctx* ctx, int dir, double* y, double a, const double* in, const x3d2c_tdsops* op) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && y && in && op, "x3d2c_tds_solve_axpy: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_axpy: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(y != in, "x3d2c_tds_solve_axpy: y and in must be different fields");
  if (!ctx->strict) {
    int rc = tds_m4(ctx, dir, 3, y, nullptr, in, y, op, op, a, dir, dir);
    const bool tma = rc != X3D2C_EUNSUPPORTED;
    if (!tma) rc = tds_pair_m3(ctx, dir, 2, y, nullptr, in, y, op, op, a);
    trace_path("tds_solve_axpy", dir, ctx->cfg.nproc_dir[dir - 1], tma ? 4 : (rc != X3D2C_EUNSUPPORTED ? 3 : 1));
    if (rc != X3D2C_EUNSUPPORTED) return rc;
  }
  int rc = ensure_scratch(ctx);
  if (rc) return rc;
  if ((rc = x3d2c_tds_solve(ctx, dir, ctx->scratch[0], in, op))) return rc;
  return x3d2c_vecadd(ctx, a, ctx->scratch[0], 1.0, y);
}

int x3d2c_transeq(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
                  const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
                  const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && du && dv && dw && u && v && w && der1st && der1st_sym && der2nd && der2nd_sym,
                "x3d2c_transeq: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_transeq: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(der1st->n_rhs == der1st->n_tds && der2nd->n_rhs == der2nd->n_tds &&
                    der1st->n_tds == der2nd->n_tds && der1st_sym->n_tds == der1st->n_tds &&
                    der2nd_sym->n_tds == der1st->n_tds, "x3d2c_transeq: operators must share n_tds == n_rhs");
  const int P = ctx->cfg.nproc_dir[dir - 1];
  if (!ctx->strict) {
    int rc = transeq_m4(ctx, dir, du, dv, dw, u, v, w, nu, der1st, der1st_sym, der2nd, der2nd_sym, dir);  // TMA tiles
    const bool tma = rc != X3D2C_EUNSUPPORTED;
    if (!tma) rc = transeq_m3(ctx, dir, du, dv, dw, u, v, w, nu, der1st, der1st_sym, der2nd, der2nd_sym);
    if (rc != X3D2C_EUNSUPPORTED) {
      trace_path("transeq", dir, P, tma ? 4 : 3);
      return rc;
    }
  }
  // argument permutation of omp/backend.f90:154,168,182: component 0 is the line-aligned velocity
  double* out[3];
  const double* in[3];
  if (dir == X3D2C_DIR_X) { out[0] = du; out[1] = dv; out[2] = dw; in[0] = u; in[1] = v; in[2] = w; }
  else if (dir == X3D2C_DIR_Y) { out[0] = dv; out[1] = du; out[2] = dw; in[0] = v; in[1] = u; in[2] = w; }
  else { out[0] = dw; out[1] = du; out[2] = dv; in[0] = w; in[1] = u; in[2] = v; }
  if (!ctx->strict) {  // walls, stretched meshes, ragged lines: generic segment-parallel kernel (tds_g.cu)
    int rc = transeq_g(ctx, dir, out, in, nu, der1st, der1st_sym, der2nd, der2nd_sym);
    if (rc != X3D2C_EUNSUPPORTED) {
      trace_path("transeq", dir, P, 5);
      return rc;
    }
  }
  trace_path("transeq", dir, P, 1);
  int rc = ensure_scratch(ctx);
  if (rc) return rc;
  const int n_pad = ctx->n_pad(dir), G = ctx->n_groups[dir];
  const dim3 block(128), grid((G + 3) / 4);
  HaloBufs h = carve(ctx);
  if (P > 1) {  // transeq_halo_exchange (omp/backend.f90:264-297)
    for (int f = 0; f < 3; ++f) {
      halo_pack_kernel<<<G, 128, 0, ctx->stream>>>(h.send_s[f], h.send_e[f], in[f], der1st->n_tds, n_pad, G);
      X3D2C_CHECK_LAUNCH(ctx);
    }
    for (int f = 0; f < 3; ++f) {
      rc = sendrecv_fields(ctx, dir, h.recv_s[f], h.recv_e[f], h.send_s[f], h.send_e[f], (size_t)SZ * 4 * G);
      if (rc) return rc;
    }
  }
  for (int comp = 0; comp < 3; ++comp) {  // omp/backend.f90:246-260
    const x3d2c_tdsops* t_du = comp == 0 ? der1st : der1st_sym;
    const x3d2c_tdsops* t_dud = comp == 0 ? der1st_sym : der1st;
    const x3d2c_tdsops* t_d2u = comp == 0 ? der2nd : der2nd_sym;
    const double *uhs = nullptr, *uhe = nullptr, *chs = nullptr, *che = nullptr;
    if (P > 1) { uhs = h.recv_s[comp]; uhe = h.recv_e[comp]; chs = h.recv_s[0]; che = h.recv_e[0]; }
    double* send = h.rsend + (size_t)comp * 6 * h.row;
    double* recv = h.rrecv + (size_t)comp * 6 * h.row;
    auto launch = [&](int phase) {
      if (ctx->strict || kM1AlwaysExact)
        transeq_m1_kernel<true><<<grid, block, 0, ctx->stream>>>(
            out[comp], ctx->scratch[0], ctx->scratch[1], in[comp], in[0], uhs, uhe, chs, che, send, recv, h.row,
            t_du->dev, t_dud->dev, t_d2u->dev, t_du->tap_mask, t_dud->tap_mask, t_d2u->tap_mask, nu, n_pad, G, phase);
      else
        transeq_m1_kernel<false><<<grid, block, 0, ctx->stream>>>(
            out[comp], ctx->scratch[0], ctx->scratch[1], in[comp], in[0], uhs, uhe, chs, che, send, recv, h.row,
            t_du->dev, t_dud->dev, t_d2u->dev, t_du->tap_mask, t_dud->tap_mask, t_d2u->tap_mask, nu, n_pad, G, phase);
    };
    if (P == 1) {
      launch(PH_ALL);
      X3D2C_CHECK_LAUNCH(ctx);
    } else {
      launch(PH_DIST);
      X3D2C_CHECK_LAUNCH(ctx);
      for (int q = 0; q < 3; ++q) {  // exec_dist.f90:163-168
        rc = sendrecv_fields(ctx, dir, recv + (2 * q) * h.row, recv + (2 * q + 1) * h.row, send + (2 * q) * h.row,
                             send + (2 * q + 1) * h.row, (size_t)SZ * G);
        if (rc) return rc;
      }
      launch(PH_SUBS);
      X3D2C_CHECK_LAUNCH(ctx);
    }
  }
  return X3D2C_OK;
}

// transeq_species (src/backend/omp/backend.f90:186-233): convection + diffusion of one scalar `spec` advected by the
// line-aligned velocity `uvw`: rhs = -1/2 (uvw d(spec) + d(spec uvw)) + nu d2(spec) with (der1st, der1st_sym, der2nd) in
// the roles (du, dud, d2u) of transeq_dist_component (:299-338). `sync`: exchange the velocity halos too (the first
// species of a direction); otherwise the halos received by the previous call are reused, as in the reference.
// Reference-order kernels (one scalar per call has a third of the arithmetic intensity of the fused momentum kernel).

int x3d2c_transeq_r_fused(x3d2c_ctx* ctx, int dir, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
                          const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym, int rdr_in) {
  if (!ctx || !der1st || !der1st_sym || !der2nd || !der2nd_sym || dir < 1 || dir > 3 || ctx->strict) return 0;
  if (!rdr_in || rdr_in % 10 != dir || rdr_in / 10 < 1 || rdr_in / 10 > 3) return 0;
  static const bool off = std::getenv("X3D2C_NO_XT") != nullptr || std::getenv("X3D2C_NO_TRANSEQ_XT") != nullptr;
  if (off) return 0;
  double* f = ctx->halo;  // any valid device address: the dry run only encodes tensor maps
  return transeq_m4(ctx, dir, f, f, f, f, f, f, 1.0, der1st, der1st_sym, der2nd, der2nd_sym, rdr_in / 10, true) == X3D2C_OK;
}

int x3d2c_transeq_r(x3d2c_ctx* ctx, int dir, double* du, double* dv, double* dw, const double* u, const double* v,
                    const double* w, double nu, const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym,
                    const x3d2c_tdsops* der2nd, const x3d2c_tdsops* der2nd_sym, int rdr_in) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && du && dv && dw && u && v && w && der1st && der1st_sym && der2nd && der2nd_sym,
                "x3d2c_transeq_r: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_transeq_r: dir must be DIR_X/Y/Z");
  if (!rdr_in) return x3d2c_transeq(ctx, dir, du, dv, dw, u, v, w, nu, der1st, der1st_sym, der2nd, der2nd_sym);
  X3D2C_REQUIRE(rdr_in % 10 == dir && rdr_in / 10 >= 1 && rdr_in / 10 <= 3 && rdr_in / 10 != dir,
                "x3d2c_transeq_r: rdr_in must be a reorder code that ends in dir");
  static const bool trace = std::getenv("X3D2C_TRACE") != nullptr;
  if (!ctx->strict) {  // y lines reading the x layout through swizzled tiles (transeq_m4.cu)
    int rc = transeq_m4(ctx, dir, du, dv, dw, u, v, w, nu, der1st, der1st_sym, der2nd, der2nd_sym, rdr_in / 10);
    if (rc != X3D2C_EUNSUPPORTED) {
      if (trace) std::fprintf(stderr, "[x3d2c] transeq_r dir=%d rdr_in=%d -> inputs through the swizzled tensor maps\n", dir, rdr_in);
      return rc;
    }
  }
  if (trace) std::fprintf(stderr, "[x3d2c] transeq_r dir=%d rdr_in=%d -> reorder + transeq\n", dir, rdr_in);
  int rc;
  const double* in[3] = {u, v, w};
  for (int f = 0; f < 3; ++f) {
    if ((rc = ensure_scratch_slot(ctx, 2 + f))) return rc;
    if ((rc = x3d2c_reorder(ctx, rdr_in, ctx->scratch[2 + f], in[f]))) return rc;
  }
  return x3d2c_transeq(ctx, dir, du, dv, dw, ctx->scratch[2], ctx->scratch[3], ctx->scratch[4], nu, der1st, der1st_sym, der2nd,
                       der2nd_sym);
}

int x3d2c_transeq_species(x3d2c_ctx* ctx, int dir, double* dspec, const double* uvw, const double* spec, double nu,
                          const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym, const x3d2c_tdsops* der2nd,
                          int sync) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && dspec && uvw && spec && der1st && der1st_sym && der2nd, "x3d2c_transeq_species: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_transeq_species: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(der1st->n_rhs == der1st->n_tds && der2nd->n_rhs == der2nd->n_tds && der1st->n_tds == der2nd->n_tds &&
                    der1st_sym->n_tds == der1st->n_tds, "x3d2c_transeq_species: operators must share n_tds == n_rhs");
  X3D2C_REQUIRE(dspec != spec && dspec != uvw, "x3d2c_transeq_species: dspec must differ from the inputs");
  const int P = ctx->cfg.nproc_dir[dir - 1];
  int rc = ensure_scratch(ctx);
  if (rc) return rc;
  const int n_pad = ctx->n_pad(dir), G = ctx->n_groups[dir];
  const dim3 block(128), grid((G + 3) / 4);
  HaloBufs h = carve(ctx);
  if (P > 1) {
    if (sync) {  // velocity halos into the u buffers
      halo_pack_kernel<<<G, 128, 0, ctx->stream>>>(h.send_s[0], h.send_e[0], uvw, der1st->n_tds, n_pad, G);
      X3D2C_CHECK_LAUNCH(ctx);
      if ((rc = sendrecv_fields(ctx, dir, h.recv_s[0], h.recv_e[0], h.send_s[0], h.send_e[0], (size_t)SZ * 4 * G))) return rc;
    }
    halo_pack_kernel<<<G, 128, 0, ctx->stream>>>(h.send_s[1], h.send_e[1], spec, der1st->n_tds, n_pad, G);  // v buffers
    X3D2C_CHECK_LAUNCH(ctx);
    if ((rc = sendrecv_fields(ctx, dir, h.recv_s[1], h.recv_e[1], h.send_s[1], h.send_e[1], (size_t)SZ * 4 * G))) return rc;
  }
  const double *shs = nullptr, *she = nullptr, *chs = nullptr, *che = nullptr;
  if (P > 1) { shs = h.recv_s[1]; she = h.recv_e[1]; chs = h.recv_s[0]; che = h.recv_e[0]; }
  double *send = h.rsend, *recv = h.rrecv;
  auto launch = [&](int phase) {
    if (ctx->strict || kM1AlwaysExact)
      transeq_m1_kernel<true><<<grid, block, 0, ctx->stream>>>(
          dspec, ctx->scratch[0], ctx->scratch[1], spec, uvw, shs, she, chs, che, send, recv, h.row, der1st->dev,
          der1st_sym->dev, der2nd->dev, der1st->tap_mask, der1st_sym->tap_mask, der2nd->tap_mask, nu, n_pad, G, phase);
    else
      transeq_m1_kernel<false><<<grid, block, 0, ctx->stream>>>(
          dspec, ctx->scratch[0], ctx->scratch[1], spec, uvw, shs, she, chs, che, send, recv, h.row, der1st->dev,
          der1st_sym->dev, der2nd->dev, der1st->tap_mask, der1st_sym->tap_mask, der2nd->tap_mask, nu, n_pad, G, phase);
  };
  if (P == 1) {
    launch(PH_ALL);
    X3D2C_CHECK_LAUNCH(ctx);
    return X3D2C_OK;
  }
  launch(PH_DIST);
  X3D2C_CHECK_LAUNCH(ctx);
  for (int q = 0; q < 3; ++q) {
    rc = sendrecv_fields(ctx, dir, recv + (2 * q) * h.row, recv + (2 * q + 1) * h.row, send + (2 * q) * h.row,
                         send + (2 * q + 1) * h.row, (size_t)SZ * G);
    if (rc) return rc;
  }
  launch(PH_SUBS);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

}  // extern "C"
