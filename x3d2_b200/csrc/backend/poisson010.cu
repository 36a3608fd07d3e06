// FFT Poisson solver with walls in y (010) of the cuda_c backend, uniform or stretched in y; single rank (the
// reference stops for non-periodic BCs on more than one rank, src/poisson_fft.f90:178-180).
// Replaces
//   poisson_010                      src/poisson_fft.f90:228-242
//   enforce / undo_periodicity_y     src/backend/omp/poisson_fft.f90:237-285, cuda/kernels/spectral_processing.f90:1062-1114
//   process_spectral_010             src/backend/omp/kernels/spectral_processing.f90:108-283
//   process_spectral_010_fw/_poisson/_bw   cuda/kernels/spectral_processing.f90:385-702 (stretched y: CUDA-Fortran only)
// The transforms are the ones of the periodic solver (poisson.cu; fft_forward_010 => fft_forward in the reference too).
//
// What is different from the reference's kernels:
//  * one kernel does the whole spectral step. The spectrum is stored as C(j, i, k) with y fastest (poisson.cu), so
//    the y lines of 32 consecutive kx are one contiguous chunk: a CTA stages it in shared memory with coalesced
//    128-bit accesses, runs normalisation + rotations, the paired-mode combination, the solve and the way back on chip,
//    and writes it once. The reference makes 3 (uniform) or 5 (stretched) passes with one thread per (kx, kz) striding
//    through y;
//  * the pentadiagonal systems of the stretched mesh are factorised ONCE at creation (multipliers, pivots' reciprocals
//    and the final upper bands, 5 doubles per row). The reference eliminates in place inside the kernel, destroying
//    the coefficient tensors, and restores all of them with device-to-device copies before every solve
//    (cuda/poisson_fft.f90:870-895). The arithmetic applied to the right-hand side is the same, in the same order.
#include "common.cuh"

namespace {

constexpr int LPC = 32;    // y lines (consecutive kx) per CTA
constexpr int NT = 128;    // threads per CTA: (line, family, re | im) in the pentadiagonal phase
constexpr double kEps = 1.e-16;

struct P010 {
  int nx, ny, nz, nxh;       // global cell dims, nx / 2 + 1
  int stretched, rows;       // 0 | 1 | 2, rows per family
  int pow2;
  double inv_n;
};

extern __shared__ __align__(16) double sm010[];

// Whole spectral step of poisson_010 on the lines (i0 .. i0 + LPC - 1, k) of C(j, i, k).
__global__ void __launch_bounds__(NT)
spectral_010_kernel(double2* __restrict__ c, const double2* __restrict__ waves, const double* __restrict__ fac_re,
                    const double* __restrict__ fac_im, const double* __restrict__ ax, const double* __restrict__ bx,
                    const double* __restrict__ ay, const double* __restrict__ by, const double* __restrict__ az,
                    const double* __restrict__ bz, const P010 p) {
  const int k = blockIdx.y, i0 = blockIdx.x * LPC;
  const int nl = min(LPC, p.nxh - i0);        // lines of this CTA
  const int ny = p.ny, LS = 2 * ny + 2;       // shared line stride in doubles (+2: neighbouring lines on other banks)
  const size_t base = (size_t)ny * (i0 + (size_t)p.nxh * k);
  const int tid = threadIdx.x;
  const double azk = az[k], bzk = bz[k];
  const bool fz = (k + 1) > p.nz / 2 + 1;
  // ---- load; normalisation; rotations in z and x (first block of process_spectral_010)
  for (int e = tid; e < nl * ny; e += NT) {
    const int line = e / ny, y0 = e - line * ny, i = i0 + line;
    const double2 v = c[base + e];
    double div_r, div_c;
    if (p.pow2) { div_r = v.x * p.inv_n; div_c = v.y * p.inv_n; }
    else { div_r = v.x / p.nx / p.ny / p.nz; div_c = v.y / p.nx / p.ny / p.nz; }
    double tr = div_r, tc = div_c;
    div_r = tr * bzk + tc * azk;
    div_c = tc * bzk - tr * azk;
    if (fz) { div_r = -div_r; div_c = -div_c; }
    tr = div_r; tc = div_c;
    const double axi = ax[i], bxi = bx[i];
    div_r = tr * bxi + tc * axi;
    div_c = tc * bxi - tr * axi;
    sm010[line * LS + 2 * y0] = div_r;
    sm010[line * LS + 2 * y0 + 1] = div_c;
  }
  __syncthreads();
  // ---- paired modes (j, ny - j + 2), j = 2 .. ny / 2 + 1 (second block)
  const int half = ny / 2;
  for (int e = tid; e < nl * half; e += NT) {
    const int line = e / half, j = 2 + (e - line * half), jr = ny - j + 2;  // 1-based
    double* L = sm010 + line * LS + 2 * (j - 1);
    double* R = sm010 + line * LS + 2 * (jr - 1);
    const double l_r = L[0], l_c = L[1], r_r = R[0], r_c = R[1];
    const double ayj = ay[j - 1], byj = by[j - 1], ayr = ay[jr - 1], byr = by[jr - 1];
    const double n_lr = 0.5 * (l_r * byj + l_c * ayj + r_r * byj - r_c * ayj);
    const double n_lc = 0.5 * (-l_r * ayj + l_c * byj + r_r * ayj + r_c * byj);
    const double n_rr = 0.5 * (r_r * byr + r_c * ayr + l_r * byr - l_c * ayr);
    const double n_rc = 0.5 * (-r_r * ayr + r_c * byr + l_r * ayr + l_c * byr);
    if (j != jr) { L[0] = n_lr; L[1] = n_lc; }  // the self-paired mode keeps the second assignment, as in the reference
    R[0] = n_rr; R[1] = n_rc;
  }
  __syncthreads();
  // ---- solve
  const bool zero_k = k == p.nz / 2;  // (i == nx / 2 + 1 .and. k == nz / 2 + 1) in 1-based indices
  if (p.stretched == 0) {
    for (int e = tid; e < nl * ny; e += NT) {
      const int line = e / ny, y0 = e - line * ny, i = i0 + line;
      const double2 w = waves[base + e];
      double* X = sm010 + line * LS + 2 * y0;
      double div_r = X[0], div_c = X[1];
      div_r = fabs(w.x) < kEps ? 0.0 : -div_r / w.x;
      div_c = fabs(w.y) < kEps ? 0.0 : -div_c / w.y;
      if (zero_k && i == p.nx / 2) { div_r = 0.0; div_c = 0.0; }
      X[0] = div_r; X[1] = div_c;
    }
  } else {
    const int F = p.stretched == 1 ? 2 : 1, n = p.rows;
    const int inc = F, per_line = 2 * F;
    if (tid < nl * per_line) {
      const int line = tid / per_line, r = tid - line * per_line, fam = r >> 1, part = r & 1;
      const int i = i0 + line;
      const double* fc = (part ? fac_im : fac_re) + ((size_t)(i + (size_t)p.nxh * k) * F + fam) * (size_t)n * 5;
      double* X = sm010 + line * LS + part;  // X[2 * y0]
      auto yidx = [&](int j) { return 2 * (inc * (j - 1) + fam); };  // row j (1-based) of the family -> offset in X
      const bool zero = zero_k && i == p.nx / 2;
      // forward elimination with the stored multipliers
      double xj = X[yidx(1)], xj1 = X[yidx(2)];
      for (int j = 1; j <= n - 2; ++j) {
        const double m1 = fc[(j - 1) * 5], m2 = fc[(j - 1) * 5 + 1];
        double xj2 = X[yidx(j + 2)];
        xj1 = xj1 - m1 * xj;
        xj2 = xj2 - m2 * xj;
        X[yidx(j)] = xj;
        xj = xj1; xj1 = xj2;
      }
      // xj = x(n-1), xj1 = x(n): the last two rows
      const double* fl = fc + (size_t)(n - 2) * 5;  // row n - 1: {., t / d, 1 / a3, a4, a4 / a3}; row n: {d, ...}
      const double d = fl[5], td = fl[1], ti = fl[2], dd = fl[4];
      double xn = fabs(d) > kEps ? xj1 / d - td * xj : 0.0;
      double xn1 = xj * ti - xn * dd;
      if (zero) { xn = 0.0; xn1 = 0.0; }
      X[yidx(n)] = xn;
      X[yidx(n - 1)] = xn1;
      // back substitution
      double x1 = xn1, x2 = xn;
      for (int j = n - 2; j >= 1; --j) {
        const double i3 = fc[(j - 1) * 5 + 2], u4 = fc[(j - 1) * 5 + 3], u5 = fc[(j - 1) * 5 + 4];
        double x0 = i3 * (X[yidx(j)] - u4 * x1 - u5 * x2);
        if (zero) x0 = 0.0;
        X[yidx(j)] = x0;
        x2 = x1; x1 = x0;
      }
    }
  }
  __syncthreads();
  // ---- paired modes, backward (fourth block)
  for (int e = tid; e < nl * half; e += NT) {
    const int line = e / half, j = 2 + (e - line * half), jr = ny - j + 2;
    double* L = sm010 + line * LS + 2 * (j - 1);
    double* R = sm010 + line * LS + 2 * (jr - 1);
    const double l_r = L[0], l_c = L[1], r_r = R[0], r_c = R[1];
    const double ayj = ay[j - 1], byj = by[j - 1], ayr = ay[jr - 1], byr = by[jr - 1];
    const double n_lr = l_r * byj - l_c * ayj + r_r * ayj + r_c * byj;
    const double n_lc = l_r * ayj + l_c * byj - r_r * byj + r_c * ayj;
    const double n_rr = r_r * byr - r_c * ayr + l_r * ayr + l_c * byr;
    const double n_rc = r_r * ayr + r_c * byr - l_r * byr + l_c * ayr;
    if (j != jr) { L[0] = n_lr; L[1] = n_lc; }
    R[0] = n_rr; R[1] = n_rc;
  }
  __syncthreads();
  // ---- inverse rotations in z and x; store (last block)
  for (int e = tid; e < nl * ny; e += NT) {
    const int line = e / ny, y0 = e - line * ny, i = i0 + line;
    double div_r = sm010[line * LS + 2 * y0], div_c = sm010[line * LS + 2 * y0 + 1];
    double tr = div_r, tc = div_c;
    div_r = tr * bzk - tc * azk;
    div_c = tc * bzk + tr * azk;
    if (fz) { div_r = -div_r; div_c = -div_c; }
    tr = div_r; tc = div_c;
    const double axi = ax[i], bxi = bx[i];
    div_r = tr * bxi - tc * axi;
    div_c = tc * bxi + tr * axi;
    c[base + e] = make_double2(div_r, div_c);
  }
}

// f_out(i, j, k) = f_in(i, perm(j), k) on padded DIR_C blocks; ENFORCE: perm(j) = 2 j - 1 (j <= ny / 2), 2 ny - 2 j + 2
// (j > ny / 2); otherwise the inverse map (omp/poisson_fft.f90:237-285; odd ny: cuda/kernels/spectral_processing.f90:1075-1083)
template <bool ENFORCE>
__global__ void __launch_bounds__(256)
periodicity_y_kernel(double* __restrict__ f_out, const double* __restrict__ f_in, const int nx, const int ny, const int nz,
                     const int nx_pad, const int ny_pad) {
  const size_t n = (size_t)nx * ny * nz;
  const int n2 = ny / 2;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % nx);
    const size_t r = t / nx;
    const int j = (int)(r % ny) + 1, k = (int)(r / ny);  // j 1-based: index in the periodised (FFT) ordering
    int jp;                                              // 1-based index in the physical ordering
    if (j <= n2) jp = 2 * j - 1;
    else if ((ny & 1) && j == n2 + 1) jp = ny;
    else jp = 2 * ny - 2 * j + 2;
    const size_t a = i + (size_t)nx_pad * ((j - 1) + (size_t)ny_pad * k), b = i + (size_t)nx_pad * ((jp - 1) + (size_t)ny_pad * k);
    if (ENFORCE) f_out[a] = f_in[b];
    else f_out[b] = f_in[a];
  }
}

// Factorises the pentadiagonal systems of one family on the host: the elimination of process_spectral_010_poisson
// (cuda/kernels/spectral_processing.f90:488-534,536-590) applied to the coefficients only.
// a: (nxh, n, nz, 5) Fortran order; out: [line = i + nxh k][F families][n rows][5], this family's slot.
void factorise(const double* a, int nxh, int n, int nz, int F, int fam, std::vector<double>& out) {
  std::vector<double> a1(n + 3), a2(n + 3), a3(n + 3), a4(n + 3), a5(n + 3);
  for (int k = 0; k < nz; ++k)
    for (int i = 0; i < nxh; ++i) {
      auto A = [&](int j, int d) { return a[(size_t)i + (size_t)nxh * ((j - 1) + (size_t)n * (k + (size_t)nz * (d - 1)))]; };
      for (int j = 1; j <= n; ++j) { a1[j] = A(j, 1); a2[j] = A(j, 2); a3[j] = A(j, 3); a4[j] = A(j, 4); a5[j] = A(j, 5); }
      double* o = out.data() + ((size_t)(i + (size_t)nxh * k) * F + fam) * (size_t)n * 5;
      for (int j = 1; j <= n - 2; ++j) {
        double t = std::fabs(a3[j]) > kEps ? a2[j + 1] / a3[j] : 0.0;
        o[(j - 1) * 5] = t;
        a3[j + 1] = a3[j + 1] - t * a4[j];
        a4[j + 1] = a4[j + 1] - t * a5[j];
        t = std::fabs(a3[j]) > kEps ? a1[j + 2] / a3[j] : 0.0;
        o[(j - 1) * 5 + 1] = t;
        a2[j + 2] = a2[j + 2] - t * a4[j];
        a3[j + 2] = a3[j + 2] - t * a5[j];
        o[(j - 1) * 5 + 2] = std::fabs(a3[j]) > kEps ? 1.0 / a3[j] : 0.0;
        o[(j - 1) * 5 + 3] = a4[j];
        o[(j - 1) * 5 + 4] = a5[j];
      }
      const double t = std::fabs(a3[n - 1]) > kEps ? a2[n] / a3[n - 1] : 0.0;
      const double d = a3[n] - t * a4[n - 1];
      const double ti = std::fabs(a3[n - 1]) > kEps ? 1.0 / a3[n - 1] : 0.0;
      double* l = o + (size_t)(n - 2) * 5;
      l[0] = t;
      l[1] = std::fabs(d) > kEps ? t / d : 0.0;
      l[2] = ti;
      l[3] = a4[n - 1];
      l[4] = a4[n - 1] * ti;
      l[5] = d;
      l[6] = l[7] = l[8] = l[9] = 0.0;
    }
}

int unsupported(const char* what) {
  x3d2c::set_error(std::string(what) + ": walls in x (100 / 110) are not implemented by the cuda_c backend (no supported "
                   "configuration uses them; the reference's OMP backend stops here too)");
  return X3D2C_EUNSUPPORTED;
}

}  // namespace

using namespace x3d2c;

extern "C" {

int x3d2c_poisson_create_010(x3d2c_ctx* ctx, const double* waves, const double* ax, const double* bx, const double* ay,
                             const double* by, const double* az, const double* bz, int stretched,
                             const double* a_odd_re, const double* a_odd_im, const double* a_even_re,
                             const double* a_even_im, x3d2c_poisson** out) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && waves && ax && bx && ay && by && az && bz && out, "x3d2c_poisson_create_010: null argument");
  X3D2C_REQUIRE(ctx->cfg.periodic[0] && !ctx->cfg.periodic[1] && ctx->cfg.periodic[2],
                "x3d2c_poisson_create_010: needs periodic x and z and walls in y");
  X3D2C_REQUIRE(ctx->cfg.nproc == 1, "Multiple ranks are not yet supported for non-periodic BCs!");
  X3D2C_REQUIRE(stretched >= 0 && stretched <= 2, "x3d2c_poisson_create_010: stretched must be 0, 1 or 2");
  X3D2C_REQUIRE(stretched == 0 || (a_odd_re && a_odd_im), "x3d2c_poisson_create_010: missing coefficient tensors");
  X3D2C_REQUIRE(stretched != 1 || (a_even_re && a_even_im), "x3d2c_poisson_create_010: missing even-mode tensors");
  X3D2C_REQUIRE(ctx->cfg.dims_cell_global[1] % 2 == 0 && ctx->cfg.dims_cell_global[1] >= 8,
                "x3d2c_poisson_create_010: the number of cells in y must be even (and >= 8)");
  x3d2c_poisson* p = nullptr;
  int rc = poisson_create_common(ctx, 10, waves, ax, bx, ay, by, az, bz, &p);
  if (rc) return rc;
  struct Guard {
    x3d2c_ctx* c;
    x3d2c_poisson* p;
    ~Guard() { if (p) x3d2c_poisson_destroy(c, p); }
  } guard{ctx, p};
  p->stretched = stretched;
  if (stretched) {
    const int F = stretched == 1 ? 2 : 1, n = stretched == 1 ? p->ny / 2 : p->ny;
    p->penta_rows = n;
    const size_t total = (size_t)p->nxh * p->nz * F * n * 5, per = (size_t)p->nxh * n * p->nz * 5;
    const bool same = std::memcmp(a_odd_re, a_odd_im, per * sizeof(double)) == 0 &&
                      (F == 1 || std::memcmp(a_even_re, a_even_im, per * sizeof(double)) == 0);
    std::vector<double> fac(total);
    factorise(a_odd_re, p->nxh, n, p->nz, F, 0, fac);
    if (F == 2) factorise(a_even_re, p->nxh, n, p->nz, F, 1, fac);
    X3D2C_CHECK_CUDA(cudaMalloc(&p->fac_re, sizeof(double) * total));
    X3D2C_CHECK_CUDA(cudaMemcpy(p->fac_re, fac.data(), sizeof(double) * total, cudaMemcpyHostToDevice));
    if (same) {
      p->fac_im = p->fac_re;  // every complex coefficient of the reference is (1 + i) x
    } else {
      factorise(a_odd_im, p->nxh, n, p->nz, F, 0, fac);
      if (F == 2) factorise(a_even_im, p->nxh, n, p->nz, F, 1, fac);
      X3D2C_CHECK_CUDA(cudaMalloc(&p->fac_im, sizeof(double) * total));
      X3D2C_CHECK_CUDA(cudaMemcpy(p->fac_im, fac.data(), sizeof(double) * total, cudaMemcpyHostToDevice));
    }
  }
  const size_t smem = sizeof(double) * (size_t)LPC * (2 * p->ny + 2);
  X3D2C_REQUIRE(smem <= 227 * 1024, "x3d2c_poisson_create_010: ny too large for the on-chip y lines");
  X3D2C_CHECK_CUDA(cudaFuncSetAttribute(spectral_010_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  guard.p = nullptr;
  *out = p;
  return X3D2C_OK;
}

int x3d2c_fft_postprocess_010(x3d2c_ctx* ctx, x3d2c_poisson* p) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p, "x3d2c_fft_postprocess_010: null argument");
  X3D2C_REQUIRE(p->bc_case == 10, "x3d2c_fft_postprocess_010: the solver was not created with x3d2c_poisson_create_010");
  P010 q;
  q.nx = p->nx; q.ny = p->ny; q.nz = p->nz; q.nxh = p->nxh;
  q.stretched = p->stretched; q.rows = p->penta_rows;
  const long long N = (long long)p->nx * p->ny * p->nz;
  q.pow2 = (N & (N - 1)) == 0;
  q.inv_n = 1.0 / (double)N;
  const size_t smem = sizeof(double) * (size_t)LPC * (2 * p->ny + 2);
  const dim3 grid((p->nxh + LPC - 1) / LPC, p->nz);
  spectral_010_kernel<<<grid, NT, smem, ctx->stream>>>((double2*)p->B, (const double2*)p->waves, p->fac_re, p->fac_im,
                                                       p->ax, p->bx, p->ay, p->by, p->az, p->bz, q);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_enforce_periodicity_y(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p && f_out && f_in && f_out != f_in, "x3d2c_enforce_periodicity_y: bad argument");
  periodicity_y_kernel<true><<<1184, 256, 0, ctx->stream>>>(f_out, f_in, p->nx, p->ny, p->nz_loc, ctx->nx_pad, ctx->ny_pad);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}
int x3d2c_undo_periodicity_y(x3d2c_ctx* ctx, x3d2c_poisson* p, double* f_out, const double* f_in) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && p && f_out && f_in && f_out != f_in, "x3d2c_undo_periodicity_y: bad argument");
  periodicity_y_kernel<false><<<1184, 256, 0, ctx->stream>>>(f_out, f_in, p->nx, p->ny, p->nz_loc, ctx->nx_pad, ctx->ny_pad);
  X3D2C_CHECK_LAUNCH(ctx);
  return X3D2C_OK;
}

int x3d2c_fft_forward_100(x3d2c_ctx*, x3d2c_poisson*, const double*) { return unsupported("fft_forward_100"); }
int x3d2c_fft_forward_110(x3d2c_ctx*, x3d2c_poisson*, const double*) { return unsupported("fft_forward_110"); }
int x3d2c_fft_backward_100(x3d2c_ctx*, x3d2c_poisson*, double*) { return unsupported("fft_backward_100"); }
int x3d2c_fft_backward_110(x3d2c_ctx*, x3d2c_poisson*, double*) { return unsupported("fft_backward_110"); }
int x3d2c_fft_postprocess_100(x3d2c_ctx*, x3d2c_poisson*) { return unsupported("fft_postprocess_100"); }
int x3d2c_fft_postprocess_110(x3d2c_ctx*, x3d2c_poisson*) { return unsupported("fft_postprocess_110"); }
int x3d2c_enforce_periodicity_x(x3d2c_ctx*, x3d2c_poisson*, double*, const double*) { return unsupported("enforce_periodicity_x"); }
int x3d2c_undo_periodicity_x(x3d2c_ctx*, x3d2c_poisson*, double*, const double*) { return unsupported("undo_periodicity_x"); }
int x3d2c_enforce_periodicity_xy(x3d2c_ctx*, x3d2c_poisson*, double*, const double*) { return unsupported("enforce_periodicity_xy"); }
int x3d2c_undo_periodicity_xy(x3d2c_ctx*, x3d2c_poisson*, double*, const double*) { return unsupported("undo_periodicity_xy"); }

}  // extern "C"
