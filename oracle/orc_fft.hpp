// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.hpp).
// The reference's OMP Poisson solver calls the third-party 2DECOMP&FFT v2.0.3.1 (not vendored in
// /root/reference: cmake/decomp2d/downloadBuild2decomp.cmake.in:10-11). Call sites:
// src/backend/omp/poisson_fft.f90:72-73,95,135 (decomp_2d_fft_init(PHYSICAL_IN_X), decomp_2d_fft_3d).
// Restated published semantics: forward = unnormalised DFT with exp(-i w x): r2c along x keeping
// nx/2+1 modes, then c2c along y, then c2c along z; backward = unnormalised inverse (c2c z, c2c y,
// c2r x). tests/verification/test_fft.f90:156-169 pins backward(forward(f)) == N*f.
// Direct spectrum values are parity-unpinned (no golden vectors exist in the reference).
#pragma once
#include "orc_common.hpp"

namespace orc {

using cplx = std::complex<double>;

struct FftPlan {
  int n = 0;
  bool pow2 = false;
  std::vector<cplx> tw;  // exp(-2 pi i k / n), k = 0..n-1
  std::vector<int> rev;
  explicit FftPlan(int n_) : n(n_) {
    pow2 = n > 0 && (n & (n - 1)) == 0;
    tw.resize(n);
    for (int k = 0; k < n; ++k) {
      double a = -2.0 * pi * k / n;
      tw[k] = cplx(std::cos(a), std::sin(a));
    }
    if (pow2) {
      rev.resize(n);
      int lg = 0;
      while ((1 << lg) < n) ++lg;
      for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int b = 0; b < lg; ++b)
          if (i & (1 << b)) r |= 1 << (lg - 1 - b);
        rev[i] = r;
      }
    }
  }
  // in-place transform of a contiguous buffer x[0..n); sign = -1 forward, +1 backward (unnormalised)
  void exec(cplx* x, int sign, cplx* work) const {
    if (n == 1) return;
    if (pow2) {
      for (int i = 0; i < n; ++i)
        if (rev[i] > i) std::swap(x[i], x[rev[i]]);
      for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len)
          for (int k = 0; k < half; ++k) {
            cplx w = tw[k * step];
            if (sign > 0) w = std::conj(w);
            cplx a = x[s + k], b = x[s + k + half] * w;
            x[s + k] = a + b;
            x[s + k + half] = a - b;
          }
      }
    } else {  // plain O(n^2) DFT; only small non power-of-two sizes reach this
      for (int k = 0; k < n; ++k) {
        cplx acc = 0;
        for (int j = 0; j < n; ++j) {
          cplx w = tw[(int)(((long long)j * k) % n)];
          if (sign > 0) w = std::conj(w);
          acc += x[j] * w;
        }
        work[k] = acc;
      }
      for (int k = 0; k < n; ++k) x[k] = work[k];
    }
  }
};

// 3-D forward: real f(nx, ny, nz) (x fastest, leading dims ldx, ldy) -> spec(nxh, ny, nz), nxh = nx/2+1
inline void fft3d_forward(const double* f, int ldx, int ldy, int nx, int ny, int nz, cplx* spec) {
  const int nxh = nx / 2 + 1;
  FftPlan px(nx), py(ny), pz(nz);
#pragma omp parallel
  {
    std::vector<cplx> buf(std::max(nx, std::max(ny, nz))), work(buf.size());
#pragma omp for collapse(2)
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j) {
        const double* row = f + (size_t)ldx * (j + (size_t)ldy * k);
        for (int i = 0; i < nx; ++i) buf[i] = row[i];
        px.exec(buf.data(), -1, work.data());
        cplx* out = spec + (size_t)nxh * (j + (size_t)ny * k);
        for (int i = 0; i < nxh; ++i) out[i] = buf[i];
      }
#pragma omp for collapse(2)
    for (int k = 0; k < nz; ++k)
      for (int i = 0; i < nxh; ++i) {
        cplx* base = spec + i + (size_t)nxh * ny * k;
        for (int j = 0; j < ny; ++j) buf[j] = base[(size_t)nxh * j];
        py.exec(buf.data(), -1, work.data());
        for (int j = 0; j < ny; ++j) base[(size_t)nxh * j] = buf[j];
      }
#pragma omp for collapse(2)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nxh; ++i) {
        cplx* base = spec + i + (size_t)nxh * j;
        for (int k = 0; k < nz; ++k) buf[k] = base[(size_t)nxh * ny * k];
        pz.exec(buf.data(), -1, work.data());
        for (int k = 0; k < nz; ++k) base[(size_t)nxh * ny * k] = buf[k];
      }
  }
}

// 3-D backward (destroys spec): spec(nxh, ny, nz) -> real f(nx, ny, nz), unnormalised
inline void fft3d_backward(cplx* spec, int nx, int ny, int nz, double* f, int ldx, int ldy) {
  const int nxh = nx / 2 + 1;
  FftPlan px(nx), py(ny), pz(nz);
#pragma omp parallel
  {
    std::vector<cplx> buf(std::max(nx, std::max(ny, nz))), work(buf.size());
#pragma omp for collapse(2)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nxh; ++i) {
        cplx* base = spec + i + (size_t)nxh * j;
        for (int k = 0; k < nz; ++k) buf[k] = base[(size_t)nxh * ny * k];
        pz.exec(buf.data(), +1, work.data());
        for (int k = 0; k < nz; ++k) base[(size_t)nxh * ny * k] = buf[k];
      }
#pragma omp for collapse(2)
    for (int k = 0; k < nz; ++k)
      for (int i = 0; i < nxh; ++i) {
        cplx* base = spec + i + (size_t)nxh * ny * k;
        for (int j = 0; j < ny; ++j) buf[j] = base[(size_t)nxh * j];
        py.exec(buf.data(), +1, work.data());
        for (int j = 0; j < ny; ++j) base[(size_t)nxh * j] = buf[j];
      }
#pragma omp for collapse(2)
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j) {
        const cplx* in = spec + (size_t)nxh * (j + (size_t)ny * k);
        // c2r: Hermitian extension X[nx-i] = conj(X[i]); imaginary parts of X[0] and X[nx/2] drop out
        for (int i = 0; i < nxh; ++i) buf[i] = in[i];
        for (int i = nxh; i < nx; ++i) buf[i] = std::conj(in[nx - i]);
        px.exec(buf.data(), +1, work.data());
        double* row = f + (size_t)ldx * (j + (size_t)ldy * k);
        for (int i = 0; i < nx; ++i) row[i] = buf[i].real();
      }
  }
}

}  // namespace orc
