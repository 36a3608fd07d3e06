// Host-side constants of the m3 fast path and the rank-boundary exchange of its distributed variant.
//
// For a rank-split periodic direction the main kernels (tds_m3.cu, transeq_m3.cu) need, per line,
//   * the 4 rows before / after the local line (the stencil halo; same data as copy_into_buffers +
//     sendrecv_fields of the reference, omp/backend.f90:264-297,714-737), and
//   * the carries of the neighbouring ranks' three nearest segments for every recurrence.
// Both are produced here: one pack kernel + one NCCL exchange for all halos, one edge kernel + one NCCL
// exchange for all carries (the reference's distributed solve needs the halo exchange plus one exchange of the
// 2x2 reduced-system rows per recurrence, exec_dist.f90:163-168).
#include "m3_common.cuh"

namespace x3d2c {
int sendrecv_fields(x3d2c_ctx* ctx, int dir, double* recv_s, double* recv_e, const double* send_s,
                    const double* send_e, size_t count);
}

namespace m3 {

namespace {

struct PackParams {
  const double* f[3];
  int n, n_pad, nf;
};

// first / last four rows of nf directional fields -> (SZ, 4, nf, G)
__global__ void __launch_bounds__(128) halo_pack_kernel(double* __restrict__ send_s, double* __restrict__ send_e,
                                                        const __grid_constant__ PackParams p) {
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5, g = blockIdx.x, f = blockIdx.y;
  const double* ug = p.f[f] + (size_t)SZ * p.n_pad * g + lane;
  const size_t o = (((size_t)g * p.nf + f) * 4 + row) * SZ + lane;
  send_s[o] = ug[(size_t)row * SZ];
  send_e[o] = ug[(size_t)(p.n - 4 + row) * SZ];
}

// One thread per (lane, edge segment): segments 0..2 and nseg-3..nseg-1 of every line; blockIdx.y = recurrence.
__global__ void __launch_bounds__(32 * 2 * DMAX) edge_kernel(const __grid_constant__ EdgeParams p) {
  const int lane = threadIdx.x, e = threadIdx.y, g = blockIdx.x, s = blockIdx.y;
  const int q = e < DMAX ? e : p.nseg - 2 * DMAX + e;
  const int j0 = q * S;
  const Op& o = p.ops[p.op[s]];
  const size_t base = (size_t)g * p.n_pad * SZ + lane;
  const double* f = p.f[s] + base;
  const double* c = p.c[s] ? p.c[s] + base : nullptr;
  const size_t hf = ((size_t)g * p.nf + p.ff[s]) * 4 * SZ + lane, hc = ((size_t)g * p.nf + p.cf[s]) * 4 * SZ + lane;
  auto val = [&](int row) -> double {
    double v;
    if (row < 0) {
      v = p.halo_s[hf + (size_t)(row + 4) * SZ];
      if (c) v *= p.halo_s[hc + (size_t)(row + 4) * SZ];
    } else if (row >= p.n) {
      v = p.halo_e[hf + (size_t)(row - p.n) * SZ];
      if (c) v *= p.halo_e[hc + (size_t)(row - p.n) * SZ];
    } else {
      v = f[(size_t)row * SZ];
      if (c) v *= c[(size_t)row * SZ];
    }
    return v;
  };
  double wf[9], z[S];
#pragma unroll
  for (int t = 0; t < 8; ++t) wf[t] = val(j0 - 4 + t);
  double pz = 0.0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    wf[8] = val(j0 + 4 + k);
    pz = fma(o.a, pz, sten<0x1FFu>(o.cfw, wf));
    z[k] = pz;
#pragma unroll
    for (int t = 0; t < 8; ++t) wf[t] = wf[t + 1];
  }
  const size_t ob = (((size_t)g * p.ns + s) * EXP_ROWS) * SZ + lane;
  if (e >= DMAX) {  // ze of the last three segments, for the next rank
    p.to_next[ob + (size_t)(e - DMAX) * SZ] = z[S - 1];
    if (e == DMAX) {  // unused rows of the equally sized buffer
      p.to_next[ob + (size_t)3 * SZ] = 0.0;
      p.to_next[ob + (size_t)4 * SZ] = 0.0;
    }
    return;
  }
  if (e < DMAX - 1) p.to_prev[ob + (size_t)(DMAX + e) * SZ] = z[S - 1];  // ze of the first two segments
  double y = 0.0;
#pragma unroll
  for (int k = S - 1; k >= 0; --k) y = fma(o.cb, y, z[k]);
  p.to_prev[ob + (size_t)e * SZ] = y;  // ys of the first three segments
}

}  // namespace

bool same_tables(const x3d2c_tdsops* a, const x3d2c_tdsops* b) {
  if (a->n_tds != b->n_tds || a->n_rhs != b->n_rhs) return false;
  if (std::memcmp(a->dev.coeffs, b->dev.coeffs, sizeof a->dev.coeffs)) return false;
  const int m = a->n_tds / 2;
  return a->h_fw[m] == b->h_fw[m] && a->h_bw[m] == b->h_bw[m] && a->h_af[m] == b->h_af[m];
}

// constant set of one operator; false when the operator does not qualify for the fast path.
// dist: the operator belongs to a rank-split periodic direction (halo boundary rows on both sides).
bool make_op(const x3d2c_tdsops* t, double scale, bool dist, Op* o) {
  const int n = t->n_tds;
  if ((!t->periodic && !dist) || t->n_rhs != n || n < 4 * S || n % S) return false;
  if (dist && n < 2 * DMAX * S) return false;
  if (t->has_stretch || t->has_stretch_correct) return false;
  const int m = n / 2;
  const double fw = t->h_fw[m], bw = t->h_bw[m], al = t->h_af[m];
  // the factors must have converged to their Toeplitz limit over the whole central region
  for (int j = 40; j < n - 40; ++j) {
    if (std::fabs(t->h_fw[j] - fw) > 4e-16 * std::fabs(fw) || std::fabs(t->h_bw[j] - bw) > 4e-16 * std::fabs(bw) ||
        t->h_af[j] != al)
      return false;
  }
  for (int k = 0; k < 9; ++k) o->cfw[k] = scale * fw * t->dev.coeffs[k];
  o->mask = t->tap_mask;
  o->a = -fw * al;
  o->cb = -bw;
  if (std::pow(std::fabs(o->a), S * DMAX) > 1e-18 || std::pow(std::fabs(o->cb), S * DMAX) > 1e-18) return false;
  for (int d = 0; d < DMAX; ++d) { o->zw[d] = std::pow(o->a, S * d); o->yw[d] = std::pow(o->cb, S * d); }
  double W[S + 1];
  W[S] = 0.0;
  for (int k = S - 1; k >= 0; --k) W[k] = std::pow(o->a, k + 1) + o->cb * W[k + 1];
  for (int k = 0; k < S; ++k) { o->W[k] = W[k]; o->Cp[k] = std::pow(o->cb, S - k); }
  for (int m2 = 0; m2 < 2 * DMAX - 1; ++m2) o->om[m2] = 0.0;
  for (int d = 1; d <= DMAX; ++d)
    for (int dp = 1; dp <= DMAX; ++dp) o->om[d - dp + DMAX - 1] += W[0] * o->yw[d - 1] * o->zw[dp - 1];
  return true;
}

int num_sms(const x3d2c_ctx* ctx) {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  return sms > 0 ? sms : 148;
}

// every rank must take the same decision: equal split of a periodic direction, enough segments per rank
bool dist_supported(const x3d2c_ctx* ctx, int dir, int n) {
  const int d = dir - 1, P = ctx->cfg.nproc_dir[d];
  if (P <= 1 || !ctx->cfg.periodic[d] || ctx->strict) return false;
  if (!ctx->nccl_comm) return false;
  if (ctx->cfg.dims_vert_global[d] != n * P || ctx->cfg.dims_vert[d] != n) return false;
  return n % S == 0 && n >= 2 * DMAX * S;
}

DistBufs carve_dist(x3d2c_ctx* ctx) {
  int ng = ctx->n_groups[1] > ctx->n_groups[2] ? ctx->n_groups[1] : ctx->n_groups[2];
  if (ctx->n_groups[3] > ng) ng = ctx->n_groups[3];
  const size_t row = (size_t)SZ * ng;
  DistBufs b;
  double* p = ctx->halo;
  b.halo_send_s = p; p += 12 * row;
  b.halo_send_e = p; p += 12 * row;
  b.halo_recv_s = p; p += 12 * row;
  b.halo_recv_e = p; p += 12 * row;
  b.carr_to_prev = p; p += 9 * EXP_ROWS * row;
  b.carr_to_next = p; p += 9 * EXP_ROWS * row;
  b.carr_from_prev = p; p += 9 * EXP_ROWS * row;
  b.carr_from_next = p;
  return b;
}

int exchange_edges(x3d2c_ctx* ctx, int dir, const double* const* fields, int nf, EdgeParams& ep, const DistBufs& b) {
  const int G = ctx->n_groups[dir];
  PackParams pp;
  for (int f = 0; f < 3; ++f) pp.f[f] = fields[f < nf ? f : 0];
  pp.n = ep.n;
  pp.n_pad = ep.n_pad;
  pp.nf = nf;
  halo_pack_kernel<<<dim3(G, nf), 128, 0, ctx->stream>>>(b.halo_send_s, b.halo_send_e, pp);
  X3D2C_CHECK_LAUNCH(ctx);
  int rc = x3d2c::sendrecv_fields(ctx, dir, b.halo_recv_s, b.halo_recv_e, b.halo_send_s, b.halo_send_e,
                                  (size_t)SZ * 4 * nf * G);
  if (rc) return rc;
  ep.G = G;
  ep.nf = nf;
  ep.halo_s = b.halo_recv_s;
  ep.halo_e = b.halo_recv_e;
  ep.to_prev = b.carr_to_prev;
  ep.to_next = b.carr_to_next;
  edge_kernel<<<dim3(G, ep.ns), dim3(32, 2 * DMAX), 0, ctx->stream>>>(ep);
  X3D2C_CHECK_LAUNCH(ctx);
  return x3d2c::sendrecv_fields(ctx, dir, b.carr_from_prev, b.carr_from_next, b.carr_to_prev, b.carr_to_next,
                                (size_t)SZ * EXP_ROWS * ep.ns * G);
}

}  // namespace m3
