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

// flags of the peer exchange: slot = ((dir - 1) * 2 + phase) * 2 + side; side 0: written by the previous rank, 1: by the next
__global__ void edge_signal_kernel(unsigned long long* to_prev, unsigned long long* to_next, const unsigned long long v) {
  __threadfence_system();  // the stores of the preceding kernel on this stream are complete
  *reinterpret_cast<volatile unsigned long long*>(threadIdx.x == 0 ? to_prev : to_next) = v;
  __threadfence_system();
}
__global__ void edge_wait_kernel(const unsigned long long* from_prev, const unsigned long long* from_next,
                                 const unsigned long long v) {
  const volatile unsigned long long* f = threadIdx.x == 0 ? from_prev : from_next;
  while (*f < v) __nanosleep(100);
  __threadfence_system();
}

// first / last four rows of nf directional fields -> (SZ, 4, nf, G). A block of 256 threads moves the eight rows of
// one field of kPackGroups line groups (the destination may be a neighbour's buffer: 64 x 256-byte rows per block keep
// enough stores in flight for NVLink; one group per block was launch-bound, 0.30 ms at 1024 x 1024 lines).
constexpr int kPackGroups = 8;
__global__ void __launch_bounds__(256) halo_pack_kernel(double* __restrict__ send_s, double* __restrict__ send_e,
                                                        const int G, const __grid_constant__ PackParams p) {
  const int lane = threadIdx.x & 31, row = (threadIdx.x >> 5) & 3, end = threadIdx.x >> 7, f = blockIdx.y;
  const double* src = p.f[f] + (size_t)(end ? p.n - 4 + row : row) * SZ + lane;
  double* dst = end ? send_e : send_s;
  double v[kPackGroups];
  const int g0 = blockIdx.x * kPackGroups;
#pragma unroll
  for (int i = 0; i < kPackGroups; ++i)
    if (g0 + i < G) v[i] = __ldcs(src + (size_t)SZ * p.n_pad * (g0 + i));
#pragma unroll
  for (int i = 0; i < kPackGroups; ++i)
    if (g0 + i < G) dst[(((size_t)(g0 + i) * p.nf + f) * 4 + row) * SZ + lane] = v[i];
}

// One thread per (lane, edge segment): segments 0..2 and nseg-3..nseg-1 of every line; blockIdx.y = field.
// NR recurrences per field. TRANSEQ (NR = 3): ops[0] on f, ops[1] on f * f[0], ops[2] on f. Otherwise recurrence k
// of field fi applies ops[fi * NR + k] to f (tds_solve: one field, one operator; pair kernels: two fields with one
// operator each, or one field with two operators).
// The 24-row window is loaded up front (rows outside the line come from the received halos), then swept.
template <int NR, bool TRANSEQ>
__global__ void __launch_bounds__(32 * 2 * DMAX) edge_kernel(const __grid_constant__ EdgeParams p) {
  const int lane = threadIdx.x, e = threadIdx.y, g = blockIdx.x, fi = blockIdx.y;
  const int q = e < DMAX ? e : p.nseg - 2 * DMAX + e;
  const int j0 = q * S;
  const size_t base = (size_t)g * p.n_pad * SZ + lane;
  const size_t hf = ((size_t)g * p.nf + fi) * 4 * SZ + lane, hc = (size_t)g * p.nf * 4 * SZ + lane;
  const double* f = p.f[fi] + base + (size_t)j0 * SZ;
  const double* c = p.f[0] + base + (size_t)j0 * SZ;
  double w[S + 8], wc[S + 8];
#pragma unroll
  for (int t = 0; t < S + 8; ++t) {
    const int r = t - 4;  // row relative to j0
    const double *pf, *pc;
    if (t < 4 && e == 0) { pf = p.halo_s + hf + (size_t)t * SZ; pc = p.halo_s + hc + (size_t)t * SZ; }
    else if (t >= S + 4 && e == 2 * DMAX - 1) { pf = p.halo_e + hf + (size_t)(t - S - 4) * SZ; pc = p.halo_e + hc + (size_t)(t - S - 4) * SZ; }
    else { pf = f + (ptrdiff_t)r * SZ; pc = c + (ptrdiff_t)r * SZ; }
    w[t] = *pf;
    if (TRANSEQ) wc[t] = *pc;
  }
  if (TRANSEQ) {
#pragma unroll
    for (int t = 0; t < S + 8; ++t) wc[t] *= w[t];
  }
  __shared__ double sh[NR][2 * DMAX - 1][32];  // ys(0..2), ze(0..1) of the first segments
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const Op& o = p.ops[TRANSEQ ? k : fi * NR + k];
    const double* src = (TRANSEQ && k == 1) ? wc : w;
    double z[S], pz = 0.0;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      double win[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) win[t] = src[i + t];
      pz = fma(o.a, pz, o.exact ? o.fs * sten_exact<0x1FFu>(o.cfw, win) : sten<0x1FFu>(o.cfw, win));
      z[i] = pz;
    }
    if (e >= DMAX) {  // ze of the last three segments, for the next rank
      p.to_next[(((size_t)g * p.ns + fi * NR + k) * EXP_ROWS + (e - DMAX)) * SZ + lane] = z[S - 1];
    } else {
      if (e < DMAX - 1) sh[k][DMAX + e][lane] = z[S - 1];
      double y = 0.0;
#pragma unroll
      for (int i = S - 1; i >= 0; --i) y = fma(o.cb, y, z[i]);
      sh[k][e][lane] = y;
    }
  }
  __syncthreads();
  if (e < DMAX) {
    // What this rank's first segments add to yin of the previous rank's segment nseg' - 3 + e (carries() in
    // m3_common.cuh with q + d >= nseg', q + m >= nseg'): sum_{d >= 3-e} yw[d-1] ys(e+d-3) + sum_{m >= 3-e} om[m+2] ze(e+m-3)
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const Op& o = p.ops[TRANSEQ ? k : fi * NR + k];
      double acc = 0.0;
#pragma unroll
      for (int d = 1; d <= DMAX; ++d)
        if (e + d - DMAX >= 0) acc = fma(o.yw[d - 1], sh[k][e + d - DMAX][lane], acc);
#pragma unroll
      for (int m = 1; m <= DMAX - 1; ++m)
        if (e + m - DMAX >= 0) acc = fma(o.om[m + DMAX - 1], sh[k][DMAX + e + m - DMAX][lane], acc);
      p.to_prev[(((size_t)g * p.ns + fi * NR + k) * EXP_ROWS + e) * SZ + lane] = acc;
    }
  }
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
bool make_op(const x3d2c_tdsops* t, double scale, bool dist, Op* o, bool exact) {
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
  for (int k = 0; k < 9; ++k) o->cfw[k] = exact ? t->dev.coeffs[k] : scale * fw * t->dev.coeffs[k];
  o->fs = exact ? scale * fw : 1.0;
  o->exact = exact ? 1 : 0;
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
  static int sms_dev[x3d2c::kMaxDevices] = {};
  int& sms = sms_dev[ctx->device];
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  return sms > 0 ? sms : 148;
}

// every rank must take the same decision: equal split of a periodic direction, enough segments per rank
bool dist_supported(const x3d2c_ctx* ctx, int dir, int n) {
  const int d = dir - 1, P = ctx->cfg.nproc_dir[d];
  if ((P <= 1 && !ctx->force_dist) || !ctx->cfg.periodic[d] || ctx->strict) return false;
  if (P > 1 && !ctx->nccl_comm) return false;
  if (ctx->cfg.dims_vert_global[d] != n * P || ctx->cfg.dims_vert[d] != n) return false;
  return n % S == 0 && n >= 2 * DMAX * S;
}

DistBufs carve_dist(x3d2c_ctx* ctx, int recv_set) {
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
  b.carr_from_next = p; p += 9 * EXP_ROWS * row;
  if (recv_set) {  // the alternate receive buffers live behind the first set (common.cuh: kHaloRowsRecv2)
    p = ctx->halo + (size_t)(x3d2c::kHaloRows - x3d2c::kHaloRowsInline - x3d2c::kHaloRowsRecv2) * row;
    b.halo_recv_s = p; p += 12 * row;
    b.halo_recv_e = p; p += 12 * row;
    b.carr_from_prev = p; p += 9 * EXP_ROWS * row;
    b.carr_from_next = p; p += 9 * EXP_ROWS * row;
  }
  if ((size_t)(p - ctx->halo) > ctx->halo_doubles) { std::fprintf(stderr, "x3d2c: halo buffer overflow in carve_dist()\n"); std::abort(); }
  return b;
}

int exchange_edges(x3d2c_ctx* ctx, int dir, const double* const* fields, int nf, EdgeParams& ep, DistBufs& b,
                   InlineCarries* inl) {
  const int G = ctx->n_groups[dir];
  static const bool no_inline = std::getenv("X3D2C_NO_INLINE_CARRIES") != nullptr;
  // Peer-store path (P > 1, buffers mapped): the pack kernel writes this rank's first / last four rows straight into
  // the neighbours' receive buffers, the edge kernel its carries; each phase is announced with one flag per neighbour
  // and awaited with a two-thread kernel. The receive buffers alternate between two sets (see common.cuh).
  const bool peer = ctx->halo_p2p && ctx->cfg.nproc_dir[dir - 1] > 1;
  double *dst_halo_s = b.halo_send_s, *dst_halo_e = b.halo_send_e, *dst_to_prev = b.carr_to_prev, *dst_to_next = b.carr_to_next;
  unsigned long long epoch = 0;
  unsigned long long *sig_prev[2] = {nullptr, nullptr}, *sig_next[2] = {nullptr, nullptr}, *my_flag[2][2] = {{nullptr}};
  epoch = ++ctx->edge_epoch;
  const int set = (int)(epoch & 1);
  const bool use_inline = inl && !no_inline && (peer || ctx->cfg.nproc_dir[dir - 1] == 1);
  if (inl) *inl = InlineCarries{};
  if (use_inline) {
    int ng = ctx->n_groups[1] > ctx->n_groups[2] ? ctx->n_groups[1] : ctx->n_groups[2];
    if (ctx->n_groups[3] > ng) ng = ctx->n_groups[3];
    const size_t row = (size_t)SZ * ng;
    double* base = ctx->halo + (size_t)(x3d2c::kHaloRows - x3d2c::kHaloRowsInline) * row + (size_t)set * 2 * 9 * EXP_ROWS * row;
    inl->from_prev = base;
    inl->from_next = base + 9 * EXP_ROWS * row;
    const int prev = ctx->cfg.pprev[dir - 1], next = ctx->cfg.pnext[dir - 1];
    inl->to_prev = peer ? ctx->peer_halo[prev] + (inl->from_next - ctx->halo) : inl->from_next;
    inl->to_next = peer ? ctx->peer_halo[next] + (inl->from_prev - ctx->halo) : inl->from_prev;
  }
  if (peer) {
    b = carve_dist(ctx, set);
    const int prev = ctx->cfg.pprev[dir - 1], next = ctx->cfg.pnext[dir - 1];
    auto remote = [&](int r, const double* local) { return ctx->peer_halo[r] + (local - ctx->halo); };
    dst_halo_s = remote(prev, b.halo_recv_e);   // my first rows are the previous rank's "after the line" rows
    dst_halo_e = remote(next, b.halo_recv_s);
    dst_to_prev = remote(prev, b.carr_from_next);
    dst_to_next = remote(next, b.carr_from_prev);
    for (int ph = 0; ph < 2; ++ph) {
      const int slot = ((dir - 1) * 2 + ph) * 2;
      sig_prev[ph] = ctx->peer_halo_flags[prev] + slot + 1;  // I am prev's next rank
      sig_next[ph] = ctx->peer_halo_flags[next] + slot + 0;
      my_flag[ph][0] = ctx->halo_flags + slot + 0;
      my_flag[ph][1] = ctx->halo_flags + slot + 1;
    }
  }
  // X3D2C_TRACE_TIMES=1: device time of the four phases on stderr (debugging aid; synchronises)
  static const bool timing = std::getenv("X3D2C_TRACE_TIMES") != nullptr;
  cudaEvent_t ev[5];
  auto mark = [&](int i) {
    if (timing) { if (i == 0) for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[i], ctx->stream); }
  };
  mark(0);
  PackParams pp;
  for (int f = 0; f < 3; ++f) pp.f[f] = fields[f < nf ? f : 0];
  pp.n = ep.n;
  pp.n_pad = ep.n_pad;
  pp.nf = nf;
  halo_pack_kernel<<<dim3((G + kPackGroups - 1) / kPackGroups, nf), 256, 0, ctx->stream>>>(dst_halo_s, dst_halo_e, G, pp);
  X3D2C_CHECK_LAUNCH(ctx);
  mark(1);
  int rc = X3D2C_OK;
  if (peer) {
    edge_signal_kernel<<<1, 2, 0, ctx->stream>>>(sig_prev[0], sig_next[0], epoch);
    X3D2C_CHECK_LAUNCH(ctx);
    edge_wait_kernel<<<1, 2, 0, ctx->stream>>>(my_flag[0][0], my_flag[0][1], epoch);
    X3D2C_CHECK_LAUNCH(ctx);
  } else {
    rc = x3d2c::sendrecv_fields(ctx, dir, b.halo_recv_s, b.halo_recv_e, b.halo_send_s, b.halo_send_e,
                                (size_t)SZ * 4 * nf * G);
    if (rc) return rc;
  }
  mark(2);
  if (use_inline) {
    if (timing) {
      cudaEventSynchronize(ev[2]);
      float t[2];
      for (int i = 0; i < 2; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
      std::fprintf(stderr, "[x3d2c] rank %d edges dir=%d ns=%d: pack %.3f ms, halo exchange %.3f, carries in the main kernel\n",
                   ctx->cfg.rank, dir, ep.ns, t[0], t[1]);
      for (auto& e : ev) cudaEventDestroy(e);
    }
    return X3D2C_OK;
  }
  ep.G = G;
  ep.nf = nf;
  ep.halo_s = b.halo_recv_s;
  ep.halo_e = b.halo_recv_e;
  ep.to_prev = dst_to_prev;
  ep.to_next = dst_to_next;
  const dim3 eg(G, nf), eb(32, 2 * DMAX);
  if (ep.transeq)
    edge_kernel<3, true><<<eg, eb, 0, ctx->stream>>>(ep);
  else if (ep.ns == 2 * nf)
    edge_kernel<2, false><<<eg, eb, 0, ctx->stream>>>(ep);
  else
    edge_kernel<1, false><<<eg, eb, 0, ctx->stream>>>(ep);
  X3D2C_CHECK_LAUNCH(ctx);
  mark(3);
  if (peer) {
    edge_signal_kernel<<<1, 2, 0, ctx->stream>>>(sig_prev[1], sig_next[1], epoch);
    X3D2C_CHECK_LAUNCH(ctx);
    edge_wait_kernel<<<1, 2, 0, ctx->stream>>>(my_flag[1][0], my_flag[1][1], epoch);
    X3D2C_CHECK_LAUNCH(ctx);
  } else {
    rc = x3d2c::sendrecv_fields(ctx, dir, b.carr_from_prev, b.carr_from_next, b.carr_to_prev, b.carr_to_next,
                                (size_t)SZ * EXP_ROWS * ep.ns * G);
  }
  mark(4);
  if (timing) {
    cudaEventSynchronize(ev[4]);
    float t[4];
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
    std::fprintf(stderr, "[x3d2c] rank %d edges dir=%d ns=%d: pack %.3f ms, halo exchange %.3f, edge kernel %.3f, "
                 "carry exchange %.3f\n", ctx->cfg.rank, dir, ep.ns, t[0], t[1], t[2], t[3]);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return rc;
}

}  // namespace m3
