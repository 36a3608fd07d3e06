// Internal declarations of the cuda_c backend (libx3d2c.so). Not part of the ABI.
//
// Private directional layouts (SURVEY.md F2 lets a backend choose; SZ = 32 lanes):
//   DIR_X (i_l, j, g): y = yb*SZ + i_l, x = j, g = yb + nyb*z      idx = i_l + SZ*(x + nx_pad*(yb + nyb*z))
//   DIR_Y (i_l, j, g): x = xb*SZ + i_l, y = j, g = xb + nxb*z      idx = i_l + SZ*(y + ny_pad*(xb + nxb*z))
//   DIR_Z (i_l, j, g): x = xb*SZ + i_l, z = j, g = xb + nxb*y      idx = i_l + SZ*(z + nz*(xb + nxb*y))
//   DIR_C            : idx = x + nx_pad*(y + ny_pad*z)             (the only externally defined layout)
// Y, Z and C keep x fastest, so Y<->Z<->C are 256-byte granule permutations; anything involving X is a
// 32x32 tile transpose.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/x3d2c.h"

#define SZ X3D2C_SZ

namespace x3d2c {

void set_error(const std::string& msg);

#define X3D2C_CHECK_CUDA(expr)                                                                      \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      x3d2c::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                       std::to_string(__LINE__));                                                   \
      return X3D2C_ECUDA;                                                                           \
    }                                                                                               \
  } while (0)

#define X3D2C_CHECK_CUFFT(expr)                                                                     \
  do {                                                                                              \
    cufftResult r__ = (expr);                                                                       \
    if (r__ != CUFFT_SUCCESS) {                                                                     \
      x3d2c::set_error(std::string(#expr) + ": cufft error " + std::to_string((int)r__) + " at " +  \
                       __FILE__ + ":" + std::to_string(__LINE__));                                  \
      return X3D2C_ECUDA;                                                                           \
    }                                                                                               \
  } while (0)

#define X3D2C_REQUIRE(cond, msg)       \
  do {                                 \
    if (!(cond)) {                     \
      x3d2c::set_error(msg);           \
      return X3D2C_EINVAL;             \
    }                                  \
  } while (0)

// Every entry point makes the context's device current first: the host application (torch, another Sim on another
// GPU of the same process) may have switched devices between two calls.
#define X3D2C_ENTER(ctx)                                  \
  do {                                                    \
    if (ctx) X3D2C_CHECK_CUDA(cudaSetDevice((ctx)->device)); \
  } while (0)

#define X3D2C_CHECK_LAUNCH(ctx)                      \
  do {                                               \
    (ctx)->launches++;                               \
    X3D2C_CHECK_CUDA(cudaGetLastError());            \
  } while (0)

struct NcclApi;  // dlopen'ed entry points (nccl.cu)

// rows of SZ * max(n_groups) doubles in ctx->halo: the reference-order path carves 3 fields x 4 buffers x 4 rows + 2 x 18
// reduced-system rows (tds_m1.cu: carve), the rank-split fast path 4 x (3 x 4) halo rows + 4 x (9 x 3) carry rows
// (m3_common.cuh: kDistRows, checked there by a static_assert)
constexpr int kHaloRowsRef = 3 * 4 * 4 + 2 * 18;
constexpr int kHaloRowsDist = 4 * (3 * 4) + 4 * (9 * 3);
// + a second set of the RECEIVE buffers of the rank-split path (2 x 12 halo rows, 2 x 27 carry rows): with peer stores the
// neighbours write the next operator's rows while this rank still reads the current ones (m3_edge.cu)
constexpr int kHaloRowsRecv2 = 2 * (3 * 4) + 2 * (9 * 3);
// + the receive buffers of the in-kernel carry exchange (m3_common.cuh: InlineCarries): 2 sets x (from prev, from next) x
// 27 carry rows, kept at the sentinel value between uses
constexpr int kHaloRowsInline = 2 * 2 * (9 * 3);
constexpr int kHaloRows = (kHaloRowsDist > kHaloRowsRef ? kHaloRowsDist : kHaloRowsRef) + kHaloRowsRecv2 + kHaloRowsInline;
// A carry slot that has not been written yet holds this NaN (no computation produces this payload); see m3_common.cuh
constexpr unsigned long long kCarrySentinel = 0xFFFA5A5AFFFA5A5Aull;

constexpr int kMaxDevices = 64;  // per-device caches of launch attributes / occupancy are indexed by the device ordinal

}  // namespace x3d2c

// Device-side view of one tdsops_t (passed by value to kernels)
struct TdsDev {
  int n_tds, n_rhs;
  double coeffs[9];
  double coeffs_s[4][9];
  double coeffs_e[4][9];
  const double *fw, *bw, *sa, *sc, *af, *stretch, *stretch_correct;  // device arrays, 0-based
};

struct x3d2c_tdsops {
  int n_tds, n_rhs, move, periodic;
  TdsDev dev;
  double* d_block = nullptr;  // one allocation holding all seven arrays
  std::vector<double> h_fw, h_bw, h_sa, h_sc, h_af, h_stretch, h_stretch_correct;
  // fast-path tables (m3): generic first-order recurrences z_j = A_j z_{j-1} + B_j r_j, y_j = C_j y_{j+1} + E_j z_j
  // generic segment-parallel tables (tds_g.cu): per-row recurrences z_j = A_j z_{j-1} + B_j r_j, y_j = C_j y_{j+1} + E_j z_j,
  // carry weights, substitution vectors; built on first use. gen_state: 0 not built, 1 ready, -1 not eligible
  double* d_m3 = nullptr;
  double* d_stc = nullptr;  // stretch_correct padded to the processed rows (transeq)
  int gen_state = 0, gen_nseg = 0;
  std::vector<double> gen_scalars;
  int has_stretch = 0, has_stretch_correct = 0;
  unsigned tap_mask = 0x1ff;  // non-zero bulk taps
};

struct x3d2c_ctx {
  x3d2c_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t lane[3] = {nullptr, nullptr, nullptr};  // I/O lanes (x3d2c.h): [0] = stream, [1] upload, [2] download; created on use
  cudaEvent_t lane_ev[16] = {nullptr};
  int strict = 0;
  int force_dist = 0;  // X3D2C_FORCE_DIST: run the rank-split fast path on a single rank (self exchange); tests, profiling
  int nx_pad = 0, ny_pad = 0, nz_pad = 0;
  int n_groups[4] = {0, 0, 0, 0};  // [dir]
  long long ngrid = 0;
  long long launches = 0;
  // scratch
  // [0], [1]: dud / d2u blocks of transeq (omp/backend.f90:319-320), temporaries of the fused tds fallbacks;
  // [2..5]: reordered inputs / outputs of the *_r fallbacks (tds_fused.cu). Allocated on first use.
  double* scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* halo = nullptr;                   // packed halo / reduced-row exchange buffers
  size_t halo_doubles = 0;
  double* red = nullptr;       // device reduction scratch
  double* red_host = nullptr;  // pinned
  int red_blocks = 0;
  // multi-rank
  void* nccl_comm = nullptr;
  x3d2c::NcclApi* nccl = nullptr;
  // neighbours' exchange buffers mapped with CUDA IPC: halo and carry rows are stored straight into them by the pack /
  // edge kernels and announced with flags (m3_edge.cu); the flags live behind the halo rows in the same allocation
  bool halo_p2p = false;
  double* peer_halo[8] = {nullptr};
  unsigned long long* halo_flags = nullptr;
  unsigned long long* peer_halo_flags[8] = {nullptr};
  unsigned long long edge_epoch = 0;  // number of rank-split exchanges so far (all directions: the buffers are shared)

  int n_pad(int dir) const { return dir == X3D2C_DIR_X ? nx_pad : (dir == X3D2C_DIR_Y ? ny_pad : nz_pad); }
};

struct x3d2c_poisson {
  int nx, ny, nz;          // global cell dims
  int nxh;                 // nx/2 + 1
  int nz_loc, ny_loc;      // slab extents (physical z-slab, spectral y-slab)
  cufftHandle plan_r2c = 0, plan_c2r = 0, plan_y = 0, plan_z = 0;
  bool have_plans = false;
  cufftDoubleComplex *A = nullptr, *B = nullptr;  // A(nxh, ny, nz_loc) ; B(ny, nxh, nz) working layout
  double* waves = nullptr;                        // interleaved re/im in B's layout
  double *ax = nullptr, *bx = nullptr, *ay = nullptr, *by = nullptr, *az = nullptr, *bz = nullptr;
  double* compact = nullptr;  // un-padded real buffer when the DIR_C block is padded
  // multi-rank: the slab exchange writes straight into the peers' buffers over NVLink (CUDA IPC mappings)
  bool p2p = false;
  cufftDoubleComplex* peerA[8] = {nullptr};
  cufftDoubleComplex* peerB[8] = {nullptr};
  double* bar_word = nullptr;  // device word of the all-reduce barriers
  // pipelined exchange (P > 1, poisson.cu): the planes of the slab are transformed and sent chunk by chunk; the copies
  // (peer stores on a second stream) overlap the transforms of the next chunk; ranks signal each other with flags in peer
  // memory instead of all-reduce barriers
  bool pipe = false;
  int nch = 1;                                   // chunks of nz_loc / nch planes
  cufftDoubleComplex* Cx = nullptr;              // C(j_loc, i, k): destination of the forward exchange, spectral buffer
  cufftDoubleComplex* peerC[8] = {nullptr};
  unsigned long long* flags = nullptr;           // behind Cx in the same allocation: [0..7] forward done by rank r,
  unsigned long long* peerFlags[8] = {nullptr};  //   [8 + 16 r + c] backward chunk c delivered by rank r
  unsigned long long epoch = 0;
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_chunk[16] = {nullptr}, ev_z = nullptr;
  cufftHandle plan_r2c_c = 0, plan_c2r_c = 0, plan_y_c = 0;
  // non-periodic y (poisson010.cu)
  int bc_case = 0;      // 0: 000, 10: 010
  int stretched = 0;    // 0 uniform, 1 odd / even families, 2 one family ('bottom')
  int penta_rows = 0;   // rows per family
  double *fac_re = nullptr, *fac_im = nullptr;  // factorised pentadiagonal systems [line][family][row][5]; may alias
};

// internal launch helpers implemented across the .cu files
namespace x3d2c {
int launch_reorder(x3d2c_ctx* ctx, int dir_from, int dir_to, double* dst, const double* src, bool accumulate);
int ensure_scratch(x3d2c_ctx* ctx, int count = 2);
int ensure_scratch_slot(x3d2c_ctx* ctx, int i);
int get_dims_dataloc(const x3d2c_ctx* ctx, int data_loc, int dims[3], bool global);
int setup_peer_halo(x3d2c_ctx* ctx);    // nccl.cu: maps the peers' exchange buffers (P > 1)
void release_peer_halo(x3d2c_ctx* ctx);
// generic segment-parallel kernels for single-rank directions, any operator (tds_g.cu); EUNSUPPORTED -> tds_m1.cu
int tds_g(x3d2c_ctx* ctx, int dir, double* du, const double* u, const x3d2c_tdsops* ops);
int transeq_g(x3d2c_ctx* ctx, int dir, double* const out[3], const double* const in[3], double nu,
              const x3d2c_tdsops* der1st, const x3d2c_tdsops* der1st_sym, const x3d2c_tdsops* der2nd,
              const x3d2c_tdsops* der2nd_sym);
// shared part of x3d2c_poisson_create / x3d2c_poisson_create_010 (poisson.cu)
int poisson_create_common(x3d2c_ctx* ctx, int bc_case, const double* waves, const double* ax, const double* bx,
                          const double* ay, const double* by, const double* az, const double* bz, x3d2c_poisson** out);
}  // namespace x3d2c
