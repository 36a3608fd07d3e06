// Host-side helpers of the TMA-based kernels (m4_common.cuh).
#include "m4_common.cuh"

namespace m4 {

namespace {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

// (lane, segment, row in segment, group) view of a directional field: a box of (L, nseg, 16, 1) is one tile
}  // namespace

bool make_line_map(CUtensorMap* m, const double* field, int L, int nseg, int n_pad, int groups) {
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)SZ, (cuuint64_t)nseg, (cuuint64_t)S, (cuuint64_t)groups};
  const cuuint64_t strides[3] = {(cuuint64_t)S * SZ * 8, (cuuint64_t)SZ * 8, (cuuint64_t)n_pad * SZ * 8};
  const cuuint32_t box[4] = {(cuuint32_t)L, (cuuint32_t)nseg, (cuuint32_t)S, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  static int promo = -1;  // X3D2C_TMA_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B (tuning knob)
  if (promo < 0) {
    const char* e = std::getenv("X3D2C_TMA_L2PROMO");
    promo = e ? std::atoi(e) & 3 : 0;
  }
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(field), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


bool make_map5(CUtensorMap* m, const double* field, int layout_dir, int line_dir, int L, int nseg,
               const x3d2c_ctx* ctx, int* nb) {
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t R = (cuuint64_t)SZ * 8;  // one row of 32 lanes
  const cuuint64_t nxp = ctx->nx_pad, nyp = ctx->ny_pad, nz = ctx->nz_pad, nxb = nxp / SZ, nyb = nyp / SZ;
  cuuint64_t sk, s3, s4, d3, d4;  // byte strides of one line point, c3, c4; extents of c3, c4
  if (line_dir == X3D2C_DIR_X) {
    if (layout_dir != X3D2C_DIR_X) return false;
    d3 = nyb; d4 = nz; sk = R; s3 = nxp * R; s4 = nyb * nxp * R;
  } else if (line_dir == X3D2C_DIR_Y) {  // group = xb + nxb * z
    d3 = nxb; d4 = nz;
    if (layout_dir == X3D2C_DIR_Y) { sk = R; s3 = nyp * R; s4 = nxb * nyp * R; }
    else if (layout_dir == X3D2C_DIR_Z) { sk = R * nz * nxb; s3 = R * nz; s4 = R; }
    else if (layout_dir == X3D2C_DIR_C) { sk = nxp * 8; s3 = R; s4 = nxp * nyp * 8; }
    else return false;
  } else {  // Z lines, group = xb + nxb * y
    d3 = nxb; d4 = nyp;
    if (layout_dir == X3D2C_DIR_Z) { sk = R; s3 = nz * R; s4 = nxb * nz * R; }
    else if (layout_dir == X3D2C_DIR_Y) { sk = R * nyp * nxb; s3 = R * nyp; s4 = R; }
    else if (layout_dir == X3D2C_DIR_C) { sk = nxp * nyp * 8; s3 = R; s4 = nxp * 8; }
    else return false;
  }
  *nb = (int)d3;
  const cuuint64_t dims[5] = {(cuuint64_t)SZ, (cuuint64_t)nseg, (cuuint64_t)S, d3, d4};
  const cuuint64_t strides[4] = {S * sk, sk, s3, s4};
  const cuuint32_t box[5] = {(cuuint32_t)L, (cuuint32_t)nseg, (cuuint32_t)S, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(field), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tile_shape(int n, int* L, int* NT) {
  switch (n) {
    case 64: *L = 32; *NT = 128; return true;
    case 128: *L = 16; *NT = 128; return true;
    case 256: *L = 8; *NT = 128; return true;
    case 512: *L = 4; *NT = 128; return true;
    case 1024: *L = 4; *NT = 256; return true;
    default: return false;
  }
}

bool make_map_xt(CUtensorMap* m, const double* field, int layout_dir, int line_dir, int L, int nseg, const x3d2c_ctx* ctx) {
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t R = (cuuint64_t)SZ * 8;  // one row of 32 points
  const cuuint64_t nxp = ctx->nx_pad, nyp = ctx->ny_pad, nz = ctx->nz_pad;
  cuuint64_t lanes, blocks;  // extent of the lane index, 32-point blocks along the line
  if (line_dir == X3D2C_DIR_X && layout_dir == X3D2C_DIR_Y) {         // idx = x_l + 32 * (y + ny_pad * (xb + nxb * z))
    lanes = nyp; blocks = nxp / SZ;
  } else if (line_dir == X3D2C_DIR_Y && layout_dir == X3D2C_DIR_X) {  // idx = y_l + 32 * (x + nx_pad * (yb + nyb * z))
    lanes = nxp; blocks = nyp / SZ;
  } else {
    return false;
  }
  if ((cuuint64_t)nseg != 2 * blocks) return false;  // the line is the whole padded extent
  const cuuint64_t dims[5] = {16, lanes, 2, blocks, nz};
  const cuuint64_t strides[4] = {R, 128, lanes * R, blocks * lanes * R};
  const cuuint32_t box[5] = {16, (cuuint32_t)L, 2, (cuuint32_t)blocks, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(field), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace m4
