// Fast path placeholder: filled in by the segment-parallel kernels. Until then the
// reference-order kernels of tds_m1.cu serve every shape.
#include "common.cuh"
namespace x3d2c {
int tds_solve_m3(x3d2c_ctx*, int, double*, const double*, const x3d2c_tdsops*) { return X3D2C_EUNSUPPORTED; }
int transeq_m3(x3d2c_ctx*, int, double*, double*, double*, const double*, const double*, const double*, double,
               const x3d2c_tdsops*, const x3d2c_tdsops*, const x3d2c_tdsops*, const x3d2c_tdsops*) {
  return X3D2C_EUNSUPPORTED;
}
}  // namespace x3d2c
