// Shared pieces of the TMA-based fast-path kernels (transeq_m4.cu, tds_m4.cu): PTX wrappers for bulk tensor copies
// and mbarriers, and the tensor map that presents a directional field as (lane, segment, row in segment, group).
#pragma once
#include <cuda.h>

#include "m3_common.cuh"

namespace m4 {
using namespace m3;

extern __shared__ __align__(1024) double smem4[];

__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, unsigned src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, unsigned src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// window element t (row j0 - 4 + t, t = 0..S+7): rows 12..15 of the previous segment, own rows, rows 0..3 of the next
template <int NT>
__device__ __forceinline__ int woff4(int t, int bm, int b0, int bp) {
  return t < 4 ? bm + (S - 4 + t) * NT : (t < S + 4 ? b0 + (t - 4) * NT : bp + (t - S - 4) * NT);
}

// ------------------------------------------------------------------------------------------------ host side
bool make_line_map(CUtensorMap* m, const double* field, int L, int nseg, int n_pad, int groups);
// tile geometry for a line of n points: L lanes x nseg segments = NT threads (128, or 256 for n = 1024)
bool tile_shape(int n, int* L, int* NT);
// The tiles of line direction `line_dir` (L lanes x all segments x 16 rows of group g = c3 + nb * c4) inside a field
// stored in layout `layout_dir`, as a 5-D tensor (lane, segment, row in segment, c3, c4). Y and Z lines can address
// fields stored as DIR_Y, DIR_Z or DIR_C (all keep 32 consecutive x in a row): this is how a kernel reads or writes
// "through" a reorder. *nb returns the extent of c3 (tile coordinates: c3 = g % nb, c4 = g / nb).
bool make_map5(CUtensorMap* m, const double* field, int layout_dir, int line_dir, int L, int nseg,
               const x3d2c_ctx* ctx, int* nb);

// Line tiles whose line runs along the CONTIGUOUS index of the field's layout: x lines of a DIR_Y field (16 x per
// 128-byte row, L lanes in y) or y lines of a DIR_X field (16 y per row, L lanes in x), as a 5-D tensor
// (point in segment, lane, half of the 32-point row, block, z) with 128-byte swizzle (tds_m4.cu: XT kernels)
bool make_map_xt(CUtensorMap* m, const double* field, int layout_dir, int line_dir, int L, int nseg, const x3d2c_ctx* ctx);

}  // namespace m4
