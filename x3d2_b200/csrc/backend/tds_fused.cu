// tds_solve combinations that read or write "through" a reorder (x3d2c_tds_solve_r / _sum_r / _dual_r).
//
// Each call equals: reorder the input(s) with rdr_in, apply the operator(s) in direction `dir`, reorder the output(s)
// with rdr_out (rdr = 0: no reorder). On the fast path the TMA kernels of tds_m4.cu store into the foreign layout
// directly through a 5-D tensor map (Y, Z and C layouts keep 32 consecutive x per row, so a Y- or Z-line tile is a
// box in all three): the output reorder passes (16 B per point each) disappear. Everywhere else the call runs as
// exactly the sequence above through scratch fields, so the values are those of the reference's call sequence
// (divergence_v2c / gradient_c2v, src/vector_calculus.f90:142-332; poisson_fft, src/solver.f90:741-775).
#include "common.cuh"

namespace x3d2c {
int tds_m4(x3d2c_ctx* ctx, int dir, int mode, double* out_a, double* out_b, const double* in_a, const double* in_b,
           const x3d2c_tdsops* ta, const x3d2c_tdsops* tb, double scale_a, int lay_in, int lay_out);
}

using namespace x3d2c;

namespace {

// validates the reorder codes against `dir`; returns the layouts of the inputs and outputs
int layouts(int dir, int rdr_in, int rdr_out, int* lay_in, int* lay_out) {
  *lay_in = dir;
  *lay_out = dir;
  if (rdr_in) {
    if (rdr_in % 10 != dir || rdr_in / 10 < 1 || rdr_in / 10 > 4 || rdr_in / 10 == dir) {
      set_error("fused tds_solve: rdr_in must be a reorder code that ends in dir");
      return X3D2C_EINVAL;
    }
    *lay_in = rdr_in / 10;
  }
  if (rdr_out) {
    if (rdr_out / 10 != dir || rdr_out % 10 < 1 || rdr_out % 10 > 4 || rdr_out % 10 == dir) {
      set_error("fused tds_solve: rdr_out must be a reorder code that starts from dir");
      return X3D2C_EINVAL;
    }
    *lay_out = rdr_out % 10;
  }
  return X3D2C_OK;
}

// mode 0 single, 1 sum, 2 dual (tds_m4.cu numbering)
int run(x3d2c_ctx* ctx, const char* what, int dir, int mode, double* out_a, double* out_b, const double* in_a,
        const double* in_b, const x3d2c_tdsops* op_a, const x3d2c_tdsops* op_b, int rdr_in, int rdr_out) {
  int lay_in, lay_out;
  int rc = layouts(dir, rdr_in, rdr_out, &lay_in, &lay_out);
  if (rc) return rc;
  static const bool trace = std::getenv("X3D2C_TRACE") != nullptr;
  auto report = [&](const char* how) {
    if (trace) std::fprintf(stderr, "[x3d2c] %s dir=%d rdr_in=%d rdr_out=%d -> %s\n", what, dir, rdr_in, rdr_out, how);
  };
  // Inputs are reordered explicitly: reading through a tensor map was measured slower than the separate pass (a
  // c2z input has its rows one z-plane = 2 MB apart; pressure correction 12.15 ms against 12.02 ms), and rank-split
  // directions need their inputs in the direction's own layout anyway. The outputs go through the tensor map.
  const double *a = in_a, *b = in_b;
  // ... except x lines: X <-> Y is a 32 x 32 tile transpose, which the swizzled tiles of tds_m4.cu (XT) do on the way in
  // or out at no cost (divergence: x2y on the output; gradient: y2x on the input)
  if (!ctx->strict && dir == X3D2C_DIR_X && mode == 0 && (rdr_in != 0) != (rdr_out != 0)) {
    rc = tds_m4(ctx, dir, mode, out_a, out_b, a, b, op_a, op_b, 1.0, lay_in, lay_out);
    if (rc != X3D2C_EUNSUPPORTED) {
      report(rdr_in ? "input through the swizzled tensor map" : "output through the swizzled tensor map");
      return rc;
    }
  }
  if (rdr_in) {
    if ((rc = ensure_scratch_slot(ctx, 2))) return rc;
    if ((rc = x3d2c_reorder(ctx, rdr_in, ctx->scratch[2], in_a))) return rc;
    a = ctx->scratch[2];
    if (mode == 1) {
      if ((rc = ensure_scratch_slot(ctx, 3))) return rc;
      if ((rc = x3d2c_reorder(ctx, rdr_in, ctx->scratch[3], in_b))) return rc;
      b = ctx->scratch[3];
    }
  }
  if (!ctx->strict && rdr_out) {
    rc = tds_m4(ctx, dir, mode, out_a, out_b, a, b, op_a, op_b, 1.0, dir, lay_out);
    if (rc != X3D2C_EUNSUPPORTED) {
      report(rdr_in ? "input reorder pass, output through the tensor map" : "output through the tensor map");
      return rc;
    }
  }
  if (rdr_in || rdr_out) report("reorder + operator sequence");
  if (rdr_out && ((rc = ensure_scratch_slot(ctx, 4)) || (rc = ensure_scratch_slot(ctx, 5)))) return rc;
  double* oa = rdr_out ? ctx->scratch[4] : out_a;
  double* ob = rdr_out ? ctx->scratch[5] : out_b;
  if (mode == 0) rc = x3d2c_tds_solve(ctx, dir, oa, a, op_a);
  else if (mode == 1) rc = x3d2c_tds_solve_sum(ctx, dir, oa, a, op_a, b, op_b);
  else rc = x3d2c_tds_solve_dual(ctx, dir, oa, ob, a, op_a, op_b);
  if (rc) return rc;
  if (rdr_out) {
    if ((rc = x3d2c_reorder(ctx, rdr_out, out_a, oa))) return rc;
    if (mode == 2 && (rc = x3d2c_reorder(ctx, rdr_out, out_b, ob))) return rc;
  }
  return X3D2C_OK;
}

}  // namespace

extern "C" {

int x3d2c_tds_solve_r(x3d2c_ctx* ctx, int dir, double* out, const double* in, const x3d2c_tdsops* op, int rdr_in,
                      int rdr_out) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out && in && op, "x3d2c_tds_solve_r: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_r: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(out != in, "x3d2c_tds_solve_r: out and in must be different fields");
  return run(ctx, "tds_solve_r", dir, 0, out, nullptr, in, nullptr, op, op, rdr_in, rdr_out);
}

int x3d2c_tds_solve_sum_r(x3d2c_ctx* ctx, int dir, double* out, const double* in_a, const x3d2c_tdsops* op_a,
                          const double* in_b, const x3d2c_tdsops* op_b, int rdr_in, int rdr_out) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out && in_a && in_b && op_a && op_b, "x3d2c_tds_solve_sum_r: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_sum_r: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(out != in_a && out != in_b, "x3d2c_tds_solve_sum_r: out must differ from the inputs");
  return run(ctx, "tds_solve_sum_r", dir, 1, out, nullptr, in_a, in_b, op_a, op_b, rdr_in, rdr_out);
}

int x3d2c_tds_solve_dual_r(x3d2c_ctx* ctx, int dir, double* out_a, double* out_b, const double* in,
                           const x3d2c_tdsops* op_a, const x3d2c_tdsops* op_b, int rdr_in, int rdr_out) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && out_a && out_b && in && op_a && op_b, "x3d2c_tds_solve_dual_r: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_dual_r: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(out_a != in && out_b != in && out_a != out_b, "x3d2c_tds_solve_dual_r: fields must be distinct");
  return run(ctx, "tds_solve_dual_r", dir, 2, out_a, out_b, in, nullptr, op_a, op_b, rdr_in, rdr_out);
}

int x3d2c_tds_solve_axpy_r(x3d2c_ctx* ctx, int dir, double* y, double a, const double* in, const x3d2c_tdsops* op,
                           int rdr_in) {
  X3D2C_ENTER(ctx);
  X3D2C_REQUIRE(ctx && y && in && op, "x3d2c_tds_solve_axpy_r: null argument");
  X3D2C_REQUIRE(dir >= 1 && dir <= 3, "x3d2c_tds_solve_axpy_r: dir must be DIR_X/Y/Z");
  X3D2C_REQUIRE(y != in, "x3d2c_tds_solve_axpy_r: y and in must be different fields");
  if (!rdr_in) return x3d2c_tds_solve_axpy(ctx, dir, y, a, in, op);
  int lay_in, lay_out;
  int rc = layouts(dir, rdr_in, 0, &lay_in, &lay_out);
  if (rc) return rc;
  static const bool trace = std::getenv("X3D2C_TRACE") != nullptr;
  if (!ctx->strict && dir == X3D2C_DIR_X) {
    rc = tds_m4(ctx, dir, 3, y, nullptr, in, y, op, op, a, lay_in, dir);
    if (rc != X3D2C_EUNSUPPORTED) {
      if (trace) std::fprintf(stderr, "[x3d2c] tds_solve_axpy_r dir=%d rdr_in=%d -> input through the swizzled tensor map\n", dir, rdr_in);
      return rc;
    }
  }
  if (trace) std::fprintf(stderr, "[x3d2c] tds_solve_axpy_r dir=%d rdr_in=%d -> reorder + operator sequence\n", dir, rdr_in);
  if ((rc = ensure_scratch_slot(ctx, 2))) return rc;
  if ((rc = x3d2c_reorder(ctx, rdr_in, ctx->scratch[2], in))) return rc;
  return x3d2c_tds_solve_axpy(ctx, dir, y, a, ctx->scratch[2], op);
}

}  // extern "C"
