// Host layer of the cuda_c backend: constants of /root/reference/src/common.f90:12-88.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace x3d2h {

constexpr int SZ = 32;  // X3D2C_SZ, role of src/backend/cuda/common.f90:4

enum { DIR_X = 1, DIR_Y = 2, DIR_Z = 3, DIR_C = 4 };           // common.f90:27
enum { RDR_X2Y = 12, RDR_X2Z = 13, RDR_Y2X = 21, RDR_Y2Z = 23,  // common.f90:23-26
       RDR_Z2X = 31, RDR_Z2Y = 32, RDR_C2X = 41, RDR_C2Y = 42,
       RDR_C2Z = 43, RDR_X2C = 14, RDR_Y2C = 24, RDR_Z2C = 34 };
enum { VERT = 0, CELL = 1110, X_FACE = 1100, Y_FACE = 1010, Z_FACE = 110,  // common.f90:29-37
       X_EDGE = 10, Y_EDGE = 100, Z_EDGE = 1000, NULL_LOC = -1 };
enum { BC_PERIODIC = 0, BC_NEUMANN = 1, BC_DIRICHLET = 2, BC_HALO = -1 };  // common.f90:38-39

static const double pi = 4 * std::atan(1.0);  // common.f90:21

// common.f90:44-53: rdr code = 10*from + to for every pair the reference defines
inline void get_dirs_from_rdr(int& dir_from, int& dir_to, int rdr) {
  dir_from = rdr / 10;
  dir_to = rdr % 10;
}
inline int get_rdr_from_dirs(int from, int to) { return from == to ? 0 : 10 * from + to; }

// common.f90:84-88
inline int move_data_loc(int in_loc, int dir, int move) {
  int p = 1;
  for (int i = 0; i < dir; ++i) p *= 10;
  return in_loc + move * p;
}

[[noreturn]] inline void fail(const std::string& msg) { throw std::runtime_error(msg); }

}  // namespace x3d2h
