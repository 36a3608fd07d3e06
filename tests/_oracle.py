"""ctypes binding of oracle/libx3d2_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The library is the CPU restatement of the reference OMP backend (see oracle/x3d2_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libx3d2_oracle.so")

DIR_X, DIR_Y, DIR_Z, DIR_C = 1, 2, 3, 4
VERT, CELL = 0, 1110
X_FACE, Y_FACE, Z_FACE, X_EDGE, Y_EDGE, Z_EDGE = 1100, 1010, 110, 10, 100, 1000
BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_HALO = 0, 1, 2, -1
RDR = dict(X2Y=12, X2Z=13, Y2X=21, Y2Z=23, Z2X=31, Z2Y=32, C2X=41, C2Y=42, C2Z=43, X2C=14, Y2C=24, Z2C=34)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def _load():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_set_num_threads.argtypes = [C.c_int]
    lib.orc_set_num_threads.restype = None
    lib.orc_num_threads.restype = C.c_int
    lib.orc_tdsops_create.restype = C.c_void_p
    lib.orc_tdsops_create.argtypes = [C.c_int, C.c_double, C.c_char_p, C.c_char_p, C.c_int, C.c_int, _dp, _dp,
                                      C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.orc_tdsops_destroy.argtypes = [C.c_void_p]
    lib.orc_tdsops_info.argtypes = [C.c_void_p, _ip, _dp]
    lib.orc_tdsops_arrays.argtypes = [C.c_void_p] + [_dp] * 10
    lib.orc_lines_tds_solve.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, _dp, _dp]
    lib.orc_lines_transeq.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]
    lib.orc_world_create.restype = C.c_void_p
    lib.orc_world_create.argtypes = [_ip, _ip, _dp, _ip, C.c_double, C.c_double, C.c_char_p, C.c_char_p, C.c_char_p,
                                     C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), _dp]
    lib.orc_world_destroy.argtypes = [C.c_void_p]
    for name, args in dict(
        orc_world_init_tgv=[C.c_void_p],
        orc_world_set_uvw=[C.c_void_p, _dp, _dp, _dp],
        orc_world_get_uvw=[C.c_void_p, _dp, _dp, _dp],
        orc_world_step=[C.c_void_p, C.c_int],
        orc_world_set_case_channel=[C.c_void_p, C.c_double, C.c_int],
        orc_world_monitor=[C.c_void_p, _dp],
        orc_world_transeq=[C.c_void_p] + [_dp] * 6,
        orc_world_transeq_dir=[C.c_void_p, C.c_int] + [_dp] * 6,
        orc_world_transeq_lowmem=[C.c_void_p] + [_dp] * 7,
        orc_world_transeq_species=[C.c_void_p] + [_dp] * 4 + [C.c_double, _dp],
        orc_world_tds_solve=[C.c_void_p, C.c_int, C.c_char_p, C.c_int, _dp, _dp, _ip],
        orc_world_divergence=[C.c_void_p] + [_dp] * 4,
        orc_world_gradient=[C.c_void_p] + [_dp] * 4,
        orc_world_interpl_c2v=[C.c_void_p, _dp, _dp],
        orc_world_laplacian=[C.c_void_p, _dp, _dp],
        orc_world_curl=[C.c_void_p] + [_dp] * 6,
        orc_world_poisson=[C.c_void_p, _dp, _dp],
        orc_world_fft_roundtrip=[C.c_void_p, _dp, _dp, _dp],
        orc_world_spec_dims=[C.c_void_p, _ip],
        orc_world_waves=[C.c_void_p, _dp],
        orc_world_stretching_matrix=[C.c_void_p, _ip, _dp, _dp],
        orc_world_pressure_correction=[C.c_void_p],
        orc_world_reorder_chain=[C.c_void_p, _dp, _ip, C.c_int, _dp],
        orc_world_sum_intox=[C.c_void_p, C.c_int, _dp, _dp, _dp],
        orc_world_vecadd=[C.c_void_p, C.c_int, C.c_double, _dp, C.c_double, _dp, _dp],
        orc_world_scalar_product=[C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp],
        orc_world_field_max_mean=[C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp],
        orc_world_mesh_info=[C.c_void_p, C.c_int, _ip],
        orc_world_geo=[C.c_void_p, C.c_int] + [_dp] * 6,
    ).items():
        getattr(lib, name).argtypes = args
        getattr(lib, name).restype = C.c_int
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def set_num_threads(n):
    """omp_set_num_threads for the oracle (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    lib().orc_set_num_threads(int(n))


def num_threads():
    """Threads an OpenMP parallel region of the oracle really runs with."""
    return int(lib().orc_num_threads())


def _p(a):
    return a.ctypes.data_as(_dp)


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Tdsops:
    """tdsops_t built by the oracle's restatement of tdsops_init (src/tdsops.f90:63-203)."""

    def __init__(self, n_tds, delta, operation, scheme, bc_start, bc_end, stretch=None, stretch_correct=None,
                 n_halo=4, from_to=None, sym=False, c_nu=None, nu0_nu=None):
        st = _f(stretch) if stretch is not None else None
        sc = _f(stretch_correct) if stretch_correct is not None else None
        has_hv = c_nu is not None and nu0_nu is not None
        self.h = lib().orc_tdsops_create(n_tds, float(delta), operation.encode(), scheme.encode(), bc_start, bc_end,
                                         _p(st) if st is not None else None, _p(sc) if sc is not None else None,
                                         n_halo, from_to.encode() if from_to else None, int(sym), int(has_hv),
                                         float(c_nu or 0.0), float(nu0_nu or 0.0))
        if not self.h:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())
        info = (C.c_int * 4)()
        sc5 = (C.c_double * 5)()
        lib().orc_tdsops_info(self.h, info, sc5)
        self.n_tds, self.n_rhs, self.move, self.periodic = info[0], info[1], info[2], bool(info[3])
        self.alpha, self.a, self.b, self.c, self.d = list(sc5)
        n, m = self.n_rhs, self.n_tds
        self.coeffs = np.zeros(9)
        self.coeffs_s = np.zeros((4, 9))
        self.coeffs_e = np.zeros((4, 9))
        self.dist_fw, self.dist_bw, self.dist_sa, self.dist_sc, self.dist_af = (np.zeros(n) for _ in range(5))
        self.stretch = np.zeros(m)
        self.stretch_correct = np.zeros(m)
        lib().orc_tdsops_arrays(self.h, _p(self.coeffs), _p(self.coeffs_s), _p(self.coeffs_e), _p(self.dist_fw),
                                _p(self.dist_bw), _p(self.dist_sa), _p(self.dist_sc), _p(self.dist_af),
                                _p(self.stretch), _p(self.stretch_correct))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tdsops_destroy(self.h)
            self.h = None


def _handles(ops):
    arr = (C.c_void_p * len(ops))(*[o.h for o in ops])
    return arr


def lines_tds_solve(ops, u):
    """u: [P, n_lines, n_pad] -> du same shape; ops: list of P Tdsops (one per emulated rank)."""
    u = _f(u)
    P, n_lines, n_pad = u.shape
    du = np.zeros_like(u)
    _chk(lib().orc_lines_tds_solve(P, _handles(ops), n_lines, n_pad, _p(u), _p(du)))
    return du


def lines_transeq(ops_du, ops_dud, ops_d2u, nu, u, v):
    u, v = _f(u), _f(v)
    P, n_lines, n_pad = u.shape
    rhs = np.zeros_like(u)
    _chk(lib().orc_lines_transeq(P, _handles(ops_du), _handles(ops_dud), _handles(ops_d2u), float(nu), n_lines, n_pad,
                                 _p(u), _p(v), _p(rhs)))
    return rhs


class World:
    """P emulated MPI ranks of the reference OMP backend + solver. Arrays are numpy [nz, ny, nx] (x fastest)."""

    def __init__(self, dims, nproc_dir=(1, 1, 1), L=(2 * np.pi,) * 3, bcs=((0, 0), (0, 0), (0, 0)), Re=1600.0,
                 dt=1e-3, time_intg="RK3", der1st="compact6", der2nd="compact6", interpl="classic",
                 stagder="compact6", stretching=None, beta=None):
        self.dims = tuple(int(d) for d in dims)
        d3 = (C.c_int * 3)(*self.dims)
        p3 = (C.c_int * 3)(*nproc_dir)
        L3 = (C.c_double * 3)(*L)
        b6 = (C.c_int * 6)(*[b for pair in bcs for b in pair])
        st = None
        be = None
        if stretching is not None:
            st = (C.c_char_p * 3)(*[s.encode() for s in stretching])
            be = (C.c_double * 3)(*(beta or (1.0, 1.0, 1.0)))
        self.periodic = [pair[0] == BC_PERIODIC for pair in bcs]
        self.h = lib().orc_world_create(d3, p3, L3, b6, Re, dt, time_intg.encode(), der1st.encode(), der2nd.encode(),
                                        interpl.encode(), stagder.encode(), st, be)
        if not self.h:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_world_destroy(self.h)
            self.h = None

    # shapes -----------------------------------------------------------------------------------
    def shape(self, loc=VERT):
        nx, ny, nz = self.dims
        cx, cy, cz = [n if p else n - 1 for n, p in zip(self.dims, self.periodic)]
        d = {VERT: (nx, ny, nz), CELL: (cx, cy, cz), X_FACE: (nx, cy, cz), Y_FACE: (cx, ny, cz), Z_FACE: (cx, cy, nz),
             X_EDGE: (cx, ny, nz), Y_EDGE: (nx, cy, nz), Z_EDGE: (nx, ny, cz)}[loc]
        return (d[2], d[1], d[0])

    def _out(self, loc=VERT):
        return np.zeros(self.shape(loc))

    # solver -----------------------------------------------------------------------------------
    def init_tgv(self):
        _chk(lib().orc_world_init_tgv(self.h))

    def set_uvw(self, u, v, w):
        u, v, w = _f(u), _f(v), _f(w)
        _chk(lib().orc_world_set_uvw(self.h, _p(u), _p(v), _p(w)))

    def get_uvw(self):
        u, v, w = self._out(), self._out(), self._out()
        _chk(lib().orc_world_get_uvw(self.h, _p(u), _p(v), _p(w)))
        return u, v, w

    def set_case_channel(self, omega_rot=0.0, n_rotate=0):
        """case/channel.f90 hooks around every sub-stage: bulk-velocity correction, rotation forcing, wall values (zero)."""
        _chk(lib().orc_world_set_case_channel(self.h, omega_rot, n_rotate))

    def step(self, n=1):
        _chk(lib().orc_world_step(self.h, n))

    def monitor(self):
        out = np.zeros(4)
        _chk(lib().orc_world_monitor(self.h, _p(out)))
        return dict(enstrophy=out[0], ke=out[1], div_u_max=out[2], div_u_mean=out[3])

    def pressure_correction(self):
        _chk(lib().orc_world_pressure_correction(self.h))

    # ops --------------------------------------------------------------------------------------
    def transeq(self, u, v, w):
        u, v, w = _f(u), _f(v), _f(w)
        a, b, c = self._out(), self._out(), self._out()
        _chk(lib().orc_world_transeq(self.h, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c)))
        return a, b, c

    def transeq_dir(self, dir, u, v, w):
        u, v, w = _f(u), _f(v), _f(w)
        a, b, c = self._out(), self._out(), self._out()
        _chk(lib().orc_world_transeq_dir(self.h, dir, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c)))
        return a, b, c

    def transeq_lowmem(self, u, v, w):
        """solver.f90:391-505; returns (du, dv, dw, u after its round trip x -> y -> z -> x)."""
        u, v, w = _f(u), _f(v), _f(w)
        a, b, c, ub = self._out(), self._out(), self._out(), self._out()
        _chk(lib().orc_world_transeq_lowmem(self.h, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c), _p(ub)))
        return a, b, c, ub

    def transeq_species(self, u, v, w, spec, nu_s):
        u, v, w, spec = _f(u), _f(v), _f(w), _f(spec)
        d = self._out()
        _chk(lib().orc_world_transeq_species(self.h, _p(u), _p(v), _p(w), _p(spec), float(nu_s), _p(d)))
        return d

    def tds_solve(self, dir, opname, f, in_loc=VERT):
        f = _f(f)
        move = {"stagder_v2p": 1, "interpl_v2p": 1, "stagder_p2v": -1, "interpl_p2v": -1}.get(opname, 0)
        out_loc = in_loc + move * 10 ** dir
        out = self._out(out_loc)
        ol = C.c_int(0)
        _chk(lib().orc_world_tds_solve(self.h, dir, opname.encode(), in_loc, _p(f), _p(out), C.byref(ol)))
        assert ol.value == out_loc
        return out

    def divergence(self, u, v, w):
        u, v, w = _f(u), _f(v), _f(w)
        d = self._out(CELL)
        _chk(lib().orc_world_divergence(self.h, _p(u), _p(v), _p(w), _p(d)))
        return d

    def gradient(self, p):
        p = _f(p)
        a, b, c = self._out(), self._out(), self._out()
        _chk(lib().orc_world_gradient(self.h, _p(p), _p(a), _p(b), _p(c)))
        return a, b, c

    def interpl_c2v(self, p):
        """vector_calculus_t%interpl_c2v with the interpl_p2v operators (postprocess.f90:184-189): CELL -> VERT."""
        p = _f(p)
        a = self._out()
        _chk(lib().orc_world_interpl_c2v(self.h, _p(p), _p(a)))
        return a

    def laplacian(self, u):
        """vector_calculus_t%laplacian with the der2nd operators: VERT -> VERT."""
        u = _f(u)
        a = self._out()
        _chk(lib().orc_world_laplacian(self.h, _p(u), _p(a)))
        return a

    def curl(self, u, v, w):
        u, v, w = _f(u), _f(v), _f(w)
        a, b, c = self._out(), self._out(), self._out()
        _chk(lib().orc_world_curl(self.h, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c)))
        return a, b, c

    def poisson(self, f):
        f = _f(f)
        p = self._out(CELL)
        _chk(lib().orc_world_poisson(self.h, _p(f), _p(p)))
        return p

    def spec_dims(self):
        d = (C.c_int * 3)()
        lib().orc_world_spec_dims(self.h, d)
        return tuple(d)

    def waves(self):
        nx, ny, nz = self.spec_dims()
        w = np.zeros((nz, ny, nx, 2))
        lib().orc_world_waves(self.h, _p(w))
        return w[..., 0] + 1j * w[..., 1]

    def stretching_matrix(self):
        """a_odd / a_even (or a_re in a_odd for 'bottom') of poisson_fft.f90:275-652 as [5, nz, rows, nx_spec]."""
        nx, ny, nz = self.spec_dims()
        info = (C.c_int * 2)()
        _chk(lib().orc_world_stretching_matrix(self.h, info, None, None))
        rows = info[1]
        ao, ae = np.zeros((5, nz, rows, nx)), np.zeros((5, nz, rows, nx))
        _chk(lib().orc_world_stretching_matrix(self.h, info, _p(ao), _p(ae)))
        return dict(stretched=info[0], rows=rows, a_odd=ao, a_even=ae)

    def fft_roundtrip(self, f, want_spec=False):
        f = _f(f)
        out = self._out(CELL)
        spec = None
        if want_spec:
            nx, ny, nz = self.spec_dims()
            spec = np.zeros((nz, ny, nx, 2))
        _chk(lib().orc_world_fft_roundtrip(self.h, _p(f), _p(out), _p(spec) if want_spec else None))
        if want_spec:
            return out, spec[..., 0] + 1j * spec[..., 1]
        return out

    def reorder_chain(self, f, names):
        f = _f(f)
        out = self._out()
        r = (C.c_int * len(names))(*[RDR[n] for n in names])
        _chk(lib().orc_world_reorder_chain(self.h, _p(f), r, len(names), _p(out)))
        return out

    def sum_intox(self, dir_from, a, b):
        a, b = _f(a), _f(b)
        out = self._out()
        _chk(lib().orc_world_sum_intox(self.h, dir_from, _p(a), _p(b), _p(out)))
        return out

    def vecadd(self, dir, a, x, b, y):
        x, y = _f(x), _f(y)
        out = self._out()
        _chk(lib().orc_world_vecadd(self.h, dir, a, _p(x), b, _p(y), _p(out)))
        return out

    def scalar_product(self, dir, x, y, loc=VERT):
        x, y = _f(x), _f(y)
        s = C.c_double(0)
        _chk(lib().orc_world_scalar_product(self.h, dir, loc, _p(x), _p(y), C.byref(s)))
        return s.value

    def field_max_mean(self, dir, x, loc=VERT):
        x = _f(x)
        mx, mean = C.c_double(0), C.c_double(0)
        _chk(lib().orc_world_field_max_mean(self.h, dir, loc, _p(x), C.byref(mx), C.byref(mean)))
        return mx.value, mean.value

    def mesh_info(self, rank=0):
        out = (C.c_int * 18)()
        lib().orc_world_mesh_info(self.h, rank, out)
        o = list(out)
        return dict(vert_dims=o[0:3], cell_dims=o[3:6], padded=o[6:9], n_groups=o[9:12],
                    BCs=[o[12:14], o[14:16], o[16:18]])

    def geo(self, dir):
        """geo_t of emulated rank 0 along dir (0..2): mesh_content.f90:159-263 as restated in orc_world.hpp."""
        nv = self.shape()[2 - dir]
        nc = self.shape(CELL)[2 - dir]
        a = [np.zeros(nv) for _ in range(4)] + [np.zeros(nc) for _ in range(2)]
        _chk(lib().orc_world_geo(self.h, dir, *[_p(x) for x in a]))
        return dict(zip(("vert_coords", "vert_ds", "vert_ds2", "vert_d2s", "midp_coords", "midp_ds"), a))


def field_set_face(f, c_start, c_end, face="y"):
    """field_set_face_omp (omp/backend.f90:903-953) on a Cartesian [nz, ny, nx] array: Y_FACE sets the bottom row to
    c_start and the top row to c_end (the row meant by the TODO at :939 and set by the CUDA backend)."""
    if face != "y":
        raise RuntimeError("Setting X_FACE / Z_FACE is not yet supported.")
    g = np.array(f, dtype=np.float64, copy=True)
    g[:, 0, :] = c_start
    g[:, -1, :] = c_end
    return g


def field_set_face_from_field(f, f_start, c_end, face, flow_rate_diff=0.0):
    """field_set_face_from_field_omp (omp/backend.f90:954-1021) on Cartesian [nz, ny, nx] arrays."""
    g = np.array(f, dtype=np.float64, copy=True)
    if face == "y":      # :1003-1014: both walls from f_start
        g[:, 0, :] = f_start[:, 0, :]
        g[:, -1, :] = f_start[:, -1, :]
    elif face == "x":    # :981-1001: inlet from f_start, convective outflow
        g[:, :, 0] = f_start[:, :, 0]
        fd, fd1 = f[:, :, -1], f[:, :, -2]
        g[:, :, -1] = (fd - c_end * (fd - fd1)) + flow_rate_diff
    else:
        raise RuntimeError("field_set_face_from_field: only X_FACE and Y_FACE supported.")
    return g


def compute_vorticity(g):
    """compute_vorticity_omp (omp/backend.f90:616-630); g = (dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz)."""
    dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz = g
    return np.sqrt((dwdy - dvdz) * (dwdy - dvdz) + (dudz - dwdx) * (dudz - dwdx) + (dvdx - dudy) * (dvdx - dudy))


def compute_qcriterion(g):
    """compute_qcriterion_omp (omp/backend.f90:632-649)."""
    dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz = g
    return -0.5 * (dudx * dudx + dvdy * dvdy + dwdz * dwdz) - dudy * dvdx - dudz * dwdx - dvdz * dwdy


def slice_max_sum(f, dir, i_slice):
    """slice_max_sum_omp (omp/backend.f90:816-881) on a Cartesian [nz, ny, nx] array: signed max and sum of the plane
    i_slice (1-based) along `dir`."""
    pl = f[:, :, i_slice - 1] if dir == DIR_X else (f[:, i_slice - 1, :] if dir == DIR_Y else f[i_slice - 1, :, :])
    return float(pl.max()), float(pl.sum())
