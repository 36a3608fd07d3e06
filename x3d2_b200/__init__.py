"""x3d2_b200 — Python binding of the B200-native `cuda_c` backend for x3d2's per-timestep hot path.

Two in-tree shared libraries carry the product:
  libx3d2c.so  hand-written CUDA (sm_100a) kernels behind the C ABI of include/x3d2c.h — the drop-in
               for the reference's abstract backend (src/backend/backend.f90:13-62);
  libx3d2h.so  C++ host layer mirroring the reference's solver-side modules (src/solver.f90,
               src/time_integrator.f90, src/vector_calculus.f90, src/tdsops.f90, ...), include/x3d2h.h.
This module only marshals numpy arrays; there is no Python/CPU fallback for any operator: if the
extension is missing or no GPU is visible, creating a `Sim` raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import lib as _libmod
from .lib import (BC_DIRICHLET, BC_HALO, BC_NEUMANN, BC_PERIODIC, CELL, DIR_C, DIR_X, DIR_Y, DIR_Z, FLAG_BASE_OPS, FLAG_STRICT, RDR,
                  VERT, X3D2HConfig, build, load)

__all__ = ["Sim", "build", "load", "tdsops_tables", "poisson_tables_010", "decompose", "geo", "waves_000", "DIR_X", "DIR_Y", "DIR_Z", "DIR_C", "VERT",
           "CELL", "BC_PERIODIC", "BC_NEUMANN", "BC_DIRICHLET", "BC_HALO", "FLAG_STRICT", "FLAG_BASE_OPS", "RDR"]

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a, shape=None):
    """float64, C-contiguous; with `shape` the extents are checked: the host layer copies shape-many doubles."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"x3d2_b200: array of shape {a.shape} where the rank-local extents are {tuple(shape)}")
    return a


def _out_ok(a, shape):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == tuple(shape)):
        raise ValueError(f"x3d2_b200: output arrays must be C-contiguous float64 of shape {tuple(shape)}")
    return a


def _chk(rc):
    if rc != 0:
        h = load()[1]
        raise RuntimeError("x3d2_b200: " + h.x3d2h_last_error().decode())


def _config(dims, nproc_dir, L, bcs, Re, dt, time_intg, der1st, der2nd, interpl, stagder, rank, nproc, device, flags,
            nccl_id, stretching=None, beta=None):
    cfg = X3D2HConfig()
    cfg.dims_global = (C.c_int * 3)(*[int(d) for d in dims])
    cfg.nproc_dir = (C.c_int * 3)(*nproc_dir)
    cfg.L_global = (C.c_double * 3)(*L)
    cfg.bc = (C.c_int * 6)(*[b for pair in bcs for b in pair])
    cfg.Re, cfg.dt = Re, dt
    cfg.time_intg = time_intg.encode()
    cfg.der1st_scheme = der1st.encode()
    cfg.der2nd_scheme = der2nd.encode()
    cfg.interpl_scheme = interpl.encode()
    cfg.stagder_scheme = stagder.encode()
    cfg.rank, cfg.nproc, cfg.device, cfg.flags = rank, nproc, device, flags
    cfg.nccl_unique_id = C.cast(C.c_char_p(nccl_id), C.c_void_p) if nccl_id else None
    st = stretching or ("uniform",) * 3
    cfg.stretching = (C.c_char_p * 3)(*[s.encode() for s in st])
    cfg.beta = (C.c_double * 3)(*(beta or (1.0, 1.0, 1.0)))
    return cfg


def decompose(dims, nproc_dir, rank, bcs=((0, 0), (0, 0), (0, 0)), L=(2 * np.pi,) * 3):
    """mesh_t decomposition of the host layer (no GPU needed)."""
    nproc = int(np.prod(nproc_dir))
    cfg = _config(dims, nproc_dir, L, bcs, 1600.0, 1e-3, "RK3", "compact6", "compact6", "classic", "compact6", rank,
                  nproc, -1, 0, None)
    out = (C.c_int * 24)()
    _chk(load()[1].x3d2h_decompose(C.byref(cfg), out))
    o = list(out)
    return dict(vert_dims=o[0:3], cell_dims=o[3:6], n_offset=o[6:9], nrank_dir=o[9:12], pprev=o[12:15], pnext=o[15:18],
                BCs=[o[18:20], o[20:22], o[22:24]])


def geo(dims, dir, bcs=((0, 0), (0, 0), (0, 0)), L=(2 * np.pi,) * 3, stretching=None, beta=None):
    """geo_t of the host layer along `dir` (0..2) on a single rank (no GPU needed)."""
    cfg = _config(dims, (1, 1, 1), L, bcs, 1600.0, 1e-3, "RK3", "compact6", "compact6", "classic", "compact6", 0, 1, -1, 0,
                  None, stretching, beta)
    nv = dims[dir]
    nc = nv if bcs[dir][0] == BC_PERIODIC else nv - 1
    a = [np.zeros(nv) for _ in range(4)] + [np.zeros(nc) for _ in range(2)]
    _chk(load()[1].x3d2h_geo(C.byref(cfg), dir, *[_p(x) for x in a]))
    return dict(zip(("vert_coords", "vert_ds", "vert_ds2", "vert_d2s", "midp_coords", "midp_ds"), a))


def tdsops_tables(n_tds, delta, operation, scheme, bc_start, bc_end, stretch=None, stretch_correct=None, n_halo=4,
                  from_to=None, sym=False):
    """tdsops_init of the host layer (no GPU needed): the tables handed to x3d2c_tdsops_create."""
    n = n_tds + 1
    info = (C.c_int * 4)()
    sc = np.zeros(5)
    coeffs, cs, ce = np.zeros(9), np.zeros((4, 9)), np.zeros((4, 9))
    fw, bw, sa, scc, af = (np.zeros(n) for _ in range(5))
    st, stc = np.zeros(n_tds), np.zeros(n_tds)
    s_in = _f(stretch) if stretch is not None else None
    sc_in = _f(stretch_correct) if stretch_correct is not None else None
    _chk(load()[1].x3d2h_tdsops_tables(n_tds, float(delta), operation.encode(), scheme.encode(), bc_start, bc_end,
                                       _p(s_in) if s_in is not None else None,
                                       _p(sc_in) if sc_in is not None else None, n_halo,
                                       from_to.encode() if from_to else None, int(sym), info, _p(sc), _p(coeffs), _p(cs),
                                       _p(ce), _p(fw), _p(bw), _p(sa), _p(scc), _p(af), _p(st), _p(stc)))
    n_rhs = info[1]
    return dict(n_tds=info[0], n_rhs=n_rhs, move=info[2], periodic=bool(info[3]), alpha=sc[0], a=sc[1], b=sc[2], c=sc[3],
                d=sc[4], coeffs=coeffs, coeffs_s=cs, coeffs_e=ce, dist_fw=fw[:n_rhs], dist_bw=bw[:n_rhs],
                dist_sa=sa[:n_rhs], dist_sc=scc[:n_rhs], dist_af=af[:n_rhs], stretch=st, stretch_correct=stc)


def waves_000(dims, L=(2 * np.pi,) * 3, interpl="classic", stagder="compact6"):
    cfg = _config(dims, (1, 1, 1), L, ((0, 0),) * 3, 1600.0, 1e-3, "RK3", "compact6", "compact6", interpl, stagder, 0, 1,
                  -1, 0, None)
    nx, ny, nz = dims
    w = np.zeros((nz, ny, nx // 2 + 1, 2))
    _chk(load()[1].x3d2h_waves_000(C.byref(cfg), _p(w)))
    return w[..., 0] + 1j * w[..., 1]


def poisson_tables_010(dims, L, stretching="uniform", beta=1.0, interpl="classic", stagder="compact6"):
    """Host-layer tables of the Poisson solver with walls in y (no GPU needed): waves and the pentadiagonal operators."""
    cfg = _config(dims, (1, 1, 1), L, ((0, 0), (2, 2), (0, 0)), 1600.0, 1e-3, "RK3", "compact6", "compact6", interpl,
                  stagder, 0, 1, -1, 0, None, ("uniform", stretching, "uniform"), (1.0, beta, 1.0))
    nx, ny, nz = dims[0], dims[1] - 1, dims[2]
    nxh = nx // 2 + 1
    w = np.zeros((nz, ny, nxh, 2))
    rows = ny if stretching == "bottom" else ny // 2
    ao, ae = np.zeros((5, nz, rows, nxh)), np.zeros((5, nz, rows, nxh))
    info = (C.c_int * 2)()
    _chk(load()[1].x3d2h_poisson_tables_010(C.byref(cfg), info, _p(w), _p(ao), _p(ae)))
    return dict(stretched=info[0], rows=info[1], waves=w[..., 0] + 1j * w[..., 1], a_odd=ao, a_even=ae)


class Sim:
    """One rank of an x3d2 run on the cuda_c backend (solver_t + case + monitoring of the reference).

    Host arrays are numpy [nz, ny, nx] (x fastest), rank-local and un-padded.
    """

    def __init__(self, dims, nproc_dir=(1, 1, 1), L=(2 * np.pi,) * 3, bcs=((0, 0), (0, 0), (0, 0)), Re=1600.0, dt=1e-3,
                 time_intg="RK3", der1st="compact6", der2nd="compact6", interpl="classic", stagder="compact6", rank=0,
                 nproc=1, device=-1, strict=False, nccl_unique_id=None, stretching=None, beta=None, base_ops=False):
        """base_ops=True: the host layer issues the unchanged reference solver's operator graph through the base_backend_t
        entry points only (X3D2H_FLAG_BASE_OPS, the drop-in path); default: the fused extension entry points."""
        self._c, self._h = load()
        self.dims = tuple(int(d) for d in dims)
        self.periodic = [pair[0] == BC_PERIODIC for pair in bcs]
        self._nccl_id = nccl_unique_id  # keep the bytes alive
        cfg = _config(dims, nproc_dir, L, bcs, Re, dt, time_intg, der1st, der2nd, interpl, stagder, rank, nproc, device,
                      (FLAG_STRICT if strict else 0) | (FLAG_BASE_OPS if base_ops else 0), nccl_unique_id, stretching,
                      beta)
        self.h = C.c_void_p()
        rc = self._h.x3d2h_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RuntimeError("x3d2_b200: " + self._h.x3d2h_last_error().decode())
        self.ctx = self._h.x3d2h_backend(self.h)

    def close(self):
        if getattr(self, "h", None):
            self._h.x3d2h_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ------------------------------------------------------------------ shapes
    def local_dims(self, loc=VERT):
        d = (C.c_int * 3)()
        _chk(self._h.x3d2h_local_dims(self.h, loc, d))
        return tuple(d)

    def shape(self, loc=VERT):
        d = self.local_dims(loc)
        return (d[2], d[1], d[0])

    def _out(self, loc=VERT):
        return np.zeros(self.shape(loc))

    # ------------------------------------------------------------------ solver
    def init_tgv(self):
        _chk(self._h.x3d2h_init_tgv(self.h))

    def set_uvw(self, u, v, w):
        u, v, w = (_f(a, self.shape()) for a in (u, v, w))
        _chk(self._h.x3d2h_set_velocity(self.h, _p(u), _p(v), _p(w)))

    def get_uvw(self, out=None):
        u, v, w = [_out_ok(a, self.shape()) for a in out] if out is not None else (self._out(), self._out(), self._out())
        _chk(self._h.x3d2h_get_velocity(self.h, _p(u), _p(v), _p(w)))
        return u, v, w

    def set_case_channel(self, omega_rot=0.0, n_rotate=0):
        """case/channel.f90 hooks around every sub-stage of the following steps: bulk-velocity correction, rotation forcing,
        wall rows reset (zero wall values)."""
        _chk(self._h.x3d2h_set_case_channel(self.h, omega_rot, n_rotate))

    def step(self, n=1):
        _chk(self._h.x3d2h_step(self.h, n))

    def step_batches(self, n, ins, outs):
        """n independent batches, one step each: upload `ins` (u, v, w) -> step -> download into `outs`; the copies overlap
        the kernels of the neighbouring batches. Arrays: C-contiguous float64 of self.shape(), page-locked for real overlap."""
        ins = [_out_ok(a, self.shape()) for a in ins]
        outs = [_out_ok(a, self.shape()) for a in outs]
        _chk(self._h.x3d2h_step_batches(self.h, n, *[_p(a) for a in ins], *[_p(a) for a in outs]))

    def sync(self):
        _chk(self._h.x3d2h_sync(self.h))

    def monitor(self):
        out = np.zeros(4)
        _chk(self._h.x3d2h_monitor(self.h, _p(out)))
        return dict(enstrophy=out[0], ke=out[1], div_u_max=out[2], div_u_mean=out[3])

    def pressure_correction(self):
        _chk(self._h.x3d2h_pressure_correction(self.h))

    def launch_count(self):
        return int(self._c.x3d2c_launch_count(self.ctx))

    def stream(self):
        return int(self._c.x3d2c_stream(self.ctx) or 0)

    def bench_op(self, op, reps=1):
        _chk(self._h.x3d2h_bench_op(self.h, op.encode(), reps))

    # ------------------------------------------------------------------ operators on host data
    def transeq(self, u, v, w):
        u, v, w = (_f(a, self.shape()) for a in (u, v, w))
        a, b, c = self._out(), self._out(), self._out()
        _chk(self._h.x3d2h_transeq(self.h, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c)))
        return a, b, c

    def transeq_dir(self, dir, u, v, w):
        u, v, w = (_f(a, self.shape()) for a in (u, v, w))
        a, b, c = self._out(), self._out(), self._out()
        _chk(self._h.x3d2h_transeq_dir(self.h, dir, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c)))
        return a, b, c

    def transeq_lowmem(self, u, v, w):
        """solver.f90:391-505: (du, dv, dw, u after its x -> y -> z -> x round trip)."""
        u, v, w = (_f(a, self.shape()) for a in (u, v, w))
        a, b, c, ub = self._out(), self._out(), self._out(), self._out()
        _chk(self._h.x3d2h_transeq_lowmem(self.h, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c), _p(ub)))
        return a, b, c, ub

    def transeq_species(self, u, v, w, spec, nu_s):
        u, v, w, spec = (_f(a, self.shape()) for a in (u, v, w, spec))
        d = self._out()
        _chk(self._h.x3d2h_transeq_species(self.h, _p(u), _p(v), _p(w), _p(spec), float(nu_s), _p(d)))
        return d

    def derived(self, what, grads):
        """compute_vorticity / compute_qcriterion from the nine velocity gradients (dudx, dudy, ..., dwdz)."""
        g = [_f(a, self.shape()) for a in grads]
        arr = (_dp * 9)(*[_p(a) for a in g])
        out = self._out()
        _chk(self._h.x3d2h_derived(self.h, what.encode(), arr, _p(out)))
        return out

    def slice_max_sum(self, dir, x, i_slice, loc=VERT):
        x = _f(x, self.shape(loc))
        mx, sm = C.c_double(0), C.c_double(0)
        _chk(self._h.x3d2h_slice_max_sum(self.h, dir, loc, _p(x), i_slice, C.byref(mx), C.byref(sm)))
        return mx.value, sm.value

    def tds_solve(self, dir, opname, f, in_loc=VERT):
        f = _f(f, self.shape(in_loc))
        move = {"stagder_v2p": 1, "interpl_v2p": 1, "stagder_p2v": -1, "interpl_p2v": -1}.get(opname, 0)
        out = self._out(in_loc + move * 10 ** dir)
        ol = C.c_int(0)
        _chk(self._h.x3d2h_tds_solve(self.h, dir, opname.encode(), in_loc, _p(f), _p(out), C.byref(ol)))
        return out

    def tds_fused(self, mode, dir, op_a, op_b, a_in, b_in=None, a=1.0, in_loc=VERT):
        """x3d2c_tds_solve_sum / _dual / _axpy. sum: A(a_in) + B(b_in); dual: (A(a_in), B(a_in)); axpy: b_in + a A(a_in)."""
        move = {"stagder_v2p": 1, "interpl_v2p": 1, "stagder_p2v": -1, "interpl_p2v": -1}.get(op_a, 0)
        out_loc = in_loc + move * 10 ** dir
        a_in = _f(a_in)
        b_in = _f(b_in) if b_in is not None else a_in
        oa, ob = self._out(out_loc), self._out(out_loc)
        _chk(self._h.x3d2h_tds_fused(self.h, mode.encode(), dir, op_a.encode(), (op_b or op_a).encode(), in_loc, out_loc,
                                     _p(a_in), _p(b_in), a, _p(oa), _p(ob)))
        return (oa, ob) if mode == "dual" else oa

    def tds_fused_r(self, mode, dir, op_a, op_b, a_in, b_in=None, rdr_in=0, rdr_out=0, in_loc=VERT):
        """x3d2c_tds_solve_r / _sum_r / _dual_r: reorder(rdr_in) -> operator(s) along `dir` -> reorder(rdr_out)."""
        move = {"stagder_v2p": 1, "interpl_v2p": 1, "stagder_p2v": -1, "interpl_p2v": -1}.get(op_a, 0)
        out_loc = in_loc + move * 10 ** dir
        a_in = _f(a_in)
        b_in = _f(b_in) if b_in is not None else a_in
        oa, ob = self._out(out_loc), self._out(out_loc)
        _chk(self._h.x3d2h_tds_fused_r(self.h, mode.encode(), dir, op_a.encode(), (op_b or op_a).encode(), in_loc,
                                       out_loc, rdr_in, rdr_out, _p(a_in), _p(b_in), _p(oa), _p(ob)))
        return (oa, ob) if mode == "dual" else oa

    def divergence(self, u, v, w):
        u, v, w = (_f(a, self.shape()) for a in (u, v, w))
        d = self._out(CELL)
        _chk(self._h.x3d2h_divergence(self.h, _p(u), _p(v), _p(w), _p(d)))
        return d

    def interpl_c2v(self, p):
        """vector_calculus_t%interpl_c2v (cell centres -> vertices, interpl_p2v operators)."""
        p = _f(p, self.shape(CELL))
        a = self._out()
        _chk(self._h.x3d2h_interpl_c2v(self.h, _p(p), _p(a)))
        return a

    def laplacian(self, u):
        """vector_calculus_t%laplacian (der2nd operators) of a vertex field."""
        u = _f(u, self.shape())
        a = self._out()
        _chk(self._h.x3d2h_laplacian(self.h, _p(u), _p(a)))
        return a

    def gradient(self, p):
        p = _f(p, self.shape(CELL))
        a, b, c = self._out(), self._out(), self._out()
        _chk(self._h.x3d2h_gradient(self.h, _p(p), _p(a), _p(b), _p(c)))
        return a, b, c

    def curl(self, u, v, w):
        u, v, w = (_f(a, self.shape()) for a in (u, v, w))
        a, b, c = self._out(), self._out(), self._out()
        _chk(self._h.x3d2h_curl(self.h, _p(u), _p(v), _p(w), _p(a), _p(b), _p(c)))
        return a, b, c

    def poisson(self, f):
        f = _f(f, self.shape(CELL))
        p = self._out(CELL)
        _chk(self._h.x3d2h_poisson(self.h, _p(f), _p(p)))
        return p

    def fft_roundtrip(self, f, want_spec=False):
        f = _f(f)
        out = self._out(CELL)
        spec = None
        if want_spec:
            nz, ny, nx = self.shape(CELL)
            spec = np.zeros((nz, ny, nx // 2 + 1, 2))
        _chk(self._h.x3d2h_fft_roundtrip(self.h, _p(f), _p(out), _p(spec) if want_spec else None))
        if want_spec:
            return out, spec[..., 0] + 1j * spec[..., 1]
        return out

    def reorder_chain(self, f, names):
        f = _f(f, self.shape())
        out = self._out()
        r = (C.c_int * len(names))(*[RDR[n] for n in names])
        _chk(self._h.x3d2h_reorder_chain(self.h, _p(f), r, len(names), _p(out)))
        return out

    def sum_intox(self, dir_from, a, b):
        a, b = _f(a, self.shape()), _f(b, self.shape())
        out = self._out()
        _chk(self._h.x3d2h_sum_intox(self.h, dir_from, _p(a), _p(b), _p(out)))
        return out

    def vecadd(self, dir, a, x, b, y):
        x, y = _f(x, self.shape()), _f(y, self.shape())
        out = self._out()
        _chk(self._h.x3d2h_vecadd(self.h, dir, a, _p(x), b, _p(y), _p(out)))
        return out

    def scalar_product(self, dir, x, y, loc=VERT):
        x, y = _f(x, self.shape(loc)), _f(y, self.shape(loc))
        s = C.c_double(0)
        _chk(self._h.x3d2h_scalar_product(self.h, dir, loc, _p(x), _p(y), C.byref(s)))
        return s.value

    def field_max_mean(self, dir, x, loc=VERT):
        x = _f(x, self.shape(loc))
        mx, mean = C.c_double(0), C.c_double(0)
        _chk(self._h.x3d2h_field_max_mean(self.h, dir, loc, _p(x), C.byref(mx), C.byref(mean)))
        return mx.value, mean.value

    def fieldop(self, op, dir, x, y=None, a=0.0, loc=VERT, extra=None):
        """field_scale / field_shift / vecmult / veccopy / fill / volume_integral / set_face / set_face_from_field."""
        x = _f(x, self.shape(loc))
        y = _f(y, self.shape(loc)) if y is not None else None
        out = self._out(loc)
        s = (C.c_double * 2)(*(extra or (0.0, 0.0)))  # set_face*: (c_end | flow_rate_diff, face)
        _chk(self._h.x3d2h_fieldop(self.h, op.encode(), dir, loc, float(a), _p(x), _p(y) if y is not None else None,
                                   _p(out), s))
        return s[0] if op == "volume_integral" else out
