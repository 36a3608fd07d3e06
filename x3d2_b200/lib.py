"""Loading / building of the in-tree shared libraries libx3d2c.so (CUDA backend) and libx3d2h.so (host layer)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_BACKEND = os.path.join(_HERE, "libx3d2c.so")
SO_HOST = os.path.join(_HERE, "libx3d2h.so")

DIR_X, DIR_Y, DIR_Z, DIR_C = 1, 2, 3, 4
VERT, CELL = 0, 1110
BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_HALO = 0, 1, 2, -1
FLAG_STRICT = 1
FLAG_BASE_OPS = 0x100  # include/x3d2h.h: X3D2H_FLAG_BASE_OPS
RDR = dict(X2Y=12, X2Z=13, Y2X=21, Y2Z=23, Z2X=31, Z2Y=32, C2X=41, C2Y=42, C2Z=43, X2C=14, Y2C=24, Z2C=34)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class X3D2HConfig(C.Structure):  # include/x3d2h.h: x3d2h_config
    _fields_ = [("dims_global", C.c_int * 3), ("nproc_dir", C.c_int * 3), ("L_global", C.c_double * 3),
                ("bc", C.c_int * 6), ("Re", C.c_double), ("dt", C.c_double), ("time_intg", C.c_char_p),
                ("der1st_scheme", C.c_char_p), ("der2nd_scheme", C.c_char_p), ("interpl_scheme", C.c_char_p),
                ("stagder_scheme", C.c_char_p), ("rank", C.c_int), ("nproc", C.c_int), ("device", C.c_int),
                ("flags", C.c_int), ("nccl_unique_id", C.c_void_p), ("stretching", C.c_char_p * 3), ("beta", C.c_double * 3)]


def build(verbose=False):
    """Compile both libraries for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-j", str(min(16, os.cpu_count() or 4)), "-C", os.path.join(_HERE, "csrc"), "all"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building x3d2_b200 failed:\n" + r.stdout)
    if verbose:
        print(r.stdout)


_libs = None


def load():
    """Returns (backend CDLL, host CDLL). Raises if the extension is missing — there is no fallback."""
    global _libs
    if _libs is not None:
        return _libs
    for so in (SO_BACKEND, SO_HOST):
        if not os.path.exists(so):
            raise RuntimeError(f"{so} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(x3d2_b200 has no CPU fallback)")
    c = C.CDLL(SO_BACKEND, mode=C.RTLD_GLOBAL)
    h = C.CDLL(SO_HOST)
    c.x3d2c_last_error.restype = C.c_char_p
    c.x3d2c_launch_count.restype = C.c_longlong
    c.x3d2c_launch_count.argtypes = [C.c_void_p]
    c.x3d2c_stream.restype = C.c_void_p
    c.x3d2c_stream.argtypes = [C.c_void_p]
    h.x3d2h_last_error.restype = C.c_char_p
    h.x3d2h_backend.restype = C.c_void_p
    h.x3d2h_backend.argtypes = [C.c_void_p]
    cfgp = C.POINTER(X3D2HConfig)
    sig = dict(
        x3d2h_decompose=[cfgp, _ip],
        x3d2h_geo=[cfgp, C.c_int] + [_dp] * 6,
        x3d2h_tdsops_tables=[C.c_int, C.c_double, C.c_char_p, C.c_char_p, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_char_p,
                             C.c_int, _ip] + [_dp] * 11,
        x3d2h_waves_000=[cfgp, _dp],
        x3d2h_poisson_tables_010=[cfgp, _ip, _dp, _dp, _dp],
        x3d2h_create=[cfgp, C.POINTER(C.c_void_p)],
        x3d2h_destroy=[C.c_void_p],
        x3d2h_local_dims=[C.c_void_p, C.c_int, _ip],
        x3d2h_init_tgv=[C.c_void_p],
        x3d2h_set_velocity=[C.c_void_p, _dp, _dp, _dp],
        x3d2h_get_velocity=[C.c_void_p, _dp, _dp, _dp],
        x3d2h_set_case_channel=[C.c_void_p, C.c_double, C.c_int],
        x3d2h_step=[C.c_void_p, C.c_int],
        x3d2h_step_batches=[C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp],
        x3d2h_sync=[C.c_void_p],
        x3d2h_monitor=[C.c_void_p, _dp],
        x3d2h_transeq=[C.c_void_p] + [_dp] * 6,
        x3d2h_transeq_dir=[C.c_void_p, C.c_int] + [_dp] * 6,
        x3d2h_transeq_lowmem=[C.c_void_p] + [_dp] * 7,
        x3d2h_transeq_species=[C.c_void_p] + [_dp] * 4 + [C.c_double, _dp],
        x3d2h_derived=[C.c_void_p, C.c_char_p, C.POINTER(_dp), _dp],
        x3d2h_slice_max_sum=[C.c_void_p, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp],
        x3d2h_tds_solve=[C.c_void_p, C.c_int, C.c_char_p, C.c_int, _dp, _dp, _ip],
        x3d2h_tds_fused=[C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, _dp, _dp, C.c_double,
                         _dp, _dp],
        x3d2h_tds_fused_r=[C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                           _dp, _dp, _dp, _dp],
        x3d2h_divergence=[C.c_void_p] + [_dp] * 4,
        x3d2h_interpl_c2v=[C.c_void_p, _dp, _dp],
        x3d2h_laplacian=[C.c_void_p, _dp, _dp],
        x3d2h_gradient=[C.c_void_p] + [_dp] * 4,
        x3d2h_curl=[C.c_void_p] + [_dp] * 6,
        x3d2h_poisson=[C.c_void_p, _dp, _dp],
        x3d2h_pressure_correction=[C.c_void_p],
        x3d2h_fft_roundtrip=[C.c_void_p, _dp, _dp, _dp],
        x3d2h_reorder_chain=[C.c_void_p, _dp, _ip, C.c_int, _dp],
        x3d2h_sum_intox=[C.c_void_p, C.c_int, _dp, _dp, _dp],
        x3d2h_vecadd=[C.c_void_p, C.c_int, C.c_double, _dp, C.c_double, _dp, _dp],
        x3d2h_scalar_product=[C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp],
        x3d2h_field_max_mean=[C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp],
        x3d2h_bench_op=[C.c_void_p, C.c_char_p, C.c_int],
        x3d2h_fieldop=[C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _dp],
    )
    for name, args in sig.items():
        fn = getattr(h, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _libs = (c, h)
    return _libs


def abi_symbols():
    """Every function declared in include/x3d2c.h (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "x3d2c.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(x3d2c_[a-z0-9_]+)\s*\(", txt)))
