"""The Fortran side of the boundary (fortran/cuda_c_backend.f90) cannot be compiled in this image (no Fortran
compiler). This test parses it instead:
  * every `bind(c, name='...')` interface must name a function declared in include/x3d2c.h with the same number of
    arguments, and each argument must be passed the way the C prototype expects (scalars by value, pointers either
    as `type(c_ptr), value` or as a by-reference dummy);
  * every deferred procedure of the reference's base_backend_t and poisson_fft_t must be overridden;
  * every C function an override needs must be bound.
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F90 = os.path.join(ROOT, "fortran", "cuda_c_backend.f90")
HDR = os.path.join(ROOT, "include", "x3d2c.h")

# src/backend/backend.f90:34-58 and src/poisson_fft.f90:45-62 of the reference (the deferred type-bound procedures)
BACKEND_DEFERRED = ["transeq_x", "transeq_y", "transeq_z", "transeq_species", "tds_solve", "reorder", "sum_yintox",
                    "sum_zintox", "veccopy", "vecadd", "vecmult", "scalar_product", "field_max_mean", "slice_max_sum",
                    "field_scale", "field_shift", "field_volume_integral", "field_set_face",
                    "field_set_face_from_field", "compute_vorticity", "compute_qcriterion", "copy_data_to_f",
                    "copy_f_to_data", "alloc_tdsops", "init_poisson_fft"]
POISSON_DEFERRED = ["fft_forward_010", "fft_forward_100", "fft_forward_110", "fft_forward", "fft_backward_010",
                    "fft_backward_100", "fft_backward_110", "fft_backward", "fft_postprocess_000", "fft_postprocess_010",
                    "fft_postprocess_100", "fft_postprocess_110", "enforce_periodicity_x", "undo_periodicity_x",
                    "enforce_periodicity_y", "undo_periodicity_y", "enforce_periodicity_xy", "undo_periodicity_xy"]


def c_prototypes():
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|long long|void\s*\*|const char\s*\*)\s*(x3d2c_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        name, args = m.group(1), m.group(2).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                params.append("ptr" if ("*" in a or "[" in a) else ("double" if a.startswith("double") else "int"))
        protos[name] = params
    return protos


def fortran_interfaces():
    src = open(F90).read()
    src = re.sub(r"&\s*\n\s*&?", " ", src)          # join continuation lines
    src = re.sub(r"!.*", "", src)                    # strip comments
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(c,\s*name='(\w+)'\)\s*result\(\w+\)(.*?)end function", src, flags=re.S):
        fname, args, cname, body = m.group(1), m.group(2), m.group(3), m.group(4)
        assert fname == cname, (fname, cname)
        names = [a.strip() for a in args.split(",") if a.strip()]
        kinds = {}
        for line in body.splitlines():
            if "::" not in line:
                continue
            decl, vars_ = line.split("::", 1)
            decl = decl.strip().lower()
            for v in re.split(r",(?![^()]*\))", vars_):
                v = re.sub(r"\(.*\)", "", v).strip()
                if v in names:
                    by_value = "value" in decl
                    if decl.startswith("type(c_ptr)"):
                        kinds[v] = "ptr" if by_value else "ptr"      # c_ptr by value = pointer; by reference = T**
                    elif decl.startswith("real(c_double)"):
                        kinds[v] = "double" if by_value else "ptr"
                    elif decl.startswith("integer(c_int)") or decl.startswith("integer(c_long_long)"):
                        kinds[v] = "int" if by_value else "ptr"
                    elif decl.startswith("character") or decl.startswith("type(x3d2c_config)"):
                        kinds[v] = "ptr"
        out[cname] = [kinds.get(n, "?") for n in names]
    return out


def test_every_binding_matches_the_header():
    protos, binds = c_prototypes(), fortran_interfaces()
    assert len(binds) >= 50, len(binds)
    # what the shim does NOT bind: the fused extension entry points (a Fortran solver that does not know them keeps
    # calling the base operators) and the debugging / bench helpers
    extensions = {"x3d2c_transeq_r", "x3d2c_transeq_r_fused", "x3d2c_tds_solve_sum", "x3d2c_tds_solve_dual", "x3d2c_tds_solve_axpy", "x3d2c_tds_solve_r",
                  "x3d2c_tds_solve_sum_r", "x3d2c_tds_solve_dual_r", "x3d2c_tds_solve_axpy_r", "x3d2c_reorder_x2yz", "x3d2c_sum_yzintox", "x3d2c_sum_yzintox_lincomb",
                  "x3d2c_veclincomb", "x3d2c_copy_data_to_f_async", "x3d2c_copy_f_to_data_async", "x3d2c_lane_record",
                      "x3d2c_lane_wait", "x3d2c_lane_sync", "x3d2c_launch_count", "x3d2c_stream", "x3d2c_version", "x3d2c_poisson_get_spectrum"}
    assert set(protos) - set(binds) == extensions
    for name, kinds in binds.items():
        assert name in protos, f"{name} is bound in the shim but not declared in x3d2c.h"
        assert "?" not in kinds, (name, kinds)
        assert kinds == protos[name], f"{name}: Fortran passes {kinds}, C expects {protos[name]}"


def test_every_deferred_procedure_is_overridden():
    src = open(F90).read()
    for p in BACKEND_DEFERRED:
        assert re.search(rf"procedure\s*::\s*{p}\s*=>\s*\w+", src), f"base_backend_t%{p} is not overridden"
    for p in POISSON_DEFERRED:
        assert re.search(rf"procedure\s*::\s*{p}\s*=>\s*\w+", src), f"poisson_fft_t%{p} is not overridden"
    # the overrides name subroutines / functions that exist in the file
    for m in re.finditer(r"procedure\s*::\s*\w+\s*=>\s*(\w+)", src):
        assert re.search(rf"(subroutine|function)\s+{m.group(1)}\b", src), m.group(1)


def test_deferred_lists_match_the_reference_when_it_is_present():
    """In this container the reference tree is available: the lists above are the reference's own deferred bindings."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        return
    def deferred(path):
        return re.findall(r"procedure\(\w+\),\s*deferred\s*::\s*(\w+)", open(path).read())
    assert sorted(deferred(os.path.join(ref, "backend", "backend.f90"))) == sorted(BACKEND_DEFERRED)
    assert sorted(deferred(os.path.join(ref, "poisson_fft.f90"))) == sorted(POISSON_DEFERRED)


def test_operators_used_by_the_shim_are_all_exported(x3d2):
    c, _ = x3d2.load()
    for name in fortran_interfaces():
        assert hasattr(c, name), name
