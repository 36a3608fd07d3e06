"""CPU tests of the product's host layer (libx3d2h.so) and of the C ABI surface (libx3d2c.so) — no GPU compute."""
import ctypes
import os

import numpy as np
import pytest

CASES = [("first-deriv", "compact6", None), ("second-deriv", "compact6", None), ("interpolate", "classic", "v2p"),
         ("interpolate", "classic", "p2v"), ("interpolate", "optimised", "v2p"), ("interpolate", "aggressive", "p2v"),
         ("stag-deriv", "compact6", "v2p"), ("stag-deriv", "compact6", "p2v")]
BCS = [(0, 0), (-1, -1), (1, 1), (1, -1), (-1, 1), (2, 2), (2, 1), (2, -1)]


def test_abi_symbols_exported(x3d2):
    """libx3d2c.so exports every function include/x3d2c.h declares."""
    from x3d2_b200 import lib
    c, h = x3d2.load()
    names = lib.abi_symbols()
    assert len(names) >= 36
    for n in names:
        assert hasattr(c, n), n
    assert c.x3d2c_version() == 100


def test_no_cpu_fallback(x3d2):
    """Without a CUDA device the product refuses to run (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        x3d2.Sim((32, 32, 32))


def test_product_does_not_link_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for so in ("libx3d2c.so", "libx3d2h.so"):
        blob = open(os.path.join(root, "x3d2_b200", so), "rb").read()
        assert b"libx3d2_oracle" not in blob and b"orc_world" not in blob
    for dirpath, _, files in os.walk(os.path.join(root, "x3d2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "_oracle" not in txt and "oracle/" not in txt, f


@pytest.mark.parametrize("op,scheme,from_to", CASES)
@pytest.mark.parametrize("bc", BCS)
@pytest.mark.parametrize("sym", [False, True])
def test_host_tdsops_matches_oracle(x3d2, oracle, op, scheme, from_to, bc, sym):
    """The host layer's tdsops_init and the oracle's restatement of src/tdsops.f90 agree bit for bit."""
    if from_to and 2 in bc:
        with pytest.raises(RuntimeError):
            x3d2.tdsops_tables(40, 0.1, op, scheme, bc[0], bc[1], from_to=from_to, sym=sym)
        return
    rng = np.random.default_rng(3)
    st, stc = 1 + 0.1 * rng.random(40), 0.1 * rng.random(40)
    t = x3d2.tdsops_tables(40, 0.037, op, scheme, bc[0], bc[1], st, stc, from_to=from_to, sym=sym)
    o = oracle.Tdsops(40, 0.037, op, scheme, bc[0], bc[1], st, stc, from_to=from_to, sym=sym)
    assert (t["n_tds"], t["n_rhs"], t["move"], t["periodic"]) == (o.n_tds, o.n_rhs, o.move, o.periodic)
    for k in ("coeffs", "coeffs_s", "coeffs_e", "dist_fw", "dist_bw", "dist_sa", "dist_sc", "dist_af", "stretch",
              "stretch_correct"):
        assert np.array_equal(t[k], getattr(o, k)), k
    assert (t["alpha"], t["a"], t["b"], t["c"], t["d"]) == (o.alpha, o.a, o.b, o.c, o.d)


def test_distd2_factorisation_solves_the_system(x3d2):
    """preprocess_dist (src/tdsops.f90:874-931): the periodic single-rank DistD2 solve equals an exact cyclic solve
    (SURVEY.md F3). Evaluated with numpy from the host layer's tables."""
    n = 64
    t = x3d2.tdsops_tables(n, 2 * np.pi / n, "first-deriv", "compact6", 0, 0)
    rhs = np.random.default_rng(5).standard_normal(n)
    fw, bw, sa, sc, af = (t[k] for k in ("dist_fw", "dist_bw", "dist_sa", "dist_sc", "dist_af"))
    d = np.zeros(n)
    d[0], d[1] = rhs[0] * af[0], rhs[1] * af[1]
    for j in range(2, n):
        d[j] = fw[j] * (rhs[j] - af[j] * d[j - 1])
    zn = d[n - 1]
    for j in range(n - 3, 0, -1):
        d[j] -= bw[j] * d[j + 1]
    d[0] = fw[0] * (d[0] - bw[0] * d[1])
    s = (d[0] - sa[0] * zn) / (1 - sa[0] ** 2)
    e = (d[n - 1] - sc[n - 1] * d[0]) / (1 - sc[n - 1] ** 2)
    x = d - sa * s - sc * e
    x[0], x[n - 1] = s, e
    A = np.eye(n) + t["alpha"] * (np.roll(np.eye(n), 1, 1) + np.roll(np.eye(n), -1, 1))
    assert np.abs(A @ x - rhs).max() < 1e-13


@pytest.mark.parametrize("nproc_dir", [(1, 1, 1), (1, 1, 2), (1, 2, 2), (1, 1, 8)])
def test_decompose_matches_oracle(x3d2, oracle, nproc_dir):
    dims, bcs = (64, 64, 128), ((0, 0), (2, 2), (0, 0))
    W = oracle.World(dims, nproc_dir=nproc_dir, bcs=bcs)
    P = int(np.prod(nproc_dir))
    for r in range(P):
        d, m = x3d2.decompose(dims, nproc_dir, r, bcs), W.mesh_info(r)
        assert d["vert_dims"] == m["vert_dims"] and d["cell_dims"] == m["cell_dims"] and d["BCs"] == m["BCs"]
        z = r // (nproc_dir[0] * nproc_dir[1])
        assert d["nrank_dir"][2] == z and d["n_offset"][2] == z * dims[2] // nproc_dir[2]
    d = x3d2.decompose(dims, (1, 1, 4), 0)
    assert d["pprev"][2] == 3 and d["pnext"][2] == 1 and d["pprev"][0] == 0  # cyclic neighbours (mesh_content.f90:87-100)


def test_waves_match_oracle(x3d2, oracle):
    dims = (32, 16, 24)
    w = x3d2.waves_000(dims, L=(1.0, 2.0, 3.0))
    W = oracle.World(dims, L=(1.0, 2.0, 3.0))
    assert np.array_equal(w, W.waves())
    assert w[0, 0, 0] == 0 and np.all(w.real == w.imag)


@pytest.mark.parametrize("stretching,beta,bc", [("uniform", 1.0, (0, 0)), ("uniform", 1.0, (2, 2)),
                                                ("centred", 0.8, (1, 1)), ("top-bottom", 0.259065151, (2, 2)),
                                                ("bottom", 1.3, (2, 1)), ("centred", 2.0, (0, 0))])
def test_geo_matches_oracle(x3d2, oracle, stretching, beta, bc):
    """Stretched-mesh coordinates and stretching factors of the host mesh (mesh_content.f90:159-263) are the oracle's, bitwise."""
    n = 64 if bc[0] == 0 else 65
    for d in range(3):
        dims, bcs = [32, 32, 32], [(0, 0)] * 3
        st, be = ["uniform"] * 3, [1.0] * 3
        dims[d], bcs[d], st[d], be[d] = n, bc, stretching, beta
        g = x3d2.geo(tuple(dims), d, tuple(bcs), (2.0, 1.0, 3.0), st, be)
        e = oracle.World(tuple(dims), L=(2.0, 1.0, 3.0), bcs=tuple(bcs), stretching=st, beta=be).geo(d)
        for k in e:
            assert np.array_equal(g[k], e[k]), (d, k)
        if stretching != "uniform":
            assert np.ptp(g["vert_ds"]) > 0
            assert np.all(np.diff(g["vert_coords"]) > 0)


def test_bench_grid_and_reference_arm():
    import subprocess, sys, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    assert bench.grid_for(1, 512) == [512, 512, 512] and bench.grid_for(2, 512) == [512, 512, 1024]
    assert bench.grid_for(4, 512) == [512, 1024, 1024] and bench.grid_for(8, 512) == [1024, 1024, 1024]
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3", "--cpu-size", "32"], capture_output=True, text=True, check=True).stdout
    j = json.loads(out.strip().splitlines()[-1])
    assert j["impl"] == "reference" and j["cpu_baseline"]["kind"] == "port" and j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.parametrize("stretching,beta", [("uniform", 1.0), ("top-bottom", 0.259065151), ("centred", 0.8), ("bottom", 1.3)])
def test_poisson_010_tables_match_oracle(x3d2, oracle, stretching, beta):
    """Host layer's base_init for walls in y: the wave-number table and the pentadiagonal spectral operators of a
    stretched mesh (src/poisson_fft.f90:275-652) against the oracle's transcription, bit for bit."""
    dims, L = (32, 33, 16), (1.0, 2.0, 3.0)
    t = x3d2.poisson_tables_010(dims, L, stretching, beta)
    W = oracle.World(dims, L=L, bcs=((0, 0), (2, 2), (0, 0)), stretching=("uniform", stretching, "uniform"),
                     beta=(1.0, beta, 1.0))
    assert np.array_equal(t["waves"], W.waves())
    m = W.stretching_matrix()
    assert t["stretched"] == m["stretched"]
    if stretching != "uniform":
        assert t["rows"] == m["rows"]
        assert np.array_equal(t["a_odd"], m["a_odd"]) and np.abs(m["a_odd"]).max() > 0
        assert np.array_equal(t["a_even"], m["a_even"])
