"""Non-periodic boundaries, padded blocks and stretched meshes through the C ABI.

These exercise the reference-order kernels (tds_m1.cu) with Dirichlet / Neumann coefficient rows, n_rhs = n_tds + 1
operators, stretch / stretch_correct factors and allocator padding (257 -> 288 style), against the oracle.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    "channel_y_dirichlet_stretched": dict(dims=(64, 65, 48), bcs=((0, 0), (2, 2), (0, 0)), L=(4.0, 2.0, 2.0),
                                          stretching=("uniform", "top-bottom", "uniform"), beta=(1.0, 0.259065151, 1.0)),
    "x_dirichlet_neumann_y_neumann": dict(dims=(49, 40, 32), bcs=((2, 1), (1, 1), (0, 0)), L=(1.0, 1.0, 1.0)),
    "z_neumann_centred": dict(dims=(32, 32, 41), bcs=((0, 0), (0, 0), (1, 1)), L=(2.0, 2.0, 2.0),
                              stretching=("uniform", "uniform", "centred"), beta=(1.0, 1.0, 0.8)),
}


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_operators_nonperiodic(oracle, x3d2, name, strict):
    kw = CASES[name]
    sim = x3d2.Sim(strict=strict, Re=50.0, **kw)
    ref = oracle.World(Re=50.0, **kw)
    rng = np.random.default_rng(11)
    u, v, w = (rng.standard_normal(sim.shape()) for _ in range(3))
    tol = 0 if strict else 1e-12
    for d in (1, 2, 3):
        for op in ("der1st", "der1st_sym", "der2nd", "der2nd_sym", "stagder_v2p", "interpl_v2p"):
            g, e = sim.tds_solve(d, op, u), ref.tds_solve(d, op, u)
            assert g.shape == e.shape and rel(g, e) <= tol, (name, d, op)
        c = rng.standard_normal(sim.shape(1110))
        for op in ("stagder_p2v", "interpl_p2v"):
            g, e = sim.tds_solve(d, op, c, 1110), ref.tds_solve(d, op, c, 1110)
            assert g.shape == e.shape and rel(g, e) <= tol, (name, d, op)
        for g, e in zip(sim.transeq_dir(d, u, v, w), ref.transeq_dir(d, u, v, w)):
            assert rel(g, e) <= tol, (name, d)
    for g, e in zip(sim.transeq(u, v, w), ref.transeq(u, v, w)):
        assert rel(g, e) <= tol
    assert rel(sim.divergence(u, v, w), ref.divergence(u, v, w)) <= tol
    p = rng.standard_normal(sim.shape(1110))
    for g, e in zip(sim.gradient(p), ref.gradient(p)):
        assert rel(g, e) <= tol
    for g, e in zip(sim.curl(u, v, w), ref.curl(u, v, w)):
        assert rel(g, e) <= tol
    # reductions honour the un-padded extents of the data location
    for d in (1, 2, 3):
        assert abs(sim.scalar_product(d, u, v) - ref.scalar_product(d, u, v)) < 1e-10 * np.abs(u * v).sum()
        (mx, mean), (emx, emean) = sim.field_max_mean(d, u), ref.field_max_mean(d, u)
        assert mx == emx and abs(mean - emean) < 1e-13
    sim.close()


def test_elementwise_ops(x3d2):
    sim = x3d2.Sim((33, 20, 24), bcs=((2, 2), (1, 1), (0, 0)))
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(sim.shape()), rng.standard_normal(sim.shape())
    for d in (1, 2, 3):
        assert np.array_equal(sim.fieldop("scale", d, x, a=1.7), 1.7 * x)      # field_scale (omp/backend.f90:883-891)
        assert np.array_equal(sim.fieldop("shift", d, x, a=-0.3), x + -0.3)    # field_shift (:893-901)
        assert np.array_equal(sim.fieldop("vecmult", d, x, y), y * x)          # vecmult (:587-614)
        assert np.array_equal(sim.fieldop("veccopy", d, x, y), x)
        assert np.all(sim.fieldop("fill", d, x, a=2.5) == 2.5)
    s = sim.fieldop("volume_integral", 1, x)                                    # field_volume_integral (:1023-1066)
    assert abs(s - x.sum()) < 1e-10 * np.abs(x).sum()
    c = rng.standard_normal(sim.shape(1110))
    assert abs(sim.fieldop("volume_integral", 1, c, loc=1110) - c.sum()) < 1e-10 * np.abs(c).sum()
    sim.close()


def test_veclincomb_matches_vecadd_chain(x3d2):
    """x3d2c_veclincomb == the vecadd(c_k, x_k, 1.0, out) chain of the time integrators: exact in strict mode."""
    sim = x3d2.Sim((33, 20, 24), bcs=((2, 2), (1, 1), (0, 0)), strict=True)
    rng = np.random.default_rng(6)
    x, y = rng.standard_normal(sim.shape()), rng.standard_normal(sim.shape())
    a = 0.37
    exp = (-a / 2) * y + ((a * y) + x)
    for d in (1, 2, 3):
        assert np.array_equal(sim.fieldop("lincomb", d, x, y, a=a), exp)
    sim.close()
    fast = x3d2.Sim((33, 20, 24), bcs=((2, 2), (1, 1), (0, 0)))
    assert np.abs(fast.fieldop("lincomb", 1, x, y, a=a) - exp).max() < 1e-15
    fast.close()


@pytest.mark.parametrize("dims,bcs", [((33, 41, 24), ((2, 2), (2, 2), (0, 0))), ((64, 64, 32), ((0, 0), (0, 0), (0, 0)))])
def test_field_set_face(oracle, x3d2, dims, bcs):
    """field_set_face / field_set_face_from_field (channel and inflow cases) against the Cartesian restatement."""
    X_FACE, Y_FACE = 1100, 1010
    sim = x3d2.Sim(dims, bcs=bcs)
    rng = np.random.default_rng(8)
    for loc in (0, 1110):
        f, fs = rng.standard_normal(sim.shape(loc)), rng.standard_normal(sim.shape(loc))
        got = sim.fieldop("set_face", 1, f, a=1.5, loc=loc, extra=(-2.5, Y_FACE))
        assert np.array_equal(got, oracle.field_set_face(f, 1.5, -2.5))
        got = sim.fieldop("set_face_from_field", 1, f, fs, a=0.0, loc=loc, extra=(0.0, Y_FACE))
        assert np.array_equal(got, oracle.field_set_face_from_field(f, fs, 0.0, "y"))
        got = sim.fieldop("set_face_from_field", 1, f, fs, a=0.37, loc=loc, extra=(0.01, X_FACE))
        assert np.array_equal(got, oracle.field_set_face_from_field(f, fs, 0.37, "x", 0.01))
    with pytest.raises(RuntimeError, match="not yet supported"):
        sim.fieldop("set_face", 1, f, a=1.0, loc=loc, extra=(1.0, X_FACE))
    sim.close()


# ---------------------------------------------------------------------------------------------- Poisson 010 (walls in y)
STRETCH = [("uniform", 1.0), ("top-bottom", 0.259065151), ("centred", 0.8), ("bottom", 1.3)]


def _wall_y(dims, L, stretching, beta, **kw):
    return dict(L=L, bcs=((0, 0), (2, 2), (0, 0)), stretching=("uniform", stretching, "uniform"), beta=(1.0, beta, 1.0), **kw)


@pytest.mark.parametrize("stretching,beta", STRETCH)
@pytest.mark.parametrize("dims", [(64, 65, 32), (128, 129, 64), (96, 65, 48)])
def test_poisson_010_vs_oracle(oracle, x3d2, dims, stretching, beta):
    """poisson_010 (src/poisson_fft.f90:228-242) through the C ABI: enforce_periodicity_y, fft_forward,
    fft_postprocess_010 (uniform: omp/kernels/spectral_processing.f90:108-283; stretched: the pentadiagonal solve of
    cuda/kernels/spectral_processing.f90:385-702 with coefficients factorised once), fft_backward, undo_periodicity_y."""
    L = (1.0, 2.0, 1.5)
    kw = _wall_y(dims, L, stretching, beta)
    sim, ref = x3d2.Sim(dims, **kw), oracle.World(dims, **kw)
    rng = np.random.default_rng(5)
    f = rng.standard_normal(sim.shape(1110))
    f -= f.mean()
    got, exp = sim.poisson(f), ref.poisson(f)
    err = np.abs(got - exp).max() / np.abs(exp).max()
    print(dims, stretching, "poisson 010 rel err vs oracle %.2e" % err)
    assert err < 1e-12
    if stretching != "bottom":  # the property the reference's check 2 relies on (tests/verification/test_poisson_bc.f90:554-618)
        p0 = ref.poisson(f)
        rhs = sim.divergence(*sim.gradient(p0))
        p = sim.poisson(rhs)
        back = sim.divergence(*sim.gradient(p))
        assert np.abs(back - rhs).max() <= 1e-11 * np.abs(rhs).max()
    sim.close()


@pytest.mark.parametrize("n_wave,name", [(2, "COS_Y"), (3, "COS_Y"), (2, "COS_XY"), (2, "COS_XYZ")])
def test_poisson_010_known_answers_gpu(x3d2, n_wave, name):
    """The 010 row of tests/verification/test_poisson_bc.f90 on the GPU path: grid 128 x 65 x 32, L = 1, tolerance 1e-11."""
    dims = (128, 65, 32)
    sim = x3d2.Sim(dims, L=(1.0, 1.0, 1.0), bcs=((0, 0), (2, 2), (0, 0)))
    nz, ny, nx = sim.shape(1110)
    mask = {"COS_Y": (0, 1, 0), "COS_XY": (1, 1, 0), "COS_XYZ": (1, 1, 1)}[name]
    c = [(np.arange(n) + 0.5) / n for n in (nx, ny, nz)]
    f = np.ones((nz, ny, nx))
    if mask[0]:
        f = f * np.cos(n_wave * np.pi * c[0])[None, None, :]
    if mask[1]:
        f = f * np.cos(n_wave * np.pi * c[1])[None, :, None]
    if mask[2]:
        f = f * np.cos(n_wave * np.pi * c[2])[:, None, None]
    pa = -f / (sum(mask) * (n_wave * np.pi) ** 2)
    p = sim.poisson(f)
    assert np.linalg.norm((p - p[0, 0, 0]) - (pa - pa[0, 0, 0])) / p.size <= 1e-11
    assert np.linalg.norm(sim.divergence(*sim.gradient(p)) - f) / f.size <= 1e-11
    sim.close()


@pytest.mark.parametrize("stretching,beta,strict", [("uniform", 1.0, False), ("top-bottom", 0.259065151, False),
                                                    ("top-bottom", 0.259065151, True)])
def test_channel_like_steps_vs_oracle(oracle, x3d2, stretching, beta, strict):
    """Time steps of a wall-bounded flow (walls in y, Poisson 010, stretched mesh as examples/channel/input.x3d with
    the noise set to zero): velocity after two RK3 steps and the divergence of the corrected field."""
    dims, L = (64, 65, 32), (4.0, 2.0, 2.0)
    kw = _wall_y(dims, L, stretching, beta, Re=4200.0, dt=0.005)
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    nz, ny, nx = sim.shape()
    yv = ref.geo(1)["vert_coords"]
    x = (np.arange(nx) * (L[0] / nx))[None, None, :]
    z = (np.arange(nz) * (L[2] / nz))[:, None, None]
    y = yv[None, :, None]
    wall = (1 - (y - 1.0) ** 2)  # parabolic profile, zero on both walls (channel.f90:101-106 without the noise)
    u = wall * (1 + 0.1 * np.sin(2 * np.pi * x / L[0]) * np.cos(2 * np.pi * z / L[2]))
    v = 0.05 * wall ** 2 * np.cos(2 * np.pi * x / L[0]) * np.sin(2 * np.pi * z / L[2])
    w = 0.05 * wall * np.sin(4 * np.pi * x / L[0]) * np.sin(2 * np.pi * z / L[2])
    sim.set_uvw(u, v, w)
    ref.set_uvw(u, v, w)
    sim.step(2)
    ref.step(2)
    a, b = sim.get_uvw(), ref.get_uvw()
    scale = max(np.abs(q).max() for q in b)
    err = max(np.abs(p - q).max() for p, q in zip(a, b)) / scale
    m, rm = sim.monitor(), ref.monitor()
    print(stretching, "strict" if strict else "fast", "2 RK3 steps, rel err %.2e, div_u_max %.2e (oracle %.2e)" % (err, m["div_u_max"], rm["div_u_max"]))
    assert err < 1e-12
    # with walls the corrected field is solenoidal only up to the boundary closures (the oracle shows the same residual)
    assert abs(m["div_u_max"] - rm["div_u_max"]) <= 1e-6 * rm["div_u_max"] + 1e-13
    assert abs(m["enstrophy"] - rm["enstrophy"]) <= 1e-10 * rm["enstrophy"]
    sim.close()


@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_channel_case_steps_vs_oracle(oracle, x3d2, strict):
    """Two RK3 steps of the channel case (case/channel.f90 hooks: bulk-velocity correction, rotation forcing, wall rows
    reset) on the stretched channel mesh against the oracle running the same hooks."""
    dims, L = (64, 65, 32), (4.0, 2.0, 2.0)
    kw = _wall_y(dims, L, "top-bottom", 0.259065151, Re=4200.0, dt=0.005)
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    nz, ny, nx = sim.shape()
    y = ref.geo(1)["vert_coords"][None, :, None]
    x = (np.arange(nx) * (L[0] / nx))[None, None, :]
    z = (np.arange(nz) * (L[2] / nz))[:, None, None]
    wall = (1 - (y - 1.0) ** 2)
    u = wall * (1 + 0.1 * np.sin(2 * np.pi * x / L[0]) * np.cos(2 * np.pi * z / L[2]))
    v = 0.05 * wall ** 2 * np.cos(2 * np.pi * x / L[0]) * np.sin(2 * np.pi * z / L[2])
    w = 0.05 * wall * np.sin(4 * np.pi * x / L[0]) * np.sin(2 * np.pi * z / L[2]) + 0 * y
    for s in (sim, ref):
        s.set_uvw(u, v, w)
        s.set_case_channel(0.3, 2)  # rotation during the first step only
        s.step(2)
    a, b = sim.get_uvw(), ref.get_uvw()
    scale = max(np.abs(q).max() for q in b)
    err = max(np.abs(p - q).max() for p, q in zip(a, b)) / scale
    print("channel case,", "strict" if strict else "fast", "rel err %.2e" % err)
    assert err < 1e-12
    sim.close()


# ---------------------------------------------------------------------------- generic segment-parallel kernels (tds_g.cu)
@pytest.mark.parametrize("dims,stretching,beta,bcs", [
    ((64, 257, 32), "top-bottom", 0.259065151, ((0, 0), (2, 2), (0, 0))),   # channel lines (BASELINE.json configs[4])
    ((64, 129, 32), "uniform", 1.0, ((0, 0), (2, 2), (0, 0))),
    ((32, 128, 64), "centred", 0.8, ((0, 0), (1, 1), (0, 0))),               # free-slip walls: Neumann rows, sym operators
    ((129, 64, 32), "uniform", 1.0, ((2, 1), (0, 0), (0, 0))),               # walls in x: Dirichlet / Neumann mix
])
def test_generic_kernels_walls_and_stretching(oracle, x3d2, dims, stretching, beta, bcs, monkeypatch):
    """Walls, stretched meshes, 257-point lines on the fast path: the generic segment-parallel kernels against the
    oracle (1e-12) on SMOOTH wall-bounded data, where the second-derivative stencil cancels, and a check
    that they, not the one-thread-per-line kernels, served the calls (one launch per transeq instead of three)."""
    wall_dir = [i for i, b in enumerate(bcs) if b[0] != 0][0]
    st = ["uniform"] * 3
    be = [1.0] * 3
    st[wall_dir], be[wall_dir] = stretching, beta
    L = [2 * np.pi] * 3
    L[wall_dir] = 2.0
    kw = dict(L=tuple(L), bcs=bcs, stretching=tuple(st), beta=tuple(be), Re=180.0)
    fast, strict, ref = x3d2.Sim(dims, **kw), x3d2.Sim(dims, strict=True, **kw), oracle.World(dims, **kw)
    generic_transeq = os.environ.get("X3D2C_TRANSEQ_GENERIC") is not None  # read once per process by the library
    nz, ny, nx = fast.shape()
    coords = [ref.geo(d)["vert_coords"] for d in range(3)]
    x, y, z = coords[0][None, None, :], coords[1][None, :, None], coords[2][:, None, None]
    s = [x, y, z][wall_dir]
    wall = np.sin(np.pi * s / 2.0)  # vanishes on the wall at 0, smooth
    o = [q for i, q in enumerate((x, y, z)) if i != wall_dir]
    u = wall * (1 + 0.3 * np.sin(o[0]) * np.cos(o[1])) + 0 * (x + y + z)
    v = 0.2 * wall ** 2 * np.cos(o[0] + o[1]) + 0 * (x + y + z)
    w = 0.1 * wall * np.sin(2 * o[0]) * np.sin(o[1]) + 0.05 * np.cos(np.pi * s) + 0 * (x + y + z)
    d = wall_dir + 1
    worst = {}
    for op in ("der1st", "der1st_sym", "der2nd", "der2nd_sym", "stagder_v2p", "interpl_v2p"):
        worst[op] = rel(fast.tds_solve(d, op, u), ref.tds_solve(d, op, u))
    c = fast.tds_solve(d, "interpl_v2p", u)
    loc = 0 + 10 ** d
    for op in ("stagder_p2v", "interpl_p2v"):
        worst[op] = rel(fast.tds_solve(d, op, c, loc), ref.tds_solve(d, op, c, loc))
    l0 = fast.launch_count()
    got = fast.transeq_dir(d, u, v, w)
    lf = fast.launch_count() - l0
    l0 = strict.launch_count()
    sgot = strict.transeq_dir(d, u, v, w)
    ls = strict.launch_count() - l0
    exp = ref.transeq_dir(d, u, v, w)
    scale = max(np.abs(e).max() for e in exp)
    worst["transeq_dir"] = max(np.abs(g - e).max() for g, e in zip(got, exp)) / scale
    assert all(np.array_equal(a, b) for a, b in zip(sgot, exp))
    exp = ref.transeq(u, v, w)
    worst["transeq"] = max(np.abs(g - e).max() for g, e in zip(fast.transeq(u, v, w), exp)) / max(np.abs(e).max() for e in exp)
    print(dims, stretching, {k: "%.1e" % e for k, e in worst.items()}, "launches fast / strict:", lf, ls)
    assert max(worst.values()) < 1e-12, worst
    if generic_transeq:
        assert ls - lf == 2  # reference-order path: one launch per velocity component; generic kernel: one per call
    fast.close()
    strict.close()
