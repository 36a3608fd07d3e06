"""Non-periodic boundaries, padded blocks and stretched meshes through the C ABI.

These exercise the reference-order kernels (tds_m1.cu) with Dirichlet / Neumann coefficient rows, n_rhs = n_tds + 1
operators, stretch / stretch_correct factors and allocator padding (257 -> 288 style), against the oracle.
"""
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
