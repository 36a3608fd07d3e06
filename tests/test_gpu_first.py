"""First GPU parity checks: every backend op through the C ABI vs the oracle (fast and strict modes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module", params=[False, True], ids=["fast", "strict"])
def pair(request, oracle, x3d2):
    dims = (64, 32, 96)
    sim = x3d2.Sim(dims, strict=request.param)
    ref = oracle.World(dims)
    yield sim, ref, request.param
    sim.close()


def rnd(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


def test_reorders(pair):
    sim, ref, _ = pair
    f = rnd(sim.shape(), 1)
    for chain in (["C2X", "X2C"], ["C2Y", "Y2C"], ["C2Z", "Z2C"], ["C2X", "X2Y", "Y2X", "X2C"], ["C2X", "X2Z", "Z2X"],
                  ["C2X", "X2Y", "Y2Z", "Z2X"], ["C2X", "X2Z", "Z2Y", "Y2X"], ["C2Z", "Z2C", "C2Y", "Y2C"]):
        assert np.array_equal(sim.reorder_chain(f, chain), f), chain


def test_sum_intox_vecadd(pair):
    sim, ref, strict = pair
    a, b = rnd(sim.shape(), 2), rnd(sim.shape(), 3)
    for d in (2, 3):
        assert np.array_equal(sim.sum_intox(d, a, b), ref.sum_intox(d, a, b))
        assert np.all(sim.sum_intox(d, a, -a) == 0)
    for d in (1, 2, 3):
        got, exp = sim.vecadd(d, 0.3, a, -1.7, b), ref.vecadd(d, 0.3, a, -1.7, b)
        if strict:
            assert np.array_equal(got, exp)
        else:
            assert rel(got, exp) < 1e-15


def test_reductions(pair):
    sim, ref, _ = pair
    a, b = rnd(sim.shape(), 4), rnd(sim.shape(), 5)
    for d in (1, 2, 3):
        s, e = sim.scalar_product(d, a, b), ref.scalar_product(d, a, b)
        assert abs(s - e) < 1e-11 * np.abs(a * b).sum()
        (mx, mean), (emx, emean) = sim.field_max_mean(d, a), ref.field_max_mean(d, a)
        assert mx == emx and abs(mean - emean) < 1e-13


@pytest.mark.parametrize("op", ["der1st", "der2nd", "stagder_v2p", "stagder_p2v", "interpl_v2p", "interpl_p2v"])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_tds_solve(pair, op, d):
    sim, ref, strict = pair
    loc = 1110 if op.endswith("p2v") else 0  # p2v operators act on cell-centred data
    f = rnd(sim.shape(loc), 6)
    got, exp = sim.tds_solve(d, op, f, loc), ref.tds_solve(d, op, f, loc)
    if strict:
        assert np.array_equal(got, exp)
    else:
        assert rel(got, exp) < 1e-12


@pytest.mark.parametrize("d", [1, 2, 3])
def test_transeq_dir(pair, d):
    sim, ref, strict = pair
    u, v, w = rnd(sim.shape(), 7), rnd(sim.shape(), 8), rnd(sim.shape(), 9)
    got, exp = sim.transeq_dir(d, u, v, w), ref.transeq_dir(d, u, v, w)
    for g, e in zip(got, exp):
        if strict:
            assert np.array_equal(g, e)
        else:
            assert rel(g, e) < 1e-12


def test_transeq_full(pair):
    sim, ref, strict = pair
    u, v, w = rnd(sim.shape(), 10), rnd(sim.shape(), 11), rnd(sim.shape(), 12)
    got, exp = sim.transeq(u, v, w), ref.transeq(u, v, w)
    for g, e in zip(got, exp):
        if strict:
            assert np.array_equal(g, e)
        else:
            assert rel(g, e) < 1e-12


def test_div_grad_curl(pair):
    sim, ref, strict = pair
    u, v, w = rnd(sim.shape(), 13), rnd(sim.shape(), 14), rnd(sim.shape(), 15)
    tol = 0 if strict else 1e-12
    assert rel(sim.divergence(u, v, w), ref.divergence(u, v, w)) <= tol
    for g, e in zip(sim.gradient(u), ref.gradient(u)):
        assert rel(g, e) <= tol
    for g, e in zip(sim.curl(u, v, w), ref.curl(u, v, w)):
        assert rel(g, e) <= tol


def test_fft_and_poisson(pair):
    sim, ref, _ = pair
    nz, ny, nx = sim.shape()
    x, y, z = (np.arange(n) * 2 * np.pi / n for n in (nx, ny, nz))
    f = np.sin(x)[None, None, :] * np.cos(y)[None, :, None] * np.cos(z)[:, None, None] + 2 * x[None, None, :]
    out, spec = sim.fft_roundtrip(f, True)
    assert np.sqrt(np.mean((out / f.size - f) ** 2)) < 1e-10  # tests/verification/test_fft.f90:56
    _, espec = ref.fft_roundtrip(f, True)
    assert np.abs(spec - espec).max() < 1e-12 * np.abs(espec).max()
    g = rnd(sim.shape(), 16)
    g -= g.mean()
    assert rel(sim.poisson(g), ref.poisson(g)) < 1e-12


def test_tgv_steps(oracle, x3d2):
    for strict in (False, True):
        sim = x3d2.Sim((64, 64, 64), strict=strict)
        ref = oracle.World((64, 64, 64))
        sim.init_tgv()
        ref.init_tgv()
        a, b = sim.get_uvw(), ref.get_uvw()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        sim.step(1)
        ref.step(1)
        a, b = sim.get_uvw(), ref.get_uvw()
        scale = max(np.abs(y).max() for y in b)
        err = max(np.abs(x - y).max() for x, y in zip(a, b)) / scale
        assert err < 1e-12, err
        m, e = sim.monitor(), ref.monitor()
        assert abs(m["enstrophy"] - e["enstrophy"]) < 1e-10 * e["enstrophy"]
        assert abs(m["ke"] - e["ke"]) < 1e-10 * e["ke"]
        sim.close()
