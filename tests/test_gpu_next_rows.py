"""SURVEY.md 8(f) rows 3 and 4 through the C ABI: transeq_lowmem (src/solver.f90:391-505), species transport
(transeq_species, src/solver.f90:507-600 + omp/backend.f90:186-233), the output-side operators compute_vorticity /
compute_qcriterion (omp/backend.f90:616-649) and slice_max_sum (:816-881), against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rnd(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("dims,bcs", [((64, 64, 64), None), ((96, 64, 80), None), ((65, 64, 33), ((2, 2), (0, 0), (1, 1)))])
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_transeq_lowmem(oracle, x3d2, dims, bcs, strict):
    kw = dict(bcs=bcs) if bcs else {}
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    u, v, w = rnd(sim.shape(), 1), rnd(sim.shape(), 2), rnd(sim.shape(), 3)
    got, exp = sim.transeq_lowmem(u, v, w), ref.transeq_lowmem(u, v, w)
    dflt = ref.transeq(u, v, w)
    for g, e, d in zip(got[:3], exp[:3], dflt):
        assert np.array_equal(e, d)  # the low-memory sequence computes the same right-hand side (solver.f90:392)
        assert np.array_equal(g, e) if strict else rel(g, e) < TOL
    assert np.array_equal(got[3], u) and np.array_equal(exp[3], u)  # the velocity survives its x -> y -> z -> x trip
    sim.close()


@pytest.mark.parametrize("dims,bcs", [((64, 64, 64), None), ((128, 64, 96), None), ((65, 64, 33), ((2, 2), (0, 0), (1, 1)))])
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_transeq_species(oracle, x3d2, dims, bcs, strict):
    kw = dict(bcs=bcs) if bcs else {}
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    u, v, w, phi = (rnd(sim.shape(), s) for s in (4, 5, 6, 7))
    for nu_s in (1.0 / 1600.0, 0.7):
        got, exp = sim.transeq_species(u, v, w, phi, nu_s), ref.transeq_species(u, v, w, phi, nu_s)
        assert np.array_equal(got, exp) if strict else rel(got, exp) < TOL, nu_s
    sim.close()


def test_derived_fields_and_slice(oracle, x3d2):
    dims = (64, 48, 40)
    sim = x3d2.Sim(dims)
    g = [rnd(sim.shape(), 10 + k) for k in range(9)]
    assert np.array_equal(sim.derived("vorticity", g), oracle.compute_vorticity(g))
    assert np.array_equal(sim.derived("qcriterion", g), oracle.compute_qcriterion(g))
    f = rnd(sim.shape(), 30)
    for d, n in ((1, dims[0]), (2, dims[1]), (3, dims[2])):
        for i_slice in (1, n // 2, n):
            mx, sm = sim.slice_max_sum(d, f, i_slice)
            emx, esm = oracle.slice_max_sum(f, d, i_slice)
            assert mx == emx and abs(sm - esm) < 1e-12 * max(1.0, abs(esm) + np.abs(f).sum() / n), (d, i_slice)
    with pytest.raises(RuntimeError, match="i_slice out of range"):
        sim.slice_max_sum(1, f, dims[0] + 1)
    sim.close()
    # vorticity of the Taylor-Green vortex from real gradients: |curl u| through der1st solves == the curl operator
    n = 64
    sim, ref = x3d2.Sim((n, n, n)), oracle.World((n, n, n))
    sim.init_tgv()
    u, v, w = sim.get_uvw()
    grads = [sim.tds_solve(d, "der1st", f) for f in (u, v, w) for d in (1, 2, 3)]
    ox, oy, oz = ref.curl(u, v, w)
    assert rel(sim.derived("vorticity", grads), np.sqrt(ox * ox + oy * oy + oz * oz)) < TOL
    sim.close()


def test_step_batches_equals_set_step_get(x3d2):
    """Sim.step_batches (uploads / downloads on copy lanes, overlapped with the kernels of the neighbouring batches) gives
    every batch exactly what set_uvw -> step -> get_uvw gives; padded grids and Adams-Bashforth are refused."""
    n = 64
    sim = x3d2.Sim((n, n, n))
    sim.init_tgv()
    sim.step(2)
    ins = [a.copy() for a in sim.get_uvw()]
    sim.set_uvw(*ins)
    sim.step(1)
    exp = sim.get_uvw()
    outs = [np.full(sim.shape(), np.nan) for _ in range(3)]
    for nb in (1, 2, 5):
        for o in outs:
            o.fill(np.nan)
        sim.step_batches(nb, ins, outs)
        assert all(np.array_equal(o, e) for o, e in zip(outs, exp)), nb
    sim.close()
    sim = x3d2.Sim((48, 40, 36))  # nx - 1 not a multiple of 32: padded blocks
    a = [np.zeros(sim.shape()) for _ in range(3)]
    with pytest.raises(RuntimeError, match="needs padding"):
        sim.step_batches(1, a, a)
    sim.close()


@pytest.mark.parametrize("dims,bcs", [((64, 64, 64), None), ((128, 64, 32), None), ((64, 65, 32), ((0, 0), (2, 2), (0, 0)))])
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_interpl_c2v_and_laplacian_vs_oracle(oracle, x3d2, dims, bcs, strict):
    """vector_calculus_t%interpl_c2v (through-reorder solves on the fast path) and %laplacian against the oracle."""
    kw = dict(bcs=bcs) if bcs else {}
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    p = rnd(sim.shape(1110), 41)
    u = rnd(sim.shape(), 42)
    for got, exp in ((sim.interpl_c2v(p), ref.interpl_c2v(p)), (sim.laplacian(u), ref.laplacian(u))):
        assert np.array_equal(got, exp) if strict else rel(got, exp) < TOL
    sim.close()
