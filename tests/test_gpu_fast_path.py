"""Fast path (segment-parallel register-resident kernels, tds_m3.cu) vs the oracle and vs the strict kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12  # north_star: derivatives and fields within 1e-12 relative in FP64


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def rnd(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


@pytest.mark.parametrize("dims", [(64, 64, 64), (128, 96, 80), (256, 64, 128), (512, 64, 64), (64, 1024, 64)])
def test_fast_vs_oracle_and_strict(oracle, x3d2, dims):
    fast, strict, ref = x3d2.Sim(dims), x3d2.Sim(dims, strict=True), oracle.World(dims)
    u, v, w = rnd(fast.shape(), 1), rnd(fast.shape(), 2), rnd(fast.shape(), 3)
    worst = 0.0
    for d in (1, 2, 3):
        for op in ("der1st", "der2nd", "stagder_v2p", "interpl_v2p"):
            e = ref.tds_solve(d, op, u)
            assert np.array_equal(strict.tds_solve(d, op, u), e)
            worst = max(worst, rel(fast.tds_solve(d, op, u), e))
        for op in ("stagder_p2v", "interpl_p2v"):
            e = ref.tds_solve(d, op, u, 1110)
            worst = max(worst, rel(fast.tds_solve(d, op, u, 1110), e))
        exp = ref.transeq_dir(d, u, v, w)
        for g, s, e in zip(fast.transeq_dir(d, u, v, w), strict.transeq_dir(d, u, v, w), exp):
            assert np.array_equal(s, e)
            worst = max(worst, rel(g, e))
    print(dims, "worst fast-path rel err", worst)
    assert worst < TOL
    fast.close()
    strict.close()


def test_tgv_256_one_step(oracle, x3d2):
    """BASELINE.json configs[1] size: one RK3 step at 256^3, fields within 1e-12 of the OMP oracle."""
    n = 256
    sim, ref = x3d2.Sim((n, n, n)), oracle.World((n, n, n))
    sim.init_tgv()
    ref.init_tgv()
    sim.step(1)
    ref.step(1)
    a, b = sim.get_uvw(), ref.get_uvw()
    scale = max(np.abs(y).max() for y in b)
    assert max(np.abs(x - y).max() for x, y in zip(a, b)) / scale < TOL
    sim.close()


@pytest.mark.parametrize("dims", [(96, 128, 112), (512, 96, 256)])
def test_rank_split_kernels_on_one_gpu(oracle, x3d2, dims, monkeypatch):
    """The rank-split variant of the fast path (halo rows + neighbour carries from m3_edge.cu) with the rank acting
    as its own periodic neighbour (X3D2C_FORCE_DIST): same answers as the oracle, plus a TGV step."""
    monkeypatch.setenv("X3D2C_FORCE_DIST", "1")
    sim, ref = x3d2.Sim(dims), oracle.World(dims)
    monkeypatch.delenv("X3D2C_FORCE_DIST")
    plain = x3d2.Sim(dims)
    u, v, w = rnd(sim.shape(), 4), rnd(sim.shape(), 5), rnd(sim.shape(), 6)
    worst = 0.0
    for d in (1, 2, 3):
        for op in ("der1st", "der2nd", "stagder_v2p", "interpl_v2p"):
            worst = max(worst, rel(sim.tds_solve(d, op, u), ref.tds_solve(d, op, u)))
        for op in ("stagder_p2v", "interpl_p2v"):
            worst = max(worst, rel(sim.tds_solve(d, op, u, 1110), ref.tds_solve(d, op, u, 1110)))
        got, exp, one = sim.transeq_dir(d, u, v, w), ref.transeq_dir(d, u, v, w), plain.transeq_dir(d, u, v, w)
        for g, e, o in zip(got, exp, one):
            worst = max(worst, rel(g, e))
            assert rel(g, o) < 1e-13  # same segments, same carries: only the edge kernel's rounding may differ
    assert sim.launch_count() > plain.launch_count()  # pack + edge kernels + exchanges were really used
    print(dims, "worst rank-split rel err", worst)
    assert worst < TOL
    sim.close()
    plain.close()


@pytest.mark.parametrize("dims,bcs,env", [((128, 96, 112), None, {}), ((128, 96, 112), None, {"X3D2C_FORCE_DIST": "1"}),
                                          ((65, 64, 64), ((2, 2), (0, 0), (1, 1)), {})])
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_fused_tds_combinations(oracle, x3d2, dims, bcs, env, strict, monkeypatch):
    """x3d2c_tds_solve_sum / _dual / _axpy equal the reference's call sequences (two tds_solve [+ vecadd]):
    bit-exact in strict mode, 1e-12 on the fast path (single-rank and rank-split kernels, non-periodic fallback)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    kw = dict(bcs=bcs) if bcs else {}
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    tol = 0 if strict else TOL
    u, v = rnd(sim.shape(), 7), rnd(sim.shape(), 8)
    c = rnd(sim.shape(1110), 9)
    for d in (1, 2, 3):
        e = ref.tds_solve(d, "interpl_v2p", u) + ref.tds_solve(d, "stagder_v2p", v)
        assert rel(sim.tds_fused("sum", d, "interpl_v2p", "stagder_v2p", u, v), e) <= tol, ("sum", d)
        ga, gb = sim.tds_fused("dual", d, "interpl_p2v", "stagder_p2v", c, in_loc=1110)
        assert rel(ga, ref.tds_solve(d, "interpl_p2v", c, 1110)) <= tol, ("dual a", d)
        assert rel(gb, ref.tds_solve(d, "stagder_p2v", c, 1110)) <= tol, ("dual b", d)
        y = rnd(ga.shape, 10 + d)
        e = -1.0 * ref.tds_solve(d, "stagder_p2v", c, 1110) + y
        assert rel(sim.tds_fused("axpy", d, "stagder_p2v", None, c, y, a=-1.0, in_loc=1110), e) <= tol, ("axpy", d)
    sim.close()


RDR_CASES = [(2, 0, 23), (2, 0, 24), (2, 32, 0), (2, 42, 23), (3, 43, 0), (3, 23, 0), (3, 0, 34), (3, 43, 32),
             (1, 0, 12), (1, 21, 13), (1, 21, 0)]  # (dir, rdr_in, rdr_out); X lines: x2y on the output / y2x on the input
                                                   # go through swizzled tiles (tds_m4.cu XT), (1, 21, 13) as a sequence


@pytest.mark.parametrize("dims,bcs,env", [((128, 64, 256), None, {}), ((128, 128, 256), None, {"X3D2C_FORCE_DIST": "1"}),
                                          ((96, 64, 80), None, {}), ((65, 64, 64), ((2, 2), (0, 0), (1, 1)), {}),
                                          ((1024, 64, 64), None, {}), ((512, 64, 64), None, {}), ((256, 64, 64), None, {})])
@pytest.mark.parametrize("strict", [False, True], ids=["fast", "strict"])
def test_tds_through_reorders(oracle, x3d2, dims, bcs, env, strict, monkeypatch):
    """x3d2c_tds_solve_r / _sum_r / _dual_r == reorder -> operator(s) -> reorder. Host arrays are Cartesian, so the
    reorders are invisible in the result: it must equal the plain operator (tensor-map path, rank-split kernels with
    explicit input reorders, sequence fallback)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    kw = dict(bcs=bcs) if bcs else {}
    sim, ref = x3d2.Sim(dims, strict=strict, **kw), oracle.World(dims, **kw)
    tol = 0 if strict else TOL
    u, v = rnd(sim.shape(), 21), rnd(sim.shape(), 22)
    c = rnd(sim.shape(1110), 23)
    for d, rin, rout in RDR_CASES:
        e1 = ref.tds_solve(d, "interpl_v2p", u)
        assert rel(sim.tds_fused_r("single", d, "interpl_v2p", None, u, rdr_in=rin, rdr_out=rout), e1) <= tol, (d, rin, rout)
        e2 = e1 + ref.tds_solve(d, "stagder_v2p", v)
        assert rel(sim.tds_fused_r("sum", d, "interpl_v2p", "stagder_v2p", u, v, rdr_in=rin, rdr_out=rout), e2) <= tol
        ga, gb = sim.tds_fused_r("dual", d, "interpl_p2v", "stagder_p2v", c, rdr_in=rin, rdr_out=rout, in_loc=1110)
        assert rel(ga, ref.tds_solve(d, "interpl_p2v", c, 1110)) <= tol, ("dual a", d, rin, rout)
        assert rel(gb, ref.tds_solve(d, "stagder_p2v", c, 1110)) <= tol, ("dual b", d, rin, rout)
    # axpy through an input reorder (the gradient's y2x + pressure correction): y - A(c)
    for d, rin in [(1, 21), (1, 0), (3, 23)]:
        y = rnd(sim.shape(1110 - 10 ** d), 24)
        for op in ("stagder_p2v", "interpl_p2v"):
            e = y - ref.tds_solve(d, op, c, 1110)
            assert rel(sim.tds_fused_r("axpy", d, op, None, c, y, rdr_in=rin, in_loc=1110), e) <= tol, ("axpy_r", d, rin, op)
    with pytest.raises(RuntimeError, match="rdr_out must be a reorder code that starts from dir"):
        sim.tds_fused_r("single", 2, "interpl_v2p", None, u, rdr_out=34)
    sim.close()


def test_full_size_512_properties(x3d2):
    """BASELINE.json configs[2] size (TGV 512^3, the benchmarked workload): the oracle would need minutes, so the
    check is through size-independent properties of the same code paths."""
    n = 512
    sim = x3d2.Sim((n, n, n))
    sim.init_tgv()
    m0 = sim.monitor()
    # TGV at t = 0: <u^2>/2 = 1/8 exactly, enstrophy 3/8 up to the O(dx^6) error of the compact derivative
    assert abs(m0["ke"] - 0.125) < 1e-13 and abs(m0["enstrophy"] - 0.375) < 1e-9 and m0["div_u_max"] < 1e-11
    sim.step(1)
    m1 = sim.monitor()
    # the corrected field is solenoidal to rounding, and the energy decays at the viscous rate dE/dt = -2 nu Omega
    nu, dt = 1.0 / 1600.0, 1e-3
    assert m1["div_u_max"] < 1e-11
    assert abs((m0["ke"] - m1["ke"]) / dt - 2 * nu * 0.375) < 2e-3 * (2 * nu * 0.375)
    # linearity of a distributed-tridiagonal operator and exact reorder round trip on full-size random fields
    rng = np.random.default_rng(12)
    a, b = rng.standard_normal(sim.shape()), rng.standard_normal(sim.shape())
    da, db = sim.tds_solve(3, "der1st", a), sim.tds_solve(3, "der1st", b)
    dc = sim.tds_solve(3, "der1st", 0.5 * a - 2.0 * b)
    assert rel(dc, 0.5 * da - 2.0 * db) < TOL
    assert abs(da.mean()) < 1e-10 * np.abs(da).max()  # the derivative of a periodic field has zero mean
    assert np.array_equal(sim.reorder_chain(a, ["C2X", "X2Y", "Y2Z", "Z2C"]), a)
    sim.close()
