"""Pins the oracle's DistD2-TDS operators (coefficient tables of src/tdsops.f90:205-872, the factorisation of
:874-931 and the kernels of omp/kernels/distributed.f90:11-229) to an INDEPENDENT statement of the same schemes:
dense systems A f = B u assembled from the published scheme constants by ghost-point reflection / Lele's closures in
exact rational arithmetic and solved in 80-bit long double (tests/_schemes.py). Bar: 1e-13 relative - five orders
tighter than the reference's own analytic tests (1e-8) and below the 1e-12 parity bar of the GPU path - or, where the
evaluation itself is ill-conditioned (second derivative of a smooth field: the stencil cancels to O(h^2) of its
terms), 8 eps x the componentwise condition number computed in long double: the floor of ANY double
evaluation. A slip of relative size d in a coefficient moves the result by ~ d x cond, so slips above ~2e-15 show.
"""
import numpy as np
import pytest

import _schemes as S

P_, N_, D_ = S.PERIODIC, S.NEUMANN, S.DIRICHLET
BAR = 1e-13

# (oracle operation, oracle scheme, from_to, independent scheme name, admissible (bc, input parity) combinations)
OPS = [
    ("first-deriv", "compact6", None, "der1st"),
    ("second-deriv", "compact6", None, "der2nd"),
    ("stag-deriv", "compact6", "v2p", "stagder"),
    ("stag-deriv", "compact6", "p2v", "stagder"),
    ("interpolate", "classic", "v2p", "interpl"),
    ("interpolate", "classic", "p2v", "interpl"),
]
BCS = [(P_, P_), (N_, N_), (D_, D_), (D_, N_), (N_, D_)]


def _fields(n_in, n_lines, h, kind):
    rng = np.random.default_rng(11)
    if kind == "random":
        return rng.standard_normal((n_lines, n_in))
    x = np.arange(n_in) * h
    ph = rng.random((n_lines, 1)) * 2 * np.pi
    return np.sin(x[None, :] + ph) + 0.3 * np.cos(3 * x[None, :] - ph)  # smooth, generic (neither even nor odd)


@pytest.mark.parametrize("operation,scheme,from_to,name", OPS)
@pytest.mark.parametrize("bc", BCS)
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("data", ["random", "smooth"])
def test_oracle_operator_vs_independent_longdouble(oracle, operation, scheme, from_to, name, bc, sym, data):
    O = oracle
    stag = from_to is not None
    if stag and D_ in bc:
        pytest.skip("midpoint operators take Neumann closures on Dirichlet walls (solver.f90:236-245)")
    if stag and sym:
        pytest.skip("the parity of midpoint operators is fixed by the operator, `sym` is ignored")
    # parity of the input about a free-slip wall (tdsops.f90: 'sym is always ...' comments restated as physics: the
    # interpolated quantity is even, the staggered derivative acts on an odd field on the way to the cells and on an
    # even one on the way back)
    even = sym if not stag else (name == "interpl" or from_to == "p2v")
    n_vert, nl = 48, 16
    periodic = bc == (P_, P_)
    h = 2 * np.pi / n_vert if periodic else np.pi / (n_vert - 1)
    A, B, n_out, n_in = S.build(name, n_vert, bc[0], bc[1], from_to, even)
    u = _fields(n_in, nl, h, data)
    exp, cond = S.solve_longdouble(A, B, u, h, S.SCHEMES[name]["order"], want_cond=True)
    # the oracle: one rank, the reference's calling convention (tests/verification/test_omp_tridiag.f90:365-404)
    n_tds = n_out
    op = O.Tdsops(n_tds, h, operation, scheme, bc[0], bc[1], from_to=from_to, sym=sym)
    n_pad = max(op.n_rhs, n_in)
    up = np.zeros((1, nl, n_pad))
    up[0, :, :n_in] = u
    got = O.lines_tds_solve([op], up)[0, :, :n_out]
    scale = np.abs(exp).max()
    err = float(np.abs(got - exp).max() / scale)
    assert err < max(BAR, 8 * 2.2e-16 * cond), (operation, from_to, bc, sym, data, err, cond)


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("operation,scheme,from_to,name", OPS[:2] + OPS[2:3] + OPS[5:6])
def test_oracle_multi_rank_vs_independent(oracle, P, operation, scheme, from_to, name):
    """The P-rank emulation (halo rows + 2x2 reduced systems, exec_dist.f90:16-65) against the same dense solve."""
    O = oracle
    n_vert, nl = 32 * P, 16
    h = 2 * np.pi / n_vert
    A, B, n_out, n_in = S.build(name, n_vert, P_, P_, from_to, True)
    u = _fields(n_in, nl, h, "smooth")
    exp, cond = S.solve_longdouble(A, B, u, h, S.SCHEMES[name]["order"], want_cond=True)
    n = n_vert // P
    ops = [O.Tdsops(n, h, operation, scheme, 0 if P == 1 else O.BC_HALO, 0 if P == 1 else O.BC_HALO, from_to=from_to)
           for _ in range(P)]
    up = np.stack([u[:, r * n:(r + 1) * n] for r in range(P)])
    got = np.concatenate(list(O.lines_tds_solve(ops, up)), axis=1)
    err = float(np.abs(got - exp).max() / np.abs(exp).max())
    assert err < max(BAR, 8 * 2.2e-16 * cond), (operation, from_to, P, err, cond)
