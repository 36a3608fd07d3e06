"""CPU statement of the generic segment-parallel DistD2 algorithm that x3d2_b200/csrc/backend/tds_g.cu implements
(per-row recurrences, per-segment carry weights, substitution with the line's y_1 and z_n), evaluated with numpy from
the tables of tdsops_init and compared with the oracle's reference-order solve. It documents the table construction of
build_tables() and catches algorithmic slips without a GPU; the kernels themselves are tested in tests/test_gpu_*.py."""
import numpy as np
import pytest

S, DMAX = 16, 3


def build_tables(t):
    n, n_rhs = t.n_tds, t.n_rhs
    nseg = (max(n, n_rhs) + S - 1) // S
    npad = nseg * S
    fw, bw, sa, sc, af = t.dist_fw, t.dist_bw, t.dist_sa, t.dist_sc, t.dist_af
    A, B, C, E, SA, SC, ST = (np.zeros(npad) for _ in range(7))
    for j in range(1, n + 1):
        r = j - 1
        if j <= 2:
            A[r], B[r] = 0.0, af[r]
        else:
            A[r], B[r] = -fw[r] * af[r], fw[r]
        if j == 1:
            C[r], E[r] = -fw[0] * bw[0], fw[0]
        else:
            C[r], E[r] = (-bw[r] if j <= n - 2 else 0.0), 1.0
        if 2 <= j <= n - 1:
            SA[r], SC[r] = sa[r], sc[r]
        ST[r] = t.stretch[r]
    W, Cp = np.zeros(npad), np.zeros(npad)
    alpha, g, beta = np.zeros(nseg), np.zeros(nseg), np.zeros(nseg)
    for q in range(nseg):
        pf = np.cumprod(A[q * S:(q + 1) * S])
        alpha[q] = pf[-1]
        y, cp = 0.0, 1.0
        for k in range(S - 1, -1, -1):
            r = q * S + k
            y = C[r] * y + E[r] * pf[k]
            W[r] = y
            cp *= C[r]
            Cp[r] = cp
        g[q], beta[q] = W[q * S], Cp[q * S]

    def ZW(q, d):
        return 0.0 if q - d < 0 else float(np.prod(alpha[q - d + 1:q]))

    seg = np.zeros((nseg, 11))
    for q in range(nseg):
        for d in range(1, DMAX + 1):
            seg[q, d - 1] = ZW(q, d)
            if q + d < nseg:
                seg[q, 3 + d - 1] = float(np.prod(beta[q + 1:q + d]))
        for d in range(1, DMAX + 1):
            for e in range(1, DMAX + 1):
                if q + d < nseg and abs(d - e) <= DMAX - 1:
                    seg[q, 6 + (d - e) + DMAX - 1] += seg[q, 3 + d - 1] * g[q + d] * ZW(q + d, e)
    return dict(A=A, B=B, C=C, E=E, W=W, Cp=Cp, SA=SA, SC=SC, ST=ST, seg=seg, nseg=nseg, npad=npad)


def solve_segments(t, u):
    """u: [n_lines, >= n_rhs] -> x: [n_lines, n_tds] with the segment-parallel algorithm (periodic wrap for the halo)."""
    T = build_tables(t)
    n, n_rhs, nseg, npad = t.n_tds, t.n_rhs, T["nseg"], T["npad"]
    nl = u.shape[0]
    up = np.zeros((nl, npad))
    up[:, :n_rhs] = u[:, :n_rhs]
    rhs = np.zeros((nl, npad))
    for j in range(1, npad + 1):
        if j <= 4:
            c = t.coeffs_s[j - 1]
        elif j > n_rhs:
            continue
        elif j >= n_rhs - 3:
            c = t.coeffs_e[j - (n_rhs - 3)]
        else:
            c = t.coeffs
        idx = [(j - 1 + k - 4) % npad for k in range(9)]
        rhs[:, j - 1] = sum(c[k] * up[:, idx[k]] for k in range(9))
    z = np.zeros((nl, npad))
    fe, bs = np.zeros((nl, nseg)), np.zeros((nl, nseg))
    for q in range(nseg):
        pz = np.zeros(nl)
        for k in range(S):
            r = q * S + k
            pz = T["A"][r] * pz + T["B"][r] * rhs[:, r]
            z[:, r] = pz
        fe[:, q] = pz
        y = np.zeros(nl)
        for k in range(S - 1, -1, -1):
            r = q * S + k
            y = T["C"][r] * y + T["E"][r] * z[:, r]
            z[:, r] = y
        bs[:, q] = y
    yfin = np.zeros_like(z)
    for q in range(nseg):
        sw = T["seg"][q]
        zin, yin = np.zeros(nl), np.zeros(nl)
        for d in range(1, DMAX + 1):
            zin += sw[d - 1] * fe[:, max(q - d, 0)]
            yin += sw[3 + d - 1] * bs[:, min(q + d, nseg - 1)]
        for m in range(-(DMAX - 1), DMAX):
            yin += sw[6 + m + DMAX - 1] * fe[:, min(max(q + m, 0), nseg - 1)]
        for k in range(S):
            r = q * S + k
            yfin[:, r] = z[:, r] + T["W"][r] * zin + T["Cp"][r] * yin
    y1, zn = yfin[:, 0], yfin[:, n - 1]
    sa1, scn = t.dist_sa[0], t.dist_sc[n - 1]
    s = (y1 - sa1 * zn) / (1 - sa1 * sa1)
    e = (zn - scn * y1) / (1 - scn * scn)
    x = (yfin - T["SA"] * s[:, None] - T["SC"] * e[:, None])
    x[:, 0], x[:, n - 1] = s, e
    return (x * T["ST"])[:, :n]


CASES = [("first-deriv", "compact6", None), ("second-deriv", "compact6", None), ("stag-deriv", "compact6", "v2p"),
         ("stag-deriv", "compact6", "p2v"), ("interpolate", "classic", "v2p"), ("interpolate", "classic", "p2v")]


@pytest.mark.parametrize("operation,scheme,from_to", CASES)
@pytest.mark.parametrize("bc", [(0, 0), (1, 1), (2, 2), (2, 1)])
@pytest.mark.parametrize("n_vert", [64, 65, 257])
@pytest.mark.parametrize("sym", [False, True])
def test_segment_algorithm_matches_reference_order(oracle, operation, scheme, from_to, bc, n_vert, sym):
    O = oracle
    if from_to and 2 in bc:
        bc = tuple(1 if b == 2 else b for b in bc)  # midpoint operators take Neumann closures on Dirichlet walls
    periodic = bc == (0, 0)
    if periodic and n_vert % S:
        pytest.skip("a periodic line wraps inside the tile: n must be a multiple of 16")
    if from_to and sym:
        pytest.skip("sym is ignored by the midpoint operators")
    n_cell = n_vert if periodic else n_vert - 1
    n_tds = n_cell if from_to == "v2p" else n_vert
    rng = np.random.default_rng(8)
    st, stc = 1 + 0.3 * rng.random(n_tds), None
    t = O.Tdsops(n_tds, 0.05, operation, scheme, bc[0], bc[1], st, stc, from_to=from_to, sym=sym)
    n_pad = ((max(t.n_rhs, n_tds) + 31) // 32) * 32
    u = np.zeros((1, 16, n_pad))
    n_in = t.n_rhs if from_to != "p2v" else (n_vert if periodic else n_vert - 1)
    u[0, :, :n_in] = rng.standard_normal((16, n_in))
    exp = O.lines_tds_solve([t], u)[0, :, :n_tds]
    got = solve_segments(t, u[0])
    err = np.abs(got - exp).max() / np.abs(exp).max()
    assert err < 1e-13, err
