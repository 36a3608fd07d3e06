"""CPU statement of the segment decomposition ACROSS ranks that the fast-path kernels use for a periodic direction split
over P ranks (x3d2_b200/csrc/backend/m3_common.cuh: carries(), carries_ext(), push_carries_inline(); m3_edge.cu:
make_op(), edge_kernel): every rank sweeps its 16-point segments locally from zero, hands three rows of boundary carries
to each neighbour, and adds the carry terms. Evaluated with numpy and compared with the dense solution of the same
global periodic recurrences z_j = a z_{j-1} + r_j, y_j = cb y_{j+1} + z_j. It documents the constants (zw, yw, om, W, Cp)
and the meaning of the exchanged rows; the kernels themselves are tested on GPUs (tests/test_gpu_*.py,
tools/mgpu_check.py)."""
import numpy as np
import pytest

S, DMAX = 16, 3


def make_op(a, cb):
    zw = np.array([a ** (S * d) for d in range(DMAX)])
    yw = np.array([cb ** (S * d) for d in range(DMAX)])
    W = np.zeros(S + 1)
    for k in range(S - 1, -1, -1):
        W[k] = a ** (k + 1) + cb * W[k + 1]
    Cp = np.array([cb ** (S - k) for k in range(S)])
    om = np.zeros(2 * DMAX - 1)  # index m + DMAX - 1, m = d - d'
    for d in range(1, DMAX + 1):
        for dp in range(1, DMAX + 1):
            om[d - dp + DMAX - 1] += W[0] * yw[d - 1] * zw[dp - 1]
    return dict(a=a, cb=cb, zw=zw, yw=yw, W=W[:S], Cp=Cp, om=om)


def local_sweeps(o, r):
    """Per segment: forward and backward sweep from zero. Returns the local solution, ze (end of the forward sweep) and
    ys (start of the backward sweep of the local z)."""
    nseg = len(r) // S
    yl, ze, ys = np.zeros(len(r)), np.zeros(nseg), np.zeros(nseg)
    for q in range(nseg):
        z, p = np.zeros(S), 0.0
        for k in range(S):
            p = o["a"] * p + r[q * S + k]
            z[k] = p
        ze[q] = p
        y = 0.0
        for k in range(S - 1, -1, -1):
            y = o["cb"] * y + z[k]
            yl[q * S + k] = y
        ys[q] = yl[q * S]
    return yl, ze, ys


def pushes(o, ze, ys):
    """What a rank sends: to the next rank ze of its last three segments; to the previous rank, per row e, what its first
    segments add to yin of the previous rank's segment nseg' - 3 + e (push_carries_inline / edge_kernel)."""
    nseg = len(ze)
    to_next = ze[nseg - DMAX:].copy()
    to_prev = np.zeros(DMAX)
    for e in range(DMAX):
        acc = 0.0
        for d in range(1, DMAX + 1):
            if e + d - DMAX >= 0:
                acc += o["yw"][d - 1] * ys[e + d - DMAX]
        for m in range(1, DMAX):
            if e + m - DMAX >= 0:
                acc += o["om"][m + DMAX - 1] * ze[e + m - DMAX]
        to_prev[e] = acc
    return to_prev, to_next


def finish(o, yl, ze, ys, from_prev, from_next):
    """carries() with DIST = true: segments beyond the line ends belong to the neighbours."""
    nseg = len(ze)
    out = yl.copy()
    for q in range(nseg):
        zv = np.zeros(2 * DMAX)  # ze(q - DMAX .. q + DMAX - 1)
        for t in range(2 * DMAX):
            s = q - DMAX + t
            zv[t] = from_prev[s + DMAX] if s < 0 else (ze[s] if s < nseg else 0.0)
        zin = sum(o["zw"][d - 1] * zv[DMAX - d] for d in range(1, DMAX + 1))
        r = q - (nseg - DMAX)
        yin = from_next[r] if r >= 0 else 0.0
        for d in range(1, DMAX + 1):
            if q + d < nseg:
                yin += o["yw"][d - 1] * ys[q + d]
        for m in range(-(DMAX - 1), DMAX):
            yin += o["om"][m + DMAX - 1] * zv[DMAX + m]
        out[q * S:(q + 1) * S] += o["W"] * zin + o["Cp"] * yin
    return out


@pytest.mark.parametrize("P,n_loc", [(1, 96), (2, 96), (2, 128), (4, 96), (8, 128)])
@pytest.mark.parametrize("a,cb", [(-0.2679491924311227, -0.2679491924311227), (0.33, -0.41), (-0.45, 0.1)])
def test_rank_split_segments_reproduce_the_global_periodic_solve(P, n_loc, a, cb):
    n = P * n_loc
    rng = np.random.default_rng(P * 1000 + n_loc)
    r = rng.standard_normal(n)
    shift_down = np.roll(np.eye(n), 1, axis=0)   # (shift_down z)_j = z_{j-1}, periodic
    z = np.linalg.solve(np.eye(n) - a * shift_down, r)
    y = np.linalg.solve(np.eye(n) - cb * shift_down.T, z)
    o = make_op(a, cb)
    parts = [local_sweeps(o, r[p * n_loc:(p + 1) * n_loc]) for p in range(P)]
    sent = [pushes(o, ze, ys) for (_, ze, ys) in parts]
    got = np.concatenate([finish(o, *parts[p], from_prev=sent[(p - 1) % P][1], from_next=sent[(p + 1) % P][0])
                          for p in range(P)])
    assert np.abs(got - y).max() <= 2e-15 * np.abs(y).max()
