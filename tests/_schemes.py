"""Independent statement of the compact finite-difference operators (TEST INFRASTRUCTURE ONLY).

Nothing here is transcribed from the reference's coefficient tables (src/tdsops.f90) or from the oracle / host-layer
code: every operator is built as a dense system  A f = B u  from

  * the PUBLISHED interior scheme constants (Lele 1992):
        first derivative    alpha = 1/3,  a = 14/9 / (2 h),  b = 1/9 / (4 h)
        second derivative   alpha = 2/11, a = 12/11 / h^2,   b = 3/11 / (4 h^2)
        staggered deriv.    alpha = 9/62, a = 63/62 / h,     b = 17/62 / (3 h)
        midpoint interp.    alpha = 3/10, a = 3/4 (x 1/2 per point pair ... written below as weights), b = 1/20
  * ghost points eliminated by REFLECTION about the wall (even / odd extension of the field, the derivative of an even
    field being odd and vice versa) for free-slip (Neumann) boundaries,
  * Lele's one-sided closures for Dirichlet boundaries,
  * periodic wrap-around otherwise,

in exact rational arithmetic (fractions.Fraction), then solved in 80-bit long double. The oracle (a line-by-line
restatement of the Fortran) has to reproduce these solutions to 1e-13: a transcription slip in any coefficient row of
src/tdsops.f90 that both the oracle and the host layer might share shows up here.
"""
from fractions import Fraction as F

import numpy as np

PERIODIC, NEUMANN, DIRICHLET = 0, 1, 2

# interior schemes: alpha and the right-hand-side weights as {offset: weight}; offsets are in units of h measured
# from the OUTPUT point; staggered operators have half-integer offsets. Weights carry 1/h^p separately (`order`).
SCHEMES = {
    # f'_i:  a (u_{i+1} - u_{i-1}) / (2h) + b (u_{i+2} - u_{i-2}) / (4h),  a = 14/9, b = 1/9
    "der1st": dict(alpha=F(1, 3), order=1, stag=False,
                   w={F(1): F(14, 9) / 2, F(-1): -F(14, 9) / 2, F(2): F(1, 9) / 4, F(-2): -F(1, 9) / 4}),
    # f''_i: a (u_{i+1} - 2u_i + u_{i-1}) / h^2 + b (u_{i+2} - 2u_i + u_{i-2}) / (4 h^2),  a = 12/11, b = 3/11
    "der2nd": dict(alpha=F(2, 11), order=2, stag=False,
                   w={F(1): F(12, 11), F(-1): F(12, 11), F(2): F(3, 11) / 4, F(-2): F(3, 11) / 4,
                      F(0): -2 * F(12, 11) - 2 * F(3, 11) / 4}),
    # staggered f'_{i+1/2}: a (u_{i+1} - u_i) / h + b (u_{i+2} - u_{i-1}) / (3h),  a = 63/62, b = 17/62
    "stagder": dict(alpha=F(9, 62), order=1, stag=True,
                    w={F(1, 2): F(63, 62), F(-1, 2): -F(63, 62), F(3, 2): F(17, 62) / 3, F(-3, 2): -F(17, 62) / 3}),
    # midpoint interpolation f_{i+1/2}: a (u_{i+1} + u_i) / 2 + b (u_{i+2} + u_{i-1}) / 2,  a = 3/2, b = 1/10
    "interpl": dict(alpha=F(3, 10), order=0, stag=True,
                    w={F(1, 2): F(3, 2) / 2, F(-1, 2): F(3, 2) / 2, F(3, 2): F(1, 10) / 2, F(-3, 2): F(1, 10) / 2}),
}


def positions(kind, n_vert, periodic):
    """Coordinates (in units of h) of the vertex grid and of the cell (midpoint) grid of a line with n_vert vertices."""
    nc = n_vert if periodic else n_vert - 1
    return [F(i) for i in range(n_vert)], [F(i) + F(1, 2) for i in range(nc)]


def build(op, n_vert, bc_start, bc_end, from_to=None, even=True):
    """Dense (A, B, n_out, n_in) with Fraction entries (B without the 1/h^order factor).

    op: der1st | der2nd | stagder | interpl; from_to: None | 'v2p' | 'p2v'; even: parity of the INPUT field about a
    free-slip wall (True = symmetric / cos type). Vertices sit at 0..n_vert-1 (walls at 0 and n_vert-1 when the line is
    not periodic; period n_vert otherwise)."""
    sc = SCHEMES[op]
    periodic = bc_start == PERIODIC and bc_end == PERIODIC
    verts, cells = positions(op, n_vert, periodic)
    if not sc["stag"]:
        xin, xout = verts, verts
    elif from_to == "v2p":
        xin, xout = verts, cells
    else:
        xin, xout = cells, verts
    n_in, n_out = len(xin), len(xout)
    period = F(n_vert)
    wall0, wall1 = F(0), F(n_vert - 1)
    flips = sc["order"] % 2 == 1  # an odd derivative changes the parity
    even_out = (not even) if flips else even
    idx_in = {x: j for j, x in enumerate(xin)}
    idx_out = {x: j for j, x in enumerate(xout)}

    def fold(x, is_even, idx):
        """(index, sign) of the in-domain point that represents coordinate x, or None if the value is pinned to 0."""
        sign = 1
        for _ in range(4):
            if periodic:
                x = x % period
            if x in idx:
                return idx[x], sign
            if x < wall0 and bc_start == NEUMANN:
                x, sign = 2 * wall0 - x, sign * (1 if is_even else -1)
            elif x > wall1 and bc_end == NEUMANN:
                x, sign = 2 * wall1 - x, sign * (1 if is_even else -1)
            else:
                return None
        return None

    A = [[F(0)] * n_out for _ in range(n_out)]
    B = [[F(0)] * n_in for _ in range(n_out)]
    for i, xo in enumerate(xout):
        at_wall0 = (not periodic) and xo == wall0
        at_wall1 = (not periodic) and xo == wall1
        near0 = (not periodic) and bc_start == DIRICHLET and xo - wall0 < 4 and not sc["stag"]
        near1 = (not periodic) and bc_end == DIRICHLET and wall1 - xo < 4 and not sc["stag"]
        if near0 or near1:
            # Lele's boundary closures, written for the left wall and mirrored for the right one
            k = int(xo - wall0) if near0 else int(wall1 - xo)
            s = 1 if near0 else -1  # mirror: offsets change sign; odd derivatives change the sign of the weights
            wsign = s if flips else 1
            if op == "der1st":
                if k == 0:    # f'_1 + 2 f'_2 = (-5 u_1 + 4 u_2 + u_3) / (2h)
                    lhs, rhs = {0: F(1), 1: F(2)}, {0: F(-5, 2), 1: F(2), 2: F(1, 2)}
                elif k == 1:  # 1/4 f'_1 + f'_2 + 1/4 f'_3 = 3/4 (u_3 - u_1) / h
                    lhs, rhs = {-1: F(1, 4), 0: F(1), 1: F(1, 4)}, {-1: F(-3, 4), 1: F(3, 4)}
                else:
                    lhs = rhs = None
            else:  # der2nd
                if k == 0:    # f''_1 + 11 f''_2 = (13 u_1 - 27 u_2 + 15 u_3 - u_4) / h^2
                    lhs, rhs = {0: F(1), 1: F(11)}, {0: F(13), 1: F(-27), 2: F(15), 3: F(-1)}
                elif k == 1:  # 1/10 f''_1 + f''_2 + 1/10 f''_3 = 6/5 (u_3 - 2 u_2 + u_1) / h^2
                    lhs, rhs = {-1: F(1, 10), 0: F(1), 1: F(1, 10)}, {-1: F(6, 5), 0: F(-12, 5), 1: F(6, 5)}
                else:         # the sixth-order interior scheme (alpha = 2/11) on rows 3 and 4 whatever the interior is
                    lhs = rhs = None
            if lhs is not None:
                for off, v in lhs.items():
                    A[i][i + s * off] += v
                for off, v in rhs.items():
                    B[i][idx_in[xo + s * off]] += wsign * v
                continue
        if (at_wall0 and bc_start == NEUMANN or at_wall1 and bc_end == NEUMANN) and not even_out:
            A[i][i] = F(1)  # an odd quantity vanishes on the wall: f = 0
            continue
        # left-hand side alpha f_{i-1} + f_i + alpha f_{i+1}
        A[i][i] += F(1)
        for d in (-1, 1):
            r = fold(xo + d, even_out, idx_out)
            if r is not None:
                A[i][r[0]] += r[1] * sc["alpha"]
        # right-hand side
        for off, wgt in sc["w"].items():
            r = fold(xo + off, even, idx_in)
            if r is None:
                continue
            B[i][r[0]] += r[1] * wgt
            # the staggered derivative of an odd field takes the odd extension about the WALL VALUE (the field need
            # not vanish there: u(-x) = 2 u(0) - u(x)), i.e. a linear extrapolation through the wall
            if op == "stagder" and from_to == "v2p" and not even and r[1] < 0:
                xw = wall0 if xo + off < wall0 else wall1
                B[i][idx_in[xw]] += 2 * wgt
    return A, B, n_out, n_in


def solve_longdouble(A, B, u, h, order, want_cond=False):
    """f = A^{-1} (B u) / h^order in 80-bit long double (Gaussian elimination with partial pivoting); u: [lines, n_in].
    want_cond: also return max(|A^{-1}| |B| |u| / h^order) / max|f|, the componentwise condition number of the
    evaluation: no double-precision evaluation order can be expected to do better than ~ eps * cond."""
    LD = np.longdouble
    n = len(A)
    Am = np.array([[LD(x.numerator) / LD(x.denominator) for x in row] for row in A], dtype=LD)
    Bm = np.array([[LD(x.numerator) / LD(x.denominator) for x in row] for row in B], dtype=LD)
    rhs = (Bm @ u.astype(LD).T) / LD(h) ** order  # [n, lines]
    if want_cond:
        absrhs = (np.abs(Bm) @ np.abs(u.astype(LD)).T) / LD(h) ** order
        rhs = np.concatenate([rhs, np.eye(n, dtype=LD)], axis=1)
    M = np.concatenate([Am, rhs], axis=1)
    for k in range(n):
        p = k + int(np.argmax(np.abs(M[k:, k])))
        if p != k:
            M[[k, p]] = M[[p, k]]
        M[k] = M[k] / M[k, k]
        rows = np.nonzero(M[:, k])[0]
        rows = rows[rows != k]
        M[rows] -= np.outer(M[rows, k], M[k])
    if want_cond:
        nl = u.shape[0]
        f, Ainv = M[:, n:n + nl], M[:, n + nl:]
        cond = float((np.abs(Ainv) @ absrhs).max() / np.abs(f).max())
        return f.T, cond
    return M[:, n:].T  # [lines, n]
