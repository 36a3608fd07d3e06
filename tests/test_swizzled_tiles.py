"""The index arithmetic behind the swizzled x / y line tiles (tds_m4.cu "XT", transeq_m4.cu "XTIN"; host side:
m4_host.cu::make_map_xt), stated in numpy and checked on the CPU:

* the 5-D tensor (point in segment, lane, half, block, z) with the strides make_map_xt passes addresses exactly the
  elements of a DIR_Y field along x (and of a DIR_X field along y) - allocator layouts of DESIGN.md section 3;
* a box (16, L, 2, blocks, 1) lands in shared memory as [segment][lane][16], i.e. row r = q * L + l = threadIdx.x;
* with the 128-byte swizzle (16-byte chunk c of row r at position c ^ (r & 7)) the device-side offset function
  sw_off() finds every window element, and the eight threads of a quarter warp hit eight different chunk positions for
  every access of the kernels (own row, previous segment's row, next segment's row), for all tile shapes in use.
"""
import numpy as np
import pytest

SZ, S = 32, 16


def idx_y(x, y, z, nxp, nyp):
    return (x % SZ) + SZ * (y + nyp * ((x // SZ) + (nxp // SZ) * z))


def idx_x(x, y, z, nxp, nyp):
    return (y % SZ) + SZ * (x + nxp * ((y // SZ) + (nyp // SZ) * z))


@pytest.mark.parametrize("nxp,nyp,nz", [(64, 32, 3), (128, 64, 2), (512, 96, 2)])
def test_tensor_map_strides_address_the_layouts(nxp, nyp, nz):
    R = SZ * 8
    # x lines of a DIR_Y field: dims (k, y, h, xb, z), byte strides (8, R, 128, nyp R, nxb nyp R)
    nxb = nxp // SZ
    for (k, y, h, xb, z) in [(0, 0, 0, 0, 0), (5, 7, 1, nxb - 1, nz - 1), (15, nyp - 1, 0, 1, 1), (9, 3, 1, 0, 2 % nz)]:
        byte = 8 * k + R * y + 128 * h + nyp * R * xb + nxb * nyp * R * z
        assert byte == 8 * idx_y(SZ * xb + S * h + k, y, z, nxp, nyp)
    # y lines of a DIR_X field: dims (k, x, h, yb, z), byte strides (8, R, 128, nxp R, nyb nxp R)
    nyb = nyp // SZ
    for (k, x, h, yb, z) in [(0, 0, 0, 0, 0), (5, 7, 1, nyb - 1, nz - 1), (15, nxp - 1, 0, 0, 1), (9, 33, 1, nyb - 1, 0)]:
        byte = 8 * k + R * x + 128 * h + nxp * R * yb + nyb * nxp * R * z
        assert byte == 8 * idx_x(x, SZ * yb + S * h + k, z, nxp, nyp)


def swizzled_tile(L, nseg, line):
    """Shared-memory image (in doubles) of a box (16, L, 2, nseg / 2, 1): smem order = dim 0 fastest, 128-byte rows,
    chunk c of row r stored at chunk c ^ (r & 7). line[l][j] is point j of lane l."""
    NT = L * nseg
    tile = np.full(NT * S, np.nan)
    for q in range(nseg):
        for l in range(L):
            r = q * L + l
            for kk in range(S):
                tile[r * S + (((kk >> 1) ^ (r & 7)) << 1) + (kk & 1)] = line[l][q * S + kk]
    return tile


def sw_off(t, r, q, L, NT):  # transeq_m4.cu::sw_off / tds_m4.cu::window_sw
    nseg = NT // L
    rr = (r + NT - L if q == 0 else r - L) if t < 4 else (r if t < S + 4 else (r - (NT - L) if q == nseg - 1 else r + L))
    kk = S - 4 + t if t < 4 else (t - 4 if t < S + 4 else t - S - 4)
    return rr * S + (((kk >> 1) ^ (rr & 7)) << 1) + (kk & 1), rr, kk


@pytest.mark.parametrize("L,NT", [(32, 128), (32, 256), (16, 256), (16, 128), (8, 256), (8, 128), (8, 512), (4, 256), (4, 128)])
def test_window_offsets_and_bank_conflicts(L, NT):
    nseg = NT // L
    n = nseg * S
    rng = np.random.default_rng(L * 1000 + NT)
    line = rng.standard_normal((L, n))
    tile = swizzled_tile(L, nseg, line)
    assert not np.isnan(tile).any()  # the swizzle is a permutation of every row
    for r in range(NT):
        q, l = r // L, r % L
        for t in range(S + 8):  # periodic window: points j0 - 4 .. j0 + 19
            off, _, _ = sw_off(t, r, q, L, NT)
            assert tile[off] == line[l][(q * S - 4 + t) % n]
    # 128-bit accesses: the 8 threads of a quarter warp must touch 8 different 16-byte positions within their rows
    for r0 in range(0, NT, 8):
        for t in range(0, S + 8, 2):
            pos = set()
            for r in range(r0, r0 + 8):
                off, _, _ = sw_off(t, r, r // L, L, NT)
                pos.add((off % S) >> 1)
            assert len(pos) == 8, (L, NT, r0, t)
