"""CPU check of the likelihood kernels' cell-index arithmetic (likelihood.cu) against the reference's.

The kernels never convert a coordinate: the particle position is taken relative to the window origin and biased
once by M = 1.5 * 2^12; the two FMAs of a coordinate then produce M + t, and the HIGH WORD of that double is
K + floor(t * 2^8).  Here that chain is emulated exactly (an FMA is one correctly rounded evaluation of a*b + c:
fractions.Fraction -> float) and compared with the reference's  int((x + r cos(theta + a) - ox) / res)
(pu:126-129, all float64) on random particles and beams of the reference's maps."""
import struct
from fractions import Fraction

import numpy as np

E = 12
M = 1.5 * 2.0 ** E
K = ((1023 + E) << 20) + (1 << 19)


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))      # exact product and sum, one rounding


def hi_word(d):
    return struct.unpack("<q", struct.pack("<d", d))[0] >> 32


def test_high_word_of_biased_fma_chain_is_the_reference_cell():
    rs = np.random.RandomState(7)
    res, ox, oy = 0.05, -10.0, -10.0
    wofx, wofy = 142, 133                  # map_world window origin (free-space box - 1)
    n = 60000
    x = rs.uniform(-2.9, 2.9, n)
    y = rs.uniform(-3.3, 1.7, n)
    th = rs.uniform(-np.pi, np.pi, n)
    r = rs.uniform(0.12, 3.5, n)
    a = np.float32(rs.randint(0, 360, n) * (2 * np.pi / 360)).astype(np.float64)     # float32 beam angles (node:346)
    # reference (pu:126-129)
    ref_x = ((x + r * np.cos(th + a) - ox) / res).astype(np.int64)
    ref_y = ((y + r * np.sin(th + a) - oy) / res).astype(np.int64)
    # kernel arithmetic
    bx, by = r * np.cos(a) / res, r * np.sin(a) / res         # per-scan beam table (host)
    c, s = np.cos(th), np.sin(th)
    px, py = (x - ox) / res, (y - oy) / res
    PX, PY = (px - wofx) + M, (py - wofy) + M
    mism, closest = 0, 1.0
    for i in range(n):
        TX = fma(c[i], bx[i], fma(-s[i], by[i], PX[i]))
        TY = fma(s[i], bx[i], fma(c[i], by[i], PY[i]))
        ix = ((hi_word(TX) - K) >> 8) + wofx
        iy = ((hi_word(TY) - K) >> 8) + wofy
        mism += int(ix != ref_x[i]) + int(iy != ref_y[i])
        fx = (TX - M) % 1.0
        closest = min(closest, fx, 1.0 - fx)
    # 120 000 coordinates: the nearest one to a cell boundary is ~1e-5 cell away, the arithmetic differs from the
    # reference's by ~1e-12 cell -> no flip (fp32 arithmetic would flip ~7 of them, SURVEY 7 hard part 1)
    assert mism == 0, mism
    assert closest > 1e-9


def test_biased_coordinate_keeps_40_fraction_bits():
    """M + t is rounded to 2^-40 cell for |t| < 2^11, and the high word is monotone in t across the sign."""
    ts = [-1000.25, -1.0, -2.0 ** -40, 0.0, 2.0 ** -40, 0.999999999999, 1.0, 255.5, 2047.0]
    prev = None
    for t in ts:
        v = M + t
        assert abs((v - M) - t) <= 2.0 ** -41
        f = hi_word(v) - K
        assert f == int(np.floor(t * 256.0)), (t, f)
        assert prev is None or f >= prev
        prev = f
