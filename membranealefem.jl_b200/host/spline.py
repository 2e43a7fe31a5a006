"""B-spline machinery of the reference's input layer (src/input/Spline.jl), restated for the host side.

Algorithms A2.1-A2.3 of Piegl & Tiller in their 0-based textbook form; the floating-point operations and their
order are those of the reference, so the tables come out bit-identical to `LineGpBasisFns` (tests compare them
with the oracle's literal restatement).
"""
import math

import numpy as np

from .enums import CLAMPED, CLOSED, Curve

_RTOL = math.sqrt(np.finfo(float).eps)


def _isapprox(x, y):
    """Julia isapprox for Float64 scalars: rtol = sqrt(eps), atol = 0."""
    if x == y:
        return True
    if not (math.isfinite(x) and math.isfinite(y)):
        return False
    return abs(x - y) <= _RTOL * max(abs(x), abs(y))


class KnotVector:
    """KnotVector (Spline.jl:20-96): `KnotVector(zs, poly, curve)` or `KnotVector(nel, poly, curve)`."""

    def __init__(self, zs_or_nel, poly, curve=CLAMPED):
        curve = Curve(curve)
        if isinstance(zs_or_nel, (int, np.integer)):
            nel = int(zs_or_nel)
            num = nel + 2 * poly + 1
            zs = [0.0] * num
            if curve == CLAMPED:                      # Spline.jl:80-85
                for idx in range(poly + 1, num - poly - 1):
                    zs[idx] = (idx - poly) / nel
                for idx in range(num - poly - 1, num):
                    zs[idx] = 1.0
            else:                                     # Spline.jl:86-91
                for idx in range(num):
                    zs[idx] = (idx - poly) / nel
        else:
            zs = [float(z) for z in zs_or_nel]
        n = len(zs)
        if curve == CLAMPED:                          # Spline.jl:48-51
            assert all(zs[i] <= zs[i + 1] for i in range(n - 1)), "knots must be non-decreasing"
            assert n >= 2 * poly + 2 and all(z == zs[0] for z in zs[:poly + 1]), "need poly+1 repeated knots at start"
            assert all(z == zs[-1] for z in zs[n - poly - 1:]), "need poly+1 repeated knots at end"
        elif curve == CLOSED:                         # Spline.jl:52-53
            d = np.diff(np.asarray(zs))
            ref = (zs[1] - zs[0]) * np.ones(n - 1)
            assert np.linalg.norm(d - ref) <= _RTOL * max(np.linalg.norm(d), np.linalg.norm(ref))
        else:
            raise AssertionError(f"knot vector for {curve} curve not implemented")
        self.zs = np.asarray(zs, dtype=np.float64)
        self.nel = n - 2 * poly - 1
        self.poly = poly
        self.curve = curve

    def __eq__(self, other):                          # Spline.jl:101-104
        return (np.array_equal(self.zs, other.zs) and self.poly == other.poly and self.curve == other.curve)

    def __hash__(self):
        return hash((self.zs.tobytes(), self.poly, int(self.curve)))


def get_fine_zs(nel, poly):
    """Knots concentrated in the centre (Spline.jl:122-183); returned list is 0-based."""
    assert nel >= 18, "fine mesh requires at least 18 1-D elements"
    num = nel + 2 * poly + 1
    z = [0.0] * (num + 1)            # 1-based scratch, like the reference
    nw1 = 6
    nf1 = nel - 2 * nw1 - 1
    zw1 = 1 / 3
    zf1 = 1.0 - 2 * zw1
    for idx in range(poly + 2, poly + nw1 + 2):
        dz = zw1 / nw1
        z[idx] = (idx - poly - 1) * dz
    for idx in range(num - poly - nw1, num - poly):
        dz = zw1 / nw1
        z[idx] = 1.0 - zw1 + (idx - num + poly + nw1) * dz
    nw2 = nf1 // 4
    nf2 = nf1 - 2 * nw2
    zw2 = 1 / 9
    zf2 = zf1 - 2 * zw2
    if 2 * nw2 <= nw1:
        dz = zf1 / (nf1 + 1)
        for idx in range(poly + nw1 + 2, num - poly - nw1):
            z[idx] = zw1 + (idx - poly - nw1 - 1) * dz
    else:
        dz = zw2 / nw2
        for idx in range(nw1 + poly + 2, nw1 + poly + 2 + nw2):
            z[idx] = zw1 + (idx - nw1 - poly - 1) * dz
        for idx in range(num - poly - nw1 - nw2, num - poly - nw1):
            z[idx] = 1.0 - zw1 - zw2 + (idx - num + poly + nw1 + nw2) * dz
        dz = zf2 / (nf2 + 1)
        for idx in range(nw1 + poly + 2 + nw2, num - poly - nw1 - nw2):
            z[idx] = zw1 + zw2 + (idx - nw1 - poly - 1 - nw2) * dz
    for idx in range(num - poly, num + 1):
        z[idx] = 1.0
    return z[1:]


def get_knot_span_index(kv, zeta):
    """A2.1 extended to CLOSED curves (Spline.jl:198-248). Returns the reference's 1-based span index."""
    zs = kv.zs.copy()
    nk = len(zs)
    if kv.curve == CLOSED:
        zs[:kv.poly] = zs[kv.poly]
        zs[nk - kv.poly:] = zs[nk - kv.poly - 1]
    assert zeta >= zs[0], "ζ smaller than first active knot"
    assert zeta <= zs[-1], "ζ larger than last active knot"
    def z1(i):                         # the reference's 1-based view of the knot list
        return zs[i - 1]
    m = 1
    while z1(m) == z1(1):              # index of the second-smallest knot
        m += 1
    n = nk
    while z1(n) == z1(nk):             # index of the second-largest knot
        n -= 1
    if zeta == z1(n + 1):
        return n
    low, high = m - 1, n + 1
    mid = (low + high) // 2
    while zeta < z1(mid) or zeta >= z1(mid + 1):
        if zeta < z1(mid):
            high = mid
        else:
            low = mid
        mid = (low + high) // 2
    return mid


def get_bspline_vals(kv, zeta):
    """A2.2 (Spline.jl:264-296)."""
    assert zeta >= kv.zs[0], "ζ is less than smallest knot"
    assert zeta <= kv.zs[-1], "ζ is greater than largest knot"
    p, zs = kv.poly, kv.zs
    i = get_knot_span_index(kv, zeta) - 1
    left, right, N = [0.0] * (p + 1), [0.0] * (p + 1), [0.0] * (p + 1)
    N[0] = 1.0
    for j in range(1, p + 1):
        left[j] = zeta - zs[i + 1 - j]
        right[j] = zs[i + j] - zeta
        saved = 0.0
        for r in range(j):
            temp = N[r] / (right[r + 1] + left[j - r])
            N[r] = saved + right[r + 1] * temp
            saved = left[j - r] * temp
        N[j] = saved
    return np.asarray(N)


def get_bspline_ders(kv, zeta, num_ders):
    """A2.3 (Spline.jl:319-422): (poly+1) x (num_ders+1), column k = k-th derivative."""
    assert zeta >= kv.zs[0], "ζ is less than smallest knot"
    assert zeta <= kv.zs[-1], "ζ is greater than largest knot"
    p, zs = kv.poly, kv.zs
    assert num_ders >= 0, "cannot have fewer than zero derivatives"
    assert num_ders <= p, "basis functions have only `poly` derivatives"
    i = get_knot_span_index(kv, zeta) - 1
    left, right = [0.0] * (p + 1), [0.0] * (p + 1)
    ndu = [[0.0] * (p + 1) for _ in range(p + 1)]
    ders = [[0.0] * (num_ders + 1) for _ in range(p + 1)]
    ndu[0][0] = 1.0
    for j in range(1, p + 1):
        left[j] = zeta - zs[i + 1 - j]
        right[j] = zs[i + j] - zeta
        saved = 0.0
        for r in range(j):
            ndu[j][r] = right[r + 1] + left[j - r]
            temp = ndu[r][j - 1] / ndu[j][r]
            ndu[r][j] = saved + right[r + 1] * temp
            saved = left[j - r] * temp
        ndu[j][j] = saved
    for j in range(p + 1):
        ders[j][0] = ndu[j][p]
    a = [[0.0] * (p + 1) for _ in range(2)]
    for r in range(p + 1):
        s1, s2 = 0, 1
        a[0][0] = 1.0
        for k in range(1, num_ders + 1):
            d = 0.0
            rk, pk = r - k, p - k
            if r >= k:
                a[s2][0] = a[s1][0] / ndu[pk + 1][rk]
                d = a[s2][0] * ndu[rk][pk]
            j1 = 1 if rk >= -1 else -rk
            j2 = k - 1 if r - 1 <= pk else p - r
            for j in range(j1, j2 + 1):
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j]
                d += a[s2][j] * ndu[rk + j][pk]
            if r <= pk:
                a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r]
                d += a[s2][k] * ndu[r][pk]
            ders[r][k] = d
            s1, s2 = s2, s1
    r = p
    for k in range(1, num_ders + 1):
        for j in range(p + 1):
            ders[j][k] *= r
        r *= (p - k)
    return np.asarray(ders)


def get_bspline_indices(kv, zeta_or_span):
    """Global (1-based) indices of the non-zero B-splines (Spline.jl:441-481). A float is a ζ, an int a span id."""
    ks = zeta_or_span if isinstance(zeta_or_span, (int, np.integer)) else get_knot_span_index(kv, zeta_or_span)
    ids = np.arange(1, kv.poly + 2) + (ks - kv.poly - 1)
    if kv.curve == CLOSED:
        ids = (ids - 1) % kv.nel + 1
    return ids


def collocate_zeta(kv):
    """Collocation points (Spline.jl:580-626)."""
    p = kv.poly
    if kv.curve == CLOSED:
        return (kv.zs[p:-p - 1] + kv.zs[p + 1:len(kv.zs) - p]) / 2
    u = []
    for z in kv.zs:
        if z not in u:
            u.append(z)
    assert p > 1, "interpolation for poly ≥ 2 only when clamped"
    assert p < 4, "interpolation for poly ≤ 3 only when clamped"
    assert len(u) == len(kv.zs) - 2 * p, "no repeated interior knots"
    nb = len(kv.zs) - p - 1
    zl = np.zeros(nb)
    zl[0], zl[-1] = u[0], u[-1]
    if p == 2:
        for i in range(1, nb - 1):
            zl[i] = (u[i - 1] + u[i]) / 2
    else:
        zl[1] = (u[0] + u[1]) / 2
        zl[-2] = (u[-2] + u[-1]) / 2
        for i in range(2, nb - 2):
            zl[i] = u[i - 1]
    return zl


def get_1d_bspline_cps(kv, x):
    """Control points reproducing x(ζ) at the collocation points (Spline.jl:510-528)."""
    zl = collocate_zeta(kv)
    nb = len(zl)
    mat = np.zeros((nb, nb))
    for j in range(nb):
        mat[j, get_bspline_indices(kv, float(zl[j])) - 1] = get_bspline_vals(kv, float(zl[j]))
    return np.linalg.solve(mat, np.array([x(float(z)) for z in zl]))


def get_2d_bspline_cps(kv1, kv2, x, separable=None):
    """Control points of a scalar function on the tensor-product patch (Spline.jl:540-567).

    The reference solves one dense numnp x numnp system. The collocation matrix is the Kronecker product of the two
    1-D collocation matrices, so the same control points follow from two families of 1-D solves -- O(numnp) memory,
    which is what makes 10^6-element patches possible. Mathematically identical; rounding differs at 1e-16.
    """
    z1, z2 = collocate_zeta(kv1), collocate_zeta(kv2)
    n1, n2 = len(z1), len(z2)
    m1, m2 = np.zeros((n1, n1)), np.zeros((n2, n2))
    for j in range(n1):
        m1[j, get_bspline_indices(kv1, float(z1[j])) - 1] = get_bspline_vals(kv1, float(z1[j]))
    for k in range(n2):
        m2[k, get_bspline_indices(kv2, float(z2[k])) - 1] = get_bspline_vals(kv2, float(z2[k]))
    xv = x(z1[None, :], z2[:, None]) if separable is None else separable
    xv = np.broadcast_to(np.asarray(xv, dtype=np.float64), (n2, n1))
    c = np.linalg.solve(m1, xv.T)            # solve along direction 1: (n1, n2)
    c = np.linalg.solve(m2, c.T)             # then along direction 2: (n2, n1)
    return c.reshape(-1)                     # index i1 + (i2-1)*num1


def get_unique_1d_elements(kv):
    """Unique 1-D elements (Spline.jl:640-671): (uel_num, num_el, uel_ids[1-based], uel_list)."""
    p, zs = kv.poly, kv.zs
    nk = len(zs)
    prev = [0.0] * (2 * p + 1)
    num_el = nk - 2 * p - 1
    uel_ids = np.zeros(num_el, dtype=np.int64)
    uel_num = 0
    uel_list = []
    for k in range(p, nk - p - 1):           # 0-based knot index of the span start
        ctx = [float(zs[k - p + 1 + q] - zs[k - p + q]) for q in range(2 * p + 1)]
        if not all(_isapprox(c, q) for c, q in zip(ctx, prev)):
            uel_num += 1
            uel_list.append((float(zs[k]), float(zs[k + 1])))
        uel_ids[k - p] = uel_num
        prev = ctx
    return uel_num, num_el, uel_ids, uel_list
