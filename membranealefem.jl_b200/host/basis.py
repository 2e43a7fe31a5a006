"""Gauss points (src/input/GaussPoint.jl) and Gauss-point basis tables (src/input/GpBasisFn.jl) of the reference.

The 1-D tables `LineGpBasisFns` are what crosses the C ABI; the 2-D tables `GpBasisFnsζα` (one product of two
1-D entries per value) are formed on the device and, here, on demand for tests and the reference-API accessors.
"""
import math

import numpy as np

from .enums import BOTTOM, CLAMPED, CLOSED, NDERS, POLY, TOP
from .spline import get_bspline_ders, get_unique_1d_elements


class GaussPointsXi:
    """GaussPointsξ (GaussPoint.jl:18-29,106-171)."""

    def __init__(self, ngp):
        assert ngp <= 4, "have at most 4 1-D Gauss points"
        self.ngp = ngp
        if ngp == 1:
            self.xs, self.ws = np.array([-1.0]), np.array([0.0])
        elif ngp == 3:
            self.xs = np.array([-math.sqrt(3 / 5), 0.0, math.sqrt(3 / 5)])
            self.ws = np.array([5 / 9, 8 / 9, 5 / 9])
        elif ngp == 4:
            a = math.sqrt(3 / 7 + math.sqrt(6 / 5) * 2 / 7)
            b = math.sqrt(3 / 7 - math.sqrt(6 / 5) * 2 / 7)
            self.xs = np.array([-a, -b, b, a])
            self.ws = np.array([(18 - math.sqrt(30)) / 36, (18 + math.sqrt(30)) / 36, (18 + math.sqrt(30)) / 36,
                                (18 - math.sqrt(30)) / 36])
        else:
            raise AssertionError(f"ξs for {ngp} Gauss points not implemented")


class GaussPointsZeta:
    """GaussPointsζ (GaussPoint.jl:67-78): points/weights mapped to [lo, hi]."""

    def __init__(self, gps_xi, lo, hi):
        self.ngp = gps_xi.ngp
        self.zs = gps_xi.xs * (hi - lo) / 2 + (hi + lo) / 2
        self.ws = gps_xi.ws * (hi - lo) / 2


def gp_basis_fns_1d(w, zeta, kv):
    """GpBasisFnsζ (GpBasisFn.jl:20-55) as a length-10 row (w, N[3], dN[3], ddN[3])."""
    assert kv.poly == POLY
    d = get_bspline_ders(kv, float(zeta), NDERS)
    return np.concatenate(([w], d[:, 0], d[:, 1], d[:, 2]))


def gp_basis_fns_2d(f1, f2):
    """GpBasisFnsζα (GpBasisFn.jl:96-112): dict with w, N[9], dN[9,2], ddN[9,3] (columns 11, 22, 12)."""
    N1, d1, dd1 = f1[1:4], f1[4:7], f1[7:10]
    N2, d2, dd2 = f2[1:4], f2[4:7], f2[7:10]

    def outer(u, v):                     # index id1 + 3*(id2-1)
        return (v[:, None] * u[None, :]).reshape(-1)
    return {"w": f1[0] * f2[0], "N": outer(N1, N2),
            "dN": np.stack([outer(d1, N2), outer(N1, d2)], axis=1),
            "ddN": np.stack([outer(dd1, N2), outer(N1, dd2), outer(d1, d2)], axis=1)}


class LineGpBasisFns:
    """LineGpBasisFns (GpBasisFn.jl:143-202). `ufns` is (nuel, ngp, 10); `uel_ids` is 1-based."""

    def __init__(self, kv, ngp):
        nuel, nel, uel_ids, uel_list = get_unique_1d_elements(kv)
        gx = GaussPointsXi(ngp)
        self.nel, self.uel_ids = nel, uel_ids
        self.ufns = np.zeros((nuel, ngp, 10))
        for ue, (lo, hi) in enumerate(uel_list):
            g = GaussPointsZeta(gx, lo, hi)
            for k in range(ngp):
                self.ufns[ue, k] = gp_basis_fns_1d(g.ws[k], g.zs[k], kv)
        if kv.curve == CLAMPED:
            self.zmin_fns = gp_basis_fns_1d(1.0, kv.zs[0], kv)
            self.zmax_fns = gp_basis_fns_1d(1.0, kv.zs[-1], kv)
        elif kv.curve == CLOSED:
            self.zmin_fns = gp_basis_fns_1d(1.0, kv.zs[kv.poly], kv)
            self.zmax_fns = gp_basis_fns_1d(1.0, kv.zs[-kv.poly - 1], kv)
        else:
            raise AssertionError(f"edge basis functions for {kv.curve} curve not implemented")

    @property
    def edge(self):
        return np.stack([self.zmin_fns, self.zmax_fns])


class BdryGpBasisFns:
    """BdryGpBasisFns (GpBasisFn.jl:231-277), evaluated lazily."""

    def __init__(self, line, perp_edge_fns, bdry):
        self.nel, self.uel_ids, self.bdry = line.nel, line.uel_ids, bdry
        self._line, self._perp = line, perp_edge_fns

    def ufn(self, uel, gp):              # 1-based like ufns[uel, gp]
        f = self._line.ufns[uel - 1, gp - 1]
        return gp_basis_fns_2d(f, self._perp) if self.bdry in (BOTTOM, TOP) else gp_basis_fns_2d(self._perp, f)


class AreaGpBasisFns:
    """AreaGpBasisFns (GpBasisFn.jl:300-356), evaluated lazily (uel = uel1 + (uel2-1)*nuel1, gp = gp1 + 3(gp2-1))."""

    def __init__(self, line1, line2):
        self.nel = line1.nel * line2.nel
        self._l1, self._l2 = line1, line2
        self.nuel1, self.nuel2 = line1.ufns.shape[0], line2.ufns.shape[0]

    @property
    def uel_ids(self):
        u1, u2 = self._l1.uel_ids, self._l2.uel_ids
        return (u1[None, :] + (u2[:, None] - 1) * self.nuel1).reshape(-1)

    def uel_of(self, el):
        e1, e2 = (el - 1) % self._l1.nel, (el - 1) // self._l1.nel
        return int(self._l1.uel_ids[e1] + (self._l2.uel_ids[e2] - 1) * self.nuel1)

    def ufn(self, uel, gp):
        u1, u2 = (uel - 1) % self.nuel1, (uel - 1) // self.nuel1
        g1, g2 = (gp - 1) % 3, (gp - 1) // 3
        return gp_basis_fns_2d(self._l1.ufns[u1, g1], self._l2.ufns[u2, g2])
