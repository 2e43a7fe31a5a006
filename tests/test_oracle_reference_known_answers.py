"""Pins the CPU oracle against every known answer the reference's own test/ directory holds for the
hot path's inputs and Gauss-point kernel:

  /root/reference/test/input/spline_tests.jl
  /root/reference/test/input/gauss_pt_tests.jl
  /root/reference/test/input/gp_basis_fn_tests.jl
  /root/reference/test/analysis/geo_dyn_stress_tests.jl

Each test cites the Julia lines it restates. Nothing here reads /root/reference at run time.
"""
import numpy as np
import pytest

from oracle import oracle as orc

EPS = np.finfo(float).eps


def eps(x):
    return np.spacing(x)


# ---------------------------------------------------------------- spline_tests.jl:3-46
def test_knot_span():
    kv0 = orc.KnotVector.from_list([0., 0., 0., 0.3, 0.5, 0.5, 0.6, 1., 1., 1.], 0)
    assert kv0.span(0.0) == 3 and kv0.span(0.5) == 6 and kv0.span(1.0) == 7
    kv1 = orc.KnotVector.from_list([0., 0., 0., 1., 2., 3., 4., 4., 4.], 2)
    for z, s in [(0.0, 3), (0.7, 3), (1.0, 4), (1.3, 4), (2.0, 5), (2.1, 5), (3.0, 6), (3.9, 6), (4.0, 6)]:
        assert kv1.span(z) == s
    with pytest.raises(AssertionError):
        kv1.span(-0.1)
    with pytest.raises(AssertionError):
        kv1.span(4.1)
    for bad, poly, curve in [([2., 2., 2., 3., 4., 4., 4.], 3, orc.CLAMPED), ([2., 2., 2., 3., 4., 4.], 2, orc.CLAMPED),
                             ([2., 2., 2., 3., 2.5, 4., 4., 4.], 2, orc.CLAMPED),
                             ([2., 2., 2., 3., 4., 4., 4.], 2, orc.CLOSED)]:
        with pytest.raises(AssertionError):
            orc.KnotVector.from_list(bad, poly, curve)
    kv2 = orc.KnotVector.uniform(4, 3, orc.CLOSED)
    for z in (-0.75, 1.1, 1.75):
        with pytest.raises(AssertionError):
            kv2.span(z)
    for z, s in [(0.0, 4), (0.2, 4), (0.4, 5), (0.5, 6), (0.75, 7), (1.0, 7)]:
        assert kv2.span(z) == s


# ---------------------------------------------------------------- spline_tests.jl:48-71
def test_bspline_functions():
    kv1 = orc.KnotVector.from_list([0., 0., 0., 1., 2., 3., 4., 4., 5., 5., 5.], 2)
    assert np.linalg.norm(kv1.vals(2.5, 2) - [1 / 8, 6 / 8, 1 / 8]) < eps(8.)
    kv2 = orc.KnotVector.from_list([0., 0., 0., 1 / 3, 2 / 3, 1., 1., 1.], 2)
    assert np.linalg.norm(kv2.vals(0.0, 2) - [1., 0., 0.]) < eps(8.)
    assert np.linalg.norm(kv2.vals(1 / 3 - EPS, 2) - [0., .5, .5]) < eps(8.)
    assert np.linalg.norm(kv2.vals(1 / 3, 2) - [.5, .5, 0.]) < eps(8.)
    assert np.array_equal(kv2.zs, orc.KnotVector.uniform(3, 2).zs)  # kv2 == KnotVector(3, 2, CLAMPED)
    zs3 = [0., 0., 0., 0.3, 0.5, 0.5, 0.6, 1., 1., 1.]
    kv3 = orc.KnotVector.from_list(zs3, 1)
    for z, v in [(0.00, [1., 0.]), (0.15, [.5, .5]), (0.35, [.75, .25]), (0.53, [.7, .3]), (0.60, [1., 0.]),
                 (0.90, [.25, .75])]:
        assert np.linalg.norm(kv3.vals(z, 1) - v) < eps(8.)


# ---------------------------------------------------------------- spline_tests.jl:73-114
def test_bspline_derivatives():
    kv1 = orc.KnotVector.from_list([0., 0., 0., 1., 2., 3., 4., 4., 5., 5., 5.], 2)
    assert np.linalg.norm(kv1.ders(2.5, 2, 2)[:, 0] - [1 / 8, 6 / 8, 1 / 8]) < eps(8.)
    kv2 = orc.KnotVector.from_list([0., 0., 0., 1 / 3, 2 / 3, 1., 1., 1.], 2)
    assert np.linalg.norm(kv2.ders(0.0, 0, 2)[:, 0] - [1., 0., 0.]) < eps(8.)
    assert np.linalg.norm(kv2.ders(1 / 3 - EPS, 0, 2)[:, 0] - [0., .5, .5]) < eps(8.)
    assert np.linalg.norm(kv2.ders(1 / 3, 1, 2)[:, 0] - [.5, .5, 0.]) < eps(8.)
    assert np.linalg.norm(kv2.ders(0.5, 1, 2) - np.array([[0.125, -1.5], [0.75, 0], [0.125, 1.5]])) < eps(8.)
    zs3 = [0., 0., 0., 0.3, 0.5, 0.5, 0.6, 1., 1., 1.]
    kv3 = orc.KnotVector.from_list(zs3, 1)
    for z, d, tol in [(0.00, [-10 / 3, 10 / 3], 8.), (0.15, [-10 / 3, 10 / 3], 8.), (0.35, [-5, 5], 8.),
                      (0.53, [-10, 10], 16.), (0.60, [-2.5, 2.5], 8.), (0.90, [-2.5, 2.5], 8.)]:
        assert np.linalg.norm(kv3.ders(z, 1, 1)[:, 1] - d) < eps(tol)
    kv4 = orc.KnotVector.from_list(zs3, 2)
    for z, d, tol in [(0.00, [200 / 9, -320 / 9, 40 / 3], 64.), (0.15, [200 / 9, -320 / 9, 40 / 3], 64.),
                      (0.35, [20, -70, 50], 64.), (0.53, [200, -240, 40], 2e3), (0.60, [10, -22.5, 12.5], 128.),
                      (0.90, [10, -22.5, 12.5], 128.), (1.00, [10, -22.5, 12.5], 128.)]:
        assert np.linalg.norm(kv4.ders(z, 2, 2)[:, 2] - d) < eps(tol)
    for z, nd in [(0.5, -1), (0.5, 3), (-0.01, 2), (1.01, 2)]:
        with pytest.raises(AssertionError):
            kv4.ders(z, nd, 2) if nd >= 0 else kv4.ders(z, nd, 2)


# ---------------------------------------------------------------- spline_tests.jl:116-144
def test_global_indices():
    zs1 = [0., 0., 0., 0.3, 0.5, 0.5, 0.6, 1., 1., 1.]
    kv1 = orc.KnotVector.from_list(zs1, 2)
    for z, ids in [(0.00, [1, 2, 3]), (0.29, [1, 2, 3]), (0.30, [2, 3, 4]), (0.49, [2, 3, 4]), (0.50, [4, 5, 6]),
                   (0.59, [4, 5, 6]), (0.60, [5, 6, 7]), (0.69, [5, 6, 7]), (1.00, [5, 6, 7])]:
        assert kv1.indices(z, 2).tolist() == ids
    kv2 = orc.KnotVector.from_list(zs1, 1)
    for z, ids in [(0.00, [2, 3]), (0.29, [2, 3]), (0.30, [3, 4]), (0.49, [3, 4]), (0.50, [5, 6]), (0.59, [5, 6]),
                   (0.60, [6, 7]), (0.69, [6, 7]), (1.00, [6, 7])]:
        assert kv2.indices(z, 1).tolist() == ids


# ---------------------------------------------------------------- spline_tests.jl:146-162
def test_control_point_calculation():
    kvs = [(orc.KnotVector.from_list([0., 0., 0., 1 / 3, 2 / 3, 1., 1., 1.], 2), 2),
           (orc.KnotVector.from_list([0., 0., 0., 0., 1 / 3, 2 / 3, 1., 1., 1., 1.], 3), 3)]
    for kv, poly in kvs:
        for f in (lambda x: x, lambda x: 3 * x + 1, lambda x: 1 - 4 * (x - 0.5) ** 2):
            cps = kv.cps_1d(f)
            for z in np.arange(0.0, 1.0001, 0.1):
                z = min(z, 1.0)
                val = np.dot(kv.vals(z, poly), cps[kv.indices(z, poly) - 1])
                assert np.isclose(val, f(z), rtol=np.sqrt(EPS), atol=1e-14)
    for zs, poly in [([1., 1., 2., 3., 3.], 1), ([1., 1., 1., 1., 1., 2., 3., 3., 3., 3., 3.], 4),
                     ([1., 1., 1., 2., 2., 3., 3., 3.], 2)]:
        with pytest.raises(AssertionError):
            orc.KnotVector.from_list(zs, poly).cps_1d(lambda x: x)


# ---------------------------------------------------------------- spline_tests.jl:164-220
def test_unique_1d_elements():
    n, nel, ids, lst = orc.KnotVector.uniform(2, 2).unique_1d()
    assert (n, nel, ids.tolist(), lst) == (2, 2, [1, 2], [(0.0, 1 / 2), (1 / 2, 1.0)])
    n, nel, ids, lst = orc.KnotVector.uniform(5, 2).unique_1d()
    assert (n, nel, ids.tolist()) == (5, 5, [1, 2, 3, 4, 5])
    assert lst == [(0.0, 1 / 5), (1 / 5, 2 / 5), (2 / 5, 3 / 5), (3 / 5, 4 / 5), (4 / 5, 1.0)]
    n, nel, ids, lst = orc.KnotVector.uniform(8, 2).unique_1d()
    assert (n, nel, ids.tolist()) == (5, 8, [1, 2, 3, 3, 3, 3, 4, 5])
    assert lst == [(0.0, 1 / 8), (1 / 8, 2 / 8), (2 / 8, 3 / 8), (6 / 8, 7 / 8), (7 / 8, 1.0)]
    z4 = [0., 0., 0., 1., 2., 3., 4., 5., 6., 7., 8., 8.1, 8.2, 8.3, 8.4, 8.5, 8.6, 8.7, 8.8, 8.9, 9., 10., 11., 12.,
          13., 14., 15., 16., 17., 17., 17.]
    n, nel, ids, lst = orc.KnotVector.from_list(z4, 2).unique_1d()
    assert (n, nel) == (15, 26)
    assert ids.tolist() == [1, 2, 3, 3, 3, 3, 4, 5, 6, 7, 8, 8, 8, 8, 8, 8, 9, 10, 11, 12, 13, 13, 13, 13, 14, 15]
    assert lst == [(0., 1.), (1., 2.), (2., 3.), (6., 7.), (7., 8.), (8., 8.1), (8.1, 8.2), (8.2, 8.3), (8.8, 8.9),
                   (8.9, 9.), (9., 10.), (10., 11.), (11., 12.), (15., 16.), (16., 17.)]
    n, nel, ids, lst = orc.KnotVector.uniform(3, 3).unique_1d()
    assert (n, nel, ids.tolist(), lst) == (3, 3, [1, 2, 3], [(0.0, 1 / 3), (1 / 3, 2 / 3), (2 / 3, 1.0)])
    n, nel, ids, lst = orc.KnotVector.uniform(6, 3).unique_1d()
    assert (n, nel, ids.tolist()) == (6, 6, [1, 2, 3, 4, 5, 6])
    assert lst == [(0.0, 1 / 6), (1 / 6, 2 / 6), (2 / 6, 3 / 6), (3 / 6, 4 / 6), (4 / 6, 5 / 6), (5 / 6, 1.0)]
    n, nel, ids, lst = orc.KnotVector.uniform(9, 3).unique_1d()
    assert (n, nel, ids.tolist()) == (7, 9, [1, 2, 3, 4, 4, 4, 5, 6, 7])
    assert lst == [(0.0, 1 / 9), (1 / 9, 2 / 9), (2 / 9, 3 / 9), (3 / 9, 4 / 9), (6 / 9, 7 / 9), (7 / 9, 8 / 9),
                   (8 / 9, 1.0)]
    z9 = [0., 0., 0., 0., 1., 2., 3., 4., 5., 6., 7., 8., 8.1, 8.2, 8.3, 8.4, 8.5, 8.6, 8.7, 8.8, 8.9, 9., 10., 11.,
          12., 13., 14., 15., 16., 17., 17., 17., 17.]
    n, nel, ids, lst = orc.KnotVector.from_list(z9, 3).unique_1d()
    assert (n, nel) == (21, 26)
    assert ids.tolist() == [1, 2, 3, 4, 4, 5, 6, 7, 8, 9, 10, 11, 11, 11, 11, 12, 13, 14, 15, 16, 17, 18, 18, 19, 20,
                            21]
    assert lst == [(0., 1.), (1., 2.), (2., 3.), (3., 4.), (5., 6.), (6., 7.), (7., 8.), (8., 8.1), (8.1, 8.2),
                   (8.2, 8.3), (8.3, 8.4), (8.7, 8.8), (8.8, 8.9), (8.9, 9.), (9., 10.), (10., 11.), (11., 12.),
                   (12., 13.), (14., 15.), (15., 16.), (16., 17.)]
    n, nel, ids, lst = orc.KnotVector.uniform(5, 2, orc.CLOSED).unique_1d()
    assert (n, nel, ids.tolist()) == (1, 5, [1, 1, 1, 1, 1])
    assert lst == [(0, 0.2)]


# ---------------------------------------------------------------- spline_tests.jl:222-225
def test_collocate():
    assert np.allclose(orc.KnotVector.uniform(5, 2, orc.CLOSED).collocate(), [0.1, 0.3, 0.5, 0.7, 0.9])
    assert np.allclose(orc.KnotVector.uniform(5, 3, orc.CLOSED).collocate(), [0.1, 0.3, 0.5, 0.7, 0.9])


# ---------------------------------------------------------------- gauss_pt_tests.jl:3-23
def test_gauss_points():
    xs, ws = orc.gauss_xi(3)
    assert xs.tolist() == [-np.sqrt(3 / 5), 0, np.sqrt(3 / 5)]
    assert ws.tolist() == [5 / 9, 8 / 9, 5 / 9]
    x4, w4 = orc.gauss_xi(4)
    a, b = np.sqrt(3 / 7 + 2 / 7 * np.sqrt(6 / 5)), np.sqrt(3 / 7 - 2 / 7 * np.sqrt(6 / 5))
    assert np.linalg.norm(x4 - [-a, -b, b, a]) < eps(1.)
    assert w4.tolist() == [(18 - np.sqrt(30)) / 36, (18 + np.sqrt(30)) / 36, (18 + np.sqrt(30)) / 36,
                           (18 - np.sqrt(30)) / 36]
    z3, w3 = orc.gauss_zeta(3, 0.5, 3.5)
    assert np.linalg.norm(z3 - (xs * 1.5 + 2.0)) < eps(1.) and np.linalg.norm(w3 - ws * 1.5) < eps(1.)
    z4, w4z = orc.gauss_zeta(4, 1.0, 6.0)
    assert np.linalg.norm(z4 - (x4 * 2.5 + 3.5)) < eps(1.) and np.linalg.norm(w4z - w4 * 2.5) < eps(1.)


# ---------------------------------------------------------------- gp_basis_fn_tests.jl:4-14
def test_gp_basis_fns_1d():
    kv1 = orc.KnotVector.uniform(4, 2)
    f = kv1.fn1(0.1, 0.4)
    d = kv1.ders(0.4, 2, 2)
    assert f[0] == 0.1
    assert np.array_equal(f[1:4], d[:, 0]) and np.array_equal(f[4:7], d[:, 1]) and np.array_equal(f[7:10], d[:, 2])
    for z in (-0.1, 1.1):
        with pytest.raises(AssertionError):
            kv1.fn1(0.1, z)


# ---------------------------------------------------------------- gp_basis_fn_tests.jl:17-37
def test_gp_basis_fns_2d_tensor_layout():
    kv3 = orc.KnotVector.from_list([0., 0., 0., 1., 3., 3.5, 6., 6., 6.], 2)
    kv4 = orc.KnotVector.from_list([2., 2., 2., 4., 5., 5., 5.], 2)
    rng = np.random.default_rng(0)
    for z1 in np.arange(0.0, 6.01, 1.5):
        for z2 in np.arange(2.0, 5.01, 1.5):
            w1, w2 = rng.random(), rng.random()
            f1, f2 = kv3.fn1(w1, z1), kv4.fn1(w2, z2)
            fa = orc.fn2(f1, f2)
            assert fa["w"] == f1[0] * f2[0]
            for i1 in range(3):
                for i2 in range(3):
                    a = i1 + 3 * i2
                    assert fa["N"][a] == f1[1 + i1] * f2[1 + i2]
                    assert fa["dN"][a, 0] == f1[4 + i1] * f2[1 + i2]
                    assert fa["dN"][a, 1] == f1[1 + i1] * f2[4 + i2]
                    assert fa["ddN"][a, 0] == f1[7 + i1] * f2[1 + i2]
                    assert fa["ddN"][a, 1] == f1[1 + i1] * f2[7 + i2]
                    assert fa["ddN"][a, 2] == f1[4 + i1] * f2[4 + i2]


# ---------------------------------------------------------------- gp_basis_fn_tests.jl:40-74
def test_line_basis_fns_unique_tables():
    kshort = orc.KnotVector.from_list([1., 1., 1., 2., 3., 4., 5., 6., 7., 8., 8., 8.], 2)
    klong = orc.KnotVector.from_list([1., 1., 1., 2., 3., 4., 5., 6., 7., 8., 9., 10., 11., 12., 12., 12.], 2)
    ids_s, tab_s = kshort.line_fns(3)
    ids_l, tab_l = klong.line_fns(3)
    assert kshort.nel == 7 and klong.nel == 11 and ids_s.max() == 5 and ids_l.max() == 5
    assert np.array_equal(tab_s, tab_l)
    g01, _ = orc.gauss_zeta(3, 0.0, 1.0)
    assert np.array_equal(tab_s[0, 0, 1:4], kshort.ders(g01[0] + 1, 0, 2)[:, 0])
    assert np.array_equal(tab_s[1, 1, 4:7], klong.ders(g01[1] + 2, 1, 2)[:, 1])
    assert np.array_equal(tab_s[3, 2, 1:4], klong.ders(g01[2] + 8, 0, 2)[:, 0])
    assert np.array_equal(tab_s[3, 2, 7:10], klong.ders(g01[2] + 8, 2, 2)[:, 2])
    assert np.array_equal(tab_s[4, 0, 1:4], klong.ders(g01[0] + 11, 0, 2)[:, 0])
    kv2 = orc.KnotVector.from_list([0., 0., 0., 3., 4., 4.5, 6.5, 9., 9.6, 10.1, 10.1, 10.1], 2)
    ids2, tab2 = kv2.line_fns(3)
    assert ids2.max() == 7
    assert not np.array_equal(tab2[0, 0, 1:4], tab2[6, 2, 1:4])
    assert not np.array_equal(tab2[0, 0, 1:4], tab2[6, 0, 1:4])
    g03, _ = orc.gauss_zeta(3, 0.0, 3.0)
    assert np.array_equal(tab2[0, 0, 1:4], kv2.ders(g03[0], 0, 2)[:, 0])
    g96, _ = orc.gauss_zeta(3, 9.0, 9.6)
    assert np.array_equal(tab2[5, 1, 1:4], kv2.ders(g96[1], 0, 2)[:, 0])


# ---------------------------------------------------------------- geo_dyn_stress_tests.jl:5-55
def test_geo_dyn_stress_curved_patch():
    m = orc.Mesh(motion=orc.STATIC, scenario=orc.F_CAVI, num1el=2, num2el=3)
    A, B = 2.2, 1.7
    kv1, kv2 = m.kv(1), m.kv(2)
    x_cps = kv1.cps_1d(lambda z: 4 * z + 2)
    y_cps = kv2.cps_1d(lambda z: 2 * z - 3)
    z_cps = orc.cps_2d(kv1, kv2, lambda z1, z2: A * (4 * z1 - 2) ** 2 + B * (2 * z2 - 1) ** 2)
    xms = np.zeros((m.numnp, 3))
    for node in range(m.numnp):
        xms[node, 0] = x_cps[node % m.num1np]
        xms[node, 1] = y_cps[node // m.num1np]
        xms[node, 2] = z_cps[node]
    cps = np.zeros((m.numnp, m.ndf))
    IX = m.IX
    for el in range(1, m.numel + 1):
        xe, ce = xms[IX[:, el - 1] - 1, :], cps[IX[:, el - 1] - 1, :]
        for gp in range(1, 10):
            g = m.geo_dyn_stress(el, gp, xe, ce)
            x, y, z = g["x"]
            assert np.isclose(z, A * (x - 4) ** 2 + B * (y + 2) ** 2, rtol=np.sqrt(EPS), atol=0)
            ref = np.array([[4., 0.], [0., 2.], [8 * A * (x - 4), 4 * B * (y + 2)]])
            assert np.linalg.norm(g["a_"] - ref) < eps(4.e2)
            aco = g["a_"].T @ g["a_"]
            assert np.linalg.norm(g["acon"] @ aco - np.eye(2)) < eps(1.e2)
            assert np.linalg.norm(aco @ g["acon"] - np.eye(2)) < eps(1.e2)
            assert abs(np.trace(g["acon"] @ g["b"]) / 2 - g["H"]) < eps(1.e1)
            assert abs(np.linalg.det(g["acon"] @ g["b"]) - g["K"]) < eps(1.e1)


# ---------------------------------------------------------------- geo_dyn_stress_tests.jl:57-119
def test_geo_dyn_stress_flat_patch_couette_poiseuille():
    m = orc.Mesh(motion=orc.STATIC, scenario=orc.F_CAVI, num1el=3, num2el=2)
    kv1, kv2 = m.kv(1), m.kv(2)
    x_cps = kv1.cps_1d(lambda z: 4 * z + 2)
    y_cps = kv2.cps_1d(lambda z: 2 * z - 3)
    om, U = 12., 16.
    c_vs = kv2.cps_1d(lambda z: 2 * om * z)
    p_vs = kv1.cps_1d(lambda z: -4 * U * (z ** 2 - z))
    xms = np.zeros((m.numnp, 3))
    cps_c, cps_p = np.zeros((m.numnp, m.ndf)), np.zeros((m.numnp, m.ndf))
    for node in range(m.numnp):
        xms[node, 0] = x_cps[node % m.num1np]
        xms[node, 1] = y_cps[node // m.num1np]
        cps_c[node, 0] = c_vs[node // m.num1np]
        cps_p[node, 1] = p_vs[node % m.num1np]
    IX = m.IX
    zv = 1.0
    area = 0.0
    for el in range(1, m.numel + 1):
        idx = IX[:, el - 1] - 1
        for gp in range(1, 10):
            gc = m.geo_dyn_stress(el, gp, xms[idx], cps_c[idx])
            gq = m.geo_dyn_stress(el, gp, xms[idx], cps_p[idx])
            area += gq["J"] * m.area_fns(el, gp)["w"]
            x = gc["x"][0]
            assert np.linalg.norm(gc["sig"] - zv * np.array([0., 0., 2 * om / 8])) < eps(2.e2)
            assert np.linalg.norm(gq["sig"] - zv * np.array([0., 0., U * (4 - x) / 8])) < eps(3.e2)
    assert np.isclose(area, 8.0, rtol=np.sqrt(EPS), atol=0)
