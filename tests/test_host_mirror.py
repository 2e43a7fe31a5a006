"""Host-side mirror of the reference's Input layer (membranealefem.jl_b200/host) against the oracle's literal
restatement: DOF numbering, connectivity and basis tables must be BIT-EXACT (BASELINE.json north_star)."""
import numpy as np
import pytest

import mafb200 as maf
from oracle import oracle as orc

CONFIGS = [(maf.LAG, maf.F_PULL, 5, 4, {"pull_speed": 0.5}), (maf.EUL, maf.F_PULL, 17, 17, {"pull_speed": 0.5}),
           (maf.ALEV, maf.F_PULL, 19, 20, {"pull_speed": 0.5}), (maf.ALEVB, maf.F_PULL, 17, 17, {"pull_speed": 0.5}),
           (maf.ALEVB, maf.F_PULL, 30, 23, {"pull_speed": 0.5}), (maf.STATIC, maf.F_CAVI, 6, 5, {}),
           (maf.STATIC, maf.F_COUE, 4, 7, {}), (maf.STATIC, maf.F_POIS, 5, 5, {}),
           (maf.LAG, maf.F_BEND, 4, 3, {"bend_tm": 2.0, "bend_mf": 0.5}),
           (maf.ALEV, maf.F_BEND, 5, 4, {"bend_tm": 2.0, "bend_mf": 0.5})]


@pytest.mark.parametrize("motion,scen,n1,n2,args", CONFIGS)
def test_tables_bit_exact(motion, scen, n1, n2, args):
    p = maf.Params(motion=motion, scenario=scen, num1el=n1, num2el=n2, length=8.0, output=False)
    hm = maf.Mesh(p, **args)
    om = orc.Mesh(motion=int(motion), scenario=int(scen), num1el=n1, num2el=n2, length=8.0,
                  pull_speed=args.get("pull_speed", 0.0), bend_mf=args.get("bend_mf", 0.0),
                  bend_tm=args.get("bend_tm", 1.0))
    assert (hm.numel, hm.numnp, hm.ndf, hm.nmdf) == (om.numel, om.numnp, om.ndf, om.nmdf)
    assert np.array_equal(hm.IX, om.IX)
    assert np.array_equal(hm.ID, om.ID)
    assert np.array_equal(hm.LM, om.LM)
    n_o, d_o = om.ID_inv
    assert np.array_equal(hm.ID_inv[0], n_o) and np.array_equal(hm.ID_inv[1], d_o)
    assert np.array_equal(hm.dofs8(), om.dofs)
    assert np.array_equal(hm.kv1.zs, om.kv(1).zs) and np.array_equal(hm.kv2.zs, om.kv(2).zs)
    for d, line in ((1, hm.line_gp_fns1), (2, hm.line_gp_fns2)):
        ids, tab, edge = om.line(d)
        assert np.array_equal(ids, line.uel_ids)
        assert np.array_equal(tab, line.ufns)          # bit-exact 1-D basis tables
        assert np.array_equal(edge, line.edge)
    assert np.array_equal(hm.area_gp_fns.uel_ids, om.area_uel_ids)
    for b in (1, 2, 3, 4):
        assert np.array_equal(hm.bdry_elems[maf.Boundary(b)], om.bdry_elems(b))
    dirs, neus = om.bcs
    assert [(int(u), int(n), float(v)) for (u, n, v) in hm.inh_dir_bcs] == dirs
    assert [(int(b), int(t), float(v)) for (b, t, v) in hm.inh_neu_bcs] == neus
    # 2-D tables through the reference's accessors (Mesh.jl:311-469)
    for el in (1, hm.numel // 2 + 1, hm.numel):
        for gp in (1, 5, 9):
            fo, fh = om.area_fns(el, gp), maf.get_basis_fns(el, gp, hm)
            for k in ("w", "N", "dN", "ddN"):
                assert np.array_equal(fo[k], fh[k])
    for b in (1, 2, 3, 4):
        el = int(om.bdry_elems(b)[len(om.bdry_elems(b)) // 2])
        fo, fh = om.bdry_fns(b, el, 2), maf.get_basis_fns(maf.Boundary(b), el, 2, hm)
        for k in ("w", "N", "dN", "ddN"):
            assert np.array_equal(fo[k], fh[k])


def test_prepare_input_matches_oracle_flat_state():
    for motion in (maf.LAG, maf.EUL, maf.ALEVB):
        p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=9, num2el=8, output=False)
        mesh, xms, cps = maf.prepare_input(p, pull_speed=0.5, dts=[0.5], t0=0.0, t0_id=0)
        om = orc.Mesh(motion=int(motion), scenario=orc.F_PULL, num1el=9, num2el=8, pull_speed=0.5)
        xo, co = om.flat_state()
        assert np.abs(xms - xo).max() < 1e-12 and np.array_equal(cps, co)


def test_2d_collocation_reproduces_dense_reference_solve():
    # tensor-product control points == the reference's dense numnp x numnp solve (Spline.jl:540-567)
    kv1 = maf.KnotVector(5, 2)
    kv2 = maf.KnotVector(4, 2)
    f = lambda z1, z2: 2.2 * (4 * z1 - 2) ** 2 + 1.7 * (2 * z2 - 1) ** 2 + 0.3 * z1 * z2
    ours = maf.pkg.host.spline.get_2d_bspline_cps(kv1, kv2, f)
    ref = orc.cps_2d(orc.KnotVector.uniform(5, 2), orc.KnotVector.uniform(4, 2), f)
    assert np.abs(ours - ref).max() < 1e-12


def test_check_params_asserts_like_reference():
    ok = dict(dts=[0.5], t0=0.0, t0_id=0)
    maf.check_params(maf.Params(output=False), pull_speed=0.5, **ok)
    with pytest.raises(AssertionError):
        maf.check_params(maf.Params(output=False), **ok)                                  # need pull_speed
    with pytest.raises(AssertionError):
        maf.check_params(maf.Params(scenario=maf.F_COUE, motion=maf.LAG, output=False), **ok)
    with pytest.raises(AssertionError):
        maf.check_params(maf.Params(scenario=maf.F_BEND, motion=maf.ALEVB, length=1.0, output=False), bend_tm=1.0,
                         bend_mf=0.5, **ok)
    with pytest.raises(AssertionError):
        maf.check_params(maf.Params(output=False), pull_speed=0.5, t0=0.0, t0_id=0)       # need Δts


def test_synthetic_state_is_rank_independent():
    p = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=21, num2el=21, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    x1, c1 = maf.synthetic_state(mesh, p)
    x2, c2 = maf.synthetic_state(mesh, p)
    assert np.array_equal(x1, x2) and np.array_equal(c1, c2)
    assert np.abs(c1).max() <= 0.1 + 0.25 + 1e-12 and np.abs(x1[:, 2]).max() > 0.01 * p.length
