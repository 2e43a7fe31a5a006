"""GPU parity tests proper: the CUDA path, called through the C ABI (include/maf.h via ctypes), against the CPU
oracle on the same seeded inputs. Bar (BASELINE.json north_star): sparsity pattern and DOF numbering exact, residual
and tangent within 1e-11 relative in FP64 (rule in cases.compare / DESIGN.md 'tolerance'), Newton histories equal."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import mafb200 as maf
from cases import (DEFAULT_17, EXTRA_NEUMANN, SMALL, active_unknowns, check_pattern_contract, check_strict, compare,
                   make_case)

pytestmark = pytest.mark.gpu
RTOL = 1e-11   # stated by BASELINE.json north_star


def _K(asm, nz):
    colptr, rowval = asm.pattern()
    return sp.csc_matrix((nz, rowval - 1, colptr - 1), shape=(asm.nmdf, asm.nmdf))


@pytest.mark.parametrize("name", SMALL + list(DEFAULT_17))
def test_assembly_matches_oracle(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    asm = maf.Assembler(hm, p)
    # DOF numbering / pattern container
    colptr, rowval = asm.pattern()
    assert colptr[0] == 1 and colptr[-1] == asm.nnz + 1 and rowval.min() >= 1 and rowval.max() <= hm.nmdf
    truth = None
    for mode in (maf.SCATTER_ATOMIC, maf.SCATTER_DETERMINISTIC):
        r, nz, rn = asm.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0), scatter_mode=mode)
        K = _K(asm, nz)
        er, ek = compare(r, K, r_o, K_o, active_unknowns(om, cps))
        assert er < RTOL and ek < RTOL, (name, mode, er, ek)
        # strict rule on EVERY entry against the extended-precision truth (cases.strict_errors):
        #   |x - truth| <= 1e-11 |truth| + eps E,  E = the reference algorithm's own rounding-error bound for that entry
        uk, ur, truth = check_strict(name, r, nz, colptr, rowval, xms, cps, time, dt, f"gpu mode {mode}", truth)
        assert abs(rn - float(r @ r)) <= 1e-12 * max(float(r @ r), 1e-300)
        check_pattern_contract(K, K_o, generic="flat" not in name)
    asm.close()


@pytest.mark.parametrize("name", ["lag_bend_4x3", "alevb_bend_pn_4x3", "eul_bend_3x4"])
def test_shear_and_top_bottom_moment(name):
    """SHEAR on every side and MOMENT on TOP / BOTTOM (FiniteElement.jl:374-380; no scenario of Bc.jl sets them up):
    injected through maf_mesh_desc.neu_* and the oracle's inh_neu_bcs, both scatter paths, oracle + strict rule."""
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    hm.inh_neu_bcs = list(EXTRA_NEUMANN)
    om.set_neumann(EXTRA_NEUMANN)
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    asm = maf.Assembler(hm, p)
    colptr, rowval = asm.pattern()
    truth = None
    for mode in (maf.SCATTER_ATOMIC, maf.SCATTER_DETERMINISTIC):
        r, nz, _ = asm.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0), scatter_mode=mode)
        er, ek = compare(r, _K(asm, nz), r_o, K_o)
        assert er < RTOL and ek < RTOL, (name, mode, er, ek)
        _, _, truth = check_strict(name, r, nz, colptr, rowval, xms, cps, time, dt, f"gpu mode {mode}", truth,
                                   neumann=EXTRA_NEUMANN)
    asm.close()


@pytest.mark.parametrize("name", ["alevb_pull_17x17", "lag_pull_17x17", "eul_pull_5x4"])
def test_deterministic_path_is_bitwise_reproducible(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    asm = maf.Assembler(hm, p)
    outs = [asm.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC) for _ in range(3)]
    for r, nz, rn in outs[1:]:
        assert np.array_equal(r, outs[0][0]) and np.array_equal(nz, outs[0][1]) and rn == outs[0][2]
    # and the atomics path agrees with it to round-off
    r, nz, _ = asm.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_ATOMIC)
    assert np.abs(nz - outs[0][1]).max() <= 1e-13 * np.abs(nz).max()
    assert np.abs(r - outs[0][0]).max() <= 1e-13 * max(np.abs(r).max(), 1e-300) + 1e-13


def test_sym_pattern_is_full_lm_union():
    p, hm, om, xms, cps, time, dt, args = make_case("alevb_pull_17x17")
    asm = maf.Assembler(hm, p, pattern_mode=maf.PATTERN_SYM)
    assert asm.nnz == 393492          # SURVEY.md section 8 size table (ALEVB 17x17 symbolic nnz)
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    r, nz, _ = asm.assemble(xms, cps, time, dt)
    er, ek = compare(r, _K(asm, nz), r_o, K_o, active_unknowns(om, cps))
    assert er < RTOL and ek < RTOL


def test_device_entry_point_and_element_ranges():
    import torch
    p, hm, om, xms, cps, time, dt, args = make_case("alevb_pull_17x17")
    asm = maf.Assembler(hm, p)
    r_ref, nz_ref, _ = asm.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    dx = torch.from_numpy(np.ascontiguousarray(xms.T)).cuda()      # column-major numnp x 3
    dc = torch.from_numpy(np.ascontiguousarray(cps.T)).cuda()
    parts_r, parts_k = [], []
    cut = 8 * hm.num1el
    for (a, b) in ((1, cut), (cut + 1, hm.numel)):
        asm.set_element_range(a, b)
        dr = torch.zeros(hm.nmdf, dtype=torch.float64, device="cuda")
        dk = torch.zeros(asm.nnz, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        asm.assemble_device(dx.data_ptr(), dc.data_ptr(), time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC,
                            d_r=dr.data_ptr(), d_nzval=dk.data_ptr())
        asm.sync()
        parts_r.append(dr.cpu().numpy())
        parts_k.append(dk.cpu().numpy())
    assert np.abs(parts_r[0] + parts_r[1] - r_ref).max() <= 1e-13 * np.abs(r_ref).max()
    assert np.abs(parts_k[0] + parts_k[1] - nz_ref).max() <= 1e-13 * np.abs(nz_ref).max()


@pytest.mark.parametrize("motion", [maf.LAG, maf.EUL, maf.ALEVB])
def test_newton_history_matches_oracle(motion):
    """Configs 1-3 of BASELINE.json (tether pull from the flat patch): two time steps of time_step! with the GPU
    assembly vs. with the oracle assembly, same host solver. Histories must agree iterate by iterate."""
    p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=9, num2el=9, output=False)
    args = dict(pull_speed=0.5, dts=[0.5, 0.5], t0=0.0, t0_id=0)
    mesh, xms, cps = maf.prepare_input(p, **args)
    from oracle import oracle as orc
    om = orc.Mesh(motion=int(motion), scenario=orc.F_PULL, num1el=9, num2el=9, pull_speed=0.5)
    xo, co = xms.copy(), cps.copy()
    hist_gpu = maf.run_analysis(mesh, xms, cps, p, **args)
    # oracle-driven Newton loop (FiniteElement.jl:11-63, Analysis.jl:60-93)
    n_inv, d_inv = om.ID_inv
    mmo = maf.pkg.host.mesh.get_m_motion_order(motion, mesh.dofs)
    hist_ref = []
    t = 0.0
    for dt in args["dts"]:
        t += dt
        maf.update_xms(motion, xo, co, dt, mesh.dofs)
        eps = []
        for it in range(14):
            r, K = om.calc_r_K(xo, co, t, dt)
            du = -spla.splu(sp.csc_matrix(K)).solve(r)
            dc = np.zeros_like(co)
            dc[n_inv - 1, d_inv - 1] = du
            co += dc
            maf.update_xms(motion, xo, dc, dt, mesh.dofs)
            eps.append(np.linalg.norm(du) / om.nmdf)
            if eps[-1] < p.enr:
                break
        hist_ref.append(eps)
    assert [len(h) for h in hist_gpu] == [len(h) for h in hist_ref]
    for hg, hr in zip(hist_gpu, hist_ref):
        for eg, er_ in zip(hg[:-1], hr[:-1]):          # the last entry is at round-off level (< 1e-12)
            assert abs(eg - er_) <= 1e-6 * er_, (hist_gpu, hist_ref)
        assert hg[-1] < p.enr and hr[-1] < p.enr
    assert np.abs(xms - xo).max() <= 1e-9 and np.abs(cps - co).max() <= 1e-9


def _load_newton_golden(name):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"newton_17x17_{name}.npz"))
    hist, k = [], 0
    for n in z["lens"]:
        hist.append(z["eps"][k:k + int(n)].tolist())
        k += int(n)
    return z, hist


def _same_history(h_a, h_b, enr):
    """Iterate by iterate: same number of iterations per step; every eps above the round-off level (1e-9: the last
    iterate of a quadratically converging step is noise below enr = 1e-12) agrees to 1e-6 relative."""
    assert [len(h) for h in h_a] == [len(h) for h in h_b], (h_a, h_b)
    for ha, hb in zip(h_a, h_b):
        for ea, eb in zip(ha, hb):
            if eb > 1e-9:
                assert abs(ea - eb) <= 1e-6 * eb, (ha, hb)
            else:
                assert ea < max(1e-9, 100 * eb) or ea < enr
        assert ha[-1] < enr and hb[-1] < enr


@pytest.mark.parametrize("name,motion", [("lag", maf.LAG), ("eul", maf.EUL), ("alevb", maf.ALEVB)])
@pytest.mark.parametrize("resident", [False, True])
def test_newton_history_17x17_32_steps(name, motion, resident):
    """BASELINE.json configs 1-3 at the reference's own sizes: 17 x 17 patch, length 64, pull_speed 0.5,
    dts = fill(0.5, 32) (docs/src/index.md:149-152, Params.jl:39-40). All 32 steps through the library (host loop and
    device-resident loop) against the committed oracle-driven histories (tests/golden/make_newton_golden.py), plus a
    fresh oracle run of the first 3 steps."""
    from helpers import newton_history
    from oracle import oracle as orc
    p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=17, num2el=17, output=False)
    args = dict(pull_speed=0.5, dts=[0.5] * 32, t0=0.0, t0_id=0)
    mesh, xms, cps = maf.prepare_input(p, **args)
    x0, c0 = xms.copy(), cps.copy()
    hist = maf.run_analysis(mesh, xms, cps, p, resident=resident, **args)
    z, hist_g = _load_newton_golden(name)
    _same_history(hist, hist_g, p.enr)
    assert np.abs(xms - z["xms"]).max() <= 1e-8 and np.abs(cps - z["cps"]).max() <= 1e-8
    assert abs(xms[:, 2].max() - 8.0) < 1e-9          # the pulled nodes moved pull_speed * 16 (Bc.jl:228-243)
    if not resident:
        om = orc.Mesh(motion=int(motion), scenario=orc.F_PULL, num1el=17, num2el=17, pull_speed=0.5)
        h3 = newton_history(lambda x, c, t, dt: om.calc_r_K(x, c, t, dt, nthreads=8), motion, mesh.dofs, om.ID_inv,
                            om.nmdf, x0, c0, args["dts"][:3])
        _same_history(hist[:3], h3, p.enr)
    maf.pkg.host.analysis.close_assemblers(mesh)


def test_headline_config_spot_parity():
    """BASELINE.json config 5, the configuration the bench numbers are quoted on (1001 x 1001 F_PULL patch, centre-
    refined knots = 625 unique elements, ALEVB, the bench's synthetic state): > 200 single elements -- the pulled
    element and its two rings, corners, edges incl. the mid-edge Dirichlet nodes, the knot-spacing transitions, random
    interior ones -- through maf_set_element_range(el, el) against the oracle's element routine scattered through LM
    and against the extended-precision truth (tests/spot_parity.py). DOF numbering exact, every written slot inside
    the element's LM x LM block, strict rule on every entry."""
    import torch
    from oracle import oracle as orc
    from spot_parity import check_elements, select_elements
    p = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=1001, num2el=1001, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    xms, cps = maf.synthetic_state(mesh, p)
    kw = dict(motion=orc.ALEVB, scenario=orc.F_PULL, num1el=1001, num2el=1001, length=p.length, pull_speed=0.5)
    om, ot = orc.Mesh(**kw), orc.Mesh(kind="truth", **kw)
    assert np.array_equal(om.ID, mesh.ID) and om.nmdf == mesh.nmdf and (om.nuel1, om.nuel2) == (25, 25)
    asm = maf.Assembler(mesh, p)
    dx = torch.from_numpy(np.ascontiguousarray(xms.T)).cuda()
    dc = torch.from_numpy(np.ascontiguousarray(cps.T)).cuda()
    els = select_elements(mesh)
    assert len(els) >= 200
    worst = check_elements(asm, mesh, om, ot, dx.data_ptr(), dc.data_ptr(), xms, cps, 0.5, 0.5, els)
    print("headline spot parity:", worst)
    assert worst["n"] == len(els) and worst["K_rel_oracle"] < RTOL
    asm.close()


def test_translate_ale_emulation():
    """BASELINE.json config 4 (translate-ale, video only in the reference): converged ALEVB tether state, then an
    in-plane Dirichlet velocity on the pulled nodes; same ALE assembly path on a sheared state. Parity unpinned by the
    reference (scenario absent from its source) -- checked against the oracle like every other state."""
    p = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=9, num2el=9, output=False)
    args = dict(pull_speed=0.5, dts=[0.5, 0.5, 0.5], t0=0.0, t0_id=0)
    mesh, xms, cps = maf.prepare_input(p, **args)
    maf.run_analysis(mesh, xms, cps, p, **args)
    U = maf.Dof.Unknown
    for (unk, node, val) in mesh.inh_dir_bcs:
        cps[node - 1, mesh.dofs[U.vx] - 1] = 0.2
        cps[node - 1, mesh.dofs[U.vmx] - 1] = 0.2
    maf.update_xms(p.motion, xms, cps, 0.5, mesh.dofs)
    from oracle import oracle as orc
    om = orc.Mesh(motion=orc.ALEVB, scenario=orc.F_PULL, num1el=9, num2el=9, pull_speed=0.5)
    r_o, K_o = om.calc_r_K(xms, cps, 2.0, 0.5)
    r, K = maf.calc_r_K(mesh, xms, cps, 2.0, 0.5, p, dropzeros=False)
    er, ek = compare(r, K, r_o, K_o, active_unknowns(om, cps))
    assert er < RTOL and ek < RTOL


def test_error_behaviour():
    p, hm, om, xms, cps, time, dt, args = make_case("lag_pull_3x3")
    bad = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=3, num2el=3, length=4.0, output=False)
    with pytest.raises(maf.MafError):
        maf.Assembler(hm, bad)                        # dof set does not match the motion
    asm = maf.Assembler(hm, p)
    with pytest.raises(maf.MafError):
        asm.assemble(xms, cps, time, dt, scatter_mode=7)
    with pytest.raises(maf.MafError):
        asm.set_element_range(0, 5)
    # still usable after an error
    r, nz, _ = asm.assemble(xms, cps, time, dt)
    assert np.isfinite(r).all() and np.isfinite(nz).all()


def test_synthetic_large_patch_properties():
    """Size-independent properties at a size the oracle cannot reach in test time (301 x 301 = 90 601 elements):
    (i) the two scatter paths agree; (ii) strips sum to the whole; (iii) the tangent is the derivative of the
    residual: r(u + h du) - r(u - h du) = 2 h K du (central difference through the library itself)."""
    p = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=301, num2el=301, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    xms, cps = maf.synthetic_state(mesh, p)
    asm = maf.Assembler(mesh, p)
    dt = 0.5
    r0, nz0, _ = asm.assemble(xms, cps, dt, dt, scatter_mode=maf.SCATTER_ATOMIC)
    r1, nz1, _ = asm.assemble(xms, cps, dt, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    assert np.abs(nz0 - nz1).max() <= 1e-12 * np.abs(nz1).max()
    assert np.abs(r0 - r1).max() <= 1e-12 * np.abs(r1).max()
    K = _K(asm, nz1)
    rng = np.random.default_rng(3)
    du = rng.standard_normal(mesh.nmdf)
    node_of, dof_of = mesh.ID_inv
    h = 1e-6
    rr = []
    for s in (+1, -1):
        c2, x2 = cps.copy(), xms.copy()
        d = np.zeros_like(cps)
        d[node_of - 1, dof_of - 1] = s * h * du
        c2 += d
        maf.update_xms(p.motion, x2, d, dt, mesh.dofs)
        rr.append(asm.assemble(x2, c2, dt, dt)[0])
    fd = (rr[0] - rr[1]) / (2 * h)
    kd = K @ du
    assert np.abs(fd - kd).max() <= 1e-6 * np.abs(kd).max()


def test_chunk_plan_override_changes_the_schedule_not_the_result(monkeypatch):
    """MAF_PLAN replaces the static plan of the contraction phase (tools/tune_plan.py): any valid plan assembles the
    same K (only the order of the atomic additions moves); an invalid one is refused by maf_create."""
    p, hm, om, xms, cps, time, dt, args = make_case("alevb_pull_5x4")
    asm = maf.Assembler(hm, p)
    text = asm.chunk_plan()
    warps = [[int(c) for c in w.split(",") if c] for w in text.split("/")]
    assert sorted(c for w in warps for c in w) == list(range(sum(len(w) for w in warps)))
    r0, nz0, _ = asm.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    # everything on the first warp, in reverse order
    rev = ",".join(str(c) for c in sorted((c for w in warps for c in w), reverse=True))
    monkeypatch.setenv("MAF_PLAN", rev)
    asm2 = maf.Assembler(hm, p)
    assert asm2.chunk_plan().split("/")[0] == rev
    r1, nz1, _ = asm2.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    assert np.array_equal(nz0, nz1) and np.array_equal(r0, r1)   # deterministic path: bitwise
    ra, nza, _ = asm2.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_ATOMIC)
    assert np.abs(nza - nz0).max() <= 1e-13 * np.abs(nz0).max()
    monkeypatch.setenv("MAF_PLAN", "0,0,1")
    with pytest.raises(maf.MafError, match="invalid chunk plan"):
        maf.Assembler(hm, p)


def test_registered_host_buffers_give_the_same_result():
    """maf_host_register / maf_host_unregister: outputs written into page-locked caller buffers equal the ones
    written into pageable ones; a second registration of the same range is reported, not fatal."""
    p, hm, om, xms, cps, time, dt, args = make_case("eul_pull_5x4")
    asm = maf.Assembler(hm, p)
    r0, nz0, rn0 = asm.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    r1, nz1 = np.full(asm.nmdf, np.nan), np.full(asm.nnz, np.nan)
    maf.host_register(r1)
    maf.host_register(nz1)
    try:
        with pytest.raises(maf.MafError):
            maf.host_register(nz1)
        asm.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC, r=r1, nzval=nz1)
    finally:
        maf.host_unregister(r1)
        maf.host_unregister(nz1)
    assert np.array_equal(r0, r1) and np.array_equal(nz0, nz1)


def test_pipelined_host_path_on_a_small_mesh(monkeypatch):
    """maf_assemble's strip-pipelined path (default from 32768 elements: finished nzval / r ranges are copied to the
    host while the next strip of element rows is assembled) forced on a 17 x 17 mesh: bit-identical to the
    deterministic path up to the order of the atomic additions, and within the strict rule of the truth."""
    name = "alevb_pull_17x17"
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    asm = maf.Assembler(hm, p)
    colptr, rowval = asm.pattern()
    r0, nz0, rn0 = asm.assemble(xms, cps, time, dt)
    launches0 = asm.launch_count()
    monkeypatch.setenv("MAF_PIPELINE_MIN_ELEMS", "1")
    r1, nz1, rn1 = asm.assemble(xms, cps, time, dt)
    assert asm.launch_count() - launches0 >= 8            # one area launch per strip: the pipelined path did run
    assert np.abs(nz1 - nz0).max() <= 1e-13 * np.abs(nz0).max() and np.abs(r1 - r0).max() <= 1e-13 * np.abs(r0).max()
    assert abs(rn1 - rn0) <= 1e-12 * rn0
    check_strict(name, r1, nz1, colptr, rowval, xms, cps, time, dt, "pipelined host path")
    rd, nzd, _ = asm.assemble_resident(time, dt)        # and through the resident entry point
    assert np.abs(nzd - nz0).max() <= 1e-13 * np.abs(nz0).max()
    asm.close()


def test_graph_replay_equals_plain_launches(monkeypatch):
    """Small meshes replay their launch sequence from a CUDA graph (captured once per scatter mode / dt / Neumann
    values): same numbers as plain launches (bitwise on the deterministic path), across changes of dt, of the element
    range and of the time-dependent MOMENT value."""
    name = "alevb_bend_pn_4x3"          # F_BEND: the Neumann value depends on time (FiniteElement.jl:379)
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    monkeypatch.setenv("MAF_NO_GRAPH", "1")
    plain = maf.Assembler(hm, p)
    monkeypatch.delenv("MAF_NO_GRAPH")
    graph = maf.Assembler(hm, p)
    bt = args["bend_tm"]
    for (t, d) in ((0.1, 0.37), (0.1, 0.37), (0.2, 0.37), (0.2, 0.5), (0.1, 0.37)):
        for mode in (maf.SCATTER_DETERMINISTIC, maf.SCATTER_ATOMIC):
            r0, k0, n0 = plain.assemble(xms, cps, t, d, bend_tm=bt, scatter_mode=mode)
            r1, k1, n1 = graph.assemble(xms, cps, t, d, bend_tm=bt, scatter_mode=mode)
            if mode == maf.SCATTER_DETERMINISTIC:
                assert np.array_equal(r0, r1) and np.array_equal(k0, k1) and n0 == n1
            else:
                assert np.abs(k0 - k1).max() <= 1e-13 * np.abs(k0).max() and np.abs(r0 - r1).max() <= 1e-13 * np.abs(r0).max()
    assert plain.kernel_info()["graph_replays"] == 0 and graph.kernel_info()["graph_replays"] == 10
    graph.set_element_range(1, 6)
    plain.set_element_range(1, 6)
    r0, k0, _ = plain.assemble(xms, cps, 0.1, 0.37, bend_tm=bt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    r1, k1, _ = graph.assemble(xms, cps, 0.1, 0.37, bend_tm=bt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    i = graph.range_info()
    sl = slice(i["slots"][0] - 1, i["slots"][1])
    assert np.array_equal(k0[sl], k1[sl])
    plain.close()
    graph.close()


@pytest.mark.parametrize("name,rows", [("alevb_pull_17x17", 2), ("alevb_pull_17x17", 5), ("lag_pull_17x17", 3),
                                       ("alevb_pull_fine_19x18", 4)])
def test_banded_deterministic_staging(monkeypatch, name, rows):
    """Large ranges stage the deterministic path band by band of element rows (ring of three bands, ~8 % of nzval
    instead of 3.4 x nzval): forced on small meshes with MAF_BAND_ROWS, bitwise equal to the unbanded path (same
    ascending-element-id sums), also on a strip of a 2-strip partition."""
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    ref = maf.Assembler(hm, p)
    r0, k0, n0 = ref.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    monkeypatch.setenv("MAF_BAND_ROWS", str(rows))
    band = maf.Assembler(hm, p)
    for _ in range(2):
        r1, k1, n1 = band.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
        assert np.array_equal(r0, r1) and np.array_equal(k0, k1) and n0 == n1
    assert band.launch_count() > ref.launch_count() + 4        # the bands did run
    ref.close()
    band.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["alevb_pull_17x17", "lag_pull_flat_7x7", "alev_bend_4x4_pn"])
def test_gather_through_pair_classes_equals_the_scanning_gather(monkeypatch, name):
    """The deterministic gather looks the staged rows of a node pair up in precomputed contribution classes; the
    scanning form (element lists of the column node, row node searched in each: the definition, and the fallback for
    meshes with more than 65535 classes) must give the same bits."""
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    fast = maf.Assembler(hm, p)
    r0, k0, n0 = fast.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    monkeypatch.setenv("MAF_NO_PAIR_CLASSES", "1")
    scan = maf.Assembler(hm, p)
    r1, k1, n1 = scan.assemble(xms, cps, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    assert np.array_equal(r0, r1) and np.array_equal(k0, k1) and n0 == n1
    fast.close()
    scan.close()
