"""the Newton update on a device-resident state and the pull force.

CPU part: the same __host__ __device__ entry functions the kernels run (maf_state.cuh, element phases) are executed
by tests/emu and compared with the host loop / the oracle. GPU part: through the C ABI."""
import numpy as np
import pytest

import mafb200 as maf
from cases import PULL, make_case
from emu_driver import Emu
from oracle import oracle as orc

PULL_CASES = list(PULL)


def _host_update(mesh, p, xms, cps, du, dt):
    """time_step!'s update, written as the reference does (FiniteElement.jl:41-46)."""
    node_of, dof_of = mesh.ID_inv
    dcps = np.zeros_like(cps)
    dcps[node_of - 1, dof_of - 1] = du
    cps += dcps
    maf.update_xms(p.motion, xms, dcps, dt, mesh.dofs)


@pytest.mark.parametrize("name", ["lag_pull_5x4", "eul_pull_5x4", "alevb_pull_5x4", "static_cavi_4x5"])
def test_emulated_state_update_is_bitwise_the_host_loop(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    rng = np.random.default_rng(3)
    du = rng.standard_normal(hm.nmdf) * 1e-2
    emu = Emu(hm, p)
    xe, ce = xms.copy(order="F"), cps.copy(order="F")
    xh, ch = xms.copy(order="F"), cps.copy(order="F")
    emu.state_predict(dt, xe, ce)
    maf.update_xms(p.motion, xh, ch, dt, hm.dofs)
    assert np.array_equal(xe, xh)
    emu.state_update(du, dt, xe, ce)
    _host_update(hm, p, xh, ch, du, dt)
    assert np.array_equal(xe, xh) and np.array_equal(ce, ch)
    if p.motion != maf.STATIC:
        assert not np.array_equal(xe, xms)      # the mesh did move


@pytest.mark.parametrize("name", PULL_CASES)
def test_emulated_pull_force_matches_oracle(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    adj, maps = maf.get_adj_maps(hm.num1el, hm.numel, hm.IX, p.poly)
    assert len(adj) == 25 and adj[12] == maf.get_pull_el_id(hm.numel)
    rv = Emu(hm, p).elem_v_residuals(xms, cps, adj)
    for k, e in enumerate(adj):     # element by element against calc_elem_dof_residuals of the oracle
        rv_o = om.elem_dof_residuals(int(e), xms, cps)[0]
        assert np.abs(rv[k] - rv_o).max() <= 1e-11 * max(np.abs(rv_o).max(), 1e-300)
    f = maf.pkg.host.pullforce.sum_pull_force(rv, maps)
    f_o = om.calc_pull_force(xms, cps)
    # |f| is a small difference of O(max |rv|) terms: measure against what is summed
    assert np.abs(f - f_o).max() <= 1e-11 * np.abs(rv).max()


def test_adj_maps_follow_the_reference():
    p = maf.Params(motion=maf.LAG, scenario=maf.F_PULL, num1el=7, num2el=7, output=False)
    hm = maf.Mesh(p, pull_speed=0.5)
    adj, maps = maf.get_adj_maps(7, 49, hm.IX, 2)
    assert maf.get_pull_el_id(49) == 25 and list(adj[:5]) == [9, 10, 11, 12, 13] and adj[-1] == 41
    # the pulled element shares all 9 nodes with itself, one node with the far corners
    assert maps[12][0].size == 27 and np.array_equal(maps[12][0], maps[12][1])
    assert maps[0][0].size == 3 and maps[24][0].size == 3
    # corner element 9 = (e1, e2) - (2, 2): its last node is the pulled element's first
    assert list(maps[0][0]) == [0, 1, 2] and list(maps[0][1]) == [24, 25, 26]


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lag_pull_5x4", "eul_pull_5x4", "alevb_pull_5x4"])
def test_gpu_state_update_is_bitwise_the_host_loop(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    du = np.random.default_rng(3).standard_normal(hm.nmdf) * 1e-2
    asm = maf.Assembler(hm, p)
    asm.state_set(xms, cps)
    x0, c0 = asm.state_get()
    assert np.array_equal(x0, xms) and np.array_equal(c0, cps)
    xh, ch = xms.copy(order="F"), cps.copy(order="F")
    asm.state_predict(dt)
    maf.update_xms(p.motion, xh, ch, dt, hm.dofs)
    asm.state_update(du, dt)
    _host_update(hm, p, xh, ch, du, dt)
    xg, cg = asm.state_get()
    assert np.array_equal(xg, xh) and np.array_equal(cg, ch)
    # assembling the resident state = assembling the same state passed from the host (deterministic mode: bitwise)
    r1, k1, n1 = asm.assemble_resident(time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    r2, k2, n2 = asm.assemble(xh, ch, time, dt, scatter_mode=maf.SCATTER_DETERMINISTIC)
    assert np.array_equal(r1, r2) and np.array_equal(k1, k2) and n1 == n2
    asm.close()


@pytest.mark.gpu
def test_gpu_resident_state_requires_a_state():
    p, hm, om, xms, cps, time, dt, args = make_case("lag_pull_3x3")
    asm = maf.Assembler(hm, p)
    with pytest.raises(maf.MafError, match="no resident state"):
        asm.state_predict(0.5)
    with pytest.raises(maf.MafError, match="no resident state"):
        asm.assemble_resident(0.0, 0.5)
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("motion", [maf.LAG, maf.ALEVB])
def test_gpu_resident_newton_loop_equals_host_loop(motion):
    """run_analysis with resident=True (state on the device between iterations) and the default loop: same
    histories, same final state (deterministic scatter: bit for bit)."""
    p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=9, num2el=9, output=False)
    args = dict(pull_speed=0.5, dts=[0.5, 0.5], t0=0.0, t0_id=0, scatter_mode=maf.SCATTER_DETERMINISTIC)
    mesh, xa, ca = maf.prepare_input(p, **args)
    xb, cb = xa.copy(order="F"), ca.copy(order="F")
    fa, fb = [], []
    ha = maf.run_analysis(mesh, xa, ca, p, f_pulls=fa, **args)
    hb = maf.run_analysis(mesh, xb, cb, p, resident=True, f_pulls=fb, **args)
    assert ha == hb and np.array_equal(xa, xb) and np.array_equal(ca, cb)
    assert len(fa) == 2 and all(np.array_equal(u[2], v[2]) for u, v in zip(fa, fb))
    # physics: pulling up needs an upward force that grows as the tether forms
    assert fa[0][2][2] > 0 and fa[1][2][2] > fa[0][2][2] and abs(fa[1][2][0]) < 1e-9 * fa[1][2][2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", PULL_CASES)
def test_gpu_pull_force_matches_oracle(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    adj, maps = maf.get_adj_maps(hm.num1el, hm.numel, hm.IX, p.poly)
    asm = maf.Assembler(hm, p)
    asm.state_set(xms, cps)
    rv = asm.elem_v_residuals(adj)
    f = maf.calc_pull_force(hm, xms, cps, adj, maps, p)
    f_o = om.calc_pull_force(xms, cps)
    assert np.abs(f - f_o).max() <= 1e-11 * np.abs(rv).max()
    with pytest.raises(maf.MafError, match="outside 1..numel"):
        asm.elem_v_residuals([0])
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["alevb_pull_fine_19x18", "lag_bend_4x3", "static_pois_5x3"])
def test_generate_output_on_the_device(name):
    """maf_generate_output (Output.jl:32-118): positions and unknowns at every area / boundary Gauss
    point and corner, against the statement-by-statement numpy restatement on the host mirror's tables."""
    from cases import make_case
    from ref_numpy import generate_output_ref
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    asm = maf.Assembler(hm, p)
    asm.state_set(xms, cps)
    xo, uo = asm.generate_output()
    xr, ur = generate_output_ref(hm, xms, cps)
    assert xo.shape == xr.shape and uo.shape == ur.shape
    assert np.abs(xo - xr).max() <= 1e-14 * np.abs(xr).max() and np.abs(uo - ur).max() <= 1e-14 * max(np.abs(ur).max(), 1.0)
    # the corners of the flat-in-parameter patch are interpolated exactly by the clamped splines
    assert np.allclose(xo[0, 0, :2], xms[0, :2]) and np.allclose(xo[-1, -1, :2], xms[-1, :2])
    asm.close()
