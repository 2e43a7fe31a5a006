"""Oracle self-checks (SURVEY.md section 8c 'extra known answers'): the C++ restatement against an independent
numpy formulation, the flat F_PULL equilibrium, central finite differences, sizes of the symbolic pattern,
Julia sparse-insertion semantics with 1 and several tasks, and Newton convergence with the assembled tangent."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import ref_numpy as rn
from helpers import P_of, deformed_state
from oracle import oracle as orc


def _tabs(m, el):
    return [m.area_fns(el, gp) for gp in range(1, 10)]


@pytest.mark.parametrize("motion", [orc.LAG, orc.EUL, orc.ALEV, orc.ALEVB])
def test_element_r_K_vs_numpy(motion):
    m = orc.Mesh(motion=motion, scenario=orc.F_PULL, num1el=5, num2el=4, length=8.0, pull_speed=0.5, pn=0.0)
    xms, cps = deformed_state(m, seed=motion)
    IX, ID, dofs, ndf = m.IX, m.ID, m.dofs, m.ndf
    mmo = {orc.LAG: dofs[0:3], orc.STATIC: [0, 0, 0]}.get(motion, dofs[3:6])
    dt = 0.37
    for el in (1, 8, 13, 20):
        idx = IX[:, el - 1] - 1
        active = (ID[:, idx] != 0).T
        r_o, K_o = m.elem_r_K(el, xms, cps, dt)
        tabs = _tabs(m, el)
        r_n, K_n = rn.elem_r_K(lambda xe, ce: rn.area_residual(tabs, xe, ce, dofs, ndf, P_of(m), motion),
                               xms[idx], cps[idx], active, mmo, dt, ndf)
        sr, sk = np.abs(r_n).max(), np.abs(K_n).max()
        assert np.abs(r_o - r_n).max() <= 1e-13 * sr
        assert np.abs(K_o - K_n).max() <= 1e-12 * sk


def test_element_pn_and_static_vs_numpy():
    # normal pressure term (FiniteElement.jl:295-297) and the 2-D STATIC dof set (Bc.jl:65)
    m = orc.Mesh(motion=orc.STATIC, scenario=orc.F_COUE, num1el=4, num2el=4, length=4.0, pn=0.7)
    xms, cps = deformed_state(m, seed=3)
    idx = m.IX[:, 5] - 1
    active = (m.ID[:, idx] != 0).T
    r_o, K_o = m.elem_r_K(6, xms, cps, 0.5)
    tabs = _tabs(m, 6)
    r_n, K_n = rn.elem_r_K(lambda xe, ce: rn.area_residual(tabs, xe, ce, m.dofs, m.ndf, P_of(m), orc.STATIC),
                           xms[idx], cps[idx], active, [0, 0, 0], 0.5, m.ndf)
    assert np.abs(r_o - r_n).max() <= 1e-13 * np.abs(r_n).max()
    assert np.abs(K_o - K_n).max() <= 1e-12 * np.abs(K_n).max()


@pytest.mark.parametrize("motion", [orc.LAG, orc.EUL, orc.ALEVB])
def test_flat_pull_state_is_equilibrium(motion):
    # flat patch, lambda = kb/4, STRETCH = kb/4 => r = 0 on all active rows (before the predictor)
    m = orc.Mesh(motion=motion, scenario=orc.F_PULL, num1el=7, num2el=7, pull_speed=0.5)
    xms, cps = m.flat_state()
    r, K = m.calc_r_K(xms, cps, 0.5, 0.5)
    assert np.abs(r).max() < 1e-13


@pytest.mark.parametrize("motion,nnz_sym,nmdf", [(orc.LAG, 104345, 1273), (orc.EUL, 338457, 2321),
                                                 (orc.ALEV, 409788, 2474), (orc.ALEVB, 393492, 2410)])
def test_symbolic_sizes_17x17(motion, nnz_sym, nmdf):
    # SURVEY section 8 size table: nmdf and the full LM x LM union at the reference's default mesh
    m = orc.Mesh(motion=motion, scenario=orc.F_PULL, pull_speed=0.5)
    assert (m.numel, m.numnp, m.nmdf) == (289, 361, nmdf)
    LM = m.LM
    keys = set()
    for e in range(m.numel):
        act = LM[:, e][LM[:, e] != 0]
        keys.update((int(r) << 32 | int(c)) for r in act for c in act)
    assert len(keys) == nnz_sym


def test_julia_sparse_semantics_and_threads():
    m = orc.Mesh(motion=orc.ALEVB, scenario=orc.F_PULL, num1el=6, num2el=5, length=8.0, pull_speed=0.5)
    xms, cps = deformed_state(m, seed=11)
    r1, K1 = m.calc_r_K(xms, cps, 0.5, 0.5, nthreads=1)
    r3, K3 = m.calc_r_K(xms, cps, 0.5, 0.5, nthreads=3)
    assert np.abs(r1 - r3).max() <= 1e-13 * np.abs(r1).max()
    d = (K1 - K3)
    assert abs(d).max() <= 1e-12 * abs(K1).max()
    # structurally-zero dof blocks never get an entry: (v,pm) (lam,pm) (m,v) (m,lam) (p,lam)
    n_inv, d_inv = m.ID_inv
    dofs = m.dofs
    fld = {}
    for u, c in enumerate(dofs):
        fld[int(c)] = "vvvmmmlp"[u]
    Kc = K1.tocoo()
    zero_blocks = {("v", "p"), ("l", "p"), ("m", "v"), ("m", "l"), ("p", "l")}
    for rr, cc, vv in zip(Kc.row, Kc.col, Kc.data):
        assert (fld[int(d_inv[rr])], fld[int(d_inv[cc])]) not in zero_blocks
    # sorted rows within each column (SparseMatrixCSC invariant)
    for c in range(K1.shape[1]):
        rows = K1.indices[K1.indptr[c]:K1.indptr[c + 1]]
        assert np.all(np.diff(rows) > 0)


def test_tangent_vs_central_differences():
    # SURVEY 8c(v)
    m = orc.Mesh(motion=orc.ALEVB, scenario=orc.F_PULL, num1el=4, num2el=4, length=4.0, pull_speed=0.5)
    xms, cps = deformed_state(m, seed=5)
    dt = 0.5
    r0, K = m.calc_r_K(xms, cps, dt, dt)
    K = K.toarray()
    n_inv, d_inv = m.ID_inv
    dofs = m.dofs
    rng = np.random.default_rng(1)
    h = 1e-6
    for u in rng.choice(m.nmdf, size=12, replace=False):
        node, dof = int(n_inv[u]) - 1, int(d_inv[u])
        comp = {int(dofs[3]): 0, int(dofs[4]): 1, int(dofs[5]): 2}.get(dof)
        rr = []
        for s in (+1, -1):
            c2, x2 = cps.copy(), xms.copy()
            c2[node, dof - 1] += s * h
            if comp is not None:
                x2[node, comp] += s * h * dt
            rr.append(m.calc_r_K(x2, c2, dt, dt)[0])
        fd = (rr[0] - rr[1]) / (2 * h)
        assert np.abs(fd - K[:, u]).max() <= 2e-7 * max(1.0, np.abs(K[:, u]).max())


@pytest.mark.parametrize("motion", [orc.LAG, orc.EUL, orc.ALEVB])
def test_newton_converges_quadratically(motion):
    # time_step! (FiniteElement.jl:11-63) on the first F_PULL step after the predictor
    m = orc.Mesh(motion=motion, scenario=orc.F_PULL, num1el=9, num2el=9, pull_speed=0.5)
    xms, cps = m.flat_state()
    dt = 0.5
    dofs = m.dofs
    mmo = dofs[0:3] if motion == orc.LAG else dofs[3:6]
    for j in range(3):
        xms[:, j] += dt * cps[:, mmo[j] - 1]                 # update_xms! predictor (Analysis.jl:70)
    n_inv, d_inv = m.ID_inv
    eps_hist = []
    for it in range(14):
        r, K = m.calc_r_K(xms, cps, dt, dt)
        du = -spla.spsolve(sp.csc_matrix(K), r)
        dcps = np.zeros_like(cps)
        dcps[n_inv - 1, d_inv - 1] = du
        cps += dcps
        for j in range(3):
            xms[:, j] += dt * dcps[:, mmo[j] - 1]
        eps_hist.append(np.linalg.norm(du) / m.nmdf)
        if eps_hist[-1] < 1e-12:
            break
    assert eps_hist[-1] < 1e-12 and len(eps_hist) <= 6, eps_hist
    # quadratic: each error is below C * previous^2 once in the asymptotic regime
    assert eps_hist[2] < 0.05 * eps_hist[1]


def test_static_scenarios_are_linear():
    # STATIC scenarios converge in one correction
    for scen in (orc.F_COUE, orc.F_POIS, orc.F_CAVI):
        m = orc.Mesh(motion=orc.STATIC, scenario=scen, num1el=6, num2el=6, length=4.0)
        xms, cps = m.flat_state()
        cps[:, m.dofs[6] - 1] = 0.0
        n_inv, d_inv = m.ID_inv
        hist = []
        for it in range(3):
            r, K = m.calc_r_K(xms, cps, 1.0, 1.0)
            du = -spla.spsolve(sp.csc_matrix(K), r)
            cps[n_inv - 1, d_inv - 1] += du
            hist.append(np.linalg.norm(du) / m.nmdf)
        assert hist[1] < 1e-12 * max(1.0, hist[0]), (scen, hist)
