"""Shared helpers for tests (oracle-side state builders and comparison rules)."""
import numpy as np

from oracle import oracle as orc


def deformed_state(m, seed=0, zamp=0.05, noise=0.1):
    """Flat patch + smooth out-of-plane bump + in-plane shear + seeded noise on every dof column,
    so that no tangent entry is a trivial zero. Works on oracle Mesh or host Mesh (needs flat_state())."""
    xms, cps = m.flat_state()
    xms = np.array(xms, order="F")
    cps = np.array(cps, order="F")
    L = xms[:, 0].max() - xms[:, 0].min()
    x, y = xms[:, 0].copy(), xms[:, 1].copy()
    xms[:, 2] = zamp * L * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L + 0.3)
    xms[:, 0] = x + 0.02 * y + 0.01 * L * np.sin(np.pi * y / L)
    xms[:, 1] = y - 0.015 * x
    rng = np.random.default_rng(seed)
    cps += noise * rng.uniform(-1, 1, size=cps.shape)
    return np.asfortranarray(xms), np.asfortranarray(cps)


def tol_compare(a, b, rtol=1e-11, scale=None):
    """|a-b| <= rtol * max(|a|,|b|) + rtol * scale   (scale defaults to max|b|; see DESIGN.md 'tolerance')."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if scale is None:
        scale = np.max(np.abs(b)) if b.size else 0.0
    err = np.abs(a - b)
    bound = rtol * np.maximum(np.abs(a), np.abs(b)) + rtol * scale
    worst = np.max(err - bound) if err.size else -1.0
    return worst <= 0.0, float(np.max(err / (np.maximum(np.abs(a), np.abs(b)) + scale)) if err.size else 0.0)


def P_of(m):
    p = m.params
    return {"kb": p.kb, "kg": p.kg, "zv": p.zv, "pn": p.pn, "adb": p.adb, "am": p.am, "bend_tm": p.bend_tm}


MOTION_NAMES = {orc.STATIC: "STATIC", orc.EUL: "EUL", orc.LAG: "LAG", orc.ALEV: "ALEV", orc.ALEVB: "ALEVB"}


def newton_history(assemble, motion, dofs, ID_inv, nmdf, xms, cps, dts, t0=0.0, enr=1e-12, solve=None):
    """time_step! inside run_analysis! (FiniteElement.jl:11-63, Analysis.jl:60-93) around an arbitrary assembly:
    `assemble(xms, cps, time, dt) -> (r, K)` (K scipy sparse). Mutates xms / cps; returns the eps = |du|_2 / nmdf of
    every Newton iterate of every step."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import mafb200 as maf
    n_inv, d_inv = ID_inv
    hist, t = [], t0
    for dt in dts:
        t += dt
        maf.update_xms(motion, xms, cps, dt, dofs)                  # predictor, Analysis.jl:70
        eps = []
        for it in range(14):                                        # iter < 15, FiniteElement.jl:29
            r, K = assemble(xms, cps, t, dt)
            du = -(solve(K, r) if solve else spla.splu(sp.csc_matrix(K)).solve(r))
            dc = np.zeros_like(cps)
            dc[n_inv - 1, d_inv - 1] = du
            cps += dc
            maf.update_xms(motion, xms, dc, dt, dofs)
            eps.append(float(np.linalg.norm(du) / nmdf))
            if eps[-1] < enr:
                break
        assert eps[-1] < enr, ("did not reach Newton-Raphson tolerance", t, eps)
        hist.append(eps)
    return hist
