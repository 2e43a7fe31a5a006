"""Parity cases shared by the emulated (CPU) and the GPU tests: every motion, every scenario, uniform and
centre-refined knots, normal pressure, flat / deformed states, edge-sized meshes."""
import numpy as np

import mafb200 as maf
from helpers import deformed_state
from oracle import oracle as orc

# name -> (motion, scenario, num1el, num2el, length, extra Params, args, state kind)
CASES = {
    "lag_pull_5x4": (maf.LAG, maf.F_PULL, 5, 4, 8.0, {}, {"pull_speed": 0.5}, "deformed"),
    "eul_pull_5x4": (maf.EUL, maf.F_PULL, 5, 4, 8.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alev_pull_5x4": (maf.ALEV, maf.F_PULL, 5, 4, 8.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alevb_pull_5x4": (maf.ALEVB, maf.F_PULL, 5, 4, 8.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alevb_pull_flat_7x7": (maf.ALEVB, maf.F_PULL, 7, 7, 64.0, {}, {"pull_speed": 0.5}, "flat"),
    "lag_pull_flat_7x7": (maf.LAG, maf.F_PULL, 7, 7, 64.0, {}, {"pull_speed": 0.5}, "flat"),
    "alevb_pull_fine_19x18": (maf.ALEVB, maf.F_PULL, 19, 18, 16.0, {}, {"pull_speed": 0.5}, "deformed"),
    "lag_pull_3x3": (maf.LAG, maf.F_PULL, 3, 3, 4.0, {}, {"pull_speed": 0.25}, "deformed"),
    "static_coue_4x4_pn": (maf.STATIC, maf.F_COUE, 4, 4, 4.0, {"pn": 0.7}, {}, "deformed"),
    "static_pois_5x3": (maf.STATIC, maf.F_POIS, 5, 3, 4.0, {}, {}, "deformed"),
    "static_cavi_4x5": (maf.STATIC, maf.F_CAVI, 4, 5, 4.0, {}, {}, "deformed"),
    "lag_bend_4x3": (maf.LAG, maf.F_BEND, 4, 3, 1.0, {}, {"bend_tm": 2.0, "bend_mf": 0.5}, "deformed"),
    "eul_bend_3x4": (maf.EUL, maf.F_BEND, 3, 4, 1.0, {}, {"bend_tm": 2.0, "bend_mf": 0.5}, "deformed"),
    "alev_bend_4x4_pn": (maf.ALEV, maf.F_BEND, 4, 4, 1.0, {"pn": 0.3}, {"bend_tm": 0.4, "bend_mf": 0.5}, "deformed"),
    "alevb_bend_pn_4x3": (maf.ALEVB, maf.F_BEND, 4, 3, 1.0, {"pn": 0.25}, {"bend_tm": 0.4, "bend_mf": 0.5}, "deformed"),
    "lag_bend_pn_3x3": (maf.LAG, maf.F_BEND, 3, 3, 1.0, {"pn": -0.4}, {"bend_tm": 2.0, "bend_mf": 0.5}, "deformed"),
}
SMALL = list(CASES)

DEFAULT_17 = {
    "lag_pull_17x17": (maf.LAG, maf.F_PULL, 17, 17, 64.0, {}, {"pull_speed": 0.5}, "deformed"),
    "eul_pull_17x17": (maf.EUL, maf.F_PULL, 17, 17, 64.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alev_pull_17x17": (maf.ALEV, maf.F_PULL, 17, 17, 64.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alevb_pull_17x17": (maf.ALEVB, maf.F_PULL, 17, 17, 64.0, {}, {"pull_speed": 0.5}, "deformed"),
}


# meshes with two rings of elements around the pulled one (calc_pull_force, PullForce.jl:36-37)
PULL = {
    "lag_pull_7x7": (maf.LAG, maf.F_PULL, 7, 7, 16.0, {}, {"pull_speed": 0.5}, "deformed"),
    "eul_pull_7x5": (maf.EUL, maf.F_PULL, 7, 5, 16.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alevb_pull_7x7": (maf.ALEVB, maf.F_PULL, 7, 7, 16.0, {}, {"pull_speed": 0.5}, "deformed"),
    "alevb_pull_fine_19x19": (maf.ALEVB, maf.F_PULL, 19, 19, 16.0, {}, {"pull_speed": 0.5}, "deformed"),
}


def oracle_kwargs(name):
    motion, scen, n1, n2, L, extra, args, kind = {**CASES, **DEFAULT_17, **PULL}[name]
    return dict(motion=int(motion), scenario=int(scen), num1el=n1, num2el=n2, length=L, pn=extra.get("pn", 0.0),
                pull_speed=args.get("pull_speed", 0.0), bend_mf=args.get("bend_mf", 0.0),
                bend_tm=args.get("bend_tm", 1.0))


def truth_mesh(name, kind="truth"):
    """The oracle's own source evaluated in extended precision (oracle/Makefile: libmaf_truth.so, long double;
    "truthq": __float128) -- same inputs, same algorithm, rounded to double at the very end."""
    return orc.Mesh(kind=kind, **oracle_kwargs(name))


def make_case(name, seed=7):
    motion, scen, n1, n2, L, extra, args, kind = {**CASES, **DEFAULT_17, **PULL}[name]
    p = maf.Params(motion=motion, scenario=scen, num1el=n1, num2el=n2, length=L, output=False, **extra)
    hm = maf.Mesh(p, **args)
    om = orc.Mesh(**oracle_kwargs(name))
    if kind == "flat":
        xms, cps = om.flat_state()
    else:
        xms, cps = deformed_state(om, seed=seed)
    time, dt = 0.3, 0.37
    return p, hm, om, np.asfortranarray(xms), np.asfortranarray(cps), time, dt, args


def active_unknowns(om, cps):
    """Vector of the active unknowns (cps gathered through ID_inv, FiniteElement.jl:41-42)."""
    n_inv, d_inv = om.ID_inv
    return np.asarray(cps)[n_inv - 1, d_inv - 1]


def compare(r, K, r_o, K_o, u=None):
    """The parity rule (DESIGN.md 'tolerance'): errors are measured relative to the magnitude of what is summed,
    not of what survives cancellation:
      K: max |K - K_o| / max |K_o|
      r: max |r - r_o| / max(|r_o|_inf, | |K_o| |u| |_inf)   (the residual of an equilibrium state is ~1e-15)
    Returns (r error, K error); the tests require both < 1e-11 (typically they are ~1e-15)."""
    sr = np.abs(r_o).max()
    if u is not None:
        sr = max(sr, (abs(K_o) @ np.abs(u)).max())
    er = np.abs(r - r_o).max() / max(sr, 1e-300)
    D = (K - K_o).tocsc()
    sk = abs(K_o).max()
    ek = abs(D).max() / sk if D.nnz else 0.0
    return er, ek


# ---- the strict entrywise rule, measured against the extended-precision truth ------------------------------------
# north_star: "residual and tangent entries within 1e-11 relative". An entry is a sum over elements and Gauss points
# of terms that cancel (entries 1e-5 of the largest one are residues of terms 1e4 times their size; on the flat patch
# whole blocks vanish analytically), so no double evaluation -- the reference's included -- can hold 1e-11 relative
# to what SURVIVES the cancellation. The rule therefore adds, entry by entry, the rounding error the REFERENCE
# ALGORITHM ITSELF may commit in double precision:
#       |x_ij - truth_ij|  <=  1e-11 |truth_ij|  +  E_MULT * eps * E_ij
# truth = the oracle's own source evaluated in long double / __float128 (oracle/Makefile), E_ij = the first-order
# running error bound of that same operation sequence in units of eps, propagated through every +,-,*,/,sqrt of the
# evaluation (oracle/maf_oracle.cpp, struct cd of the ORC_TRUTH builds) -- i.e. "sum of |terms|" with the
# cancellations inside the terms counted too. Nothing global enters: an entry 1e-9 of max|K| whose terms are that
# small is held to 1e-9-sized errors. Measured (tests/test_truth_oracle.py prints the table): the double oracle sits
# at 0.01-0.07 eps E, the kernels at 0.01-0.10 eps E (E is a worst-case bound, real errors add up like a random
# walk); eps E / |x| has a median of ~1e-14, so for a well-conditioned entry the floor is three orders of magnitude
# tighter than the 1e-11 term.
EPS = float(np.finfo(float).eps)
E_MULT = 1.0


def strict_errors(x, x_t, mag, rtol=1e-11, mult=E_MULT):
    """(worst |x - x_t| / (rtol |x_t| + mult eps E),  worst |x - x_t| / (eps E)) over all entries; an entry whose
    evaluation involves no rounding at all (E = 0: every term an exact zero) must be reproduced exactly."""
    x, x_t, mag = np.asarray(x, float), np.asarray(x_t, float), np.asarray(mag, float)
    d = np.abs(x - x_t)
    bound = rtol * np.abs(x_t) + mult * EPS * mag
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(d == 0.0, 0.0, d / bound)
        in_ulps = np.where(d == 0.0, 0.0, d / (EPS * mag))
    return float(ratio.max()) if ratio.size else 0.0, float(in_ulps.max()) if in_ulps.size else 0.0


def truth_on_pattern(name, xms, cps, time, dt, colptr, rowval, kind="truth", nthreads=8, neumann=None):
    """(r, nzval, E_r, E_nz) of the truth on the library's 1-based CSC pattern (colptr, rowval)."""
    ot = truth_mesh(name, kind)
    if neumann is not None:
        ot.set_neumann(neumann)
    return ot.calc_r_K_on_pattern(xms, cps, time, dt, np.asarray(colptr) - 1, np.asarray(rowval) - 1,
                                  nthreads=nthreads)


# Neumann conditions that no scenario of the reference's Bc.jl sets up but calc_bdry_element_residual implements
# (FiniteElement.jl:374-380): SHEAR on every side, MOMENT on TOP / BOTTOM. Injected through mesh.inh_neu_bcs (the
# library reads them from maf_mesh_desc.neu_*) and orc.Mesh.set_neumann. (boundary, type, value); F_BEND meshes only
# (MOMENT asserts otherwise).
EXTRA_NEUMANN = [(1, 1, 0.7), (3, 3, 0.3), (4, 1, -0.4), (2, 2, 0.5), (1, 3, 0.2), (3, 1, 0.6), (2, 1, 0.25),
                 (4, 3, -0.15)]


def check_strict(name, r, nz, colptr, rowval, xms, cps, time, dt, what="", truth=None, neumann=None):
    """Assert the strict rule for r and every entry of nzval; returns the two errors in units of eps E (and the
    truth tuple, to be passed back in when several results are checked against the same state)."""
    if truth is None:
        truth = truth_on_pattern(name, xms, cps, time, dt, colptr, rowval, neumann=neumann)
    r_t, nz_t, r_m, nz_m = truth
    qk, uk = strict_errors(nz, nz_t, nz_m)
    qr, ur = strict_errors(r, r_t, r_m)
    assert qk <= 1.0 and qr <= 1.0, (name, what, "K", qk, uk, "r", qr, ur)
    return uk, ur, truth


def pattern_of(K):
    K = K.tocsc()
    K.sort_indices()
    return K.indptr.copy(), K.indices.copy()


def check_pattern_contract(K_full, K_o, generic):
    """Sparsity contract (hard part 1; DESIGN.md 'pattern'):
      1. every entry Julia stores (oracle, insertion semantics) is a slot of the library's symbolic pattern;
      2. on a generic (deformed) state the two patterns are IDENTICAL once exact zeros are dropped;
      3. on special states (flat patch) they may differ only in entries that are round-off residues
         (<= 1e-13 of the largest entry) of quantities that vanish analytically."""
    import scipy.sparse as sp
    ours_struct = sp.csc_matrix((np.ones_like(K_full.data), K_full.indices, K_full.indptr), shape=K_full.shape)
    Ko = K_o.tocsc()
    ref_struct = sp.csc_matrix((np.ones_like(Ko.data), Ko.indices, Ko.indptr), shape=Ko.shape)
    outside = ref_struct - ref_struct.multiply(ours_struct)
    outside.eliminate_zeros()
    assert outside.nnz == 0, "reference stores an entry outside the library's symbolic pattern"
    Kd, Kr = K_full.copy(), Ko.copy()
    Kd.eliminate_zeros()
    Kr.eliminate_zeros()
    pd, pr = pattern_of(Kd), pattern_of(Kr)
    same = np.array_equal(pd[0], pr[0]) and np.array_equal(pd[1], pr[1])
    if generic:
        assert same, "patterns differ on a generic state"
        return
    if not same:
        sd = (Kd != 0).astype(np.int8) - (Kr != 0).astype(np.int8)
        sd = sd.tocoo()
        scale = abs(Ko).max()
        for i, j in zip(sd.row, sd.col):
            assert abs(K_full[i, j]) <= 1e-13 * scale and abs(Ko[i, j]) <= 1e-13 * scale
