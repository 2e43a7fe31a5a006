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


def make_case(name, seed=7):
    motion, scen, n1, n2, L, extra, args, kind = {**CASES, **DEFAULT_17, **PULL}[name]
    p = maf.Params(motion=motion, scenario=scen, num1el=n1, num2el=n2, length=L, output=False, **extra)
    hm = maf.Mesh(p, **args)
    om = orc.Mesh(motion=int(motion), scenario=int(scen), num1el=n1, num2el=n2, length=L, pn=p.pn,
                  pull_speed=args.get("pull_speed", 0.0), bend_mf=args.get("bend_mf", 0.0),
                  bend_tm=args.get("bend_tm", 1.0))
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


def entrywise_rel_error(K, K_o, rtol=1e-11, ulps=16):
    """Entrywise criterion on EVERY entry:  |K_ij - K_o_ij| <= rtol |K_o_ij| + ulps eps max|K_o|,  returned as the
    largest |K_ij - K_o_ij| / (|K_o_ij| + ulps eps max|K_o| / rtol), to be compared with rtol.

    The absolute term is a few units in the last place of the largest entry: an element sum of terms of size
    max|K| cannot be reproduced more accurately than that by ANY other summation order (the reference's own value
    of such an entry moves by the same amount with its Julia thread count). Measured: the largest absolute
    difference to the oracle is 4-6 eps max|K|; entries of size 1e-5 max|K| therefore differ by ~1e-11 relative."""
    K, K_o = K.tocsc(), K_o.tocsc()
    D = abs(K - K_o).tocoo()
    if D.nnz == 0:
        return 0.0
    ref = np.abs(np.asarray(K_o[D.row, D.col]).ravel())
    floor = ulps * np.finfo(float).eps * abs(K_o).max() / rtol
    return float((D.data / (ref + floor)).max())


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
