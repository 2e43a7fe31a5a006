"""The extended-precision truth (oracle/libmaf_truth.so = the oracle's own source in long double, libmaf_truthq.so in
__float128) and the strict entrywise parity rule built on it (tests/cases.py::strict_errors):

    |x_ij - truth_ij| <= 1e-11 |truth_ij| + eps E_ij,   E_ij = running first-order rounding-error bound of the reference
                                                          algorithm for entry ij (units of eps)

CPU part: the truth is self-consistent (long double vs __float128), the double oracle sits a few eps M from it, and
the kernels' own code (tests/emu) meets the strict rule on every entry of every case. The GPU part of the same rule is
in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from cases import EPS, EXTRA_NEUMANN, SMALL, check_strict, compare, make_case, strict_errors, truth_mesh
from emu_driver import Emu


def _pattern0(K):
    K = K.tocsc()
    K.sort_indices()
    return K.indptr.astype(np.int64), K.indices.astype(np.int64)


@pytest.mark.parametrize("name", ["alevb_pull_5x4", "lag_bend_4x3"])
def test_long_double_truth_agrees_with_float128(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    e = Emu(hm, p)
    colptr, rowval = e.pattern()
    out = {}
    for kind in ("truth", "truthq"):
        out[kind] = truth_mesh(name, kind).calc_r_K_on_pattern(xms, cps, time, dt, colptr - 1, rowval - 1,
                                                               nthreads=8)
    r1, k1, rm1, km1 = out["truth"]
    r2, k2, rm2, km2 = out["truthq"]
    # both are rounded to double at the end (one unit in the last place of the entry itself); the long-double sum
    # carries 11 more bits than double, i.e. it is exact to ~eps M / 200 (each term is itself ~100 roundings) -- far below the rule's floor
    assert np.all(np.abs(k1 - k2) <= EPS * np.abs(k2) + 1e-2 * EPS * km2)
    assert np.all(np.abs(r1 - r2) <= EPS * np.abs(r2) + 1e-2 * EPS * rm2)
    assert np.all(np.abs(km1 - km2) <= 4 * EPS * km2) and np.all(np.abs(rm1 - rm2) <= 4 * EPS * rm2)
    assert np.all(km2 >= np.abs(k2) * (1 - 1e-12))      # a sum of magnitudes bounds the sum


@pytest.mark.parametrize("name", SMALL)
def test_oracle_and_emulated_kernels_against_truth(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    e = Emu(hm, p)
    colptr, rowval = e.pattern()
    r, K = e.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0))
    uk, ur, truth = check_strict(name, r, K.data, colptr, rowval, xms, cps, time, dt, "emulated kernels")
    # the double oracle on the same pattern: its own distance from the truth, same units
    r_o, nz_o = om.calc_r_K_fast(xms, cps, time, dt, colptr - 1, rowval - 1)
    r_t, nz_t, r_m, nz_m = truth
    qo, uo = strict_errors(nz_o, nz_t, nz_m)
    assert qo <= 1.0, (name, "oracle", qo, uo)
    nzt = np.abs(nz_t) > 0
    cond = EPS * nz_m[nzt] / np.abs(nz_t[nzt])
    print(f"{name:24s} |K - truth| / (eps E): oracle {uo:5.2f}  kernels {uk:5.2f}   r: kernels {ur:5.2f}   "
          f"eps E / |K|: median {np.median(cond):.1e}, above 1e-11 on {100.0 * np.mean(cond > 1e-11):.1f} % of the entries")


def test_truth_of_neumann_boundary_elements():
    """Element level, every Neumann type incl. the SHEAR and TOP/BOTTOM MOMENT branches (FiniteElement.jl:374-380)."""
    name = "lag_bend_4x3"
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    ot = truth_mesh(name)
    for bdry in (1, 2, 3, 4):
        el = int(om.bdry_elems(bdry)[1])
        for ntype in (1, 2, 3):
            r, K, _, _ = om.elem_r_K_mag(el, xms, cps, time, dt, bdry=bdry, ntype=ntype, nval=0.7)
            r_t, K_t, r_m, K_m = ot.elem_r_K_mag(el, xms, cps, time, dt, bdry=bdry, ntype=ntype, nval=0.7)
            qk, uk = strict_errors(K, K_t, K_m)
            qr, ur = strict_errors(r, r_t, r_m)
            assert qk <= 1.0 and qr <= 1.0, (bdry, ntype, qk, uk, qr, ur)
            assert np.abs(K_t).max() > 0


@pytest.mark.parametrize("name", ["lag_bend_4x3", "alevb_bend_pn_4x3", "eul_bend_3x4"])
def test_shear_and_top_bottom_moment_through_the_kernels(name):
    """The same branches through the library's boundary kernel code (CPU emulation; GPU: test_gpu_parity.py): the
    conditions are injected into mesh.inh_neu_bcs, i.e. they reach the kernels through maf_mesh_desc.neu_*."""
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    hm.inh_neu_bcs = list(EXTRA_NEUMANN)
    om.set_neumann(EXTRA_NEUMANN)
    e = Emu(hm, p)
    colptr, rowval = e.pattern()
    r, K = e.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0))
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    er, ek = compare(r, K, r_o, K_o)
    assert er < 1e-12 and ek < 1e-12, (er, ek)
    # the boundary terms are a visible part of the result (not lost in the area terms)
    om.set_neumann([])
    r_0, K_0 = om.calc_r_K(xms, cps, time, dt)
    assert np.abs(r_o - r_0).max() > 1e-3 * np.abs(r_o).max() and abs(K_o - K_0).max() > 1e-4 * abs(K_o).max()
    check_strict(name, r, K.data, colptr, rowval, xms, cps, time, dt, "emulated kernels", neumann=EXTRA_NEUMANN)
