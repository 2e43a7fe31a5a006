"""CPU-only validation of the CUDA library's logic: the kernels' phase functions (maf_element.cuh, maf_boundary.cuh,
maf_gather.cuh), the symbolic phase and the scatter maps are compiled for the host by tests/emu and compared with the
oracle. The GPU tests (test_gpu_parity.py) run the same cases through the real C ABI."""
import numpy as np
import pytest

import mafb200 as maf
from cases import SMALL, active_unknowns, check_pattern_contract, compare, make_case
from emu_driver import Emu


@pytest.mark.parametrize("name", SMALL)
def test_emulated_assembly_matches_oracle(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    e = Emu(hm, p)
    r, K = e.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0))
    assert not np.isnan(r).any() and not np.isnan(K.data).any()
    er, ek = compare(r, K, r_o, K_o, active_unknowns(om, cps))
    assert er < 1e-12 and ek < 1e-12, (er, ek)      # (the strict entrywise rule: tests/test_truth_oracle.py)
    check_pattern_contract(K, K_o, generic="flat" not in name)


@pytest.mark.parametrize("name", ["alevb_pull_5x4", "eul_pull_5x4", "lag_bend_4x3", "static_coue_4x4_pn"])
def test_emulated_deterministic_path(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    e = Emu(hm, p)
    r0, K0 = e.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0), mode=0)
    r1, K1 = e.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0), mode=1)
    assert not np.isnan(r1).any() and not np.isnan(K1.data).any()      # every slot written exactly once
    # the atomics path adds in processing (Z-curve) order, the deterministic path in ascending element id
    assert np.abs(r0 - r1).max() <= 1e-14 * max(np.abs(r1).max(), 1.0)
    assert np.abs(K0.data - K1.data).max() <= 1e-14 * np.abs(K1.data).max()
    r2, K2 = e.assemble(xms, cps, time, dt, bend_tm=args.get("bend_tm", 1.0), mode=1)
    assert np.array_equal(r1, r2) and np.array_equal(K1.data, K2.data)
    # the deterministic sum is the reference's own order (one Julia thread): compare with the oracle tightly
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    assert abs(K1 - K_o).max() <= 1e-14 * abs(K_o).max()


@pytest.mark.parametrize("name", ["alevb_pull_5x4", "lag_pull_5x4"])
def test_emulated_sym_pattern_and_ranges(name):
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    r_o, K_o = om.calc_r_K(xms, cps, time, dt)
    e = Emu(hm, p, pattern_mode=maf.PATTERN_SYM)
    # full LM x LM union
    LM = hm.LM
    keys = set()
    for el in range(hm.numel):
        act = LM[:, el][LM[:, el] != 0]
        keys.update((int(c) << 32 | int(r)) for r in act for c in act)
    assert e.nnz == len(keys)
    r, K = e.assemble(xms, cps, time, dt)
    er, ek = compare(r, K, r_o, K_o)
    assert er < 1e-12 and ek < 1e-12
    # strip partition (multi-GPU sharding): the sum over disjoint element ranges equals the whole
    cut = (hm.num2el // 2) * hm.num1el
    ra, Ka = e.assemble(xms, cps, time, dt, el_first=1, el_last=cut, mode=1)
    rb, Kb = e.assemble(xms, cps, time, dt, el_first=cut + 1, el_last=hm.numel, mode=1)
    assert np.abs(ra + rb - r).max() <= 1e-13 * np.abs(r).max()
    assert abs(Ka + Kb - K).max() <= 1e-13 * abs(K).max()


def test_pattern_is_sorted_csc_and_state_independent():
    p, hm, om, xms, cps, time, dt, args = make_case("alevb_pull_5x4")
    e = Emu(hm, p)
    colptr, rowval = e.pattern()
    assert colptr[0] == 1 and colptr[-1] == e.nnz + 1
    for c in range(hm.nmdf):
        rows = rowval[colptr[c] - 1:colptr[c + 1] - 1]
        assert np.all(np.diff(rows) > 0) and rows.min() >= 1 and rows.max() <= hm.nmdf


def test_config_validation():
    p, hm, om, xms, cps, time, dt, args = make_case("lag_pull_3x3")
    bad = maf.Params(motion=maf.ALEVB, scenario=maf.F_PULL, num1el=3, num2el=3, length=4.0, output=False)
    with pytest.raises(RuntimeError):
        Emu(hm, bad)          # ALEVB params on a mesh that carries LAG dofs
