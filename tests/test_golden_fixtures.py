"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle; the
reference itself ships none for this path and cannot run here -- see the script's header):

  CPU : the oracle of today reproduces them (drift guard), the seeded inputs regenerate bit for bit, and the
        emulated kernel logic matches them;
  GPU : the CUDA path through the C ABI matches them to the stated 1e-11."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

import mafb200 as maf
from cases import compare, make_case
from emu_driver import Emu

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                  if not os.path.basename(f).startswith("newton_"))   # (Newton histories: tests/test_newton_golden.py)
RTOL = 1e-11


def _load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    n = int(z["nmdf"])
    K = sp.csc_matrix((z["nzval"], z["rowval"] - 1, z["colptr"] - 1), shape=(n, n))
    return z, K


def test_fixtures_exist():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_golden(name):
    z, K_g = _load(name)
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    assert np.array_equal(xms, z["xms"]) and np.array_equal(cps, z["cps"])      # seeded inputs are reproducible
    assert np.array_equal(om.ID, z["ID"]) and om.nmdf == int(z["nmdf"])         # DOF numbering: exact
    r, K = om.calc_r_K(xms, cps, time, dt)
    K = K.tocsc()
    K.sort_indices()
    assert np.array_equal(K.indptr + 1, z["colptr"]) and np.array_equal(K.indices + 1, z["rowval"])   # pattern: exact
    assert np.abs(r - z["r"]).max() <= 1e-13 * max(np.abs(z["r"]).max(), 1e-300)
    assert np.abs(K.data - z["nzval"]).max() <= 1e-13 * np.abs(z["nzval"]).max()


@pytest.mark.parametrize("name", FIXTURES)
def test_emulated_kernels_match_golden(name):
    z, K_g = _load(name)
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    r, K = Emu(hm, p).assemble(z["xms"], z["cps"], float(z["time"]), float(z["dt"]), bend_tm=float(z["bend_tm"]))
    er, ek = compare(r, K, z["r"], K_g)
    assert er < RTOL and ek < RTOL, (name, er, ek)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_matches_golden(name):
    z, K_g = _load(name)
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    asm = maf.Assembler(hm, p)
    colptr, rowval = asm.pattern()
    for mode in (maf.SCATTER_ATOMIC, maf.SCATTER_DETERMINISTIC):
        r, nz, _ = asm.assemble(z["xms"], z["cps"], float(z["time"]), float(z["dt"]), bend_tm=float(z["bend_tm"]),
                                scatter_mode=mode)
        K = sp.csc_matrix((nz, rowval - 1, colptr - 1), shape=(asm.nmdf, asm.nmdf))
        er, ek = compare(r, K, z["r"], K_g)
        assert er < RTOL and ek < RTOL, (name, mode, er, ek)
        # the golden (reference-semantics) pattern is contained in the library's symbolic pattern
        Kb, Gb = K.copy(), K_g.copy()
        Kb.data[:] = 1.0
        Gb.data[:] = 1.0
        assert (Gb - Gb.multiply(Kb)).nnz == 0
    asm.close()
