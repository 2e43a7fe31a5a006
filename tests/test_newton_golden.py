"""Committed Newton histories of BASELINE.json configs 1-3 (tests/golden/newton_17x17_*.npz: 17 x 17 patch, 32 steps of
0.5, made by tests/golden/make_newton_golden.py from the oracle): the oracle of today reproduces their first steps,
and so do the kernels' phase functions (tests/emu) inside the same Newton loop. The GPU test of all 32 steps is
tests/test_gpu_parity.py::test_newton_history_17x17_32_steps."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import mafb200 as maf
from emu_driver import Emu
from helpers import newton_history
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(name):
    z = np.load(os.path.join(HERE, "golden", f"newton_17x17_{name}.npz"))
    hist, k = [], 0
    for n in z["lens"]:
        hist.append(z["eps"][k:k + int(n)].tolist())
        k += int(n)
    return z, hist


@pytest.mark.parametrize("name,motion", [("lag", maf.LAG), ("eul", maf.EUL), ("alevb", maf.ALEVB)])
def test_golden_histories_are_complete_and_quadratic(name, motion):
    z, hist = _golden(name)
    assert len(hist) == 32 and np.allclose(z["dts"], 0.5)
    for h in hist:
        assert 2 <= len(h) <= 14 and h[-1] < 1e-12                  # FiniteElement.jl:29, :54
        for e0, e1 in zip(h[:-2], h[1:-1]):
            assert e1 < 0.2 * e0                                    # contraction before the round-off floor
    assert abs(z["xms"][:, 2].max() - 8.0) < 1e-9                   # pull_speed * sum(dts) / ... : 0.5 * 16


@pytest.mark.parametrize("name,motion", [("lag", maf.LAG), ("alevb", maf.ALEVB)])
def test_oracle_and_emulated_kernels_reproduce_the_first_steps(name, motion):
    z, hist_g = _golden(name)
    p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=17, num2el=17, output=False)
    args = dict(pull_speed=0.5, dts=[0.5] * 2, t0=0.0, t0_id=0)
    mesh, xms, cps = maf.prepare_input(p, **args)
    om = orc.Mesh(motion=int(motion), scenario=orc.F_PULL, num1el=17, num2el=17, pull_speed=0.5)
    xo, co = xms.copy(), cps.copy()
    h_o = newton_history(lambda x, c, t, dt: om.calc_r_K(x, c, t, dt, nthreads=8), motion, mesh.dofs, om.ID_inv,
                         om.nmdf, xo, co, args["dts"])
    e = Emu(mesh, p)

    def asm(x, c, t, dt):
        r, K = e.assemble(x, c, t, dt)
        K = sp.csc_matrix(K)
        K.eliminate_zeros()
        return r, K
    h_e = newton_history(asm, motion, mesh.dofs, mesh.ID_inv, mesh.nmdf, xms, cps, args["dts"])
    for h, what in ((h_o, "oracle"), (h_e, "emulated kernels")):
        assert [len(q) for q in h] == [len(q) for q in hist_g[:2]], what
        for ha, hb in zip(h, hist_g[:2]):
            for ea, eb in zip(ha, hb):
                if eb > 1e-9:
                    assert abs(ea - eb) <= 1e-6 * eb, (what, ha, hb)
    assert np.abs(xms - xo).max() <= 1e-10 and np.abs(cps - co).max() <= 1e-10
