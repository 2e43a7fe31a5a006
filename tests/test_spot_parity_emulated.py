"""CPU run of the single-element spot check (tests/spot_parity.py) that the GPU suite applies to the 1001 x 1001 bench
patch: same selection and comparison code, the kernels' phase functions through tests/emu instead of the device."""
import numpy as np

import mafb200 as maf
from cases import make_case, truth_mesh
from emu_driver import Emu
from spot_parity import check_elements, select_elements


class EmuAsm:
    """The slice of capi.Assembler's interface check_elements uses, on top of the CPU emulation."""

    def __init__(self, mesh, p):
        self.e = Emu(mesh, p)
        self.mesh = mesh
        self._colptr, self._rowval = self.e.pattern()
        self.range = (1, mesh.numel)

    def colptr(self):
        return self._colptr

    def pattern_columns(self, c0, c1, colptr=None):
        return self._rowval[self._colptr[c0 - 1] - 1:self._colptr[c1] - 1]

    def set_element_range(self, a, b):
        self.range = (a, b)

    def assemble_device(self, xms, cps, time, dt):
        self.r, K = self.e.assemble(xms, cps, time, dt, el_first=self.range[0], el_last=self.range[1])
        self.nz = K.data

    def range_info(self):
        return {"rows": (1, self.mesh.nmdf), "slots": (1, self.e.nnz)}

    def download(self, r_first, r_count, nz_first, nz_count):
        return (self.r[r_first - 1:r_first - 1 + r_count].copy(), self.nz[nz_first - 1:nz_first - 1 + nz_count].copy())


def test_spot_check_on_the_fine_knot_patch():
    name = "alevb_pull_fine_19x19"
    p, hm, om, xms, cps, time, dt, args = make_case(name)
    els = select_elements(hm, n_random=20)
    assert len(els) > 100 and len(set(els)) == len(els)
    els = els[::3]        # (the GPU suite runs the whole selection on the 1001 x 1001 patch)
    worst = check_elements(EmuAsm(hm, p), hm, om, truth_mesh(name), xms, cps, xms, cps, time, dt, els)
    assert worst["n"] == len(els) and worst["K_strict"] <= 1.0 and worst["r_strict"] <= 1.0
    assert worst["K_rel_oracle"] < 1e-12
    print(worst)
