"""Small driver for ncu: a few device-resident assemblies of an n x n synthetic patch."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mafb200 as maf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=301)
ap.add_argument("--motion", default="ALEVB")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--scatter", type=int, default=0)
a = ap.parse_args()
import torch  # noqa: E402

p = maf.Params(motion=getattr(maf, a.motion), scenario=maf.F_PULL, num1el=a.n, num2el=a.n, output=False)
mesh = maf.Mesh(p, pull_speed=0.5)
xms, cps = maf.synthetic_state(mesh, p)
asm = maf.Assembler(mesh, p, device=0)
dx = torch.from_numpy(np.ascontiguousarray(xms.T)).cuda()
dc = torch.from_numpy(np.ascontiguousarray(cps.T)).cuda()
for _ in range(a.reps):
    asm.assemble_device(dx.data_ptr(), dc.data_ptr(), 0.5, 0.5, scatter_mode=a.scatter)
    asm.sync()
    print(asm.timings())
print(asm.kernel_info(), "numel", mesh.numel, "nnz", asm.nnz)

# profiling builds only (-DMAF_PHASE_TIMING): where the warps of a CTA spend their cycles
import ctypes  # noqa: E402
L = maf.pkg.capi.load_library()
if hasattr(L, "maf_debug_phase_cycles"):
    buf = (ctypes.c_ulonglong * 96)()
    if L.maf_debug_phase_cycles(buf, 96) == 0:
        names = ["wait0", "interp", "wait1", "gauss", "wait2", "gather", "res+tan", "-"]
        tot = sum(buf[q] for q in range(8))
        for w in range(4):
            print("warp", w, " ".join(f"{n}={100.0 * buf[8 * w + q] / tot:5.1f}%" for q, n in enumerate(names[:7])))
        print("chunk cycles (% of one warp's total):", [round(100.0 * buf[32 + q] / tot, 1) for q in range(48) if buf[32 + q]])
