"""Small assemblies for compute-sanitizer (memcheck / racecheck / synccheck): every motion, both scatter modes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mafb200 as maf  # noqa: E402

for motion, n in ((maf.ALEVB, 19), (maf.LAG, 9), (maf.EUL, 7), (maf.ALEV, 7), (maf.STATIC, 5)):
    scen = maf.F_PULL if motion != maf.STATIC else maf.F_CAVI
    p = maf.Params(motion=motion, scenario=scen, num1el=n, num2el=n, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    xms, cps = maf.synthetic_state(mesh, p)
    asm = maf.Assembler(mesh, p)
    for mode in (maf.SCATTER_ATOMIC, maf.SCATTER_DETERMINISTIC):
        r, nz, rn = asm.assemble(xms, cps, 0.5, 0.5, scatter_mode=mode)
        assert np.isfinite(r).all() and np.isfinite(nz).all()
    asm.state_set(xms, cps)
    asm.state_predict(0.5)
    asm.state_update(np.zeros(mesh.nmdf), 0.5)
    if scen == maf.F_PULL and n >= 7:
        adj, maps = maf.get_adj_maps(mesh.num1el, mesh.numel, mesh.IX, p.poly)
        asm.elem_v_residuals(adj)
    asm.close()
    print("ok", int(motion), n, float(rn))
