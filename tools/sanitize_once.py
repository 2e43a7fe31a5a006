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
    # round 2: strips behind the C ABI (flag kernels, peer pull-add, sliced tables), banded deterministic staging
    if n >= 7:
        strips = [maf.Assembler(mesh, p, strip=(k, 2)) for k in range(2)]
        strips[0].peer_attach_local(None, strips[1])
        strips[1].peer_attach_local(strips[0], None)
        for mode in (maf.SCATTER_ATOMIC, maf.SCATTER_DETERMINISTIC):
            for s_ in strips:
                s_.state_set(xms, cps)
            for s_ in strips:
                s_.assemble_strip(None, None, 0.5, 0.5, scatter_mode=mode)
            for s_ in strips:
                s_.sync()
        own = [s_.strip_info()["own_slots"] for s_ in strips]
        nz2 = np.concatenate([s_.download(1, 0, o[0], o[1] - o[0] + 1)[1] for s_, o in zip(strips, own)])
        assert np.abs(nz2 - nz).max() <= 1e-12 * np.abs(nz).max()
        for s_ in strips:
            s_.close()
        os.environ["MAF_BAND_ROWS"] = "2"
        os.environ["MAF_NO_GRAPH"] = "1"
        band = maf.Assembler(mesh, p)
        rb, nzb, _ = band.assemble(xms, cps, 0.5, 0.5, scatter_mode=maf.SCATTER_DETERMINISTIC)
        assert np.array_equal(nzb, nz)
        band.close()
        del os.environ["MAF_BAND_ROWS"]
        # banded staging inside a captured graph (gathers on the second stream), sub-strips with the pipelined copy-out
        del os.environ["MAF_NO_GRAPH"]
        os.environ["MAF_BAND_ROWS"] = "2"
        band = maf.Assembler(mesh, p)
        for _ in range(2):
            rb, nzb, _ = band.assemble(xms, cps, 0.5, 0.5, scatter_mode=maf.SCATTER_DETERMINISTIC)
        assert np.array_equal(nzb, nz)
        band.close()
        del os.environ["MAF_BAND_ROWS"]
        if n >= 17:
            os.environ["MAF_PIPELINE_MIN_ELEMS"] = "1"
            pipe = maf.Assembler(mesh, p)
            rp, nzp, _ = pipe.assemble(xms, cps, 0.5, 0.5)
            assert np.abs(nzp - nz).max() <= 1e-12 * np.abs(nz).max()
            pipe.close()
            del os.environ["MAF_PIPELINE_MIN_ELEMS"]
    print("ok", int(motion), n, float(rn))
