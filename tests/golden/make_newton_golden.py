"""Newton convergence histories of BASELINE.json configs 1-3 on the reference's default mesh: 17 x 17 F_PULL patch,
length 64, pull_speed 0.5, dts = fill(0.5, 32) (docs/src/index.md:149-152, Params.jl:39-40), LAG / EUL / ALEVB.

Like tests/golden/make_golden.py these are outputs of the CPU ORACLE (the reference is a Julia package and cannot run
in this image): every eps = |du|_2 / nmdf of every Newton iterate of all 32 steps, and the final state. They freeze the
oracle-driven loop so that the GPU test (tests/test_gpu_parity.py::test_newton_history_17x17_32_steps) compares the
library both with a fresh oracle run of the first steps and with committed numbers for all 32.

    python tests/golden/make_newton_golden.py        # rewrites tests/golden/newton_17x17_*.npz (a few minutes)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import mafb200 as maf  # noqa: E402
from helpers import newton_history  # noqa: E402
from oracle import oracle as orc  # noqa: E402

NAMES = {"lag": maf.LAG, "eul": maf.EUL, "alevb": maf.ALEVB}


def main():
    for name, motion in NAMES.items():
        p = maf.Params(motion=motion, scenario=maf.F_PULL, num1el=17, num2el=17, output=False)
        args = dict(pull_speed=0.5, dts=[0.5] * 32, t0=0.0, t0_id=0)
        mesh, xms, cps = maf.prepare_input(p, **args)
        om = orc.Mesh(motion=int(motion), scenario=orc.F_PULL, num1el=17, num2el=17, pull_speed=0.5)
        hist = newton_history(lambda x, c, t, dt: om.calc_r_K(x, c, t, dt, nthreads=8), motion, mesh.dofs, om.ID_inv,
                              om.nmdf, xms, cps, args["dts"])
        lens = np.array([len(h) for h in hist], dtype=np.int64)
        flat = np.array([e for h in hist for e in h])
        np.savez_compressed(os.path.join(HERE, f"newton_17x17_{name}.npz"), lens=lens, eps=flat, xms=xms, cps=cps,
                            dts=np.array(args["dts"]), pull_speed=0.5)
        print(name, "iterations per step", lens.tolist(), "first", hist[0], "z of the pulled node", xms[:, 2].max())


if __name__ == "__main__":
    main()
