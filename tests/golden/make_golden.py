"""Generates the golden fixtures of tests/golden/*.npz.

The reference (a Julia package) cannot run in this image and ships no residual / tangent vectors, so these are NOT
reference outputs: they are outputs of the CPU oracle (oracle/maf_oracle.cpp, the restatement of the reference
algorithm that is pinned to the reference's own known answers by tests/test_oracle_reference_known_answers.py),
frozen so that (a) a drift of the oracle itself is noticed and (b) the CUDA path is also checked against committed
numbers. Inputs are regenerated from the seeds in tests/cases.py; the fixtures hold them too.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import make_case  # noqa: E402

GOLDEN_CASES = ["lag_pull_3x3", "eul_bend_3x4", "alev_bend_4x4_pn", "alevb_pull_5x4", "static_cavi_4x5",
                "alevb_bend_pn_4x3"]


def main():
    for name in GOLDEN_CASES:
        p, hm, om, xms, cps, time, dt, args = make_case(name)
        r, K = om.calc_r_K(xms, cps, time, dt)
        K = K.tocsc()
        K.sort_indices()
        out = dict(xms=xms, cps=cps, time=time, dt=dt, bend_tm=args.get("bend_tm", 1.0), r=r,
                   colptr=K.indptr.astype(np.int64) + 1, rowval=K.indices.astype(np.int64) + 1, nzval=K.data,
                   ID=om.ID, nmdf=om.nmdf)
        if name == "alevb_pull_5x4":
            pass
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "nmdf", om.nmdf, "nnz", K.nnz, "|r|", float(np.abs(r).max()))


if __name__ == "__main__":
    main()
