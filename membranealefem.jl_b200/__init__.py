"""membranealefem.jl_b200 -- B200-native residual + tangent assembly (`calc_r_K`) of MembraneAleFem.jl.

Layout:  csrc/   CUDA kernels (sm_100a) and the C ABI of include/maf.h  -> libmembrane_b200.so
         capi.py ctypes binding of that ABI
         host/   Python mirror of the reference's Input/Analysis interface around the path
"""
from .capi import (PATTERN_BLK, PATTERN_SYM, SCATTER_ATOMIC, SCATTER_DETERMINISTIC, Assembler, MafError,
                   fp64_peak_tflops, host_register, host_unregister, load_library)
from .host.analysis import calc_r_K, run_analysis, time_step, update_xms
from .host.api import restart, solve
from .host.basis import (AreaGpBasisFns, BdryGpBasisFns, GaussPointsXi, GaussPointsZeta, LineGpBasisFns,
                         gp_basis_fns_1d, gp_basis_fns_2d)
from .host.enums import *  # noqa: F401,F403
from .host.input import prepare_input, synthetic_state
from .host.mesh import Mesh, get_basis_fns, get_ddN, get_dN, get_gpw, get_N
from .host import partition
from .host.pullforce import calc_pull_force, get_adj_maps, get_pull_el_id
from .host.params import Params, check_params
from .host.spline import KnotVector
