"""`prepare_input` of the reference (src/Input.jl:22-121) and the synthetic benchmark state of SURVEY.md 8(d)5."""
import math
import os

import numpy as np

from .enums import EUL, F_BEND, F_PULL, LAG, Dof
from .mesh import Mesh
from .params import get_dts
from .spline import get_2d_bspline_cps

U = Dof.Unknown


def prepare_input(p, **args):
    """Build the mesh, the flat initial positions `xms` (numnp x 3) and the unknowns `cps` (numnp x ndf).

    Control points come from the tensor-product form of the reference's 2-D collocation fit (see
    spline.get_2d_bspline_cps); restart files are read like Input.jl:35-37.
    """
    mesh = Mesh(p, **args)
    ks = args.keys()
    if "in_path" in ks and "in_xms" in ks and "in_cps" in ks:
        xms = np.asfortranarray(np.loadtxt(os.path.join(args["in_path"], args["in_xms"]), ndmin=2))
        cps = np.asfortranarray(np.loadtxt(os.path.join(args["in_path"], args["in_cps"]), ndmin=2))
    else:
        xms = np.zeros((mesh.numnp, 3), order="F")
        cps = np.zeros((mesh.numnp, mesh.ndf), order="F")
        L = p.length
        if p.scenario == F_BEND:                                              # Input.jl:45-73
            xms[:, 0] = get_2d_bspline_cps(mesh.kv1, mesh.kv2, lambda z1, z2: L * z1 + 0 * z2)
            xms[:, 1] = get_2d_bspline_cps(mesh.kv1, mesh.kv2, lambda z1, z2: L * z2 + 0 * z1)
            rng = np.random.default_rng(args.get("seed", None))
            d = mesh.dofs
            cps[:, d[U.vx] - 1] = (rng.random(mesh.numnp) - 0.5) / mesh.numel
            cps[:, d[U.vy] - 1] = (rng.random(mesh.numnp) - 0.5) / mesh.numel
            cps[:, d[U.vz] - 1] = rng.random(mesh.numnp) / mesh.numel
            if p.motion != LAG:
                cps[:, d[U.vmx] - 1] = (rng.random(mesh.numnp) - 0.5) / mesh.numel
                cps[:, d[U.vmy] - 1] = (rng.random(mesh.numnp) - 0.5) / mesh.numel
                cps[:, d[U.vmz] - 1] = rng.random(mesh.numnp) / mesh.numel
            for b in mesh.bdry_nodes:
                cps[mesh.bdry_nodes[b] - 1, :] = 0.0
            lval = p.kb / 4 / p.length ** 2 * (get_dts(args)[0] / args["bend_tm"]) ** 2
            cps[:, d[U.lam] - 1] = lval
        else:                                                                 # Input.jl:77-92
            xms[:, 0] = get_2d_bspline_cps(mesh.kv1, mesh.kv2, lambda z1, z2: L * (z1 - 0.5) + 0 * z2)
            xms[:, 1] = get_2d_bspline_cps(mesh.kv1, mesh.kv2, lambda z1, z2: L * (z2 - 0.5) + 0 * z1)
            cps[:, mesh.dofs[U.lam] - 1] = p.kb / 4
    for (unknown, node, value) in mesh.inh_dir_bcs:                           # Input.jl:103-108
        cps[node - 1, mesh.dofs[unknown] - 1] = value
        if p.scenario == F_PULL and p.motion == EUL:
            cps[node - 1, mesh.dofs[U.vmz] - 1] = value
    if p.output and "out_path" in ks:                                         # Input.jl:111-114
        np.savetxt(os.path.join(args["out_path"], f"t{args['t0_id']}-xms.txt"), xms, delimiter="\t")
        np.savetxt(os.path.join(args["out_path"], f"t{args['t0_id']}-cps.txt"), cps, delimiter="\t")
    return mesh, xms, cps


def _hash01(node, dof, seed):
    """Counter-based uniform(0,1): splitmix64 of (seed, node, dof). Reproducible on any rank without a stream."""
    x = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (node.astype(np.uint64) * np.uint64(16) + np.uint64(dof)))
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def synthetic_state(mesh, p, seed=20240117, zamp=0.05, noise=0.1, node_offset=0):
    """Deterministic perturbed state for assembly-only sweeps: flat patch plus
    z = A sin(2 pi x / L) sin(2 pi y / L), A = zamp * L, fitted by tensor-product collocation; every dof column gets
    noise * U(-1, 1) from a counter-based generator keyed by (global node id, dof); lambda = kb/4 + the same noise."""
    L = p.length
    xms = np.zeros((mesh.numnp, 3), order="F")
    with np.errstate(over="ignore"):
        xms[:, 0] = get_2d_bspline_cps(mesh.kv1, mesh.kv2, lambda z1, z2: L * (z1 - 0.5) + 0 * z2)
        xms[:, 1] = get_2d_bspline_cps(mesh.kv1, mesh.kv2, lambda z1, z2: L * (z2 - 0.5) + 0 * z1)
        xms[:, 2] = get_2d_bspline_cps(
            mesh.kv1, mesh.kv2, lambda z1, z2: zamp * L * np.sin(2 * math.pi * z1) * np.sin(2 * math.pi * z2))
        cps = np.zeros((mesh.numnp, mesh.ndf), order="F")
        nodes = np.arange(mesh.numnp, dtype=np.int64) + node_offset
        for d in range(mesh.ndf):
            cps[:, d] = noise * (2.0 * _hash01(nodes, d, seed) - 1.0)
    cps[:, mesh.dofs[U.lam] - 1] += p.kb / 4
    return xms, cps
