"""Analysis layer around the hot path, mirroring src/analysis/FiniteElement.jl and src/Analysis.jl:

  calc_r_K     FiniteElement.jl:75-200  -> the C ABI / CUDA kernels (this is the drop-in)
  time_step    FiniteElement.jl:11-63   Newton-Raphson loop, host sparse solve; resident=True keeps xms / cps on the
                                        device between the iterations (maf_state_*)
  update_xms   FiniteElement.jl:408-423
  run_analysis Analysis.jl:17-102
"""
import os
import time as _time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from ..capi import PATTERN_BLK, SCATTER_ATOMIC, Assembler
from .enums import F_PULL
from .mesh import get_m_motion_order, get_v_order
from .params import get_dts
from .pullforce import calc_pull_force, get_adj_maps, get_pull_el_id


def _assembler(mesh, p, args):
    """One library handle per (parameters, pattern mode, device), kept on the mesh. Keyed by the VALUE of the frozen
    Params (an `id` is reused after garbage collection; the handle bakes in kb, kg, zv, pn, adb, am and the motion)."""
    key = (p, args.get("pattern_mode", PATTERN_BLK), args.get("device", -1))
    cache = mesh.__dict__.setdefault("_assemblers", {})
    if key not in cache:
        cache[key] = Assembler(mesh, p, pattern_mode=args.get("pattern_mode", PATTERN_BLK),
                               device=args.get("device", -1))
    return cache[key]


def close_assemblers(mesh):
    """Release every cached handle of this mesh (each holds the device copies of the tables, r and nzval)."""
    for sol in mesh.__dict__.pop("_solvers", {}).values():
        sol.close()
    for asm in mesh.__dict__.pop("_assemblers", {}).values():
        asm.close()


def _pattern_solver(mesh, p, args):
    """The host solver that owns the value buffer of K on the fixed pattern (host/solver.py)."""
    from ..capi import host_register, host_unregister
    from .solver import PatternSolver
    key = (p, args.get("pattern_mode", PATTERN_BLK), args.get("device", -1))
    cache = mesh.__dict__.setdefault("_solvers", {})
    if key not in cache:
        colptr, rowval = _assembler(mesh, p, args).pattern()
        cache[key] = PatternSolver(colptr, rowval, mesh.nmdf, register=host_register, unregister=host_unregister)
    return cache[key]


def calc_r_K(mesh, xms, cps, time, dt, p, **args):
    """Global residual vector and tangent matrix: `r_gl, K_gl = calc_r_K(mesh, xms, cps, time, Δt, p; args...)`.

    K_gl is a scipy CSC matrix (the analogue of SparseMatrixCSC{Float64,Int64}) on the library's symbolic pattern.
    Extra keyword arguments (all optional): scatter_mode (0 atomics / 1 deterministic), dropzeros (default True:
    entries that are exactly 0.0 are removed, which is the pattern Julia's `K[i,j] += v` insertion stores),
    pattern_mode, device, bend_tm (args[:bend_tm], FiniteElement.jl:379).
    """
    asm = _assembler(mesh, p, args)
    r, nzval, _ = asm.assemble(xms, cps, float(time), float(dt), bend_tm=float(args.get("bend_tm", 1.0)),
                               scatter_mode=args.get("scatter_mode", SCATTER_ATOMIC))
    colptr, rowval = asm.pattern()
    K = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(mesh.nmdf, mesh.nmdf))
    if args.get("dropzeros", True):
        K.eliminate_zeros()
    return r, K


def update_xms(motion, xms, cps, dt, dofs):
    """Forward-Euler mesh update, in place (FiniteElement.jl:408-423)."""
    mmo = get_m_motion_order(motion, dofs)
    for mj, dof in enumerate(mmo):
        if dof != 0:
            xms[:, mj] += dt * cps[:, dof - 1]


def time_step(mesh, xms, cps, time, dt, p, **args):
    """Newton-Raphson iteration to the next time level, in place (FiniteElement.jl:11-63).

    Returns the list of ε = ‖Δu‖₂ / nmdf per iteration (the reference prints it, :50). The sparse solve stays on
    the host (SciPy SuperLU stands in for Julia's UMFPACK `\\`); its time is accumulated in args['timers'].
    solver="pattern" selects the hand-off of host/solver.py: maf_assemble writes nzval into the page-locked buffer
    the solver owns, and the column ordering is chosen once per pattern instead of once per iteration.
    """
    eps_hist = []
    timers = args.get("timers")
    log = args.get("log")
    node_of, dof_of = mesh.ID_inv
    resident = bool(args.get("resident", False))
    solver = _pattern_solver(mesh, p, args) if args.get("solver") == "pattern" else None
    if resident or solver is not None:
        asm = _assembler(mesh, p, args)
        colptr, rowval = asm.pattern()
    if resident:   # the state lives on the device for the whole step; only r / K come back and du goes in
        asm.state_set(xms, cps)
    it = 1
    while it < 15:
        t0 = _time.perf_counter()
        kw = dict(bend_tm=float(args.get("bend_tm", 1.0)), scatter_mode=args.get("scatter_mode", SCATTER_ATOMIC))
        if solver is not None:     # f2 hand-off: the values land in the buffer the solver factorises
            kw["nzval"] = solver.nzval
        if resident:
            r_gl, nzval, _ = asm.assemble_resident(float(time), float(dt), **kw)
        elif solver is not None:
            r_gl, nzval, _ = asm.assemble(xms, cps, float(time), float(dt), **kw)
        if solver is None and resident:
            K_gl = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(mesh.nmdf, mesh.nmdf))
            if args.get("dropzeros", True):
                K_gl.eliminate_zeros()
        elif solver is None:
            r_gl, K_gl = calc_r_K(mesh, xms, cps, time, dt, p, **args)
        t1 = _time.perf_counter()
        du = -(solver.solve(r_gl) if solver is not None else spla.splu(K_gl.tocsc()).solve(r_gl))
        t2 = _time.perf_counter()
        if resident:
            asm.state_update(du, float(dt))
        else:
            dcps = np.zeros_like(cps)
            dcps[node_of - 1, dof_of - 1] = du
            cps += dcps
            update_xms(p.motion, xms, dcps, dt, mesh.dofs)
        eps_hist.append(float(np.linalg.norm(du) / mesh.nmdf))
        if timers is not None:
            timers["assembly_s"] = timers.get("assembly_s", 0.0) + (t1 - t0)
            timers["solve_s"] = timers.get("solve_s", 0.0) + (t2 - t1)
            timers["iterations"] = timers.get("iterations", 0) + 1
        if log is not None:
            log.write(f"--> iteration {it}: ε = {eps_hist[-1]}\n")
        it += 1
        if eps_hist[-1] < p.enr:
            break
    if resident:   # hand the converged state back (the reference mutates xms / cps in place)
        x_new, c_new = asm.state_get()
        xms[...] = x_new
        cps[...] = c_new
    assert eps_hist[-1] < p.enr, "did not reach Newton--Raphson tolerance"
    return eps_hist


def run_analysis(mesh, xms, cps, p, **args):
    """Time loop (Analysis.jl:17-102). Returns the per-step Newton histories."""
    dts = get_dts(args)
    times = np.cumsum(dts) + args["t0"]
    out = p.output and "out_path" in args
    log = open(os.path.join(args["out_path"], args["out_file"]), "a") if out else None
    if out:
        log.write("\nStarting to run analysis...\n")
        with open(os.path.join(args["out_path"], "times.txt"), "a") as f:
            f.write(f"{len(dts) + args['t0_id']} times:\n")
            for t_id, t in enumerate(times, start=1):
                f.write(f"{t_id + args['t0_id']}\t{t}\n")
    histories = []
    pull = p.scenario == F_PULL
    if pull:   # pull force local mappings (Analysis.jl:46-56)
        assert all(get_v_order(mesh.dofs)), "f_pull needs 3-D velocity"
        adj_el_ids, adj_node_map = get_adj_maps(mesh.num1el, mesh.numel, mesh.IX, p.poly)
        append = "in_path" in args and args["out_path"] == args["in_path"]       # Analysis.jl:34
        if out and not append:
            with open(os.path.join(args["out_path"], "f-pull.txt"), "a") as f:
                f.write("time\txp\typ\tzp\tfx\tfy\tfz\n")
    f_pulls = args.get("f_pulls")
    for t_id, dt in enumerate(dts, start=1):
        if log:
            log.write(f"\n-> Time {times[t_id - 1]}\n")
        update_xms(p.motion, xms, cps, dt, mesh.dofs)                       # predictor, Analysis.jl:70
        histories.append(time_step(mesh, xms, cps, float(times[t_id - 1]), float(dt), p, log=log, **args))
        if out:
            np.savetxt(os.path.join(args["out_path"], f"t{t_id + args['t0_id']}-xms.txt"), xms, delimiter="\t")
            np.savetxt(os.path.join(args["out_path"], f"t{t_id + args['t0_id']}-cps.txt"), cps, delimiter="\t")
        if pull and (out or f_pulls is not None):   # pull force calculation (Analysis.jl:80-92)
            x_pull = xms[mesh.IX[:, get_pull_el_id(mesh.numel) - 1] - 1, :].sum(axis=0) / mesh.IX.shape[0]
            f_pull = calc_pull_force(mesh, xms, cps, adj_el_ids, adj_node_map, p, **args)
            if f_pulls is not None:
                f_pulls.append((float(times[t_id - 1]), x_pull.copy(), f_pull.copy()))
            if out:
                with open(os.path.join(args["out_path"], "f-pull.txt"), "a") as f:
                    f.write("\t".join(repr(float(v)) for v in (times[t_id - 1], *x_pull, *f_pull)) + "\n")
    if log:
        log.write("\nCompleted running analysis...\n\n")
        log.close()
    return histories
