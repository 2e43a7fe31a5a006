"""Scenario boundary conditions of the reference (src/input/Bc.jl). Each function returns
(dofs, ndf, ID, inh_dir_bcs, inh_neu_bcs) like its Julia counterpart; ID holds -1 for Dirichlet dofs and 0
elsewhere (numbering happens in Mesh.generate_scenario, Mesh.jl:276-284). Node ids are 1-based."""
import numpy as np

from .enums import (ALEV, ALEVB, BOTTOM, EUL, F_BEND, F_CAVI, F_COUE, F_POIS, F_PULL, LAG, LEFT, MOMENT, RIGHT,
                    STATIC, STRETCH, TOP, Corner, Dof)

U = Dof.Unknown


def get_dofs(motion):
    """Bc.jl:414-431."""
    if motion in (LAG, STATIC):
        return {U.vx: 1, U.vy: 2, U.vz: 3, U.lam: 4}
    if motion == EUL:
        return {U.vx: 1, U.vy: 2, U.vz: 3, U.vmx: 4, U.vmy: 5, U.vmz: 6, U.lam: 7}
    if motion in (ALEV, ALEVB):
        return {U.vx: 1, U.vy: 2, U.vz: 3, U.vmx: 4, U.vmy: 5, U.vmz: 6, U.lam: 7, U.pm: 8}
    raise AssertionError(f"{motion} motion degrees of freedom not provided")


def get_pull_el_id(numel):
    """PullForce.jl:8-13."""
    return -(-numel // 2)


def _fix(ID, dof, nodes):
    ID[dof - 1, np.asarray(nodes, dtype=np.int64) - 1] = -1


def get_f_cavi_bc_info(numnp, bdry_nodes, crnr_nodes):
    """Bc.jl:58-106."""
    dofs = {U.vx: 1, U.vy: 2, U.lam: 3}
    ID = np.zeros((3, numnp), dtype=np.int64, order="F")
    inh_dir = []
    for b in (BOTTOM, TOP, LEFT, RIGHT):
        _fix(ID, 1, bdry_nodes[b])
        _fix(ID, 2, bdry_nodes[b])
    for i in bdry_nodes[TOP]:
        if i != crnr_nodes[Corner.TOP_LEFT] and i != crnr_nodes[Corner.TOP_RIGHT]:
            inh_dir.append((U.vx, int(i), 1.0))
    center = numnp // 2 + 1
    _fix(ID, 3, [center])
    inh_dir.append((U.lam, center, 0.0))
    return dofs, 3, ID, inh_dir, []


def get_f_coue_bc_info(numnp, bdry_nodes):
    """Bc.jl:122-157."""
    dofs = {U.vx: 1, U.vy: 2, U.lam: 3}
    ID = np.zeros((3, numnp), dtype=np.int64, order="F")
    for b in (BOTTOM, TOP):
        _fix(ID, 1, bdry_nodes[b])
        _fix(ID, 2, bdry_nodes[b])
    inh_dir = [(U.vx, int(i), 3.0) for i in bdry_nodes[TOP]]
    _fix(ID, 2, bdry_nodes[LEFT])
    _fix(ID, 2, bdry_nodes[RIGHT])
    return dofs, 3, ID, inh_dir, [(LEFT, STRETCH, 4.0), (RIGHT, STRETCH, 4.0)]


def get_f_pois_bc_info(numnp, bdry_nodes):
    """Bc.jl:286-316."""
    dofs = {U.vx: 1, U.vy: 2, U.lam: 3}
    ID = np.zeros((3, numnp), dtype=np.int64, order="F")
    for b in (TOP, BOTTOM):
        _fix(ID, 1, bdry_nodes[b])
        _fix(ID, 2, bdry_nodes[b])
    for b in (LEFT, RIGHT):
        _fix(ID, 2, bdry_nodes[b])
    return dofs, 3, ID, [], [(LEFT, STRETCH, 4.0), (RIGHT, STRETCH, 8.0)]


def get_f_pull_bc_info(numnp, IX, bdry_nodes, bdry_inner_nodes, p, **args):
    """Bc.jl:188-269."""
    dofs = get_dofs(p.motion)
    ndf = len(dofs)
    ID = np.zeros((ndf, numnp), dtype=np.int64, order="F")
    inh_dir = []
    ale = p.motion in (ALEV, ALEVB)
    for b in (BOTTOM, RIGHT, TOP, LEFT):
        _fix(ID, dofs[U.vz], bdry_nodes[b])
        if ale:
            for u in (U.vmx, U.vmy, U.vmz):
                _fix(ID, dofs[u], bdry_nodes[b])
    for b in (BOTTOM, RIGHT, TOP, LEFT):
        _fix(ID, dofs[U.vz], bdry_inner_nodes[b])
        if p.motion == ALEVB:
            _fix(ID, dofs[U.vmz], bdry_inner_nodes[b])
    center = get_pull_el_id(IX.shape[1])
    for nd in IX[:, center - 1]:
        for u in (U.vx, U.vy, U.vz):
            _fix(ID, dofs[u], [nd])
        inh_dir.append((U.vz, int(nd), args["pull_speed"]))
        if p.motion != LAG:
            for u in (U.vmx, U.vmy, U.vmz):
                _fix(ID, dofs[u], [nd])
            inh_dir.append((U.vmz, int(nd), args["pull_speed"]))
    for b in (BOTTOM, RIGHT, TOP, LEFT):
        c = bdry_nodes[b][len(bdry_nodes[b]) // 2]          # bdry_nodes[bdry][floor(end/2)+1]
        _fix(ID, dofs[U.vx], [c])
        _fix(ID, dofs[U.vy], [c])
        if p.motion != LAG:
            _fix(ID, dofs[U.vmx], [c])
            _fix(ID, dofs[U.vmy], [c])
    lval = p.kb / 4
    neu = [(LEFT, STRETCH, lval), (RIGHT, STRETCH, lval), (TOP, STRETCH, lval), (BOTTOM, STRETCH, lval)]
    assert p.pn == 0.0, f"{F_PULL.name} with a normal pressure is not implemented"
    return dofs, ndf, ID, inh_dir, neu


def get_f_bend_bc_info(numnp, bdry_nodes, p, **args):
    """Bc.jl:347-393."""
    dofs = get_dofs(p.motion)
    ndf = len(dofs)
    ID = np.zeros((ndf, numnp), dtype=np.int64, order="F")
    ale = p.motion in (ALEV, ALEVB)
    for b in (TOP, BOTTOM):
        _fix(ID, dofs[U.vy], bdry_nodes[b])
        if ale:
            _fix(ID, dofs[U.vmy], bdry_nodes[b])
    for u in (U.vx, U.vy, U.vz):
        _fix(ID, dofs[u], bdry_nodes[LEFT])
    if ale:
        for u in (U.vmx, U.vmy, U.vmz):
            _fix(ID, dofs[u], bdry_nodes[LEFT])
    _fix(ID, dofs[U.vz], bdry_nodes[RIGHT])
    if ale:
        _fix(ID, dofs[U.vmz], bdry_nodes[RIGHT])
    return dofs, ndf, ID, [], [(LEFT, MOMENT, args["bend_mf"]), (RIGHT, MOMENT, args["bend_mf"])]


def get_scenario_bc_info(numnp, IX, bdry_nodes, bdry_inner_nodes, crnr_nodes, p, **args):
    """Bc.jl:15-40."""
    if p.scenario == F_CAVI:
        return get_f_cavi_bc_info(numnp, bdry_nodes, crnr_nodes)
    if p.scenario == F_COUE:
        return get_f_coue_bc_info(numnp, bdry_nodes)
    if p.scenario == F_POIS:
        return get_f_pois_bc_info(numnp, bdry_nodes)
    if p.scenario == F_BEND:
        return get_f_bend_bc_info(numnp, bdry_nodes, p, **args)
    if p.scenario == F_PULL:
        return get_f_pull_bc_info(numnp, IX, bdry_nodes, bdry_inner_nodes, p, **args)
    raise AssertionError(f"Need boundary conditions for {p.scenario} scenario")
