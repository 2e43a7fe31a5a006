"""Pull force of the F_PULL scenario, mirroring src/analysis/PullForce.jl.

The pulled nodes are Dirichlet nodes, so their residual rows are not part of r_gl; the reaction force is the sum of
the element residuals rv of the 25 elements around the pulled one, restricted to the pulled nodes. The element
residuals come from the device (maf_elem_v_residuals: the same gather / interpolate / Gauss-point phases as the
assembly); the bookkeeping below is the reference's.
"""
import numpy as np

XDIM = 3
NEN = 9


def get_pull_el_id(numel):
    """Element that is pulled (PullForce.jl:7-12): ceil(numel / 2)."""
    return -(-numel // 2)


def get_adj_maps(num1el, numel, IX, poly=2):
    """(adj_el_ids, adj_node_map) of PullForce.jl:25-53, 1-based.

    adj_node_map[k] = (indices into the 27-vector of the pulled element, indices into rv of adjacent element k),
    both 0-based here, ordered like the reference's findall loops."""
    pull_el = get_pull_el_id(numel)
    pull_nodes = IX[:, pull_el - 1]
    adj = [pull_el + i * num1el + j for i in range(-poly, poly + 1) for j in range(-poly, poly + 1)]
    maps = []
    for e in adj:
        nodes = IX[:, e - 1]
        into = [i for i in range(NEN) if pull_nodes[i] in nodes]      # findall(x -> x in adj_nodes, pull_nodes)
        frm = [i for i in range(NEN) if nodes[i] in pull_nodes]       # findall(x -> x in pull_nodes, adj_nodes)
        maps.append((np.array([XDIM * i + c for i in into for c in range(XDIM)], dtype=np.int64),
                     np.array([XDIM * i + c for i in frm for c in range(XDIM)], dtype=np.int64)))
    return np.array(adj, dtype=np.int64), maps


def sum_pull_force(rv_els, adj_node_map):
    """The accumulation of calc_pull_force (PullForce.jl:71-79) given rv of every adjacent element, (n, 27)."""
    rv_pull = np.zeros(XDIM * NEN)
    for k, (into, frm) in enumerate(adj_node_map):
        rv_pull[into] += rv_els[k][frm]
    return rv_pull.reshape(NEN, XDIM).sum(axis=0)


def calc_pull_force(mesh, xms, cps, adj_el_ids, adj_node_map, p, **args):
    """`calc_pull_force(mesh, xms, cps, adj_el_ids, adj_node_map, p)` -> 3-vector (PullForce.jl:61-80).

    xms / cps may be None to use the state the device already holds (resident Newton loop)."""
    from .analysis import _assembler
    asm = _assembler(mesh, p, args)
    if xms is not None:
        asm.state_set(xms, cps)
    return sum_pull_force(asm.elem_v_residuals(adj_el_ids), adj_node_map)
