"""`Mesh` of the reference (src/input/Mesh.jl:48-302, 477-593): same field names, 1-based Int64 index tables in
column-major (Fortran) order so that they cross the C ABI unchanged. Table construction is vectorised (numpy) so
that the 10^6-element patches of the benchmark build in seconds; results equal the reference's loops."""
import numpy as np

from .basis import AreaGpBasisFns, BdryGpBasisFns, LineGpBasisFns
from .bc import get_scenario_bc_info
from .enums import (BOTTOM, CLAMPED, F_BEND, F_CAVI, F_COUE, F_POIS, F_PULL, FLAT, GP1D, LAG, LEFT, NEN, POLY, RIGHT,
                    STATIC, TOP, Corner, Dof)
from .spline import KnotVector, get_bspline_indices, get_fine_zs

U = Dof.Unknown


def get_topology(scenario):
    """Mesh.jl:550-566."""
    if scenario in (F_CAVI, F_COUE, F_POIS, F_PULL, F_BEND):
        return FLAT
    raise AssertionError(f"topology for {scenario} scenario not specified")


def construct_IX(kv1, kv2, num1np):
    """Mesh.jl:574-593: IX[nid1 + 3(nid2-1), eid1 + (eid2-1) nel1] = ids1[nid1] + num1np (ids2[nid2] - 1)."""
    s1 = np.stack([get_bspline_indices(kv1, e + POLY) for e in range(1, kv1.nel + 1)])   # (nel1, 3)
    s2 = np.stack([get_bspline_indices(kv2, e + POLY) for e in range(1, kv2.nel + 1)])   # (nel2, 3)
    # axes: (nid2, nid1, eid2, eid1)
    ix = s1.T[None, :, None, :] + num1np * (s2.T[:, None, :, None] - 1)
    return np.asfortranarray(ix.reshape(NEN, kv1.nel * kv2.nel).astype(np.int64))


def get_v_order(dofs):
    return [dofs.get(U.vx, 0), dofs.get(U.vy, 0), dofs.get(U.vz, 0)]        # Mesh.jl:477-483


def get_m_order(dofs):
    return [dofs.get(U.vmx, 0), dofs.get(U.vmy, 0), dofs.get(U.vmz, 0)]     # Mesh.jl:491-497


def get_lam_order(dofs):
    return dofs.get(U.lam, 0)                                               # Mesh.jl:505-509  (get_λ_order)


def get_p_order(dofs):
    return dofs.get(U.pm, 0)                                                # Mesh.jl:517-521


def get_m_motion_order(motion, dofs):
    """Mesh.jl:529-542."""
    if motion == STATIC:
        return [0, 0, 0]
    if motion == LAG:
        return get_v_order(dofs)
    return get_m_order(dofs)


class Mesh:
    """Mesh(p; args...) (Mesh.jl:79, generate_mesh :94-248, generate_scenario :262-302)."""

    def __init__(self, p, **args):
        self.topology = get_topology(p.scenario)
        assert self.topology == FLAT, f"mesh construction for {self.topology} topology not implemented"
        self.num1el, self.num2el = p.num1el, p.num2el
        self.numel = p.num1el * p.num2el
        self.num1np, self.num2np = p.num1el + p.poly, p.num2el + p.poly
        self.numnp = self.num1np * self.num2np
        n1, nn, ne, e1 = self.num1np, self.numnp, self.numel, p.num1el
        ar = lambda a, b, s=1: np.arange(a, b + 1, s, dtype=np.int64)   # Julia a:s:b
        self.bdry_elems = {BOTTOM: ar(1, e1), RIGHT: ar(e1, ne, e1), TOP: ar(ne - e1 + 1, ne),
                           LEFT: ar(1, ne - e1 + 1, e1)}                                   # Mesh.jl:126-131
        self.crnr_elems = {Corner.BOTTOM_LEFT: 1, Corner.BOTTOM_RIGHT: e1, Corner.TOP_LEFT: ne - e1 + 1,
                           Corner.TOP_RIGHT: ne}
        self.bdry_nodes = {BOTTOM: ar(1, n1), RIGHT: ar(n1, nn, n1), TOP: ar(nn - n1 + 1, nn),
                           LEFT: ar(1, nn - n1 + 1, n1)}                                   # Mesh.jl:144-149
        self.bdry_inner_nodes = {BOTTOM: ar(1, n1) + n1, RIGHT: ar(n1, nn, n1) - 1, TOP: ar(nn - n1 + 1, nn) - n1,
                                 LEFT: ar(1, nn - n1 + 1, n1) + 1}                         # Mesh.jl:152-157
        self.crnr_nodes = {Corner.BOTTOM_LEFT: 1, Corner.BOTTOM_RIGHT: n1, Corner.TOP_LEFT: nn - n1 + 1,
                           Corner.TOP_RIGHT: nn}
        self.crnr_inner_nodes = {Corner.BOTTOM_LEFT: n1 + 2, Corner.BOTTOM_RIGHT: 2 * n1 - 1,
                                 Corner.TOP_LEFT: nn - 2 * n1 + 2, Corner.TOP_RIGHT: nn - n1 - 1}
        if p.scenario == F_PULL and p.num1el >= 18 and p.num2el >= 18:                      # Mesh.jl:176-181
            self.kv1 = KnotVector(get_fine_zs(p.num1el, p.poly), p.poly, CLAMPED)
            self.kv2 = KnotVector(get_fine_zs(p.num2el, p.poly), p.poly, CLAMPED)
        else:
            self.kv1 = KnotVector(p.num1el, p.poly, CLAMPED)
            self.kv2 = KnotVector(p.num2el, p.poly, CLAMPED)
        self.line_gp_fns1 = LineGpBasisFns(self.kv1, p.gp1d)                               # Mesh.jl:213-214
        self.line_gp_fns2 = LineGpBasisFns(self.kv2, p.gp1d)
        self.area_gp_fns = AreaGpBasisFns(self.line_gp_fns1, self.line_gp_fns2)            # Mesh.jl:217
        self.bdry_gp_fns = {                                                                # Mesh.jl:220-225
            BOTTOM: BdryGpBasisFns(self.line_gp_fns1, self.line_gp_fns2.zmin_fns, BOTTOM),
            RIGHT: BdryGpBasisFns(self.line_gp_fns2, self.line_gp_fns1.zmax_fns, RIGHT),
            TOP: BdryGpBasisFns(self.line_gp_fns1, self.line_gp_fns2.zmax_fns, TOP),
            LEFT: BdryGpBasisFns(self.line_gp_fns2, self.line_gp_fns1.zmin_fns, LEFT)}
        self.IX = construct_IX(self.kv1, self.kv2, self.num1np)
        # generate_scenario (Mesh.jl:262-302)
        self.dofs, self.ndf, ID, self.inh_dir_bcs, self.inh_neu_bcs = get_scenario_bc_info(
            self.numnp, self.IX, self.bdry_nodes, self.bdry_inner_nodes, self.crnr_nodes, p, **args)
        flat = ID.reshape(-1, order="F")               # node-major: for node, for dof (Mesh.jl:277)
        free = flat != -1
        flat[:] = np.where(free, np.cumsum(free), 0)
        self.ID = np.asfortranarray(flat.reshape((self.ndf, self.numnp), order="F"))
        self.nmdf = int(self.ID.max())
        node_of, dof_of = np.nonzero(self.ID.T)        # ID_inv: unknown id -> (node, dof), 1-based (Mesh.jl:291-296)
        self.ID_inv = (node_of.astype(np.int64) + 1, dof_of.astype(np.int64) + 1)
        self._LM = None
        self.motion, self.scenario = p.motion, p.scenario

    @property
    def LM(self):
        """LM = reshape(ID[:, IX], (ndf*NEN, numel)) (Mesh.jl:299); built on first use (it is 9 ndf numel Int64)."""
        if self._LM is None:
            self._LM = np.asfortranarray(
                self.ID[:, self.IX.reshape(-1, order="F") - 1].reshape((self.ndf * NEN, self.numel), order="F"))
        return self._LM

    def dofs8(self):
        """Mesh.dofs as the C ABI's int32[8] (column of vx vy vz vmx vmy vmz λ pm, 0 = absent)."""
        return np.array([self.dofs.get(u, 0) for u in U], dtype=np.int32)


# accessors of the reference (Mesh.jl:311-469); el_id / gp_id are 1-based
def get_basis_fns(*a):
    if len(a) == 3:
        el_id, gp_id, mesh = a
        assert 1 <= gp_id <= GP1D ** 2, "2-D Gauss point index out of bounds"
        return mesh.area_gp_fns.ufn(mesh.area_gp_fns.uel_of(el_id), gp_id)
    bdry, el_id, gp_id, mesh = a
    assert 1 <= gp_id <= GP1D, "1-D Gauss point index out of bounds"
    pos = np.nonzero(mesh.bdry_elems[bdry] == el_id)[0]
    assert len(pos) > 0, "element id not found on boundary"
    fns = mesh.bdry_gp_fns[bdry]
    return fns.ufn(int(fns.uel_ids[pos[0]]), gp_id)


def get_gpw(*a):
    return get_basis_fns(*a)["w"]


def get_N(*a):
    return get_basis_fns(*a)["N"]


def get_dN(*a):          # get_∂Nα
    return get_basis_fns(*a)["dN"]


def get_ddN(*a):         # get_∂∂Nαβ
    return get_basis_fns(*a)["ddN"]
