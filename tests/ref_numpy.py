"""Independent numpy (complex128) formulation of the element residuals, written in index/tensor form
(no Ba/Bb matrices, no Voigt packing) straight from the weak form that
/root/reference/src/analysis/FiniteElement.jl:253-400 discretises. It exists to cross-check the C++ oracle's
restatement, and is structurally different from it on purpose. Small meshes only (dense K).
"""
import numpy as np

STATIC, EUL, LAG, ALEV, ALEVB = 1, 2, 3, 4, 5
BOTTOM, RIGHT, TOP, LEFT = 1, 2, 3, 4
SHEAR, STRETCH, MOMENT = 1, 2, 3
XI = np.array([-np.sqrt(3 / 5), 0.0, np.sqrt(3 / 5)])


def _fields(ce, dofs):
    """dofs: length-8 array (Dof.Unknown order vx..pm) of 1-based columns or 0. Returns v(9,3) vm(9,3) lam(9) pm(9)."""
    def col(u):
        return ce[:, dofs[u] - 1] if dofs[u] else np.zeros(9, dtype=complex)
    v = np.stack([col(0), col(1), col(2)], axis=1)
    vm = np.stack([col(3), col(4), col(5)], axis=1)
    return v, vm, col(6), col(7)


def _geometry(xe, dN, ddN):
    a = np.einsum("ai,ak->ki", xe, dN)            # a[alpha, i]
    xab = np.zeros((2, 2, 3), dtype=complex)      # x_{,alpha beta}
    c = np.einsum("ai,ak->ki", xe, ddN)
    xab[0, 0], xab[1, 1], xab[0, 1], xab[1, 0] = c[0], c[1], c[2], c[2]
    aco = a @ a.T
    det = aco[0, 0] * aco[1, 1] - aco[0, 1] * aco[1, 0]
    acon = np.array([[aco[1, 1], -aco[0, 1]], [-aco[1, 0], aco[0, 0]]]) / det
    J = np.sqrt(det)
    aup = acon @ a                                 # a^alpha
    n = np.cross(a[0], a[1]) / J
    b = np.einsum("abi,i->ab", xab, n)
    bcon = acon @ b @ acon
    H = np.sum(acon * b) / 2
    K = (b[0, 0] * b[1, 1] - b[0, 1] * b[1, 0]) / det
    Gam = np.einsum("abi,mi->mab", xab, aup)       # Gamma^mu_{alpha beta}
    return a, aup, acon, J, n, bcon, H, K, Gam


def area_dof_residuals(tabs, xe, ce, dofs, P, motion):
    """tabs: list of 9 dicts (w, N, dN, ddN). Returns rv(9,3) rm(9,3) rl(9) rp(9)."""
    v_c, vm_c, l_c, p_c = _fields(ce, dofs)
    kb, kg, zv, pn, adb, am = P["kb"], P["kg"], P["zv"], P["pn"], P["adb"], P["am"]
    rv, rm = np.zeros((9, 3), complex), np.zeros((9, 3), complex)
    rl, rp = np.zeros(9, complex), np.zeros(9, complex)
    G, Hm = np.zeros((3, 9), complex), np.zeros((3, 3), complex)
    for gp, t in enumerate(tabs):
        N, dN, ddN, w = t["N"], t["dN"], t["ddN"], t["w"]
        a, aup, acon, J, n, bcon, H, K, Gam = _geometry(xe, dN, ddN)
        dv = np.einsum("ai,ak->ki", v_c, dN)
        dvm = np.einsum("ai,ak->ki", vm_c, dN)
        v, vm = v_c.T @ N, vm_c.T @ N
        lam, pm = l_c @ N, p_c @ N

        def visc(dvel):
            # pi^{ab} = zv (a^a . v_{,m} a^{mb} + a^b . v_{,m} a^{ma})
            g = np.einsum("ai,mi->am", aup, dvel) @ acon
            return zv * (g + g.T)
        bend = acon * (kb * H * H - kg * K) - 2 * kb * H * bcon
        sig = bend + acon * lam + visc(dv)
        sigm = bend + visc(dvm)
        M = acon * H * (kb + 2 * kg) - kg * bcon
        d2N = np.zeros((9, 2, 2))
        d2N[:, 0, 0], d2N[:, 1, 1], d2N[:, 0, 1], d2N[:, 1, 0] = ddN[:, 0], ddN[:, 1], ddN[:, 2], ddN[:, 2]
        covN = d2N - np.einsum("mab,pm->pab", Gam, dN)            # N_{;alpha beta}
        fs = np.einsum("ab,bi,pa->pi", sig, a, dN)                # sigma^{ab} a_b N_{,a}
        fb = np.einsum("ab,pab->p", M, covN)[:, None] * n[None, :]
        rv += (fs + fb) * J * w
        if pn != 0.0:
            rv -= np.outer(N, n) * pn * J * w
        rl += N * w * (J * np.sum(aup * dv) - adb * lam / zv)
        if motion == EUL:
            rm += np.outer(N, vm - n * np.sum(n * v)) * am * J * w
        elif motion in (ALEV, ALEVB):
            rm += np.einsum("ab,bi,pa->pi", sigm, a, dN) * J * w
            if motion == ALEVB:
                rm += fb * J * w
            rm -= np.outer(N, n) * pm * J * w
            rp -= N * w * J * np.sum(n * (vm - v))
            rp -= N * w * adb * pm / zv
        ndb = np.array([XI[gp % 3], XI[gp // 3], 1.0])
        G += np.outer(ndb, N) * w
        Hm += np.outer(ndb, ndb) * w
    T = G.T @ np.linalg.solve(Hm, G)
    rl += T @ l_c * adb / zv
    if motion in (ALEV, ALEVB):
        rp += T @ p_c * adb / zv
    return rv, rm, rl, rp


def interleave(rv, rm, rl, rp, dofs, ndf):
    r = np.zeros(9 * ndf, complex)
    for a in range(9):
        for j in range(3):
            if dofs[j]:
                r[dofs[j] - 1 + ndf * a] = rv[a, j]
            if dofs[3 + j]:
                r[dofs[3 + j] - 1 + ndf * a] = rm[a, j]
        r[dofs[6] - 1 + ndf * a] = rl[a]
        if dofs[7]:
            r[dofs[7] - 1 + ndf * a] = rp[a]
    return r


def area_residual(tabs, xe, ce, dofs, ndf, P, motion):
    return interleave(*area_dof_residuals(tabs, xe, ce, dofs, P, motion), dofs, ndf)


def bdry_residual(tabs3, bdry, ntype, nval, xe, ce, dofs, ndf, P, scenario_is_bend, time):
    rv = np.zeros((9, 3), complex)
    for t in tabs3:
        N, dN, ddN, w = t["N"], t["dN"], t["ddN"], t["w"]
        a, aup, acon, J, n, *_ = _geometry(xe, dN, ddN)
        tau = {BOTTOM: a[0], RIGHT: a[1], TOP: -a[0], LEFT: -a[1]}[bdry]
        tau = tau / np.sqrt(np.sum(tau * tau))
        nu = np.cross(tau, n)
        JG = 1 / np.sqrt(np.sum((aup @ tau) ** 2))
        if ntype in (STRETCH, SHEAR):
            f = nval * (nu if ntype == STRETCH else tau)
            rv -= np.outer(N, f) * JG * w
        elif ntype == MOMENT and scenario_is_bend:
            Mval = nval * min(time / P["bend_tm"], 1.0)
            rv -= np.outer(dN @ (aup @ nu), n) * Mval * JG * w
        else:
            raise AssertionError("Neumann boundary condition not implemented")
    z27, z9 = np.zeros((9, 3), complex), np.zeros(9, complex)
    r = np.zeros(9 * ndf, complex)
    for a in range(9):
        for j in range(3):
            if dofs[j]:
                r[dofs[j] - 1 + ndf * a] = rv[a, j]
    return r


def elem_r_K(resfn, xe, ce, active, mmo, dt, ndf, h=1e-30):
    """Complex-step element tangent with the reference's column rule (FiniteElement.jl:110-126).
    active: (9, ndf) bool; mmo: mesh-motion dof order (3 entries, 1-based or 0)."""
    xe = xe.astype(complex)
    ce = ce.astype(complex)
    r = resfn(xe, ce).real
    K = np.zeros((9 * ndf, 9 * ndf))
    for a in range(9):
        for d in range(ndf):
            if not active[a, d]:
                continue
            c2 = ce.copy()
            c2[a, d] += 1j * h
            col = resfn(xe, c2).imag / h
            if (d + 1) in list(mmo):
                comp = list(mmo).index(d + 1)
                x2 = xe.copy()
                x2[a, comp] += 1j * h
                col = col + dt * resfn(x2, ce).imag / h
            K[:, d + ndf * a] = col
    return r, K


def generate_output_ref(mesh, xms, cps):
    """Output.jl:32-118, statement by statement (area Gauss points :41-54, boundary Gauss points :57-88, corners
    :91-115), on the host mirror's mesh accessors. Test-side restatement for maf_generate_output."""
    import mafb200 as maf
    GP1D = 3
    n1, n2 = mesh.num1el * GP1D + 2, mesh.num2el * GP1D + 2
    xout = np.zeros((n1, n2, 3))
    uout = np.zeros((n1, n2, mesh.ndf))
    for el2 in range(1, mesh.num2el + 1):
        for el1 in range(1, mesh.num1el + 1):
            el = el1 + mesh.num1el * (el2 - 1)
            nodes = mesh.IX[:, el - 1] - 1
            for gp2 in range(1, GP1D + 1):
                for gp1 in range(1, GP1D + 1):
                    N = maf.get_N(el, gp1 + GP1D * (gp2 - 1), mesh)
                    o1, o2 = (el1 - 1) * GP1D + gp1, (el2 - 1) * GP1D + gp2          # 0-based of out_id = ... + 1
                    xout[o1, o2, :] = xms[nodes, :].T @ N
                    uout[o1, o2, :] = cps[nodes, :].T @ N
    for bdry in (maf.BOTTOM, maf.RIGHT, maf.TOP, maf.LEFT):
        out_id = 2
        for el in mesh.bdry_elems[bdry]:
            nodes = mesh.IX[:, el - 1] - 1
            for gp in range(1, GP1D + 1):
                N = maf.get_N(bdry, int(el), gp, mesh)
                o1, o2 = {maf.BOTTOM: (out_id, 1), maf.RIGHT: (n1, out_id), maf.TOP: (out_id, n2),
                          maf.LEFT: (1, out_id)}[bdry]
                xout[o1 - 1, o2 - 1, :] = xms[nodes, :].T @ N
                uout[o1 - 1, o2 - 1, :] = cps[nodes, :].T @ N
                out_id += 1
    l1, l2 = mesh.line_gp_fns1, mesh.line_gp_fns2
    corners = {(1, 1): (1, l1.edge[0], l2.edge[0]), (n1, 1): (mesh.num1el, l1.edge[1], l2.edge[0]),
               (1, n2): (mesh.numel - mesh.num1el + 1, l1.edge[0], l2.edge[1]), (n1, n2): (mesh.numel, l1.edge[1], l2.edge[1])}
    for (o1, o2), (el, f1, f2) in corners.items():                                   # crnr_gp_fns, Mesh.jl:228-233
        nodes = mesh.IX[:, el - 1] - 1
        N = np.array([f1[1 + a % 3] * f2[1 + a // 3] for a in range(9)])
        xout[o1 - 1, o2 - 1, :] = xms[nodes, :].T @ N
        uout[o1 - 1, o2 - 1, :] = cps[nodes, :].T @ N
    return xout, uout
