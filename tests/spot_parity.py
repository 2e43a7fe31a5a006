"""Spot parity at sizes the oracle cannot assemble as a whole (BASELINE.json config 5: 1001 x 1001 F_PULL patch, centre-
refined knots, ALEVB): single elements through the library -- maf_set_element_range(el, el), one assembly, the slots it
wrote -- against the oracle's element routine (elem_r_K = FiniteElement.jl:98-126, + the Neumann element :156-184)
scattered through LM (:129-136, 187-194), and against the extended-precision truth with the strict rule of cases.py.

Used by tests/test_gpu_parity.py (>= 200 elements) and, with a handful of elements, by bench.py's `parity_spot` key
(the oracle as checker beside the cpu_baseline leg)."""
import numpy as np

from cases import EPS, strict_errors
from oracle import oracle as orc


def select_elements(mesh, n_random=100, seed=11):
    """Element ids (1-based) that stress every special place of the F_PULL patch: the pulled element and its two
    rings, the four corners, stretches of each edge incl. the mid-edge nodes with their in-plane Dirichlet
    conditions, the rows / columns where the centre-refined knot spacing changes, and random interior elements."""
    n1, n2 = mesh.num1el, mesh.num2el
    numel = n1 * n2
    el = lambda e1, e2: e1 + (e2 - 1) * n1      # noqa: E731  (Mesh.jl:582-588)
    pick = set()
    pull = (numel + 1) // 2                     # PullForce.jl:12
    pe1, pe2 = (pull - 1) % n1 + 1, (pull - 1) // n1 + 1
    for d2 in range(-2, 3):
        for d1 in range(-2, 3):
            pick.add(el(pe1 + d1, pe2 + d2))
    for e1 in (1, n1):
        for e2 in (1, n2):
            pick.add(el(e1, e2))
    mid1, mid2 = (n1 + 1) // 2, (n2 + 1) // 2
    for k in list(range(2, 6)) + list(range(mid1 - 2, mid1 + 3)) + [n1 - 1]:
        pick.update((el(k, 1), el(k, n2), el(k, 2), el(k, n2 - 1)))
    for k in list(range(2, 6)) + list(range(mid2 - 2, mid2 + 3)) + [n2 - 1]:
        pick.update((el(1, k), el(n1, k), el(2, k), el(n1 - 1, k)))
    # knot transitions: where the unique 1-D element id changes (GpBasisFn.jl:182-188)
    u1 = np.asarray(mesh.line_gp_fns1.uel_ids)
    u2 = np.asarray(mesh.line_gp_fns2.uel_ids)
    t1 = [int(k) + 1 for k in np.nonzero(np.diff(u1))[0]]
    t2 = [int(k) + 1 for k in np.nonzero(np.diff(u2))[0]]
    for a in t1[:12] + t1[-12:]:
        for b in (t2[len(t2) // 2] if t2 else mid2, mid2, 7):
            pick.add(el(min(max(a, 1), n1), min(max(b, 1), n2)))
    for b in t2[:12] + t2[-12:]:
        for a in (t1[len(t1) // 2] if t1 else mid1, mid1, 7):
            pick.add(el(min(max(a, 1), n1), min(max(b, 1), n2)))
    rng = np.random.default_rng(seed)
    pick.update(int(v) for v in rng.integers(1, numel + 1, size=n_random))
    return sorted(e for e in pick if 1 <= e <= numel)


def _boundaries_of(om, el):
    """Boundaries (code, position in that boundary's element list) element `el` lies on (Mesh.jl:126-131)."""
    n1, n2 = om.num1el, om.num2el
    e1, e2 = (el - 1) % n1 + 1, (el - 1) // n1 + 1
    out = []
    if e2 == 1:
        out.append(orc.BOTTOM)
    if e1 == n1:
        out.append(orc.RIGHT)
    if e2 == n2:
        out.append(orc.TOP)
    if e1 == 1:
        out.append(orc.LEFT)
    return out


def expected_element(om, el, xms, cps, time, dt, neumann):
    """(r_el, K_el, E_r, E_K) of everything the reference adds for element `el`: the area element plus, for every
    inhomogeneous Neumann condition on a boundary the element lies on, its boundary element."""
    r, K, rm, Km = om.elem_r_K_mag(el, xms, cps, time, dt)
    on = _boundaries_of(om, el)
    for (bdry, ntype, nval) in neumann:
        if bdry in on:
            rb, Kb, rmb, Kmb = om.elem_r_K_mag(el, xms, cps, time, dt, bdry=bdry, ntype=ntype, nval=nval)
            r, K, rm, Km = r + rb, K + Kb, rm + rmb + np.abs(rb), Km + Kmb + np.abs(Kb)
    return r, K, rm, Km


def check_elements(asm, mesh, om, ot, xms_dev, cps_dev, xms, cps, time, dt, elements, colptr=None):
    """Runs every listed element alone through the library and compares the written slots. Returns a dict with the
    worst errors; raises AssertionError on the first violation of the pattern or of the strict rule."""
    colptr = asm.colptr() if colptr is None else colptr
    LM_of = lambda el: mesh.ID[:, mesh.IX[:, el - 1] - 1].reshape(-1, order="F")    # noqa: E731  dof + ndf (a-1), Mesh.jl:299
    _, neumann = om.bcs
    worst = {"n": 0, "K_rel_oracle": 0.0, "r_abs_oracle": 0.0, "K_strict": 0.0, "K_in_epsE": 0.0, "r_strict": 0.0,
             "oracle_in_epsE": 0.0}
    for el in elements:
        asm.set_element_range(el, el)
        asm.assemble_device(xms_dev, cps_dev, time, dt)
        info = asm.range_info()
        (r_lo, r_hi), (s_lo, s_hi) = info["rows"], info["slots"]
        r_g, nz_g = asm.download(r_lo, r_hi - r_lo + 1, s_lo, s_hi - s_lo + 1)
        lm = LM_of(el)
        act = np.nonzero(lm)[0]
        r_o, K_o, _, _ = expected_element(om, el, xms, cps, time, dt, neumann)
        r_t, K_t, E_r, E_K = expected_element(ot, el, xms, cps, time, dt, neumann)
        # residual rows
        got_r = r_g[lm[act] - r_lo]
        q, _ = strict_errors(got_r, r_t[act], E_r[act])
        worst["r_strict"] = max(worst["r_strict"], q)
        worst["r_abs_oracle"] = max(worst["r_abs_oracle"], float(np.abs(got_r - r_o[act]).max()))
        r_g[lm[act] - r_lo] = 0.0
        assert not r_g.any(), (el, "residual rows outside the element's LM were written")
        # tangent: column by column
        for ci in act:
            gc = int(lm[ci])
            rows = asm.pattern_columns(gc, gc, colptr)
            base = int(colptr[gc - 1]) - s_lo          # position of the column inside the downloaded slice
            pos = np.searchsorted(rows, lm[act])
            hit = (pos < rows.size) & (rows[np.minimum(pos, rows.size - 1)] == lm[act])
            # rows of the LM x LM union the library does not store belong to dof blocks that vanish identically
            # (MAF_PATTERN_BLK): the reference's value there must be an exact zero
            assert not np.any(K_o[act[~hit], ci]) and not np.any(K_t[act[~hit], ci]), (el, gc, "entry outside the pattern")
            got = nz_g[base + pos[hit]]
            q, u = strict_errors(got, K_t[act[hit], ci], E_K[act[hit], ci])
            qo, uo = strict_errors(K_o[act[hit], ci], K_t[act[hit], ci], E_K[act[hit], ci])
            worst["K_strict"] = max(worst["K_strict"], q)
            worst["K_in_epsE"] = max(worst["K_in_epsE"], u)
            worst["oracle_in_epsE"] = max(worst["oracle_in_epsE"], uo)
            scale = np.abs(K_o[:, ci]).max()
            if scale > 0:
                worst["K_rel_oracle"] = max(worst["K_rel_oracle"], float(np.abs(got - K_o[act[hit], ci]).max() / scale))
            nz_g[base + pos[hit]] = 0.0
        assert not nz_g.any(), (el, "slots outside the element's LM x LM block were written")
        assert worst["K_strict"] <= 1.0 and worst["r_strict"] <= 1.0, (el, worst)
        worst["n"] += 1
    asm.set_element_range(1, mesh.numel)
    return worst
