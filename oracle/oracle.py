"""ctypes binding of the CPU oracle (oracle/maf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference arm -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATIC, EUL, LAG, ALEV, ALEVB = 1, 2, 3, 4, 5
F_CAVI, F_COUE, F_POIS, F_PULL, F_BEND = 1, 2, 3, 4, 5
BOTTOM, RIGHT, TOP, LEFT = 1, 2, 3, 4
SHEAR, STRETCH, MOMENT = 1, 2, 3
CLAMPED, CLOSED = 1, 2


class OrcParams(C.Structure):
    _fields_ = [("motion", C.c_int32), ("scenario", C.c_int32), ("num1el", C.c_int32), ("num2el", C.c_int32),
                ("length", C.c_double), ("kb", C.c_double), ("kg", C.c_double), ("zv", C.c_double),
                ("pn", C.c_double), ("adb", C.c_double), ("am", C.c_double), ("ek", C.c_double),
                ("pull_speed", C.c_double), ("bend_mf", C.c_double), ("bend_tm", C.c_double)]


# builds of the ONE source maf_oracle.cpp (oracle/Makefile):
#   oracle  the restated reference algorithm in double / complex<double>
#   truth   the same algorithm in long double (64-bit mantissa), + per-entry magnitudes sum |terms|
#   truthq  the same in __float128 (113-bit mantissa): checks the long-double truth
#   native  `oracle` compiled -O3 -march=native ON THIS MACHINE (bench.py's CPU baseline)
_SO = {"oracle": "libmaf_oracle.so", "truth": "libmaf_truth.so", "truthq": "libmaf_truthq.so",
       "native": "libmaf_oracle_native.so"}
_LIBS = {}


def _cpu_tag():
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def build(force=False, kind="oracle"):
    target = _SO[kind]
    so = os.path.join(_HERE, target)
    if kind == "native":      # keyed by the CPU it was compiled for: a copy made on another machine is not reused
        so = os.path.join(_HERE, "libmaf_oracle_native_%s.so" % _cpu_tag())
    src = os.path.join(_HERE, "maf_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", target], stdout=subprocess.DEVNULL)
        if kind == "native":
            os.replace(os.path.join(_HERE, target), so)
    return so


def lib(kind="oracle"):
    if kind not in _LIBS:
        L = C.CDLL(build(kind=kind))
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_mesh_create", "orc_kv_from_list", "orc_kv_uniform", "orc_kv_of_mesh", "orc_calc_r_K"):
            getattr(L, name).restype = C.c_void_p
        L.orc_result_nnz.restype = C.c_int64
        L.orc_mesh_bdry_count.restype = C.c_int64
        _LIBS[kind] = L
    return _LIBS[kind]


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _err():
    return lib().orc_last_error().decode()


class KnotVector:
    def __init__(self, handle):
        if not handle:
            raise AssertionError(_err())
        self.h = C.c_void_p(handle)

    @classmethod
    def from_list(cls, zs, poly, curve=CLAMPED):
        z = np.ascontiguousarray(zs, dtype=np.float64)
        return cls(lib().orc_kv_from_list(_p(z), len(z), poly, curve))

    @classmethod
    def uniform(cls, nel, poly, curve=CLAMPED):
        return cls(lib().orc_kv_uniform(nel, poly, curve))

    def __del__(self):
        try:
            lib().orc_kv_destroy(self.h)
        except Exception:
            pass

    @property
    def zs(self):
        n = lib().orc_kv_len(self.h)
        o = np.empty(n)
        lib().orc_kv_knots(self.h, _p(o))
        return o

    @property
    def nel(self):
        return lib().orc_kv_nel(self.h)

    def span(self, z):
        r = lib().orc_knot_span(self.h, C.c_double(z))
        if r < 0:
            raise AssertionError(_err())
        return r

    def vals(self, z, poly):
        o = np.empty(poly + 1)
        if lib().orc_bspline_vals(self.h, C.c_double(z), _p(o)):
            raise AssertionError(_err())
        return o

    def ders(self, z, nd, poly):
        o = np.empty((poly + 1, nd + 1))
        if lib().orc_bspline_ders(self.h, C.c_double(z), nd, _p(o)):
            raise AssertionError(_err())
        return o

    def indices(self, z, poly):
        o = np.empty(poly + 1, dtype=np.int32)
        if lib().orc_bspline_indices(self.h, C.c_double(z), _p(o, C.c_int32)):
            raise AssertionError(_err())
        return o

    def collocate(self):
        o = np.empty(lib().orc_kv_len(self.h))
        n = lib().orc_collocate(self.h, _p(o))
        if n < 0:
            raise AssertionError(_err())
        return o[:n].copy()

    def cps_1d(self, f):
        z = self.collocate()
        x = np.array([f(t) for t in z], dtype=np.float64)
        o = np.empty(len(z))
        if lib().orc_cps_1d(self.h, _p(x), len(x), _p(o)):
            raise AssertionError(_err())
        return o

    def unique_1d(self):
        nel = self.nel
        ids = np.empty(nel, dtype=np.int64)
        lst = np.empty((nel, 2))
        n = lib().orc_unique_1d(self.h, _p(ids, C.c_int64), _p(lst))
        if n < 0:
            raise AssertionError(_err())
        return n, nel, ids, [tuple(r) for r in lst[:n]]

    def line_fns(self, ngp=3):
        nel = self.nel
        ids = np.empty(nel, dtype=np.int64)
        tab = np.empty((nel, ngp, 10))
        n = lib().orc_line_fns(self.h, ngp, _p(ids, C.c_int64), _p(tab))
        if n < 0:
            raise AssertionError(_err())
        return ids, tab[:n].copy()

    def fn1(self, w, z):
        o = np.empty(10)
        if lib().orc_fn1(self.h, C.c_double(w), C.c_double(z), _p(o)):
            raise AssertionError(_err())
        return o


def cps_2d(kv1, kv2, f):
    z1, z2 = kv1.collocate(), kv2.collocate()
    x = np.array([[f(a, b) for a in z1] for b in z2], dtype=np.float64).ravel()  # index j + k*num1
    o = np.empty(len(x))
    if lib().orc_cps_2d(kv1.h, kv2.h, _p(x), len(x), _p(o)):
        raise AssertionError(_err())
    return o


def fn2(a10, b10):
    o = np.empty(55)
    a10 = np.ascontiguousarray(a10)
    b10 = np.ascontiguousarray(b10)
    lib().orc_fn2(_p(a10), _p(b10), _p(o))
    return {"w": o[0], "N": o[1:10].copy(), "dN": o[10:28].reshape(2, 9).T.copy(),
            "ddN": o[28:55].reshape(3, 9).T.copy()}


def gauss_xi(ngp):
    xs, ws = np.empty(ngp), np.empty(ngp)
    if lib().orc_gauss_xi(ngp, _p(xs), _p(ws)):
        raise AssertionError(_err())
    return xs, ws


def gauss_zeta(ngp, lo, hi):
    zs, ws = np.empty(ngp), np.empty(ngp)
    if lib().orc_gauss_zeta(ngp, C.c_double(lo), C.c_double(hi), _p(zs), _p(ws)):
        raise AssertionError(_err())
    return zs, ws


def fine_zs(nel, poly=2):
    o = np.empty(nel + 2 * poly + 1)
    if lib().orc_fine_zs(nel, poly, _p(o)):
        raise AssertionError(_err())
    return o


class Mesh:
    """Oracle mesh (reference `Mesh(p; args...)`, src/input/Mesh.jl:48-80)."""

    def __init__(self, motion=ALEVB, scenario=F_PULL, num1el=17, num2el=17, length=64.0, kb=1.0, kg=-0.5, zv=1.0,
                 pn=0.0, adb=None, am=1.0, ek=1e-15, pull_speed=0.0, bend_mf=0.0, bend_tm=1.0, kind="oracle"):
        self.L = lib(kind)      # "oracle" (double), "truth" (long double), "truthq" (__float128), "native"
        self.kind = kind
        adb = length ** 2 if adb is None else adb
        self.params = OrcParams(motion, scenario, num1el, num2el, length, kb, kg, zv, pn, adb, am, ek, pull_speed,
                                bend_mf, bend_tm)
        h = self.L.orc_mesh_create(C.byref(self.params))
        if not h:
            raise AssertionError(self.L.orc_last_error().decode())
        self.h = C.c_void_p(h)
        s = np.empty(14, dtype=np.int64)
        self.L.orc_mesh_sizes(self.h, _p(s, C.c_int64))
        (self.numel, self.numnp, self.ndf, self.nmdf, self.num1el, self.num2el, self.num1np, self.num2np,
         self.nuel1, self.nuel2, self.n_dir, self.n_neu, self.nk1, self.nk2) = [int(v) for v in s]
        self.motion, self.scenario = motion, scenario

    def __del__(self):
        try:
            self.L.orc_mesh_destroy(self.h)
        except Exception:
            pass

    def _i64(self, fn, shape, *args):
        o = np.empty(int(np.prod(shape)), dtype=np.int64)
        getattr(self.L, fn)(self.h, *args, _p(o, C.c_int64))
        return o.reshape(shape, order="F")

    @property
    def dofs(self):
        o = np.empty(8, dtype=np.int32)
        self.L.orc_mesh_dofs(self.h, _p(o, C.c_int32))
        return o

    @property
    def IX(self):
        return self._i64("orc_mesh_IX", (9, self.numel))

    @property
    def ID(self):
        return self._i64("orc_mesh_ID", (self.ndf, self.numnp))

    @property
    def LM(self):
        return self._i64("orc_mesh_LM", (9 * self.ndf, self.numel))

    @property
    def ID_inv(self):
        n, d = np.empty(self.nmdf, dtype=np.int64), np.empty(self.nmdf, dtype=np.int64)
        self.L.orc_mesh_ID_inv(self.h, _p(n, C.c_int64), _p(d, C.c_int64))
        return n, d

    def kv(self, d):
        return KnotVector(self.L.orc_kv_of_mesh(self.h, d))

    def line(self, d):
        nel, nuel = (self.num1el, self.nuel1) if d == 1 else (self.num2el, self.nuel2)
        ids = np.empty(nel, dtype=np.int64)
        tab = np.empty((nuel, 3, 10))
        edge = np.empty((2, 10))
        self.L.orc_mesh_line(self.h, d, _p(ids, C.c_int64), _p(tab), _p(edge))
        return ids, tab, edge

    def area_fns(self, el, gp):
        o = np.empty(55)
        self.L.orc_mesh_area_fns(self.h, C.c_int64(el), gp, _p(o))
        return {"w": o[0], "N": o[1:10].copy(), "dN": o[10:28].reshape(2, 9).T.copy(),
                "ddN": o[28:55].reshape(3, 9).T.copy()}

    def bdry_fns(self, bdry, el, gp):
        o = np.empty(55)
        if self.L.orc_mesh_bdry_fns(self.h, bdry, C.c_int64(el), gp, _p(o)):
            raise AssertionError(_err())
        return {"w": o[0], "N": o[1:10].copy(), "dN": o[10:28].reshape(2, 9).T.copy(),
                "ddN": o[28:55].reshape(3, 9).T.copy()}

    @property
    def area_uel_ids(self):
        return self._i64("orc_mesh_area_uel_ids", (self.numel,))

    def bdry_elems(self, bdry):
        n = self.L.orc_mesh_bdry_count(self.h, bdry)
        o = np.empty(n, dtype=np.int64)
        self.L.orc_mesh_bdry_elems(self.h, bdry, _p(o, C.c_int64))
        return o

    def bdry_nodes(self, bdry):
        n = self.num1np if bdry in (BOTTOM, TOP) else self.num2np
        a, b = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        self.L.orc_mesh_bdry_nodes(self.h, bdry, _p(a, C.c_int64), _p(b, C.c_int64))
        return a, b

    @property
    def bcs(self):
        du, dn, dv = np.empty(self.n_dir, np.int32), np.empty(self.n_dir, np.int64), np.empty(self.n_dir)
        nb, nt, nv = np.empty(self.n_neu, np.int32), np.empty(self.n_neu, np.int32), np.empty(self.n_neu)
        self.L.orc_mesh_bcs(self.h, _p(du, C.c_int32), _p(dn, C.c_int64), _p(dv), _p(nb, C.c_int32),
                           _p(nt, C.c_int32), _p(nv))
        return list(zip(du.tolist(), dn.tolist(), dv.tolist())), list(zip(nb.tolist(), nt.tolist(), nv.tolist()))

    # ---- state -------------------------------------------------------------------------------
    def flat_state(self):
        """prepare_input (src/Input.jl:77-108) for the non-F_BEND scenarios: flat patch, lambda = kb/4,
        inhomogeneous Dirichlet values applied. Control points by tensor-product 1-D collocation."""
        p = self.params
        kv1, kv2 = self.kv(1), self.kv(2)
        if self.scenario == F_BEND:
            x1 = kv1.cps_1d(lambda z: p.length * z)
            y1 = kv2.cps_1d(lambda z: p.length * z)
        else:
            x1 = kv1.cps_1d(lambda z: p.length * (z - 0.5))
            y1 = kv2.cps_1d(lambda z: p.length * (z - 0.5))
        xms = np.zeros((self.numnp, 3), order="F")
        xms[:, 0] = np.tile(x1, self.num2np)
        xms[:, 1] = np.repeat(y1, self.num1np)
        cps = np.zeros((self.numnp, self.ndf), order="F")
        d = self.dofs
        if self.scenario != F_BEND:
            cps[:, d[6] - 1] = p.kb / 4
        dirs, _ = self.bcs
        for (unk, node, val) in dirs:
            cps[node - 1, d[unk - 1] - 1] = val
            if self.scenario == F_PULL and self.motion == EUL:
                cps[node - 1, d[5] - 1] = val
        return xms, cps

    def geo_dyn_stress(self, el, gp, xms_el, cps_el):
        xe = np.asfortranarray(xms_el, dtype=np.float64)
        ce = np.asfortranarray(cps_el, dtype=np.float64)
        o = np.empty(44)
        self.L.orc_geo_dyn_stress(self.h, C.c_int64(el), gp, _p(xe), _p(ce), _p(o))
        k = [0]

        def take(n, shape=None):
            v = o[k[0]:k[0] + n].copy()
            k[0] += n
            return v if shape is None else v.reshape(shape, order="F")
        return {"x": take(3), "a_": take(6, (3, 2)), "acon": take(4, (2, 2)), "aco": take(4, (2, 2)),
                "J": take(1)[0], "n": take(3), "b": take(4, (2, 2)), "H": take(1)[0], "K": take(1)[0],
                "sig": take(3), "sigm": take(3), "M": take(3), "lam": take(1)[0], "pm": take(1)[0],
                "v": take(3), "vm": take(3)}

    def elem_r_K(self, el, xms, cps, dt):
        nd = 9 * self.ndf
        r, K = np.empty(nd), np.empty((nd, nd), order="F")
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        if self.L.orc_elem_r_K(self.h, C.c_int64(el), _p(xms), _p(cps), C.c_double(dt), _p(r), _p(K)):
            raise AssertionError(_err())
        return r, K

    def elem_dof_residuals(self, el, xms, cps):
        rv, rm, rl, rp = np.empty(27), np.empty(27), np.empty(9), np.empty(9)
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        if self.L.orc_elem_dof_residuals(self.h, C.c_int64(el), _p(xms), _p(cps), _p(rv), _p(rm), _p(rl), _p(rp)):
            raise AssertionError(_err())
        return rv, rm, rl, rp

    def calc_pull_force(self, xms, cps, poly=2):
        """calc_pull_force + get_adj_maps (PullForce.jl:25-80): reaction force on the nodes of the pulled element
        = sum of rv over the (2 poly + 1)^2 elements around it, restricted to those nodes."""
        IX = self.IX
        numel = IX.shape[1]
        num1el = self.num1el
        pull_el = int(np.ceil(numel / 2))                               # get_pull_el_id, PullForce.jl:12
        pull_nodes = list(IX[:, pull_el - 1])
        rv_pull = np.zeros(27)
        for i in range(-poly, poly + 1):                                # adj_el_ids, PullForce.jl:36-37
            for e in range(pull_el + i * num1el - poly, pull_el + i * num1el + poly + 1):
                adj_nodes = list(IX[:, e - 1])
                rv_el = self.elem_dof_residuals(e, xms, cps)[0]
                for ip, n in enumerate(pull_nodes):                     # maps of PullForce.jl:41-49
                    if n in adj_nodes:
                        ia = adj_nodes.index(n)
                        rv_pull[3 * ip:3 * ip + 3] += rv_el[3 * ia:3 * ia + 3]
        return rv_pull.reshape(9, 3).sum(axis=0)                        # PullForce.jl:79

    def calc_r_K(self, xms, cps, time, dt, nthreads=1):
        """Reference calc_r_K (FiniteElement.jl:75-200). Returns (r, scipy CSC K) with the value-dependent
        stored pattern Julia's SparseMatrixCSC would hold (explicit zeros kept, never-nonzero entries absent)."""
        import scipy.sparse as sp
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        res = self.L.orc_calc_r_K(self.h, _p(xms), _p(cps), C.c_double(time), C.c_double(dt), nthreads)
        if not res:
            raise AssertionError(_err())
        res = C.c_void_p(res)
        nnz = self.L.orc_result_nnz(res)
        r = np.empty(self.nmdf)
        colptr = np.empty(self.nmdf + 1, dtype=np.int64)
        rowval = np.empty(nnz, dtype=np.int64)
        nzval = np.empty(nnz)
        self.L.orc_result_get(res, _p(r), _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval))
        self.L.orc_result_destroy(res)
        K = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(self.nmdf, self.nmdf))
        return r, K

    def calc_r_K_fast(self, xms, cps, time, dt, colptr0, rowval0, nthreads=1, e_first=1, e_last=None,
                      with_neumann=True, want_out=True):
        """CPU-baseline mode: same element algorithm/threading, accumulating into a given 0-based CSC pattern."""
        e_last = self.numel if e_last is None else e_last
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        colptr0 = np.ascontiguousarray(colptr0, dtype=np.int64)
        rowval0 = np.ascontiguousarray(rowval0, dtype=np.int64)
        nnz = len(rowval0)
        r = np.empty(self.nmdf) if want_out else None
        nz = np.empty(nnz) if want_out else None
        rc = self.L.orc_calc_r_K_fast(self.h, _p(xms), _p(cps), C.c_double(time), C.c_double(dt), nthreads,
                                     _p(colptr0, C.c_int64), _p(rowval0, C.c_int64), C.c_int64(nnz),
                                     C.c_int64(e_first), C.c_int64(e_last), int(with_neumann),
                                     _p(r) if want_out else None, _p(nz) if want_out else None)
        if rc:
            raise AssertionError(_err())
        return r, nz

    # ---- extended-precision truth (kind="truth" / "truthq") and the untested Neumann branches ---------------
    def set_neumann(self, conds):
        """Replace mesh.inh_neu_bcs by [(bdry, type, value), ...] (tests of the SHEAR / TOP-BOTTOM MOMENT branches of
        calc_bdry_element_residual, FiniteElement.jl:374-380, which no scenario in Bc.jl sets up)."""
        nb = np.array([c[0] for c in conds], dtype=np.int32)
        nt = np.array([c[1] for c in conds], dtype=np.int32)
        nv = np.array([c[2] for c in conds], dtype=np.float64)
        if self.L.orc_mesh_set_neumann(self.h, len(conds), _p(nb, C.c_int32), _p(nt, C.c_int32), _p(nv)):
            raise AssertionError(self.L.orc_last_error().decode())
        self.n_neu = len(conds)

    def elem_r_K_mag(self, el, xms, cps, time, dt, bdry=0, ntype=0, nval=0.0):
        """(r_el, K_el, r_mag, K_mag) of an area element (bdry = 0) or of the Neumann boundary element `el` of
        boundary `bdry`; the magnitudes are sum |terms| per entry (zeros unless kind is "truth"/"truthq")."""
        nd = 9 * self.ndf
        r, K = np.empty(nd), np.empty((nd, nd), order="F")
        rm, Km = np.empty(nd), np.empty((nd, nd), order="F")
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        if self.L.orc_elem_r_K_mag(self.h, C.c_int64(el), int(bdry), int(ntype), C.c_double(nval), _p(xms), _p(cps),
                                   C.c_double(time), C.c_double(dt), _p(r), _p(K), _p(rm), _p(Km)):
            raise AssertionError(self.L.orc_last_error().decode())
        return r, K, rm, Km

    def calc_r_K_on_pattern(self, xms, cps, time, dt, colptr0, rowval0, nthreads=1):
        """r, nzval and the magnitudes sum |terms| (r_mag, nz_mag) accumulated into a given 0-based CSC pattern."""
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        colptr0 = np.ascontiguousarray(colptr0, dtype=np.int64)
        rowval0 = np.ascontiguousarray(rowval0, dtype=np.int64)
        nnz = len(rowval0)
        r, nz, rm, nzm = np.empty(self.nmdf), np.empty(nnz), np.empty(self.nmdf), np.empty(nnz)
        rc = self.L.orc_calc_r_K_fast_mag(self.h, _p(xms), _p(cps), C.c_double(time), C.c_double(dt), nthreads,
                                          _p(colptr0, C.c_int64), _p(rowval0, C.c_int64), C.c_int64(nnz),
                                          C.c_int64(1), C.c_int64(self.numel), 1, _p(r), _p(nz), _p(rm), _p(nzm))
        if rc:
            raise AssertionError(self.L.orc_last_error().decode())
        return r, nz, rm, nzm
