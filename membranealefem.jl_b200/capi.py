"""ctypes binding of the C ABI in include/maf.h (libmembrane_b200.so) -- the same calls the Julia `ccall` shim of
INTEGRATION.md makes. There is no CPU fallback: if the CUDA library is missing or no device is present, every entry
point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAF_LIB", os.path.join(_HERE, "csrc", "libmembrane_b200.so"))

PATTERN_BLK, PATTERN_SYM = 0, 1
SCATTER_ATOMIC, SCATTER_DETERMINISTIC = 0, 1

_I64P = C.POINTER(C.c_int64)
_I32P = C.POINTER(C.c_int32)
_F64P = C.POINTER(C.c_double)


class MeshDesc(C.Structure):
    """maf_mesh_desc"""
    _fields_ = [("numel", C.c_int64), ("numnp", C.c_int64), ("ndf", C.c_int64), ("nmdf", C.c_int64),
                ("num1el", C.c_int64), ("num2el", C.c_int64),
                ("IX", _I64P), ("ID", _I64P), ("LM", _I64P), ("dofs", C.c_int32 * 8),
                ("nuel1", C.c_int64), ("nuel2", C.c_int64), ("uel_ids1", _I64P), ("uel_ids2", _I64P),
                ("line1", _F64P), ("line2", _F64P), ("edge1", _F64P), ("edge2", _F64P), ("xi", C.c_double * 3),
                ("n_neu", C.c_int32), ("neu_bdry", _I32P), ("neu_type", _I32P), ("neu_val", _F64P),
                ("bdry_elems", _I64P * 4), ("bdry_count", C.c_int64 * 4)]


class ParamsC(C.Structure):
    """maf_params"""
    _fields_ = [("motion", C.c_int32), ("scenario", C.c_int32), ("kb", C.c_double), ("kg", C.c_double),
                ("zv", C.c_double), ("pn", C.c_double), ("adb", C.c_double), ("am", C.c_double),
                ("pattern_mode", C.c_int32), ("device", C.c_int32)]


class MafError(RuntimeError):
    pass


_lib = None

EXPORTS = ["maf_create", "maf_destroy", "maf_last_error", "maf_nnz", "maf_pattern", "maf_assemble",
           "maf_assemble_device", "maf_device_buffers", "maf_stream", "maf_sync", "maf_timings", "maf_launch_count",
           "maf_kernel_info", "maf_chunk_plan", "maf_set_element_range", "maf_range_info", "maf_fp64_peak",
           "maf_debug_phase_cycles", "maf_state_set", "maf_state_get", "maf_state_update", "maf_state_predict",
           "maf_assemble_resident", "maf_elem_v_residuals", "maf_host_register", "maf_host_unregister",
           "maf_colptr", "maf_pattern_columns", "maf_download", "maf_create_strip", "maf_strip_info",
           "maf_peer_attach_local", "maf_peer_export", "maf_peer_attach", "maf_assemble_strip",
           "maf_assemble_strip_host", "maf_strip_timings", "maf_area_kernel_times", "maf_generate_output"]


def load_library(path=None):
    """Load libmembrane_b200.so and declare the signatures. Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise MafError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the assembly)")
    L = C.CDLL(path)
    L.maf_last_error.restype = C.c_char_p
    L.maf_last_error.argtypes = [C.c_void_p]
    L.maf_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MeshDesc), C.POINTER(ParamsC)]
    L.maf_destroy.argtypes = [C.c_void_p]
    L.maf_nnz.argtypes = [C.c_void_p, _I64P]
    L.maf_pattern.argtypes = [C.c_void_p, _I64P, _I64P]
    L.maf_create_strip.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MeshDesc), C.POINTER(ParamsC), C.c_int32, C.c_int32]
    L.maf_strip_info.argtypes = [C.c_void_p, _I64P]
    L.maf_peer_attach_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.maf_peer_export.argtypes = [C.c_void_p, C.c_void_p]
    L.maf_peer_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.maf_assemble_strip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                                     C.c_void_p]
    L.maf_assemble_strip_host.argtypes = [C.c_void_p, _F64P, _F64P, C.c_double, C.c_double, C.c_double, C.c_int,
                                          _F64P, _F64P, _F64P]
    L.maf_strip_timings.argtypes = [C.c_void_p, _F64P]
    L.maf_generate_output.argtypes = [C.c_void_p, _F64P, _F64P]
    L.maf_colptr.argtypes = [C.c_void_p, _I64P]
    L.maf_pattern_columns.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _I64P]
    L.maf_download.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _F64P, C.c_int64, C.c_int64, _F64P]
    L.maf_assemble.argtypes = [C.c_void_p, _F64P, _F64P, C.c_double, C.c_double, C.c_double, C.c_int, _F64P, _F64P,
                               _F64P]
    L.maf_assemble_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.maf_device_buffers.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 5
    L.maf_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.maf_sync.argtypes = [C.c_void_p]
    L.maf_timings.argtypes = [C.c_void_p, _F64P]
    L.maf_launch_count.argtypes = [C.c_void_p, _I64P]
    L.maf_area_kernel_times.argtypes = [C.c_void_p, _F64P, C.c_int64]
    L.maf_kernel_info.argtypes = [C.c_void_p, _I64P]
    L.maf_chunk_plan.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.maf_host_register.argtypes = [C.c_void_p, C.c_int64]
    L.maf_host_unregister.argtypes = [C.c_void_p]
    L.maf_set_element_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    L.maf_range_info.argtypes = [C.c_void_p, _I64P]
    L.maf_fp64_peak.argtypes = [C.c_int, _F64P]
    L.maf_state_set.argtypes = [C.c_void_p, _F64P, _F64P]
    L.maf_state_get.argtypes = [C.c_void_p, _F64P, _F64P]
    L.maf_state_update.argtypes = [C.c_void_p, _F64P, C.c_double]
    L.maf_state_predict.argtypes = [C.c_void_p, C.c_double]
    L.maf_assemble_resident.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, _F64P, _F64P, _F64P]
    L.maf_elem_v_residuals.argtypes = [C.c_void_p, _I64P, C.c_int64, _F64P]
    if path == LIB_PATH:
        _lib = L
    return L


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_mesh_desc(mesh, keep):
    """Fill a maf_mesh_desc from a host `Mesh` (host/mesh.py). `keep` collects the arrays that must stay alive."""
    d = MeshDesc()
    d.numel, d.numnp, d.ndf, d.nmdf = mesh.numel, mesh.numnp, mesh.ndf, mesh.nmdf
    d.num1el, d.num2el = mesh.num1el, mesh.num2el
    IX = np.asfortranarray(mesh.IX, dtype=np.int64)
    ID = np.asfortranarray(mesh.ID, dtype=np.int64)
    keep += [IX, ID]
    d.IX, d.ID, d.LM = _ptr(IX, C.c_int64), _ptr(ID, C.c_int64), None
    d.dofs = (C.c_int32 * 8)(*[int(v) for v in mesh.dofs8()])
    l1, l2 = mesh.line_gp_fns1, mesh.line_gp_fns2
    u1 = np.ascontiguousarray(l1.uel_ids, dtype=np.int64)
    u2 = np.ascontiguousarray(l2.uel_ids, dtype=np.int64)
    t1 = np.ascontiguousarray(l1.ufns, dtype=np.float64)
    t2 = np.ascontiguousarray(l2.ufns, dtype=np.float64)
    e1 = np.ascontiguousarray(l1.edge, dtype=np.float64)
    e2 = np.ascontiguousarray(l2.edge, dtype=np.float64)
    keep += [u1, u2, t1, t2, e1, e2]
    d.nuel1, d.nuel2 = t1.shape[0], t2.shape[0]
    d.uel_ids1, d.uel_ids2 = _ptr(u1, C.c_int64), _ptr(u2, C.c_int64)
    d.line1, d.line2, d.edge1, d.edge2 = (_ptr(t1, C.c_double), _ptr(t2, C.c_double), _ptr(e1, C.c_double),
                                          _ptr(e2, C.c_double))
    from .host.basis import GaussPointsXi
    d.xi = (C.c_double * 3)(*GaussPointsXi(3).xs.tolist())
    nb = np.array([int(b) for (b, _, _) in mesh.inh_neu_bcs], dtype=np.int32)
    nt = np.array([int(t) for (_, t, _) in mesh.inh_neu_bcs], dtype=np.int32)
    nv = np.array([float(v) for (_, _, v) in mesh.inh_neu_bcs], dtype=np.float64)
    keep += [nb, nt, nv]
    d.n_neu = len(nb)
    d.neu_bdry, d.neu_type, d.neu_val = _ptr(nb, C.c_int32), _ptr(nt, C.c_int32), _ptr(nv, C.c_double)
    for b in range(1, 5):
        arr = np.ascontiguousarray(mesh.bdry_elems[b], dtype=np.int64)
        keep.append(arr)
        d.bdry_elems[b - 1] = _ptr(arr, C.c_int64)
        d.bdry_count[b - 1] = len(arr)
    return d


def host_register(arr, lib=None):
    """Page-lock a numpy array the caller owns (maf_host_register): maf_assemble then copies from / into it directly
    at the full PCIe rate. Pair with host_unregister before the array is freed."""
    L = lib or load_library()
    if L.maf_host_register(arr.ctypes.data_as(C.c_void_p), arr.nbytes) != 0:
        raise MafError(L.maf_last_error(None).decode())


def host_unregister(arr, lib=None):
    L = lib or load_library()
    if L.maf_host_unregister(arr.ctypes.data_as(C.c_void_p)) != 0:
        raise MafError(L.maf_last_error(None).decode())


class Assembler:
    """Owner of one maf_handle: the device-resident replacement of the reference's calc_r_K for one (mesh, Params)."""

    def __init__(self, mesh, p, pattern_mode=PATTERN_BLK, device=-1, lib=None, strip=None):
        """strip = (rank, nranks): a strip handle (maf_create_strip) that holds only the slices of one strip of element
        rows; None: the whole mesh on one device (maf_create)."""
        self.L = lib or load_library()
        self.mesh, self.p = mesh, p
        keep = []
        desc = make_mesh_desc(mesh, keep)
        par = ParamsC(int(p.motion), int(p.scenario), p.kb, p.kg, p.zv, p.pn, p.adb, p.am, pattern_mode, device)
        h = C.c_void_p()
        self.strip = strip
        if strip is None:
            rc = self.L.maf_create(C.byref(h), C.byref(desc), C.byref(par))
        else:
            rc = self.L.maf_create_strip(C.byref(h), C.byref(desc), C.byref(par), int(strip[0]), int(strip[1]))
        if rc != 0:
            raise MafError(f"maf_create failed ({rc}): {self.L.maf_last_error(None).decode()}")
        self.h = h
        n = C.c_int64()
        self._check(self.L.maf_nnz(self.h, C.byref(n)))
        self.nnz = n.value
        self.nmdf = mesh.nmdf
        self._pattern = None

    def _check(self, rc):
        if rc != 0:
            raise MafError(f"libmembrane_b200 error {rc}: {self.L.maf_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.maf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pattern(self):
        """(colptr, rowval) of K, 1-based Int64 like SparseMatrixCSC."""
        if self._pattern is None:
            colptr = np.empty(self.nmdf + 1, dtype=np.int64)
            rowval = np.empty(self.nnz, dtype=np.int64)
            self._check(self.L.maf_pattern(self.h, _ptr(colptr, C.c_int64), _ptr(rowval, C.c_int64)))
            self._pattern = (colptr, rowval)
        return self._pattern

    def colptr(self):
        """colptr alone (1-based Int64, nmdf + 1)."""
        colptr = np.empty(self.nmdf + 1, dtype=np.int64)
        self._check(self.L.maf_colptr(self.h, _ptr(colptr, C.c_int64)))
        return colptr

    def pattern_columns(self, col_first, col_last, colptr=None):
        """rowval (1-based) of the columns [col_first, col_last] only (maf_pattern_columns)."""
        colptr = self.colptr() if colptr is None else colptr
        n = int(colptr[col_last] - colptr[col_first - 1])
        rowval = np.empty(n, dtype=np.int64)
        self._check(self.L.maf_pattern_columns(self.h, col_first, col_last, _ptr(rowval, C.c_int64)))
        return rowval

    def download(self, r_first=1, r_count=0, nz_first=1, nz_count=0, r_out=None, nz_out=None):
        """Slices (1-based starts) of the handle's device-resident r / nzval after assemble_device; r_out / nz_out:
        caller-owned (e.g. page-locked) destination arrays."""
        r = np.empty(r_count) if r_out is None else r_out
        nz = np.empty(nz_count) if nz_out is None else nz_out
        assert r.size >= r_count and nz.size >= nz_count
        self._check(self.L.maf_download(self.h, r_first, r_count, _ptr(r, C.c_double), nz_first, nz_count,
                                        _ptr(nz, C.c_double)))
        return r, nz

    def assemble(self, xms, cps, time, dt, bend_tm=1.0, scatter_mode=SCATTER_ATOMIC, r=None, nzval=None):
        """maf_assemble with host buffers: returns (r, nzval, sum(r^2))."""
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        assert xms.shape == (self.mesh.numnp, 3) and cps.shape == (self.mesh.numnp, self.mesh.ndf)
        r = np.empty(self.nmdf) if r is None else r
        nzval = np.empty(self.nnz) if nzval is None else nzval
        rn = C.c_double()
        self._check(self.L.maf_assemble(self.h, _ptr(xms, C.c_double), _ptr(cps, C.c_double), time, dt, bend_tm,
                                        scatter_mode, _ptr(r, C.c_double), _ptr(nzval, C.c_double), C.byref(rn)))
        return r, nzval, rn.value

    # ---- device-resident state (include/maf.h: maf_state_*) ----
    def state_set(self, xms, cps):
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        assert xms.shape == (self.mesh.numnp, 3) and cps.shape == (self.mesh.numnp, self.mesh.ndf)
        self._check(self.L.maf_state_set(self.h, _ptr(xms, C.c_double), _ptr(cps, C.c_double)))

    def state_get(self):
        xms = np.empty((self.mesh.numnp, 3), order="F")
        cps = np.empty((self.mesh.numnp, self.mesh.ndf), order="F")
        self._check(self.L.maf_state_get(self.h, _ptr(xms, C.c_double), _ptr(cps, C.c_double)))
        return xms, cps

    def state_update(self, du, dt):
        du = np.ascontiguousarray(du, dtype=np.float64)
        assert du.shape == (self.nmdf,)
        self._check(self.L.maf_state_update(self.h, _ptr(du, C.c_double), dt))

    def state_predict(self, dt):
        self._check(self.L.maf_state_predict(self.h, dt))

    def assemble_resident(self, time, dt, bend_tm=1.0, scatter_mode=SCATTER_ATOMIC, r=None, nzval=None):
        """maf_assemble_resident: like assemble(), on the state the device already holds."""
        r = np.empty(self.nmdf) if r is None else r
        nzval = np.empty(self.nnz) if nzval is None else nzval
        rn = C.c_double()
        self._check(self.L.maf_assemble_resident(self.h, time, dt, bend_tm, scatter_mode, _ptr(r, C.c_double),
                                                 _ptr(nzval, C.c_double), C.byref(rn)))
        return r, nzval, rn.value

    def generate_output(self):
        """generate_output (Output.jl:32-118) of the resident state: (xout[n1, n2, 3], uout[n1, n2, ndf])."""
        n1, n2 = 3 * self.mesh.num1el + 2, 3 * self.mesh.num2el + 2
        xout = np.empty((n1, n2, 3), order="F")
        uout = np.empty((n1, n2, self.mesh.ndf), order="F")
        self._check(self.L.maf_generate_output(self.h, _ptr(xout, C.c_double), _ptr(uout, C.c_double)))
        return xout, uout

    def elem_v_residuals(self, el_ids):
        """rv of calc_elem_dof_residuals for the listed elements (1-based) on the resident state: (n, 27)."""
        el_ids = np.ascontiguousarray(el_ids, dtype=np.int64)
        rv = np.empty((el_ids.size, 27))
        self._check(self.L.maf_elem_v_residuals(self.h, _ptr(el_ids, C.c_int64), el_ids.size, _ptr(rv, C.c_double)))
        return rv

    def assemble_device(self, d_xms, d_cps, time, dt, bend_tm=1.0, scatter_mode=SCATTER_ATOMIC, d_r=None,
                        d_nzval=None, d_rnorm2=None, stream=None):
        """maf_assemble_device: raw device pointers (ints); None selects the handle's own buffers."""
        self._check(self.L.maf_assemble_device(self.h, d_xms, d_cps, time, dt, bend_tm, scatter_mode, d_r, d_nzval,
                                               d_rnorm2, stream))

    def device_buffers(self):
        ps = [C.c_void_p() for _ in range(5)]
        self._check(self.L.maf_device_buffers(self.h, *[C.byref(p) for p in ps]))
        return [p.value for p in ps]

    def stream(self):
        s = C.c_void_p()
        self._check(self.L.maf_stream(self.h, C.byref(s)))
        return s.value

    def sync(self):
        self._check(self.L.maf_sync(self.h))

    def timings(self):
        o = np.zeros(7)
        self._check(self.L.maf_timings(self.h, _ptr(o, C.c_double)))
        return dict(zip(["h2d_ms", "area_ms", "bdry_ms", "gather_ms", "d2h_ms", "total_ms", "zero_ms"], o.tolist()))

    def area_kernel_times(self, n):
        """Device ms of the area kernel in each of the last n assemblies (oldest first, n <= 64)."""
        o = np.zeros(int(n))
        self._check(self.L.maf_area_kernel_times(self.h, _ptr(o, C.c_double), int(n)))
        return o

    def launch_count(self):
        n = C.c_int64()
        self._check(self.L.maf_launch_count(self.h, C.byref(n)))
        return n.value

    def kernel_info(self):
        o = np.zeros(9, dtype=np.int64)
        self._check(self.L.maf_kernel_info(self.h, _ptr(o, C.c_int64)))
        return dict(zip(["threads_per_cta", "elements_per_cta", "smem_bytes", "ctas_per_sm", "sm_count",
                         "scatter_map_classes", "graph_replays", "staging_bytes", "band_rows"], o.tolist()))

    def chunk_plan(self):
        """Plan of the contraction phase, "c,c/c,c/..." = chunk ids per warp in execution order."""
        buf = C.create_string_buffer(1024)
        self._check(self.L.maf_chunk_plan(self.h, buf, 1024))
        return buf.value.decode()

    # ---- strips over several GPUs (include/maf.h: maf_create_strip ...) ----
    def strip_info(self):
        o = np.zeros(12, dtype=np.int64)
        self._check(self.L.maf_strip_info(self.h, _ptr(o, C.c_int64)))
        return {"elements": (int(o[0]), int(o[1])), "rows": (int(o[2]), int(o[3])), "slots": (int(o[4]), int(o[5])),
                "own_rows": (int(o[6]), int(o[7])), "own_slots": (int(o[8]), int(o[9])), "rank": int(o[10]),
                "nranks": int(o[11])}

    def peer_attach_local(self, lower, upper):
        self._check(self.L.maf_peer_attach_local(self.h, lower.h if lower is not None else None,
                                                 upper.h if upper is not None else None))

    def peer_export(self):
        buf = (C.c_ubyte * 64)()
        self._check(self.L.maf_peer_export(self.h, buf))
        return bytes(buf)

    def peer_attach(self, lower64, upper64):
        lo = (C.c_ubyte * 64).from_buffer_copy(lower64) if lower64 is not None else None
        up = (C.c_ubyte * 64).from_buffer_copy(upper64) if upper64 is not None else None
        self._check(self.L.maf_peer_attach(self.h, lo, up))

    def assemble_strip(self, d_xms, d_cps, time, dt, bend_tm=1.0, scatter_mode=SCATTER_ATOMIC, d_rnorm2_partial=None):
        self._check(self.L.maf_assemble_strip(self.h, d_xms, d_cps, time, dt, bend_tm, scatter_mode, d_rnorm2_partial))

    def assemble_strip_host(self, xms, cps, time, dt, bend_tm=1.0, scatter_mode=SCATTER_ATOMIC, r_own=None,
                            nzval_own=None):
        """maf_assemble_strip_host: returns (r_own, nzval_own, partial sum(r^2) over the owned rows)."""
        info = self.strip_info()
        nr = info["own_rows"][1] - info["own_rows"][0] + 1
        nk = info["own_slots"][1] - info["own_slots"][0] + 1
        r_own = np.empty(nr) if r_own is None else r_own
        nzval_own = np.empty(nk) if nzval_own is None else nzval_own
        rn = C.c_double()
        xp = _ptr(np.asfortranarray(xms, dtype=np.float64), C.c_double) if xms is not None else None
        cp = _ptr(np.asfortranarray(cps, dtype=np.float64), C.c_double) if cps is not None else None
        self._check(self.L.maf_assemble_strip_host(self.h, xp, cp, time, dt, bend_tm, scatter_mode,
                                                   _ptr(r_own, C.c_double), _ptr(nzval_own, C.c_double), C.byref(rn)))
        return r_own, nzval_own, rn.value

    def strip_timings(self):
        o = np.zeros(2)
        self._check(self.L.maf_strip_timings(self.h, _ptr(o, C.c_double)))
        return {"exchange_ms": float(o[0]), "result_bytes": int(o[1])}

    def set_element_range(self, el_first, el_last):
        self._check(self.L.maf_set_element_range(self.h, el_first, el_last))

    def range_info(self):
        """1-based inclusive ranges this handle's element range touches."""
        o = np.zeros(8, dtype=np.int64)
        self._check(self.L.maf_range_info(self.h, _ptr(o, C.c_int64)))
        return {"elements": (int(o[0]), int(o[1])), "nodes": (int(o[2]), int(o[3])), "rows": (int(o[4]), int(o[5])),
                "slots": (int(o[6]), int(o[7]))}


def fp64_peak_tflops(device=-1):
    """Measured DFMA throughput of the device (roofline denominator)."""
    L = load_library()
    t = C.c_double()
    if L.maf_fp64_peak(device, C.byref(t)) != 0:
        raise MafError(L.maf_last_error(None).decode())
    return t.value
