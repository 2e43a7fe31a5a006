"""Test-only driver of tests/emu/libmaf_emu.so: the library's kernel phase functions executed on the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

import mafb200

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "emu", "maf_emu.cpp")
_SO = os.path.join(_HERE, "emu", "libmaf_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        csrc = os.path.join(_ROOT, "membranealefem.jl_b200", "csrc")
        deps = [_SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".h", ".cuh"))]
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                                   "-o", _SO, _SRC])
        L = C.CDLL(_SO)
        L.emu_create.restype = C.c_void_p
        L.emu_last_error.restype = C.c_char_p
        L.emu_nnz.restype = C.c_int64
        _lib = L
    return _lib


class Emu:
    def __init__(self, mesh, p, pattern_mode=0, nthreads=128):
        capi = mafb200.pkg.capi
        keep = []
        desc = capi.make_mesh_desc(mesh, keep)
        par = capi.ParamsC(int(p.motion), int(p.scenario), p.kb, p.kg, p.zv, p.pn, p.adb, p.am, pattern_mode, -1)
        h = lib().emu_create(C.byref(desc), C.byref(par), nthreads)
        if not h:
            raise RuntimeError(lib().emu_last_error().decode())
        self.h = C.c_void_p(h)
        self.mesh = mesh
        self.nnz = lib().emu_nnz(self.h)

    def __del__(self):
        try:
            lib().emu_destroy(self.h)
        except Exception:
            pass

    def info(self):
        o = np.zeros(7, dtype=np.int64)
        lib().emu_info(self.h, o.ctypes.data_as(C.POINTER(C.c_int64)))
        return dict(zip(["asize", "smem_doubles", "ntasks", "nitems", "item_rounds", "task_rounds", "nblocks"],
                        o.tolist()))

    def strip_range(self, rank, nranks):
        """The library's own strip arithmetic (maf_host.h::strip_range): 1-based inclusive (elements, rows, slots)."""
        o = np.zeros(6, dtype=np.int64)
        if lib().emu_strip_range(self.h, rank, nranks, o.ctypes.data_as(C.POINTER(C.c_int64))):
            raise RuntimeError(lib().emu_last_error().decode())
        return (int(o[0]), int(o[1])), (int(o[2]), int(o[3])), (int(o[4]), int(o[5]))

    def chunks(self):
        """Tangent schedule: ([(f, g, kind, fused, first, count)], slots[round][warp])."""
        c6 = np.zeros(6 * 48, dtype=np.int32)
        sl = np.full(16 * 8, -1, dtype=np.int32)
        n = lib().emu_chunks(self.h, c6.ctypes.data_as(C.POINTER(C.c_int32)), sl.ctypes.data_as(C.POINTER(C.c_int32)))
        rounds = self.info()["task_rounds"]
        return [tuple(c6[6 * k:6 * k + 6].tolist()) for k in range(n)], sl[:rounds * 4].reshape(rounds, 4).tolist()

    def state_update(self, du, dt, xms, cps):
        dp = C.POINTER(C.c_double)
        du = np.ascontiguousarray(du, dtype=np.float64)
        lib().emu_state_update(self.h, du.ctypes.data_as(dp), C.c_double(dt), xms.ctypes.data_as(dp),
                               cps.ctypes.data_as(dp))

    def state_predict(self, dt, xms, cps):
        dp = C.POINTER(C.c_double)
        lib().emu_state_predict(self.h, C.c_double(dt), xms.ctypes.data_as(dp), cps.ctypes.data_as(dp))

    def elem_v_residuals(self, xms, cps, el_ids):
        dp = C.POINTER(C.c_double)
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        el_ids = np.ascontiguousarray(el_ids, dtype=np.int64)
        rv = np.empty((el_ids.size, 27))
        rc = lib().emu_elem_v_residuals(self.h, xms.ctypes.data_as(dp), cps.ctypes.data_as(dp),
                                        el_ids.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(el_ids.size),
                                        rv.ctypes.data_as(dp))
        if rc:
            raise RuntimeError(lib().emu_last_error().decode())
        return rv

    def pattern(self):
        colptr = np.empty(self.mesh.nmdf + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64)
        lib().emu_pattern(self.h, colptr.ctypes.data_as(C.POINTER(C.c_int64)),
                          rowval.ctypes.data_as(C.POINTER(C.c_int64)))
        return colptr, rowval

    def assemble(self, xms, cps, time, dt, bend_tm=1.0, mode=0, el_first=1, el_last=None):
        xms = np.asfortranarray(xms, dtype=np.float64)
        cps = np.asfortranarray(cps, dtype=np.float64)
        r = np.empty(self.mesh.nmdf)
        nz = np.empty(self.nnz)
        dp = C.POINTER(C.c_double)
        rc = lib().emu_assemble(self.h, xms.ctypes.data_as(dp), cps.ctypes.data_as(dp), C.c_double(time),
                                C.c_double(dt), C.c_double(bend_tm), int(mode), C.c_int64(el_first),
                                C.c_int64(self.mesh.numel if el_last is None else el_last), r.ctypes.data_as(dp),
                                nz.ctypes.data_as(dp))
        if rc:
            raise RuntimeError(lib().emu_last_error().decode())
        colptr, rowval = self.pattern()
        K = sp.csc_matrix((nz, rowval - 1, colptr - 1), shape=(self.mesh.nmdf, self.mesh.nmdf))
        return r, K
