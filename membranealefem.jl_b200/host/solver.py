"""Host-solve hand-off on the library's FIXED sparsity pattern.

The reference solves `Δu = -K_gl \\ r_gl` with a freshly built SparseMatrixCSC in every Newton iteration
(FiniteElement.jl:35-38): UMFPACK redoes the fill-reducing ordering and the symbolic analysis each time although the
pattern of K never changes within a run. The library's pattern is symbolic (maf_pattern, once); what changes per
iteration is `nzval` only. PatternSolver exploits that on the host side:

  * `nzval` is ONE buffer owned by the solver, page-locked once (maf_host_register) and handed to maf_assemble /
    maf_assemble_resident as the output array: the device-to-host copy lands directly in the matrix the solver
    factorises -- no intermediate array, no per-iteration allocation of a 12-bytes-per-entry CSC;
  * the fill-reducing column ordering is chosen at the first solve (the candidate with the least fill) and reused:
    later iterations factorise the column-permuted matrix with permc_spec="NATURAL"; the permutation of the values
    is a fixed gather.

SciPy's SuperLU interface exposes no more of the symbolic phase than the ordering (its `Fact=SamePattern` mode needs
the previous L/U, which splu does not return); a Julia host gets the full reuse with `F = lu(K)` once and
`lu!(F, K)` afterwards, because K keeps its colptr / rowval. The solve itself is outside the hot
path (north_star: "stays on the reference's host path, timed separately")."""
import time as _time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class PatternSolver:
    def __init__(self, colptr, rowval, n, register=None, unregister=None):
        """colptr / rowval: 1-based Int64 CSC pattern (maf_pattern). register / unregister: optional callables that
        page-lock / release the value buffer (capi.host_register / host_unregister)."""
        idx = np.int32 if len(rowval) < 2 ** 31 and n < 2 ** 31 else np.int64
        self.n = int(n)
        self.indptr = (np.asarray(colptr) - 1).astype(idx)
        self.indices = (np.asarray(rowval) - 1).astype(idx)
        self.nzval = np.zeros(len(rowval))
        self._unregister = None
        if register is not None:
            try:
                register(self.nzval)
                self._unregister = unregister
            except Exception:
                self._unregister = None          # pageable buffers work too (the library stages them)
        self.K = sp.csc_matrix((self.nzval, self.indices, self.indptr), shape=(self.n, self.n), copy=False)
        assert np.shares_memory(self.K.data, self.nzval)
        self.perm_c = None
        self.timers = {"order_s": 0.0, "factor_s": 0.0, "solve_s": 0.0, "gather_s": 0.0, "solves": 0}

    def close(self):
        if self._unregister is not None:
            try:
                self._unregister(self.nzval)
            finally:
                self._unregister = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    ORDERINGS = ("NATURAL", "MMD_ATA", "COLAMD")

    def _analyse(self):
        """First solve = the symbolic decision, made once per pattern: factorise with each of SuperLU's column
        orderings, keep the one with the least fill (for the reference's node-major numbering of a structured patch
        NATURAL -- a banded matrix -- beats COLAMD by 2-3x in time), and cache its permutation and permuted pattern."""
        t0 = _time.perf_counter()
        best = None
        for spec in self.ORDERINGS:
            cand = spla.splu(self.K, permc_spec=spec)
            fill = cand.L.nnz + cand.U.nnz
            if best is None or fill < best[0]:
                best = (fill, spec, cand)
        self.fill, self.ordering, lu = best
        self.perm_c = np.asarray(lu.perm_c)
        inv = np.argsort(self.perm_c)                       # column k of the permuted matrix = column inv[k] of K
        counts = np.diff(self.indptr)[inv]
        self.p_indptr = np.concatenate(([0], np.cumsum(counts))).astype(self.indptr.dtype)
        starts = self.indptr[inv].astype(np.int64)
        self.src = (np.repeat(starts - self.p_indptr[:-1].astype(np.int64), counts) +
                    np.arange(int(self.p_indptr[-1]), dtype=np.int64))
        self.p_indices = self.indices[self.src]
        self.timers["order_s"] += _time.perf_counter() - t0
        return lu

    def solve(self, r):
        """x with K x = r for the values currently in self.nzval."""
        t0 = _time.perf_counter()
        if self.perm_c is None:
            lu = self._analyse()
            x = lu.solve(np.asarray(r, dtype=float))
            self.timers["solves"] += 1
            return x
        data = self.nzval[self.src]
        t1 = _time.perf_counter()
        Kp = sp.csc_matrix((data, self.p_indices, self.p_indptr), shape=(self.n, self.n), copy=False)
        lu = spla.splu(Kp, permc_spec="NATURAL")
        t2 = _time.perf_counter()
        y = lu.solve(np.asarray(r, dtype=float))
        x = y[self.perm_c]
        t3 = _time.perf_counter()
        self.timers["gather_s"] += t1 - t0
        self.timers["factor_s"] += t2 - t1
        self.timers["solve_s"] += t3 - t2
        self.timers["solves"] += 1
        return x
