/* ============================================================================
 * maf.h -- C ABI of libmembrane_b200.so
 *
 * B200-native (sm_100a) drop-in for ONE function of sahu-lab/MembraneAleFem.jl:
 *
 *   r_gl, K_gl = calc_r_K(mesh, xms, cps, time, dt, p; args...)
 *                                   src/analysis/FiniteElement.jl:75-200
 *
 * i.e. the per-Newton-iteration global residual vector and consistent tangent
 * (area elements :98-138, Neumann boundary elements :151-197) of the Helfrich /
 * incompressible / viscous / ALE-mesh equations on the quadratic B-spline mesh.
 * The reference has no FFI for this path; the entry points below are what a
 * Julia `ccall` shim replacing the body of calc_r_K binds.
 *
 * Conventions
 *  - all functions return 0 on success, nonzero on error; text via maf_last_error.
 *    No exceptions, aborts or signal handlers cross this boundary.
 *  - all index arrays are the reference's own: Int64, 1-based, column-major.
 *  - all floating-point data is FP64 (the reference's ComplexF64 is only its
 *    differentiation device, FiniteElement.jl:113-122; the tangent here is exact).
 *  - the library never retains a host pointer past the call that received it.
 *  - one handle is driven by one host thread at a time (not re-entrant per handle).
 * ========================================================================== */
#ifndef MAF_H
#define MAF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct maf_handle maf_handle;

/* Enum codes are the reference's (src/input/Enums.jl). */
enum { MAF_STATIC = 1, MAF_EUL = 2, MAF_LAG = 3, MAF_ALEV = 4, MAF_ALEVB = 5 };           /* Motion   :67-73  */
enum { MAF_F_CAVI = 1, MAF_F_COUE = 2, MAF_F_POIS = 3, MAF_F_PULL = 4, MAF_F_BEND = 5 };  /* Scenario :24-30  */
enum { MAF_BOTTOM = 1, MAF_RIGHT = 2, MAF_TOP = 3, MAF_LEFT = 4 };                        /* Boundary :89-94  */
enum { MAF_SHEAR = 1, MAF_STRETCH = 2, MAF_MOMENT = 3 };                                  /* Neumann  :131-135 */

/* Sparsity pattern of K (hard part 1).
 *  MAF_PATTERN_BLK: union over elements of (active LM rows x active LM cols) minus the dof blocks that are
 *                   identically zero by the equations ((v,pm) (lambda,pm) (vm,v) (vm,lambda) (pm,lambda) for ALE,
 *                   (vm,lambda) for EUL). The pattern Julia stores on a generic state is a subset of this one and
 *                   every extra entry is an exact 0.0.
 *  MAF_PATTERN_SYM: the full LM x LM union. */
enum { MAF_PATTERN_BLK = 0, MAF_PATTERN_SYM = 1 };

/* Scatter paths. Both produce the same pattern; DETERMINISTIC sums every nnz slot in ascending element-id
 * order (bitwise reproducible run to run, same order as the reference with one Julia thread). */
enum { MAF_SCATTER_ATOMIC = 0, MAF_SCATTER_DETERMINISTIC = 1 };

/* Read-only mesh tables. Replaces what calc_r_K reads from `mesh::Mesh` (src/input/Mesh.jl:48-80). */
typedef struct {
  int64_t numel, numnp, ndf, nmdf;     /* Mesh.numel, numnp, ndf, nmdf                                  */
  int64_t num1el, num2el;              /* element id = e1 + (e2-1)*num1el  (Mesh.jl:582-588)            */
  const int64_t* IX;                   /* 9 x numel   Mesh.IX  (Mesh.jl:574-593)                        */
  const int64_t* ID;                   /* ndf x numnp Mesh.ID  (Mesh.jl:276-284), 0 = Dirichlet         */
  const int64_t* LM;                   /* optional (may be NULL): 9ndf x numel, checked == ID[:,IX]     */
  int32_t dofs[8];                     /* Mesh.dofs: column of vx vy vz vmx vmy vmz lambda pm, 0=absent */
  /* 1-D basis tables (LineGpBasisFns, src/input/GpBasisFn.jl:143-202). Every entry of the reference's 2-D
   * tables (GpBasisFnsζα, :96-112) is ONE product of two entries below, so the 2-D values formed on the
   * device are bit-identical to mesh.area_gp_fns / mesh.bdry_gp_fns. Layout [uel][gp][10] =
   * (w, N[3], dN[3], ddN[3]); edge tables [2][10] = ζmin_fns, ζmax_fns (w = 1). uel ids are 1-based. */
  int64_t nuel1, nuel2;
  const int64_t* uel_ids1;             /* num1el */
  const int64_t* uel_ids2;             /* num2el */
  const double* line1;                 /* nuel1 x 3 x 10 */
  const double* line2;                 /* nuel2 x 3 x 10 */
  const double* edge1;                 /* 2 x 10 */
  const double* edge2;                 /* 2 x 10 */
  double xi[3];                        /* GaussPointsξ(3).ξs (GaussPoint.jl:116-118), used by the DB term */
  /* inhomogeneous Neumann conditions, Mesh.inh_neu_bcs in its own order, and Mesh.bdry_elems */
  int32_t n_neu;
  const int32_t* neu_bdry;             /* Boundary code per condition */
  const int32_t* neu_type;             /* Neumann code per condition  */
  const double* neu_val;
  const int64_t* bdry_elems[4];        /* indexed by Boundary code - 1; element ids, 1-based */
  int64_t bdry_count[4];
} maf_mesh_desc;

/* Scalars calc_r_K reads from `p::Params` (src/input/Params.jl:36-55). */
typedef struct {
  int32_t motion;        /* p.motion   */
  int32_t scenario;      /* p.scenario */
  double kb, kg, zv, pn; /* p.kb p.kg p.ζv p.pn */
  double adb, am;        /* p.αdb p.αm */
  int32_t pattern_mode;  /* MAF_PATTERN_* */
  int32_t device;        /* CUDA device ordinal; -1 = current device */
} maf_params;

/* One-time setup: uploads the tables, builds the symbolic CSC pattern and the scatter maps on the device. */
int maf_create(maf_handle** out, const maf_mesh_desc* mesh, const maf_params* params);
int maf_destroy(maf_handle* h);

/* Message for the last failing call on this handle (h may be NULL: last maf_create failure). */
const char* maf_last_error(const maf_handle* h);

/* Symbolic pattern of K = CSC of a SparseMatrixCSC{Float64,Int64}: colptr (nmdf+1) and rowval (nnz),
 * 1-based, rows sorted within each column. */
int maf_nnz(maf_handle* h, int64_t* nnz);
int maf_pattern(maf_handle* h, int64_t* colptr, int64_t* rowval);

/* The same pattern in pieces, for meshes whose full rowval (nnz Int64 values: 9.4 GB for the 10^6-element ALE patch)
 * is more than a caller wants to hold at once -- a host that keeps K in column slices (one per GPU), or a spot check:
 *   maf_colptr            colptr alone (nmdf+1, 1-based)
 *   maf_pattern_columns   rowval of the columns [col_first, col_last] (1-based, inclusive):
 *                         colptr[col_last+1] - colptr[col_first] values. */
int maf_colptr(maf_handle* h, int64_t* colptr);
int maf_pattern_columns(maf_handle* h, int64_t col_first, int64_t col_last, int64_t* rowval);

/* The hot path with HOST buffers (what the Julia shim calls once per Newton iteration):
 *   xms  numnp x 3   column-major   (FiniteElement.jl:77)
 *   cps  numnp x ndf column-major   (FiniteElement.jl:78)
 *   r    nmdf            out: r_gl
 *   nzval nnz            out: K_gl.nzval in the order of maf_pattern
 *   rnorm2 (optional)    out: sum(r.^2)
 * bend_tm is args[:bend_tm] (only read for MOMENT conditions, FiniteElement.jl:379).
 * From 32768 elements on (MAF_PIPELINE_MIN_ELEMS) the atomics path assembles the mesh in sub-strips of element rows
 * (8; MAF_SUBSTRIPS) and copies the finished ranges of r / nzval to the host while the next strips are being
 * assembled; page-locked destinations (maf_host_register) make those copies asynchronous. */
int maf_assemble(maf_handle* h, const double* xms, const double* cps, double time, double dt, double bend_tm,
                 int scatter_mode, double* r, double* nzval, double* rnorm2);

/* Same, with DEVICE pointers (assembly-only sweeps, device-resident Newton loops). Any of d_r / d_nzval /
 * d_rnorm2 may be NULL: the handle's own buffers are used (see maf_device_buffers). `stream` is a cudaStream_t
 * (NULL = the handle's stream). Asynchronous with respect to the host. */
int maf_assemble_device(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt,
                        double bend_tm, int scatter_mode, double* d_r, double* d_nzval, double* d_rnorm2,
                        void* stream);

/* The handle's device buffers (xms: 3 numnp, cps: ndf numnp, r: nmdf, nzval: nnz, rnorm2: 1). */
int maf_device_buffers(maf_handle* h, double** d_xms, double** d_cps, double** d_r, double** d_nzval,
                       double** d_rnorm2);
/* Copy slices of the handle's own device results to the host after maf_assemble_device (blocking): rows
 * [r_first, r_first + r_count) of r and entries [nz_first, nz_first + nz_count) of nzval, 1-based; a count of 0 skips
 * that array. With maf_range_info this returns exactly what an element range has written. */
int maf_download(maf_handle* h, int64_t r_first, int64_t r_count, double* r, int64_t nz_first, int64_t nz_count,
                 double* nzval);
/* The handle's stream (cudaStream_t) and a blocking wait on it. */
int maf_stream(maf_handle* h, void** stream);
int maf_sync(maf_handle* h);

/* Diagnostics: device-side times of the last assembly (CUDA events on the stream it ran on), ms:
 * out[0] h2d copies, out[1] area kernel, out[2] boundary kernels, out[3] gather kernels (deterministic path),
 * out[4] d2h copies, out[5] whole maf_assemble call, out[6] zero-fill of r / nzval. out[0], out[4], out[5] are
 * only set by maf_assemble. And the number of kernels this handle has launched so far. */
int maf_timings(maf_handle* h, double* out7);
int maf_launch_count(maf_handle* h, int64_t* n);
/* Device time (ms, CUDA events on the launching stream) of the area-element kernel in each of the last n assemblies,
 * oldest first (n <= 64): read once after a timed region, so the region itself needs no host synchronisation. */
int maf_area_kernel_times(maf_handle* h, double* out_ms, int64_t n);

/* Area-element kernel configuration actually used: out[0] threads per CTA, out[1] elements per CTA,
 * out[2] dynamic shared memory bytes per CTA, out[3] resident CTAs per SM, out[4] SM count, out[5] number of distinct
 * element scatter maps (elements with equal maps share one: a few hundred on a structured 10^6-element patch),
 * out[6] assemblies replayed from a captured CUDA graph so far (meshes of <= 16384 elements assembling into the
 * handle's own buffers: their ~10 launches are bound by the host's launch rate; MAF_NO_GRAPH=1 disables it),
 * out[7] bytes of staging memory the deterministic path holds (0 until it ran), out[8] its band height in element
 * rows (0 = the whole range staged at once; large ranges are staged in a ring of three bands, <= 10 % of nzval;
 * MAF_BAND_ROWS overrides the height). */
int maf_kernel_info(maf_handle* h, int64_t* out9);

/* The static plan of the area kernel's contraction phase as text: "c,c,c/c,c/..." = the chunk ids (<= 32 tangent
 * tasks of one block each) that every warp of the CTA executes, in order. The environment variable MAF_PLAN (same
 * format) replaces the built-in plan at maf_create (tools/tune_plan.py searches plans on the GPU). No reference
 * counterpart: the reference's element loop (FiniteElement.jl:98-140) is scheduled by Julia's task runtime. */
int maf_chunk_plan(maf_handle* h, char* text, int64_t cap);

/* Element ranges on a whole-mesh handle (spot checks, and the round-1 multi-GPU path where every process holds all
 * tables and buffers; the strip handles above supersede it): restrict this handle to the elements
 * [el_first, el_last] (1-based, inclusive) -- contiguous element ids are strips of element rows (Mesh.jl:582-588).
 * Because unknowns are numbered node-major (Mesh.jl:276-284) a strip touches one contiguous range of rows of r and
 * one contiguous range of nzval; maf_range_info returns them (all 1-based, inclusive):
 *   out[0..1] elements, out[2..3] nodes, out[4..5] rows of r, out[6..7] entries of nzval.
 * Only those ranges are written by an assembly. The ranges of neighbouring strips overlap exactly in the interface
 * rows/entries (the two node rows a quadratic strip boundary shares); the caller sums the overlaps with its
 * collective (NCCL send/recv between neighbours) -- nothing else crosses GPUs. */
int maf_set_element_range(maf_handle* h, int64_t el_first, int64_t el_last);
int maf_range_info(maf_handle* h, int64_t* out8);

/* ---- Strips over several GPUs behind the boundary --------------------------------------------------
 * The reference splits the element loop into contiguous chunks, one per Julia task with private r / K, and sums them
 * (FiniteElement.jl:88-89, 144-147). Here a chunk is a strip of element rows on its own GPU; one handle per strip,
 * driven by one host thread for all GPUs of a process or by one process per GPU:
 *
 *   maf_create_strip(out, mesh, params, rank, nranks)
 *        like maf_create, for strip `rank` (element rows [rank num2el / nranks, (rank + 1) num2el / nranks), at least
 *        two rows per strip). Allocates only what the strip needs: the scatter maps of its own elements and the
 *        slices of r / nzval its elements touch (memory per GPU ~ 1 / nranks; the node-major numbering of
 *        Mesh.jl:276-284 makes both slices contiguous). params->device selects the GPU.
 *   maf_strip_info(h, out12)   1-based inclusive: out[0..1] elements, [2..3] rows of r touched, [4..5] entries of
 *        nzval touched, [6..7] rows OWNED, [8..9] entries OWNED, [10] rank, [11] nranks. Neighbouring strips overlap
 *        in the two node rows they share; those belong to the UPPER strip, so the owned ranges tile 1..nmdf / 1..nnz.
 *   maf_peer_attach_local(h, lower, upper)       same process: the handles of the strips rank-1 / rank+1 (NULL at
 *        the ends); enables peer access between their devices.
 *   maf_peer_export(h, handle64) / maf_peer_attach(h, lower64, upper64)   one process per GPU: 64-byte export
 *        (a cudaIpcMemHandle_t) of a strip's result buffer, exchanged by the caller (any transport), NULL at the ends.
 *   maf_assemble_strip(h, d_xms, d_cps, time, dt, bend_tm, mode, d_rnorm2_partial)
 *        the strip's share of calc_r_K, asynchronous on the handle's stream: its elements are assembled into its
 *        slices, then the upper strip of every pair ADDS the lower strip's partial sums of the interface, reading them
 *        straight from the neighbour's memory over NVLink (peer loads; two flags per pair order it -- no NCCL
 *        payload, no host round trip). d_xms / d_cps: device pointers of the full state arrays, or NULL for the
 *        handle's resident state. d_rnorm2_partial (device, may be NULL = the handle's own word): sum(r^2) over the
 *        OWNED rows -- the caller adds the partials of all strips (one scalar all-reduce).
 *   maf_assemble_strip_host(...)   same with host buffers: uploads only the node rows the strip reads, assembles,
 *        copies the OWNED rows / entries into r_own / nzval_own (sizes from maf_strip_info). Blocking.
 *   maf_strip_timings(h, out2)  out[0] ms between the end of the strip's kernels and the end of the interface
 *        exchange of the last assembly, out[1] bytes of the strip's result allocation.
 * All strips must run the same sequence of assemblies (the flags count them); a neighbour that does not arrive within
 * ~10 s is reported as an error by the next blocking call instead of hanging the device. */
int maf_create_strip(maf_handle** out, const maf_mesh_desc* mesh, const maf_params* params, int32_t rank,
                     int32_t nranks);
int maf_strip_info(maf_handle* h, int64_t* out12);
int maf_peer_attach_local(maf_handle* h, maf_handle* lower, maf_handle* upper);
int maf_peer_export(maf_handle* h, void* handle64);
int maf_peer_attach(maf_handle* h, const void* lower64, const void* upper64);
int maf_assemble_strip(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt, double bend_tm,
                       int scatter_mode, double* d_rnorm2_partial);
int maf_assemble_strip_host(maf_handle* h, const double* xms, const double* cps, double time, double dt, double bend_tm,
                            int scatter_mode, double* r_own, double* nzval_own, double* rnorm2_partial);
int maf_strip_timings(maf_handle* h, double* out2);

/* ---- Device-resident state: the glue of time_step! between two calls of calc_r_K ----------------
 * With these the state never returns to the host inside a time step; per Newton iteration only r / nzval come back
 * and du goes in. maf_assemble also leaves the state it was given resident.
 *
 * maf_state_set / maf_state_get   copy xms (numnp x 3) and cps (numnp x ndf), column-major Float64, to / from the
 *                                 device (either pointer of _get may be NULL).
 * maf_state_update(du, dt)        dcps[ID_inv] = du; cps += dcps; update_xms!(xms, dcps, dt)
 *                                 (FiniteElement.jl:41-46, 408-423). du: nmdf values. Bit-identical to the host loop
 *                                 (the product dt * dcps is rounded before the sum, no FMA).
 * maf_state_predict(dt)           update_xms!(xms, cps, dt), the predictor of run_analysis (Analysis.jl:70).
 * maf_assemble_resident(...)      maf_assemble on the resident state (no host-to-device copy of the state). */
int maf_state_set(maf_handle* h, const double* xms, const double* cps);
int maf_state_get(maf_handle* h, double* xms, double* cps);
int maf_state_update(maf_handle* h, const double* du, double dt);
int maf_state_predict(maf_handle* h, double dt);
int maf_assemble_resident(maf_handle* h, double time, double dt, double bend_tm, int scatter_mode, double* r,
                          double* nzval, double* rnorm2);

/* rv of calc_elem_dof_residuals (FiniteElement.jl:253-330, real parts) of n elements (1-based ids) on the resident
 * state: rv[27 k + comp + 3 (a - 1)], all rows incl. those of Dirichlet dofs -- what calc_pull_force
 * (PullForce.jl:61-80) sums over the elements adjacent to the pulled one. Needs a 3-D velocity
 * (Analysis.jl:48). */
int maf_elem_v_residuals(maf_handle* h, const int64_t* el_ids, int64_t n, double* rv);

/* generate_output (src/Output.jl:32-118) on the resident state: positions and unknowns at every area Gauss point,
 * boundary Gauss point and corner of the patch. xout: (3 num1el + 2) x (3 num2el + 2) x 3, uout: ... x ndf, column-major
 * like the reference's arrays. Needs the structured patch numbering of Mesh.jl:574-593. */
int maf_generate_output(maf_handle* h, double* xout, double* uout);

/* Page-lock / release a host buffer the caller owns (cudaHostRegister): maf_assemble copies from and into
 * page-locked memory directly and at the full PCIe rate (pageable buffers work too, through staging copies). The
 * outputs of calc_r_K have a fixed size per mesh (FiniteElement.jl:75-200 allocates them anew every call), so a host
 * shim registers its r / nzval arrays once and reuses them. Errors: maf_last_error(NULL). */
int maf_host_register(void* ptr, int64_t bytes);
int maf_host_unregister(void* ptr);

/* Measured FP64 FMA throughput of the device (TFLOP/s): the denominator of the FP64 roofline fraction. */
int maf_fp64_peak(int device, double* tflops);

/* Profiling hook. In builds with -DMAF_PHASE_TIMING: cycles per (warp, phase) of the element kernel summed over
 * all CTAs since the previous call, out[8 * warp + phase], phases = wait, interpolate, wait, gauss, wait,
 * gather-next, residual+tangent, unused; returns 0. Regular builds return 1 and leave `out` untouched. */
int maf_debug_phase_cycles(unsigned long long* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* MAF_H */
