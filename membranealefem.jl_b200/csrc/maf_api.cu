// maf_api.cu -- CUDA kernels (sm_100a) and the C ABI of include/maf.h.
//
// Kernels (all FP64 CUDA-core work; the per-element products are 9 x <=6 x <=6 contractions, far too small and too
// irregular for tcgen05 tiles, and FP64 has no tensor-core advantage on this part -- see DESIGN.md):
//   area_kernel<MOTION, STAGED>  persistent CTAs, one area element per CTA iteration: gather -> interpolate ->
//                          Gauss-point tangent (forward-mode dual numbers) -> residual + factored tangent
//                          contraction in registers -> scatter (FP64 RED atomics, or staging for the
//                          deterministic path)
//   boundary_kernel        one warp per Neumann boundary element
//   gather_K / gather_r    deterministic path: every nnz slot / residual row sums its staged contributions in
//                          ascending element order
//   rnorm2_kernel          sum(r^2), fixed-order two-level reduction
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "maf_host.h"
#include "maf_state.cuh"

using namespace maf;

#ifndef MAF_MIN_CTAS
#define MAF_MIN_CTAS 3  // resident CTAs per SM the register allocation is capped for (168 registers per thread)
#endif

static_assert(sizeof(Config) <= 4000, "Config must fit the kernel parameter space");

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
struct StageSink {      // deterministic path staging buffers (NULL = atomics path)
  double* kel;          // staged elements x 81 x nij
  double* rel;          // staged elements x 72
  int nij;
  int64_t ring;         // 0: one block per element of [e0, e1); otherwise element e lives in block (e - e0) % ring
};

#ifdef MAF_PHASE_TIMING   // profiling build only: cycles per warp and phase, summed over all CTAs
__device__ unsigned long long g_phase_cycles[MAF_NT / 32][8];
__device__ __forceinline__ long long maf_clock() {   // not to be moved across barriers or memory operations
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define MAF_TICK(k)                                                   \
  {                                                                   \
    const long long t_now = maf_clock();                              \
    t_acc[k] += t_now - t_last;                                       \
    t_last = t_now;                                                   \
  }
#else
#define MAF_TICK(k)
#endif

// LAG / STATIC evaluate a smaller Gauss-point program (no mesh equations): they fit 128 registers, one more CTA per SM
#ifndef MAF_MIN_CTAS_LAG
#define MAF_MIN_CTAS_LAG 4
#endif
constexpr int min_ctas(int motion) { return (motion == M_LAG || motion == M_STATIC) ? MAF_MIN_CTAS_LAG : MAF_MIN_CTAS; }

// STAGED = deterministic path (staging rows instead of atomics): a separate instantiation, so that each kernel
// carries one copy of the tangent phase (instruction-cache footprint)
template <int MOTION, bool STAGED>
__global__ void __launch_bounds__(MAF_NT, min_ctas(MOTION))
area_kernel(const __grid_constant__ Config cfg, const Tables T, const double* __restrict__ xms,
            const double* __restrict__ cps, double dt, double* __restrict__ r_gl, double* __restrict__ nzval,
            const StageSink st, const int32_t* __restrict__ order, int64_t e0, int64_t e1) {
  extern __shared__ double smem_all[];
  const int tid = threadIdx.x;
  const int64_t ne = e1 - e0;
  const int nfront = cfg.front_doubles;
  int32_t* ids0 = reinterpret_cast<int32_t*>(smem_all + 2 * nfront);
  double* sm = smem_all + 2 * nfront + 2 * MAF_IDS_DOUBLES;   // back block: E, S, G, A
  // Software pipeline over the elements of this CTA (k, k + G, k + 2G, ...; G = gridDim.x). At the top of the
  // iteration of element k every thread issues asynchronous global -> shared copies (cp.async): the data of
  // element k + G into the other front buffer, addressed through the node / pair ids that were copied during the
  // previous iteration, and the ids of element k + 2G. They are awaited at the end of the iteration, a whole
  // element later: the gather costs neither registers nor exposed latency. (One warp gathering item by item took
  // as long as the contraction phase; all threads with batched loads still lost 10-20 % of their time waiting:
  // profiles/r1_notes.md.)
  const int64_t G = gridDim.x;
  // elements of this iteration / the next / the one after; the order is read one iteration before it is needed
  int32_t el_cur = 0, el_nxt = 0, el_ids = 0;
  gather_init(tid, cfg, smem_all);
  gather_init(tid, cfg, smem_all + nfront);
  if ((int64_t)blockIdx.x < ne) {
    el_cur = order[blockIdx.x];
    if (blockIdx.x + G < ne) el_nxt = order[blockIdx.x + G];
    if (blockIdx.x + 2 * G < ne) el_ids = order[blockIdx.x + 2 * G];
    gather_ids_async(tid, T, el_cur, ids0);
    async_wait_all();
    __syncthreads();
    gather_data_async(tid, cfg, T, ids0, xms, cps, smem_all);
    if (blockIdx.x + G < ne) gather_ids_async(tid, T, el_nxt, ids0 + MAF_IDS_INTS);
    async_wait_all();
  }
  // (staggering the resident CTAs of an SM by a fraction of an element time with __nanosleep: no change -- they are
  // not in lockstep, the phases simply add up at this occupancy)
  int cur = 0;
#ifdef MAF_PHASE_TIMING
  long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_last = maf_clock();
#endif
  for (int64_t k = blockIdx.x; k < ne; k += G, cur ^= 1) {
    const int64_t el = el_cur;
    const double* fr = smem_all + cur * nfront;
    __syncthreads();   // front buffer `cur` and the ids of the next element complete; everything else is free
    MAF_TICK(0)
    const int64_t kn = k + G;
    if (kn < ne) {
      gather_data_async(tid, cfg, T, ids0 + (cur ^ 1) * MAF_IDS_INTS, xms, cps, smem_all + (cur ^ 1) * nfront);
      if (kn + G < ne) gather_ids_async(tid, T, el_ids, ids0 + cur * MAF_IDS_INTS);
      el_cur = el_nxt;
      el_nxt = el_ids;
      if (kn + 2 * G < ne) el_ids = order[kn + 2 * G];
    }
#ifndef MAF_STUB_INTERP
    phase_interp(tid, MAF_NT, cfg, fr, sm);
#endif
    MAF_TICK(1)
    __syncthreads();
    MAF_TICK(2)
#ifndef MAF_STUB_GAUSS   // MAF_STUB_*: timing-only builds that drop one phase (wrong results; tools/build_variant.sh)
    phase_gauss<MOTION>(tid, cfg, dt, fr, sm);
#endif
    MAF_TICK(3)
    __syncthreads();
    MAF_TICK(4)
    MAF_TICK(5)
    // (forming the residual on the warp of the IT_LIN items before this barrier was measured: -19 %, that warp
    // becomes the critical path of the Gauss phase)
#ifndef MAF_STUB_RESIDUAL
    phase_residual(tid, MAF_NT, cfg, fr, sm, r_gl, STAGED ? st.rel + 72 * stage_index(el - e0, st.ring) : nullptr);
#endif
    if (!STAGED) {
      KSink sink{nzval, nullptr, 0};
#ifndef MAF_STUB_TANGENT
      phase_tangent<MOTION>(tid, cfg, fr, sm, sink);
#endif
    } else {
      KSink sink{nullptr, st.kel + (size_t)81 * st.nij * stage_index(el - e0, st.ring), st.nij};
      phase_tangent<MOTION>(tid, cfg, fr, sm, sink);
    }
    async_wait_all();
    MAF_TICK(6)
  }
#ifdef MAF_PHASE_TIMING
  if ((tid & 31) == 0)
    for (int q = 0; q < 8; ++q) atomicAdd(&g_phase_cycles[tid >> 5][q], (unsigned long long)t_acc[q]);
#endif
}

// One-time (maf_create): the scatter maps (maf_element.cuh::build_elslot), one CTA per element, in three passes that
// never materialise one map per element (2.9 KB each): MODE 0 hashes every element's map, the host groups equal
// hashes into classes, MODE 1 writes the map of each class's representative element, MODE 2 re-derives every
// element's map and compares it with its class's (a hash collision is reported, never silently accepted).
template <int MODE>
__global__ void __launch_bounds__(128)
build_elslot_kernel(const Tables T, const int32_t* __restrict__ rep_el /* MODE 1: representative element per class */,
                    int32_t* __restrict__ elslot, int64_t* __restrict__ elbase, unsigned long long* __restrict__ hash,
                    int* __restrict__ flags /* [0] overflow, [1] mismatch */) {
  __shared__ long long s_col[72];
  __shared__ long long s_base;
  __shared__ unsigned long long s_hash[128];
  const int tid = threadIdx.x;
  const int64_t el = MODE == 1 ? (int64_t)rep_el[blockIdx.x] : T.el0 + blockIdx.x;
  if (tid < 72) s_col[tid] = T.nodecol[8 * (int64_t)T.IX[9 * el + (tid >> 3)] + (tid & 7)];
  __syncthreads();
  if (tid == 0) {
    long long base = -1;
    for (int k = 0; k < 72; ++k)
      if (s_col[k] >= 0 && (base < 0 || s_col[k] < base)) base = s_col[k];
    s_base = base < 0 ? 0 : base;
    if (MODE == 0) elbase[el - T.el0] = s_base;
  }
  __syncthreads();
  const long long base = s_base;
  unsigned long long hsum = 0;
  const int32_t* cls = MODE == 2 ? T.elslot + (size_t)MAF_SLOT_INTS * T.elclass[el - T.el0] : nullptr;
  for (int k = tid; k < MAF_SLOT_INTS; k += 128) {
    int32_t v = -1;
    if (k < 729) {
      const int b = k / 81, a = (k % 81) / 9, J = k % 9;
      if (J < 8) {
        const long long c = s_col[8 * b + J];
        if (c >= 0) {
          const long long off = c - base + T.pairoff[(int64_t)T.elpair[81 * (el - T.el0) + 9 * a + b] * 8 + J];
          if (off > 0x7fffff00LL) flags[0] = 1;
          v = (int32_t)off;
        }
      }
    }
    if (MODE == 0) {   // position-dependent mix, summed over the map (order independent)
      unsigned long long x = ((unsigned long long)(unsigned)v << 20) ^ (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ull;
      x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
      hsum += x;
    } else if (MODE == 1) {
      elslot[(size_t)MAF_SLOT_INTS * blockIdx.x + k] = v;
    } else if (cls[k] != v) {
      flags[1] = 1;
    }
  }
  if (MODE == 0) {
    s_hash[tid] = hsum;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
      if (tid < w) s_hash[tid] += s_hash[tid + w];
      __syncthreads();
    }
    if (tid == 0) hash[el - T.el0] = s_hash[0];
  }
}

__global__ void __launch_bounds__(128)
boundary_kernel(const __grid_constant__ Config cfg, const Tables T, const BoundaryTables BT, int bc, double fval,
                double dt, const double* __restrict__ xms, double* __restrict__ r_gl, double* __restrict__ nzval,
                int colour, int ncolour, int64_t e0, int64_t e1) {
  __shared__ double smem[4 * B_DOUBLES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sm = smem + warp * B_DOUBLES;
  const int n = BT.offs[bc + 1] - BT.offs[bc];
  // colour < 0: all units concurrently with atomics; otherwise only units k = colour (mod ncolour), plain adds
  for (int k = blockIdx.x * 4 + warp; k < n; k += gridDim.x * 4) {
    if (colour >= 0 && (k % ncolour) != colour) continue;
    const int64_t el = BT.elems[BT.offs[bc] + k];
    if (el < e0 || el >= e1) continue;
    boundary_gather(lane, 32, cfg, T, BT, bc, el, xms, sm);
    __syncwarp();
    boundary_gauss(lane, 32, cfg, BT, bc, fval, dt, sm);
    __syncwarp();
    boundary_scatter(lane, 32, cfg, T, sm, r_gl, nzval, colour >= 0);
    __syncwarp();
  }
}

// Small meshes: one CTA per (Neumann condition, boundary element) in ONE launch. With a few dozen boundary elements
// the warp-per-element kernel above is a chain of ~25 us per launch (729 tangent entries per warp) -- longer than
// the area kernel of a 17 x 17 mesh; 128 threads per element cut it to a quarter.
struct NeumannValues { double fval[8]; };
__global__ void __launch_bounds__(128)
boundary_kernel_cta(const __grid_constant__ Config cfg, const Tables T, const BoundaryTables BT, const NeumannValues nv,
                    double dt, const double* __restrict__ xms, double* __restrict__ r_gl, double* __restrict__ nzval,
                    int64_t e0, int64_t e1) {
  __shared__ double sm[B_DOUBLES];
  const int bc = blockIdx.y;
  const int n = BT.offs[bc + 1] - BT.offs[bc];
  if ((int)blockIdx.x >= n) return;
  const int64_t el = BT.elems[BT.offs[bc] + blockIdx.x];
  if (el < e0 || el >= e1) return;
  boundary_gather(threadIdx.x, 128, cfg, T, BT, bc, el, xms, sm);
  __syncthreads();
  boundary_gauss(threadIdx.x, 128, cfg, BT, bc, nv.fval[bc], dt, sm);
  __syncthreads();
  boundary_scatter(threadIdx.x, 128, cfg, T, sm, r_gl, nzval, false);
}

// Deterministic path (maf_gather.cuh): MAF_GATHER_LANES threads per node pair, one thread per residual row.
__global__ void __launch_bounds__(128)
gather_K_kernel(const __grid_constant__ Config cfg, const Tables T, const GatherTables G,
                const double* __restrict__ kel, int nij, int64_t e0, int64_t e1, int64_t p_lo, int64_t p_hi,
                double* __restrict__ nzval) {
  const int64_t n = (p_hi - p_lo) * MAF_GATHER_LANES;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    gather_K_pair(p_lo + t / MAF_GATHER_LANES, (int)(t % MAF_GATHER_LANES), cfg, T, G, kel, nij, e0, e1, nzval);
}
__global__ void __launch_bounds__(128)
gather_r_kernel(const __grid_constant__ Config cfg, const Tables T, const GatherTables G,
                const double* __restrict__ rel, int64_t e0, int64_t e1, int64_t k_lo, int64_t k_hi,
                double* __restrict__ r_gl) {
  for (int64_t k = k_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < k_hi;
       k += (int64_t)gridDim.x * blockDim.x)
    gather_r_row(k, cfg, T, G, rel, e0, e1, r_gl);
}

// FP64 roofline denominator: independent DFMA chains, 8 per thread, no memory traffic.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int k = 0; k < iters; ++k) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;  // keeps the chains alive without a store in the common case
}

// residual of the v rows of a few elements, all rows (also those of Dirichlet dofs): the reaction force of the
// pulled nodes is summed from them (calc_pull_force, PullForce.jl:61-80). One CTA per listed element; stages the
// element residual [u][a] (u = v0 v1 v2 m0 m1 m2 l p) like the deterministic path.
template <int MOTION>
__global__ void __launch_bounds__(MAF_NT, min_ctas(MOTION))
elem_residual_kernel(const __grid_constant__ Config cfg, const Tables T, const double* __restrict__ xms,
                     const double* __restrict__ cps, const int32_t* __restrict__ els, int n, double* __restrict__ rel) {
  extern __shared__ double smem_all[];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= n) return;
  int32_t* ids = reinterpret_cast<int32_t*>(smem_all + 2 * cfg.front_doubles);
  double* sm = smem_all + 2 * cfg.front_doubles + 2 * MAF_IDS_DOUBLES;
  gather_init(tid, cfg, smem_all);
  gather_ids_async(tid, T, els[blockIdx.x], ids);
  async_wait_all();
  __syncthreads();
  gather_data_async(tid, cfg, T, ids, xms, cps, smem_all);
  async_wait_all();
  __syncthreads();
  phase_interp(tid, MAF_NT, cfg, smem_all, sm);
  __syncthreads();
  phase_gauss<MOTION>(tid, cfg, 0.0, smem_all, sm);
  __syncthreads();
  phase_residual(tid, MAF_NT, cfg, smem_all, sm, nullptr, rel + 72 * (size_t)blockIdx.x, true);
}

__global__ void state_update_kernel(const __grid_constant__ Config cfg, const Tables T, const double* __restrict__ du,
                                    double dt, double* __restrict__ xms, double* __restrict__ cps) {
  const int64_t n = T.numnp * cfg.ndf;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    state_update_entry(k, cfg, T, du, dt, xms, cps);
}
__global__ void state_predict_kernel(const __grid_constant__ Config cfg, const Tables T, double dt,
                                     double* __restrict__ xms, const double* __restrict__ cps) {
  const int64_t n = T.numnp * 3;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    state_predict_entry(k, cfg, T, dt, xms, cps);
}

// generate_output (Output.jl:32-118): positions and unknowns sampled at every area Gauss point, boundary Gauss point
// and corner of the patch -- a tensor-product grid of (3 num1el + 2) x (3 num2el + 2) points whose 1-D basis values
// are the line tables (interior points) and the edge tables (first / last point of a direction). One thread per
// (point, output column); output arrays column-major like the reference's xout[n1, n2, XDIM], uout[n1, n2, ndf].
__global__ void __launch_bounds__(256)
sample_output_kernel(const Tables T, const BoundaryTables BT, int ndf, int num2el, const double* __restrict__ xms,
                     const double* __restrict__ cps, double* __restrict__ xout, double* __restrict__ uout) {
  const int64_t n1 = 3 * (int64_t)T.num1el + 2, n2 = 3 * (int64_t)num2el + 2, npt = n1 * n2;
  const int64_t total = npt * (3 + ndf);
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pt = k % npt;
    const int col = (int)(k / npt);
    const int64_t s1 = pt % n1, s2 = pt / n1;
    int e1, e2;
    const double *f1, *f2;
    if (s1 == 0) { e1 = 0; f1 = BT.edge1; }
    else if (s1 == n1 - 1) { e1 = T.num1el - 1; f1 = BT.edge1 + 10; }
    else { e1 = (int)((s1 - 1) / 3); f1 = T.line1 + 30 * T.uel1[e1] + 10 * ((s1 - 1) % 3); }
    if (s2 == 0) { e2 = 0; f2 = BT.edge2; }
    else if (s2 == n2 - 1) { e2 = num2el - 1; f2 = BT.edge2 + 10; }
    else { e2 = (int)((s2 - 1) / 3); f2 = T.line2 + 30 * T.uel2[e2] + 10 * ((s2 - 1) % 3); }
    const int32_t* ix = T.IX + 9 * ((int64_t)e1 + (int64_t)T.num1el * e2);
    const double* src = col < 3 ? xms + T.numnp * col : cps + T.numnp * (col - 3);
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 9; ++a) acc += src[ix[a]] * (f1[1 + a % 3] * f2[1 + a / 3]);   // N[a] = N1[a1] N2[a2]
    if (col < 3) xout[pt + npt * col] = acc;
    else uout[pt + npt * (col - 3)] = acc;
  }
}

__global__ void __launch_bounds__(256) rnorm2_partial(const double* __restrict__ r, int64_t n, double* part) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    acc += r[k] * r[k];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
__global__ void rnorm2_final(const double* part, int n, double* out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) acc += part[k];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}

// ---- strips over several GPUs: the "sum over tasks" of FiniteElement.jl:144-147 for the two node rows neighbouring
// strips share. No NCCL payload: the strip that owns the interface (the upper one) reads the lower strip's partial
// sums straight out of its memory over NVLink (peer loads) and adds them; two flags per neighbour pair order it.
//   flag_kernel   one thread: publish `sval` at *signal (a word in the NEIGHBOUR's memory) after a system-scope fence,
//                 then spin until *wait (a word in OUR memory, written by the neighbour) reaches `wval`
//   pull_add      dst[i] += src[i], src = mapped peer memory
__global__ void flag_kernel(volatile long long* signal, long long sval, volatile long long* wait, long long wval,
                            long long* err, long long max_spins) {
  if (signal) {
    __threadfence_system();
    *signal = sval;
    __threadfence_system();
  }
  if (wait) {
    long long n = 0;
    while (*wait < wval) {
      if (++n > max_spins) { *err = wval; break; }   // a neighbour that never arrives must not hang the device
      __nanosleep(200);
    }
    __threadfence_system();
  }
}
__global__ void __launch_bounds__(256)
pull_add_kernel(double* __restrict__ dst_r, const double* __restrict__ src_r, int64_t nr, double* __restrict__ dst_k,
                const double* __restrict__ src_k, int64_t nk) {   // rows of r, then entries of nzval, one launch
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nr + nk; k += (int64_t)gridDim.x * blockDim.x) {
    if (k < nr) dst_r[k] += __ldcv(src_r + k);
    else dst_k[k - nr] += __ldcv(src_k + (k - nr));
  }
}

// ---------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------
struct maf_handle {
  HostModel M;
  std::string err;
  int device = 0, sm_count = 0, ctas_per_sm = 0, grid = 0;
  size_t smem_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  std::vector<void*> allocs;
  Tables T{};
  BoundaryTables BT{};
  GatherTables G{};
  bool gather_ready = false;
  int32_t* d_order = nullptr;   // processing order of the elements of [e0, e1)
  int nij = 0;
  double *d_xms = nullptr, *d_cps = nullptr, *d_r = nullptr, *d_nz = nullptr, *d_rn = nullptr, *d_part = nullptr;
  double* d_du = nullptr;        // Newton update of the resident state
  bool state_resident = false;   // d_xms / d_cps hold a state (maf_state_set or maf_assemble)
  double *d_kel = nullptr, *d_rel = nullptr;
  size_t kel_elems = 0;
  double *h_pin_in = nullptr, *h_pin_out = nullptr;  // pinned staging for the host-buffer entry point
  size_t pin_in_bytes = 0, pin_out_bytes = 0;
  int64_t e0 = 0, e1 = 0;  // element range [e0, e1) assembled by this handle
  // what that range touches: nodes [node_lo, node_hi), equations [eq_lo, eq_hi), nnz slots [slot_lo, slot_hi)
  int64_t node_lo = 0, node_hi = 0, eq_lo = 0, eq_hi = 0, slot_lo = 0, slot_hi = 0;
  bool timed_valid = false;
  // pipelined host path (maf_assemble): strips of element rows whose finished nzval/r ranges are copied to the
  // host on a second stream while the next strip is being assembled
  // (sub-)strips of element rows of the handle's range: the atomics path of a large range assembles them one after
  // the other, so that the zero-fill of the next ones (second stream) and the device-to-host copy of the finished
  // ones (maf_assemble, copy stream) overlap the kernels
  struct Strip { int64_t e0, e1, eq_lo, slot_lo, eq_hi, slot_hi; int32_t* d_order; };
  std::vector<Strip> strips;
  int64_t strips_e0 = -1, strips_e1 = -1;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> strip_ev, zero_ev;
  bool host_copy_follows = false;   // set by maf_assemble around its device part: sub-strips pay only then
  int64_t launches = 0;
  float ms[7] = {0, 0, 0, 0, 0, 0, 0};
  // ---- strip mode (maf_create_strip): this handle holds the slices of ONE strip of element rows
  bool strip = false;
  int rank = 0, nranks = 1;
  TouchedRange lower, upper;          // what the neighbouring strips touch (empty ranges at the ends)
  int64_t own_eq_lo = 0, own_eq_hi = 0, own_slot_lo = 0, own_slot_hi = 0;   // rows / entries this strip hands out
  void* strip_alloc = nullptr;        // one allocation: flags | r slice | nzval slice (one IPC handle exports it)
  size_t strip_bytes = 0;
  long long* flags = nullptr;         // [0] lower finished step, [1] upper has pulled step, [2] error (our memory)
  void* lower_base = nullptr;         // the neighbours' allocations, mapped (IPC or same-process peer access)
  void* upper_base = nullptr;
  bool lower_ipc = false, upper_ipc = false, attached = false;
  long long step = 0;                 // assemblies done (the flags carry it)
  cudaEvent_t ev_x[4] = {};
  float ms_exchange = 0.f;
  // ring of event pairs around the area kernel of the last MAF_RING assemblies: per-launch device times of a whole
  // timed region can be read afterwards, with no host synchronisation between the steps (maf_area_kernel_times)
  cudaEvent_t ring_a[64] = {}, ring_b[64] = {};
  long long ring_n = 0;
  int64_t n_slot_classes = 0;         // distinct scatter maps (Tables::elslot rows)
  size_t rk_pad = 0;                  // whole-mesh handles: d_nz = d_r + rk_pad (one allocation)
  // small meshes: the launch sequence of an assembly, captured once per key and replayed (do_assemble_device)
  struct Band { int64_t e0, e1; int32_t* d_order; };   // deterministic path: bands of element rows (ensure_stage)
  std::vector<Band> bands;
  int64_t n_pair_classes = 0;   // node-pair contribution classes of the deterministic gather (0: not built)
  std::vector<cudaEvent_t> band_staged, band_gathered;   // one pair per band (ordering only, no timing)
  cudaStream_t gather_stream = nullptr;                  // high priority: a band's gather runs beside the next band
  int64_t band_rows = 0, band_e0 = -1, band_e1 = -1;
  struct Graph { std::vector<double> key; cudaGraphExec_t exec = nullptr; int launches = 0; };
  std::vector<Graph> graphs;
  bool use_graph = true;
  int64_t graph_replays = 0;
  // the Neumann boundary kernels (atomics path) run beside the area kernel on a second stream
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_side[2] = {};
};
#define MAF_RING 64
// layout of a strip allocation: 256 bytes of flags, the r slice, the nzval slice (each padded to 256 bytes)
static size_t strip_r_offset() { return 256; }
static size_t strip_nz_offset(const TouchedRange& R) {
  return 256 + (((size_t)(R.eq_hi - R.eq_lo) * sizeof(double) + 255) / 256) * 256;
}
static size_t strip_alloc_bytes(const TouchedRange& R) {
  return strip_nz_offset(R) + (((size_t)(R.slot_hi - R.slot_lo) * sizeof(double) + 255) / 256) * 256 + 256;
}

static std::string g_create_err;
static std::mutex g_mu;

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_));                \
  } while (0)

template <class Tp> static Tp* upload(maf_handle* h, const Tp* src, size_t n) {
  Tp* d = nullptr;
  CU(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(Tp)));
  h->allocs.push_back(d);
  if (n) CU(cudaMemcpy(d, src, n * sizeof(Tp), cudaMemcpyHostToDevice));
  return d;
}
template <class Tp> static Tp* dalloc(maf_handle* h, size_t n) {
  Tp* d = nullptr;
  CU(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(Tp)));
  h->allocs.push_back(d);
  return d;
}

typedef void (*area_fn)(const Config, const Tables, const double*, const double*, double, double*, double*,
                        const StageSink, const int32_t*, int64_t, int64_t);
template <bool STAGED> static area_fn area_kernel_sel(int motion) {
  switch (motion) {
    case M_STATIC: return area_kernel<M_STATIC, STAGED>;
    case M_EUL: return area_kernel<M_EUL, STAGED>;
    case M_LAG: return area_kernel<M_LAG, STAGED>;
    case M_ALEV: return area_kernel<M_ALEV, STAGED>;
    default: return area_kernel<M_ALEVB, STAGED>;
  }
}
static area_fn area_kernel_of(int motion, bool staged = false) {
  return staged ? area_kernel_sel<true>(motion) : area_kernel_sel<false>(motion);
}

typedef void (*elres_fn)(const Config, const Tables, const double*, const double*, const int32_t*, int, double*);
static elres_fn elem_residual_kernel_of(int motion) {
  switch (motion) {
    case M_STATIC: return elem_residual_kernel<M_STATIC>;
    case M_EUL: return elem_residual_kernel<M_EUL>;
    case M_LAG: return elem_residual_kernel<M_LAG>;
    case M_ALEV: return elem_residual_kernel<M_ALEV>;
    default: return elem_residual_kernel<M_ALEVB>;
  }
}

// what the element range [e0, e1) touches: node, equation and nnz-slot ranges (each contiguous)
static void compute_ranges(maf_handle* h) {
  const HostModel& M = h->M;
  for (auto& g : h->graphs)   // captured launch sequences hold the old range and order buffer
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
  const TouchedRange R = touched_range(M, h->e0, h->e1);
  h->node_lo = R.node_lo; h->node_hi = R.node_hi;
  h->eq_lo = R.eq_lo; h->eq_hi = R.eq_hi;
  h->slot_lo = R.slot_lo; h->slot_hi = R.slot_hi;
  std::vector<int32_t> order;
  build_element_order(M.num1el, h->e0, h->e1, order);
  if (h->d_order) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(h->d_order)); h->d_order = nullptr; }
  CU(cudaMalloc(&h->d_order, std::max<size_t>(order.size(), 1) * sizeof(int32_t)));
  if (!order.empty()) CU(cudaMemcpy(h->d_order, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
}

// does the atomics path of this handle's range run in sub-strips? (large, row-aligned ranges)
static bool wants_strips(const maf_handle* h) {
  const HostModel& M = h->M;
  int64_t min_elems = 32768;
  if (const char* e = std::getenv("MAF_PIPELINE_MIN_ELEMS")) min_elems = std::atoll(e);
  return h->e1 - h->e0 >= min_elems && h->e0 % M.num1el == 0 && h->e1 % M.num1el == 0 &&
         (h->e1 - h->e0) / M.num1el >= 16;
}

static void ensure_strips(maf_handle* h, int nstrips) {
  if (!h->strips.empty() && h->strips_e0 == h->e0 && h->strips_e1 == h->e1) return;
  const HostModel& M = h->M;
  for (auto& st : h->strips) cudaFree(st.d_order);
  h->strips.clear();
  h->strips_e0 = h->e0;
  h->strips_e1 = h->e1;
  if (!h->copy_stream) CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  const int64_t row0 = h->e0 / M.num1el, nrows = (h->e1 - h->e0) / M.num1el;
  for (int q = 0; q < nstrips; ++q) {
    const int64_t r0 = row0 + ((int64_t)q * nrows) / nstrips, r1 = row0 + ((int64_t)(q + 1) * nrows) / nstrips;
    if (r1 <= r0) continue;
    maf_handle::Strip st;
    st.e0 = r0 * M.num1el;
    st.e1 = r1 * M.num1el;
    const TouchedRange R = touched_range(M, st.e0, st.e1);
    st.eq_lo = R.eq_lo;
    st.slot_lo = R.slot_lo;
    st.eq_hi = R.eq_hi;
    st.slot_hi = R.slot_hi;
    std::vector<int32_t> order;
    build_element_order(M.num1el, st.e0, st.e1, order);
    CU(cudaMalloc(&st.d_order, order.size() * sizeof(int32_t)));
    CU(cudaMemcpy(st.d_order, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->strips.push_back(st);
    if (h->strip_ev.size() < h->strips.size()) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->strip_ev.push_back(e);
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->zero_ev.push_back(e);
    }
  }
}

static void ensure_gather(maf_handle* h) {
  if (h->gather_ready) return;
  const HostModel& M = h->M;
  const Symbolic& S = M.sym;
  GatherHost GH;
  build_gather_host(M, GH);
  h->nij = GH.nij;
  h->G.nbr_ptr = upload(h, S.nbr_ptr.data(), S.nbr_ptr.size());
  h->G.nbr = upload(h, S.nbr.data(), S.nbr.size());
  h->G.pair_node = upload(h, GH.pair_node.data(), GH.pair_node.size());
  h->G.n2e_ptr = upload(h, S.n2e_ptr.data(), S.n2e_ptr.size());
  h->G.n2e = upload(h, S.n2e.data(), S.n2e.size());
  h->G.n2e_loc = upload(h, S.n2e_loc.data(), S.n2e_loc.size());
  h->G.ij_of = upload(h, GH.ij_of.data(), GH.ij_of.size());
  fill_gather_tables(GH, h->G);
  // (MAF_NO_PAIR_CLASSES=1: keep the scanning gather -- the fallback for meshes with more than 65535 classes; tests
  // compare the two on the device)
  const char* no_pc = std::getenv("MAF_NO_PAIR_CLASSES");
  if (!(no_pc && *no_pc == '1')) build_pair_classes(M, GH);
  if (!GH.pclass.empty()) {
    h->G.pclass = upload(h, GH.pclass.data(), GH.pclass.size());
    h->G.eref = upload(h, GH.eref.data(), GH.eref.size());
    h->G.ccnt = upload(h, GH.ccnt.data(), GH.ccnt.size());
    h->G.cde = upload(h, GH.cde.data(), GH.cde.size());
    h->G.crow = upload(h, GH.crow.data(), GH.crow.size());
  }
  h->n_pair_classes = (int64_t)GH.ccnt.size();
  if (GH.nij > MAF_MAX_NIJ) throw std::runtime_error("too many dof-block classes for the gather kernel");
  h->G.npairs = S.npairs;
  h->gather_ready = true;
}

// Staging of the deterministic path. A staged element takes (81 nij + 72) doubles, 3.4 times its share of nzval, so
// large ranges are staged in BANDS of element rows: a band is assembled into a ring that holds three bands, then every
// node row whose contributing element rows (n - 2 .. n) are all staged is gathered, in ascending order -- on a second,
// high-priority stream, while the next band is being assembled into the third ring slot (the gather of band b reads
// the slots of b - 1 and b; band b + 2 reuses the slot of b - 1 and waits for that gather). The ring is sized to the
// largest band height that keeps it within 10 % of the range's nzval (taller bands are faster: 7 rows 60.1 ms, 9 rows
// 59.6, 16 rows 58.0, the whole mesh at once 55.8 with 340 %); small ranges keep one block per element (a single band).
#ifndef MAF_BAND_MIN_ELEMS
#define MAF_BAND_MIN_ELEMS 65536
#endif
#define MAF_RING_BANDS 3
static void ensure_stage(maf_handle* h) {
  const HostModel& M = h->M;
  const size_t ne = (size_t)(h->e1 - h->e0);
  const bool rows_aligned = h->e0 % M.num1el == 0 && h->e1 % M.num1el == 0;
  int64_t band_rows = 0;   // 0: no banding
  if ((int64_t)ne >= MAF_BAND_MIN_ELEMS && rows_aligned) {
    const int64_t nrows = (int64_t)ne / M.num1el;
    const double per_el = ((double)81 * h->nij + 72) * sizeof(double);
    const double nz_bytes = (double)(h->slot_hi - h->slot_lo) * sizeof(double);
    band_rows = std::max<int64_t>(4, (int64_t)(0.10 * nz_bytes / (MAF_RING_BANDS * per_el * M.num1el)));
    if (band_rows * MAF_RING_BANDS >= nrows) band_rows = 0;
  }
  if (band_rows) {   // the band arithmetic relies on the structured node numbering of the reference's patches
    const int64_t num1np = M.num1el + 2;
    bool ok = M.numnp == num1np * (M.num2el + 2);
    for (int64_t e = h->e0; e < h->e1 && ok; ++e)
      for (int a = 0; a < 9; ++a)
        if (M.IX0[9 * e + a] != (e % M.num1el + a % 3) + num1np * (e / M.num1el + a / 3)) { ok = false; break; }
    if (!ok) band_rows = 0;
  }
  if (const char* e = std::getenv("MAF_BAND_ROWS")) {   // tests force bands on small meshes
    band_rows = std::atoll(e);
    if (band_rows < 2 || !rows_aligned || band_rows * 2 >= (int64_t)ne / M.num1el) band_rows = 0;
    const int64_t num1np = M.num1el + 2;
    bool ok = M.numnp == num1np * (M.num2el + 2);
    for (int64_t el = h->e0; el < h->e1 && ok; ++el)
      for (int a = 0; a < 9; ++a)
        if (M.IX0[9 * el + a] != (el % M.num1el + a % 3) + num1np * (el / M.num1el + a / 3)) { ok = false; break; }
    if (!ok) band_rows = 0;
  }
  const size_t need_elems = band_rows ? (size_t)(MAF_RING_BANDS * band_rows * M.num1el) : ne;
  if (h->d_kel && h->kel_elems >= need_elems && h->band_rows == band_rows && h->band_e0 == h->e0 && h->band_e1 == h->e1)
    return;
  if (h->d_kel) {
    cudaFree(h->d_kel);
    cudaFree(h->d_rel);
    h->d_kel = h->d_rel = nullptr;
  }
  for (auto& b : h->bands) cudaFree(b.d_order);
  h->bands.clear();
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t need = need_elems * ((size_t)81 * h->nij + 72) * sizeof(double);
  if (need + ((size_t)1 << 30) > free_b)
    throw std::runtime_error("deterministic scatter needs " + std::to_string(need >> 20) +
                             " MiB of staging memory, more than is free on the device");
  CU(cudaMalloc(&h->d_kel, need_elems * 81 * h->nij * sizeof(double)));
  CU(cudaMalloc(&h->d_rel, need_elems * 72 * sizeof(double)));
  h->kel_elems = need_elems;
  h->band_rows = band_rows;
  h->band_e0 = h->e0;
  h->band_e1 = h->e1;
  if (band_rows) {
    const int64_t r0 = h->e0 / M.num1el, r1 = h->e1 / M.num1el;
    for (int64_t r = r0; r < r1; r += band_rows) {
      maf_handle::Band b;
      b.e0 = r * M.num1el;
      b.e1 = std::min(r + band_rows, r1) * M.num1el;
      std::vector<int32_t> order;
      build_element_order(M.num1el, b.e0, b.e1, order);
      CU(cudaMalloc(&b.d_order, order.size() * sizeof(int32_t)));
      CU(cudaMemcpy(b.d_order, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      h->bands.push_back(b);
    }
    if (!h->gather_stream) {
      int least = 0, greatest = 0;
      CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      CU(cudaStreamCreateWithPriority(&h->gather_stream, cudaStreamNonBlocking, greatest));
    }
    while (h->band_staged.size() < h->bands.size()) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->band_staged.push_back(e);
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->band_gathered.push_back(e);
    }
  }
}

// timing events: while a stream is being captured into a graph they are recorded as external event nodes, so that the
// replayed graph keeps producing the per-phase device times
static void rec(cudaEvent_t e, cudaStream_t s, bool capturing) {
  if (capturing) CU(cudaEventRecordWithFlags(e, s, cudaEventRecordExternal));
  else CU(cudaEventRecord(e, s));
}

// the launches of one assembly on stream s (everything asynchronous); returns the number of kernels launched
static int enqueue_assembly(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt,
                            double bend_tm, int mode, double* d_r, double* d_nz, double* d_rn, cudaStream_t s,
                            bool capturing) {
  const HostModel& M = h->M;
  const int64_t launches0 = h->launches;
  area_fn kern = area_kernel_of(M.motion, mode == MAF_SCATTER_DETERMINISTIC);
  const int64_t ne = h->e1 - h->e0;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(ne, 1), (int64_t)h->grid);
  rec(h->ev[1], s, capturing);
  StageSink st{nullptr, nullptr, 0, 0};
  const bool substrips = mode == MAF_SCATTER_ATOMIC && !capturing && h->host_copy_follows && h->side_stream &&
                         !h->strips.empty() && h->strips_e0 == h->e0 && h->strips_e1 == h->e1;
  if (substrips) {
    // Large range: sub-strips of element rows, assembled one after the other. Only the first one waits for its
    // zero-fill; the slots of the later ones (disjoint: everything above what the earlier strips touch) are zeroed on
    // the second stream while the first strip is being assembled -- the 3 % of a step the zero-fill used to cost --
    // followed there by the Neumann boundary kernels. Every strip records an event when its ranges are final
    // (maf_assemble copies them to the host while the next strip runs).
    const size_t ns = h->strips.size();
    CU(cudaMemsetAsync(d_r + h->eq_lo, 0, sizeof(double) * (size_t)(h->eq_hi - h->eq_lo), s));
    CU(cudaMemsetAsync(d_nz + h->slot_lo, 0, sizeof(double) * (size_t)(h->strips[0].slot_hi - h->slot_lo), s));
    rec(h->ev[6], s, false);
    CU(cudaEventRecord(h->ev_side[1], s));
    CU(cudaStreamWaitEvent(h->side_stream, h->ev_side[1], 0));
    for (size_t q = 1; q < ns; ++q) {
      const int64_t lo = h->strips[q - 1].slot_hi, hi = h->strips[q].slot_hi;
      if (hi > lo) CU(cudaMemsetAsync(d_nz + lo, 0, sizeof(double) * (size_t)(hi - lo), h->side_stream));
      CU(cudaEventRecord(h->zero_ev[q], h->side_stream));
    }
    for (int bc = 0; bc < M.n_neu; ++bc) {
      const int n = M.b_offs[bc + 1] - M.b_offs[bc];
      if (n == 0) continue;
      const double fval = neumann_value(M.b_type[bc], M.b_val[bc], time, bend_tm);
      const int gb = std::min((n + 3) / 4, h->sm_count * 8);
      boundary_kernel<<<gb, 128, 0, h->side_stream>>>(M.cfg, h->T, h->BT, bc, fval, dt, d_xms, d_r, d_nz, -1, 1, h->e0, h->e1);
      CU(cudaGetLastError());
      h->launches += 1;
    }
    CU(cudaEventRecord(h->ev_side[0], h->side_stream));
    const int q0 = (int)(h->ring_n % MAF_RING);
    rec(h->ring_a[q0], s, false);
    for (size_t q = 0; q < ns; ++q) {
      const maf_handle::Strip& sp = h->strips[q];
      if (q > 0) CU(cudaStreamWaitEvent(s, h->zero_ev[q], 0));
      const int gq = (int)std::min<int64_t>(sp.e1 - sp.e0, (int64_t)h->grid);
      kern<<<gq, MAF_NT, h->smem_bytes, s>>>(M.cfg, h->T, d_xms, d_cps, dt, d_r, d_nz, st, sp.d_order, sp.e0, sp.e1);
      CU(cudaGetLastError());
      h->launches += 1;
      if (q == 0) CU(cudaStreamWaitEvent(s, h->ev_side[0], 0));   // (the boundary terms are long done by then)
      CU(cudaEventRecord(h->strip_ev[q], s));
    }
    rec(h->ring_b[q0], s, false);
    h->ring_n += 1;
    rec(h->ev[2], s, false);
    rec(h->ev[3], s, false);
    rec(h->ev[4], s, false);
    if (d_rn) {
      rnorm2_partial<<<256, 256, 0, s>>>(d_r, M.nmdf, h->d_part);
      rnorm2_final<<<1, 256, 0, s>>>(h->d_part, 256, d_rn);
      CU(cudaGetLastError());
      h->launches += 2;
    }
    return (int)(h->launches - launches0);
  }
  if (mode == MAF_SCATTER_ATOMIC) {
    // only what this element range touches (contiguous, because unknowns are numbered node-major)
    if (d_r == h->d_r && d_nz == h->d_nz && !h->strip && h->eq_lo == 0 && h->eq_hi == M.nmdf && h->slot_lo == 0 &&
        h->slot_hi == M.sym.nnz) {
      CU(cudaMemsetAsync(d_r, 0, sizeof(double) * (h->rk_pad + (size_t)M.sym.nnz), s));
    } else {
      CU(cudaMemsetAsync(d_r + h->eq_lo, 0, sizeof(double) * (size_t)(h->eq_hi - h->eq_lo), s));
      CU(cudaMemsetAsync(d_nz + h->slot_lo, 0, sizeof(double) * (size_t)(h->slot_hi - h->slot_lo), s));
    }
  } else {
    st = StageSink{h->d_kel, h->d_rel, h->nij, h->band_rows ? (int64_t)h->kel_elems : 0};
  }
  rec(h->ev[6], s, capturing);
  // atomics path: the (tiny) Neumann boundary kernels only add into r / nzval, in any order: they run on a second
  // stream beside the area kernel instead of after it
  const bool side = mode == MAF_SCATTER_ATOMIC && M.n_neu > 0 && h->side_stream;
  if (side) {
    CU(cudaEventRecord(h->ev_side[1], s));
    CU(cudaStreamWaitEvent(h->side_stream, h->ev_side[1], 0));
    int nmax = 0;
    for (int bc = 0; bc < M.n_neu; ++bc) nmax = std::max(nmax, M.b_offs[bc + 1] - M.b_offs[bc]);
    if (M.n_neu <= 8 && nmax > 0 && (int64_t)nmax * M.n_neu <= 2048) {   // few boundary elements: a CTA each, one launch
      NeumannValues nv;
      for (int bc = 0; bc < 8; ++bc)
        nv.fval[bc] = bc < M.n_neu ? neumann_value(M.b_type[bc], M.b_val[bc], time, bend_tm) : 0.0;
      boundary_kernel_cta<<<dim3((unsigned)nmax, (unsigned)M.n_neu), 128, 0, h->side_stream>>>(
          M.cfg, h->T, h->BT, nv, dt, d_xms, d_r, d_nz, h->e0, h->e1);
      CU(cudaGetLastError());
      h->launches += 1;
    } else {
      for (int bc = 0; bc < M.n_neu; ++bc) {
        const int n = M.b_offs[bc + 1] - M.b_offs[bc];
        if (n == 0) continue;
        const double fval = neumann_value(M.b_type[bc], M.b_val[bc], time, bend_tm);
        const int gb = std::min((n + 3) / 4, h->sm_count * 8);
        boundary_kernel<<<gb, 128, 0, h->side_stream>>>(M.cfg, h->T, h->BT, bc, fval, dt, d_xms, d_r, d_nz, -1, 1, h->e0, h->e1);
        CU(cudaGetLastError());
        h->launches += 1;
      }
    }
    CU(cudaEventRecord(h->ev_side[0], h->side_stream));
  }
  const bool banded = mode == MAF_SCATTER_DETERMINISTIC && h->band_rows > 0;
  if (ne > 0 && !banded) {
    // (a replayed graph keeps writing the ring slot it was captured with)
    const int q = (int)(h->ring_n % MAF_RING);
    rec(h->ring_a[q], s, capturing);
    kern<<<grid, MAF_NT, h->smem_bytes, s>>>(M.cfg, h->T, d_xms, d_cps, dt, d_r, d_nz, st, h->d_order, h->e0, h->e1);
    CU(cudaGetLastError());
    rec(h->ring_b[q], s, capturing);
    h->ring_n += 1;
    h->launches += 1;
  }
  rec(h->ev[2], s, capturing);
  if (mode == MAF_SCATTER_DETERMINISTIC && !banded) {
    const int gb = h->sm_count * 16;
    const int64_t p_lo = M.sym.nbr_ptr[h->node_lo], p_hi = M.sym.nbr_ptr[h->node_hi];
    gather_K_kernel<<<gb, 128, 0, s>>>(M.cfg, h->T, h->G, h->d_kel, h->nij, h->e0, h->e1, p_lo, p_hi, d_nz);
    CU(cudaGetLastError());
    gather_r_kernel<<<gb, 128, 0, s>>>(M.cfg, h->T, h->G, h->d_rel, h->e0, h->e1, h->node_lo * M.ndf,
                                       h->node_hi * M.ndf, d_r);
    CU(cudaGetLastError());
    h->launches += 2;
  }
  if (banded) {
    // band after band: stage the band's elements (ring of three bands), then gather the node rows that are complete
    // Node row n of the patch takes contributions from the element rows n - 2 .. n only (quadratic splines).
    GatherTables G = h->G;
    G.ring = (int64_t)h->kel_elems;
    const int64_t num1np = M.num1el + 2;          // node rows are num1el + 2 nodes long (IX of Mesh.jl:574-593)
    int64_t node_done = h->node_lo;               // first node not gathered yet
    const int gb = h->sm_count * 16;
    cudaStream_t gs = h->gather_stream;
    for (size_t b = 0; b < h->bands.size(); ++b) {
      const maf_handle::Band& B = h->bands[b];
      const int gridb = (int)std::min<int64_t>(B.e1 - B.e0, (int64_t)h->grid);
      // the ring slot of this band was last read by the gather of band b - 2
      if (b >= MAF_RING_BANDS - 1) CU(cudaStreamWaitEvent(s, h->band_gathered[b - (MAF_RING_BANDS - 1)], 0));
      // (the area kernel indexes the staging ring by el - e0 of the RANGE: pass the range start with the band's order)
      kern<<<gridb, MAF_NT, h->smem_bytes, s>>>(M.cfg, h->T, d_xms, d_cps, dt, d_r, d_nz, st, B.d_order, h->e0,
                                              h->e0 + (B.e1 - B.e0));
      CU(cudaGetLastError());
      h->launches += 1;
      CU(cudaEventRecord(h->band_staged[b], s));
      CU(cudaStreamWaitEvent(gs, h->band_staged[b], 0));
      const bool last = b + 1 == h->bands.size();
      // complete node rows: up to (first element row of the next band) - 1, i.e. all nodes below that row's first node
      const int64_t node_hi = last ? h->node_hi : std::min<int64_t>(h->node_hi, (B.e1 / M.num1el) * num1np);
      if (node_hi > node_done) {
        gather_K_kernel<<<gb, 128, 0, gs>>>(M.cfg, h->T, G, h->d_kel, h->nij, h->e0, h->e1, M.sym.nbr_ptr[node_done],
                                           M.sym.nbr_ptr[node_hi], d_nz);
        gather_r_kernel<<<gb, 128, 0, gs>>>(M.cfg, h->T, G, h->d_rel, h->e0, h->e1, node_done * M.ndf, node_hi * M.ndf, d_r);
        CU(cudaGetLastError());
        node_done = node_hi;
        h->launches += 2;
      }
      CU(cudaEventRecord(h->band_gathered[b], gs));
    }
    if (!h->bands.empty()) CU(cudaStreamWaitEvent(s, h->band_gathered[h->bands.size() - 1], 0));
  }
  rec(h->ev[3], s, capturing);
  if (side) CU(cudaStreamWaitEvent(s, h->ev_side[0], 0));
  for (int bc = 0; bc < M.n_neu && !side; ++bc) {
    const int n = M.b_offs[bc + 1] - M.b_offs[bc];
    if (n == 0) continue;
    const double fval = neumann_value(M.b_type[bc], M.b_val[bc], time, bend_tm);
    const int gb = std::min((n + 3) / 4, h->sm_count * 8);
    if (mode == MAF_SCATTER_ATOMIC) {
      boundary_kernel<<<gb, 128, 0, s>>>(M.cfg, h->T, h->BT, bc, fval, dt, d_xms, d_r, d_nz, -1, 1, h->e0, h->e1);
      CU(cudaGetLastError());
      h->launches += 1;
    } else {
      // neighbouring boundary elements share two node columns: three colours make concurrent units disjoint
      for (int c = 0; c < 3; ++c) {
        boundary_kernel<<<gb, 128, 0, s>>>(M.cfg, h->T, h->BT, bc, fval, dt, d_xms, d_r, d_nz, c, 3, h->e0, h->e1);
        CU(cudaGetLastError());
        h->launches += 1;
      }
    }
  }
  rec(h->ev[4], s, capturing);
  if (d_rn) {
    if (M.nmdf <= 32768) {   // one block sums a small system in the same fixed order every time
      rnorm2_partial<<<1, 256, 0, s>>>(d_r, M.nmdf, d_rn);
      h->launches += 1;
    } else {
      rnorm2_partial<<<256, 256, 0, s>>>(d_r, M.nmdf, h->d_part);
      rnorm2_final<<<1, 256, 0, s>>>(h->d_part, 256, d_rn);
      h->launches += 2;
    }
    CU(cudaGetLastError());
  }
  return (int)(h->launches - launches0);
}

// Small meshes (the reference's own 17 x 17 examples: ~300 elements, ~10 launches of a few microseconds each) are
// bound by the host's launch rate, not by the device: for them the launch sequence of an assembly into the handle's
// own buffers is captured ONCE per (scatter mode, dt, Neumann values) into a CUDA graph and replayed.
#ifndef MAF_GRAPH_MAX_ELEMS
#define MAF_GRAPH_MAX_ELEMS 16384
#endif
static void do_assemble_device(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt,
                               double bend_tm, int mode, double* d_r, double* d_nz, double* d_rn, cudaStream_t s,
                               bool timed) {
  const HostModel& M = h->M;
  if (!d_r) d_r = h->d_r;
  if (!d_nz) d_nz = h->d_nz;
  if (mode != MAF_SCATTER_ATOMIC && mode != MAF_SCATTER_DETERMINISTIC) throw std::runtime_error("unknown scatter mode");
  if (!(dt == dt) || !(time == time)) throw std::runtime_error("time / dt is NaN");
  (void)timed;
  h->timed_valid = false;
  if (mode == MAF_SCATTER_DETERMINISTIC) {   // allocations and uploads: never inside a capture
    ensure_gather(h);
    ensure_stage(h);
  } else if (h->host_copy_follows && wants_strips(h)) {
    // (measured on the 1001 x 1001 workload, device part alone: 1 piece 39.60 ms, 4 strips 39.64, 8 strips 40.02,
    // 16 strips 40.60 -- the hidden zero-fill is paid back in kernel tails; what pays is the copy-out of a
    // finished strip running beside the next strips' kernels)
    int ns = 8;
    if (const char* e = std::getenv("MAF_SUBSTRIPS")) ns = std::max(1, std::atoi(e));
    ensure_strips(h, ns);   // (kept across device-only calls in between; rebuilt when the element range changes)
  }
  const bool graphable = h->use_graph && !h->strip && s == h->stream && d_r == h->d_r && d_nz == h->d_nz &&
                         d_xms == h->d_xms && d_cps == h->d_cps && (d_rn == nullptr || d_rn == h->d_rn) &&
                         M.numel <= MAF_GRAPH_MAX_ELEMS && !(mode == MAF_SCATTER_ATOMIC && h->host_copy_follows && wants_strips(h));
  if (!graphable) {
    enqueue_assembly(h, d_xms, d_cps, time, dt, bend_tm, mode, d_r, d_nz, d_rn, s, false);
    h->timed_valid = true;
    return;
  }
  // key: everything that is baked into the kernel arguments
  std::vector<double> key = {(double)mode, dt, d_rn ? 1.0 : 0.0, (double)h->e0, (double)h->e1, (double)(size_t)h->d_kel};
  for (int bc = 0; bc < M.n_neu; ++bc) key.push_back(neumann_value(M.b_type[bc], M.b_val[bc], time, bend_tm));
  maf_handle::Graph* g = nullptr;
  for (auto& c : h->graphs)
    if (c.key.size() == key.size() && std::memcmp(c.key.data(), key.data(), key.size() * sizeof(double)) == 0) g = &c;
  if (!g) {
    if (h->graphs.size() >= 8) {   // a time-dependent Neumann value changes the key every step: keep the cache small
      CU(cudaGraphExecDestroy(h->graphs.front().exec));
      h->graphs.erase(h->graphs.begin());
    }
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int n = 0;
    try {
      n = enqueue_assembly(h, d_xms, d_cps, time, dt, bend_tm, mode, d_r, d_nz, d_rn, s, true);
    } catch (...) {
      cudaStreamEndCapture(s, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CU(cudaStreamEndCapture(s, &graph));
    maf_handle::Graph ng;
    ng.key = key;
    ng.launches = n;
    CU(cudaGraphInstantiate(&ng.exec, graph, 0));
    CU(cudaGraphDestroy(graph));
    h->launches -= n;   // counted per replay below
    h->graphs.push_back(ng);
    g = &h->graphs.back();
  }
  CU(cudaGraphLaunch(g->exec, s));
  h->launches += g->launches;
  h->graph_replays += 1;
  h->timed_valid = true;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
#define MAF_API_BEGIN(h)      \
  if (!(h)) return 1;         \
  try {                       \
    CU(cudaSetDevice((h)->device));
#define MAF_API_END(h)                  \
  }                                     \
  catch (std::exception & e) {          \
    (h)->err = e.what();                \
    return 2;                           \
  }                                     \
  catch (...) {                         \
    (h)->err = "unknown error";         \
    return 3;                           \
  }                                     \
  return 0;

extern "C" {

static void upload_state_rows(maf_handle* h, const double* xms, const double* cps);

const char* maf_last_error(const maf_handle* h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> lk(g_mu);
  static thread_local std::string copy;
  copy = g_create_err;
  return copy.c_str();
}

// rank < 0: the whole mesh on one device (maf_create); otherwise strip `rank` of `nranks` (maf_create_strip)
static int create_handle(maf_handle** out, const maf_mesh_desc* mesh, const maf_params* params, int rank, int nranks) {
  if (!out) return 1;
  *out = nullptr;
  maf_handle* h = new (std::nothrow) maf_handle();
  if (!h) return 1;
  try {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
      throw std::runtime_error("no CUDA device: libmembrane_b200 has no CPU fallback for the assembly");
    if (!params) throw std::runtime_error("null params");
    int dev = params->device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) throw std::runtime_error("device ordinal out of range");
    h->device = dev;
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    h->sm_count = prop.multiProcessorCount;

    build_host_model(h->M, mesh, params, MAF_NT);
    HostModel& M = h->M;
    h->e0 = 0;
    h->e1 = M.numel;
    if (rank >= 0) {
      h->strip = true;
      h->rank = rank;
      h->nranks = nranks;
      const TouchedRange R = strip_range(M, rank, nranks);
      h->e0 = R.e0;
      h->e1 = R.e1;
      if (rank > 0) h->lower = strip_range(M, rank - 1, nranks);
      if (rank + 1 < nranks) h->upper = strip_range(M, rank + 1, nranks);
      // the interface (the node rows two strips share) belongs to the UPPER strip: a strip hands out what it touches
      // below the first row / slot of the next strip
      h->own_eq_lo = R.eq_lo; h->own_slot_lo = R.slot_lo;
      h->own_eq_hi = rank + 1 < nranks ? h->upper.eq_lo : R.eq_hi;
      h->own_slot_hi = rank + 1 < nranks ? h->upper.slot_lo : R.slot_hi;
      if (rank > 0 && (h->lower.eq_hi > R.eq_hi || h->lower.slot_hi > R.slot_hi || h->lower.eq_lo > R.eq_lo))
        throw std::runtime_error("internal: strip ranges are not nested as expected");
    }
    compute_ranges(h);

    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 8; ++k) CU(cudaEventCreate(&h->ev[k]));
    for (int k = 0; k < MAF_RING; ++k) { CU(cudaEventCreate(&h->ring_a[k])); CU(cudaEventCreate(&h->ring_b[k])); }
    for (int k = 0; k < 2; ++k) CU(cudaEventCreateWithFlags(&h->ev_side[k], cudaEventDisableTiming));

    h->T.IX = upload(h, M.IX0.data(), M.IX0.size());
    h->T.ID = upload(h, M.ID0.data(), M.ID0.size());
    h->T.nodemask = upload(h, M.sym.nodemask.data(), M.sym.nodemask.size());
    h->T.uel1 = upload(h, M.uel1.data(), M.uel1.size());
    h->T.uel2 = upload(h, M.uel2.data(), M.uel2.size());
    h->T.line1 = upload(h, M.line1.data(), M.line1.size());
    h->T.line2 = upload(h, M.line2.data(), M.line2.size());
    h->T.tdb = upload(h, M.tdb.data(), M.tdb.size());
    h->T.colptr = upload(h, M.sym.colptr.data(), M.sym.colptr.size());
    // per-element tables: a strip keeps its own elements only (memory per rank ~ 1 / nranks)
    h->T.el0 = h->strip ? h->e0 : 0;
    const int64_t nel_tab = h->strip ? h->e1 - h->e0 : M.numel;
    h->T.elpair = upload(h, M.sym.elpair.data() + (size_t)81 * h->T.el0, (size_t)81 * nel_tab);
    h->T.pairoff = upload(h, M.sym.pairoff.data(), M.sym.pairoff.size());
    h->T.eq0 = upload(h, M.sym.eq0.data(), M.sym.eq0.size());
    h->T.nodecol = upload(h, M.nodecol.data(), M.nodecol.size());
    h->T.nodemask32 = upload(h, M.nodemask32.data(), M.nodemask32.size());
    h->T.utab = M.utab.empty() ? nullptr : upload(h, M.utab.data(), M.utab.size());
    h->T.numnp = M.numnp; h->T.numel = M.numel; h->T.num1el = M.num1el; h->T.nuel1 = M.nuel1;
    {   // scatter maps, built on the device from the tables above and merged into classes of equal maps
      int64_t* d_base = dalloc<int64_t>(h, (size_t)nel_tab);
      int32_t* d_class = dalloc<int32_t>(h, (size_t)nel_tab);
      int* d_flags = dalloc<int>(h, 2);
      unsigned long long* d_hash = nullptr;
      CU(cudaMalloc(&d_hash, sizeof(unsigned long long) * (size_t)std::max<int64_t>(nel_tab, 1)));
      CU(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), h->stream));
      h->T.elbase = d_base;
      h->T.elclass = d_class;
      if (nel_tab > 0)
        build_elslot_kernel<0><<<(unsigned)nel_tab, 128, 0, h->stream>>>(h->T, nullptr, nullptr, d_base, d_hash, d_flags);
      CU(cudaGetLastError());
      std::vector<unsigned long long> hh((size_t)nel_tab);
      CU(cudaMemcpyAsync(hh.data(), d_hash, sizeof(unsigned long long) * (size_t)nel_tab, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      cudaFree(d_hash);
      std::vector<int32_t> cls((size_t)nel_tab), rep;
      {
        std::unordered_map<unsigned long long, int32_t> seen;
        seen.reserve(4096);
        for (int64_t k = 0; k < nel_tab; ++k) {
          auto it = seen.find(hh[(size_t)k]);
          if (it == seen.end()) {
            it = seen.emplace(hh[(size_t)k], (int32_t)rep.size()).first;
            rep.push_back((int32_t)(h->T.el0 + k));
          }
          cls[(size_t)k] = it->second;
        }
      }
      h->n_slot_classes = (int64_t)rep.size();
      int32_t* d_rep = nullptr;
      CU(cudaMalloc(&d_rep, sizeof(int32_t) * std::max<size_t>(rep.size(), 1)));
      int32_t* d_slot = dalloc<int32_t>(h, (size_t)MAF_SLOT_INTS * std::max<size_t>(rep.size(), 1));
      h->T.elslot = d_slot;
      CU(cudaMemcpyAsync(d_rep, rep.data(), sizeof(int32_t) * rep.size(), cudaMemcpyHostToDevice, h->stream));
      CU(cudaMemcpyAsync(d_class, cls.data(), sizeof(int32_t) * cls.size(), cudaMemcpyHostToDevice, h->stream));
      if (!rep.empty()) {
        build_elslot_kernel<1><<<(unsigned)rep.size(), 128, 0, h->stream>>>(h->T, d_rep, d_slot, nullptr, nullptr, d_flags);
        build_elslot_kernel<2><<<(unsigned)nel_tab, 128, 0, h->stream>>>(h->T, nullptr, nullptr, nullptr, nullptr, d_flags);
      }
      CU(cudaGetLastError());
      int fl[2] = {0, 0};
      CU(cudaMemcpyAsync(fl, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      cudaFree(d_rep);
      if (fl[0]) throw std::runtime_error("scatter map offset exceeds 32 bits (an element spans more than 2^31 stored entries)");
      if (fl[1]) throw std::runtime_error("internal: two different scatter maps share a hash");
    }
    h->BT.edge1 = upload(h, M.edge1.data(), M.edge1.size());
    h->BT.edge2 = upload(h, M.edge2.data(), M.edge2.size());
    h->BT.elems = upload(h, M.b_elems.data(), M.b_elems.size());
    h->BT.offs = upload(h, M.b_offs.data(), M.b_offs.size());
    h->BT.bdry = upload(h, M.b_bdry.data(), M.b_bdry.size());
    h->BT.ntype = upload(h, M.b_type.data(), M.b_type.size());
    h->BT.nval = upload(h, M.b_val.data(), M.b_val.size());
    h->BT.n_neu = M.n_neu;
    // elpair is only needed on the device from here on; the staging tables of the deterministic path are
    // uploaded on first use (ensure_gather)
    h->d_xms = dalloc<double>(h, (size_t)3 * M.numnp);
    h->d_cps = dalloc<double>(h, (size_t)M.ndf * M.numnp);
    if (!h->strip) {
      // r and nzval side by side (r padded to 256 bytes): a whole-mesh assembly zeroes both with ONE memset
      h->rk_pad = (((size_t)M.nmdf * sizeof(double) + 255) / 256) * 256 / sizeof(double);
      h->d_r = dalloc<double>(h, h->rk_pad + (size_t)M.sym.nnz);
      h->d_nz = h->d_r + h->rk_pad;
    } else {
      // ONE allocation (exported to the neighbours as one IPC handle): flags | r slice | nzval slice. d_r / d_nz are
      // kept as VIRTUAL bases (slice start minus the first row / slot of the slice) so that every kernel keeps
      // addressing rows and slots by their global index; only [eq_lo, eq_hi) / [slot_lo, slot_hi) is ever touched.
      TouchedRange R;
      R.eq_lo = h->eq_lo; R.eq_hi = h->eq_hi; R.slot_lo = h->slot_lo; R.slot_hi = h->slot_hi;
      h->strip_bytes = strip_alloc_bytes(R);
      CU(cudaMalloc(&h->strip_alloc, h->strip_bytes));
      CU(cudaMemset(h->strip_alloc, 0, h->strip_bytes));
      h->flags = reinterpret_cast<long long*>(h->strip_alloc);
      h->d_r = reinterpret_cast<double*>((char*)h->strip_alloc + strip_r_offset()) - h->eq_lo;
      h->d_nz = reinterpret_cast<double*>((char*)h->strip_alloc + strip_nz_offset(R)) - h->slot_lo;
      for (int k = 0; k < 4; ++k) CU(cudaEventCreate(&h->ev_x[k]));
    }
    h->d_rn = dalloc<double>(h, 1);
    h->d_part = dalloc<double>(h, 256);
    std::vector<int32_t>().swap(M.sym.elpair);

    h->smem_bytes = (size_t)M.cfg.smem_doubles * sizeof(double);
    area_fn kern = area_kernel_of(M.motion);
    if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
      throw std::runtime_error("element needs more shared memory than the device offers");
    for (int staged = 0; staged < 2; ++staged) {
      area_fn kq = area_kernel_of(M.motion, staged != 0);
      CU(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
      CU(cudaFuncSetAttribute(kq, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, MAF_NT, h->smem_bytes));
    if (nb < 1) throw std::runtime_error("area kernel does not fit on an SM");
    h->ctas_per_sm = nb;
    h->grid = nb * h->sm_count;  // persistent grid: a multiple of the SM count
    if (const char* e = std::getenv("MAF_NO_GRAPH")) h->use_graph = !(*e && *e != '0');
  } catch (std::exception& e) {
    {
      std::lock_guard<std::mutex> lk(g_mu);
      g_create_err = e.what();
    }
    maf_destroy(h);
    return 2;
  }
  *out = h;
  return 0;
}

int maf_create(maf_handle** out, const maf_mesh_desc* mesh, const maf_params* params) {
  return create_handle(out, mesh, params, -1, 1);
}

int maf_create_strip(maf_handle** out, const maf_mesh_desc* mesh, const maf_params* params, int32_t rank,
                     int32_t nranks) {
  if (rank < 0 || nranks < 1 || rank >= nranks) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_create_err = "strip rank outside 0..nranks-1";
    if (out) *out = nullptr;
    return 2;
  }
  return create_handle(out, mesh, params, rank, nranks);
}

int maf_destroy(maf_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->lower_base && h->lower_ipc) cudaIpcCloseMemHandle(h->lower_base);
  if (h->upper_base && h->upper_ipc) cudaIpcCloseMemHandle(h->upper_base);
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (auto& b : h->bands) cudaFree(b.d_order);
  if (h->strip_alloc) cudaFree(h->strip_alloc);
  if (h->side_stream) { cudaStreamSynchronize(h->side_stream); cudaStreamDestroy(h->side_stream); }
  for (int k = 0; k < MAF_RING; ++k) {
    if (h->ring_a[k]) cudaEventDestroy(h->ring_a[k]);
    if (h->ring_b[k]) cudaEventDestroy(h->ring_b[k]);
  }
  for (int k = 0; k < 2; ++k)
    if (h->ev_side[k]) cudaEventDestroy(h->ev_side[k]);
  for (int k = 0; k < 4; ++k)
    if (h->ev_x[k]) cudaEventDestroy(h->ev_x[k]);
  for (void* p : h->allocs) cudaFree(p);
  if (h->d_order) cudaFree(h->d_order);
  for (auto& st : h->strips) cudaFree(st.d_order);
  for (auto& e : h->strip_ev) cudaEventDestroy(e);
  for (auto& e : h->zero_ev) cudaEventDestroy(e);
  for (auto& e : h->band_staged) cudaEventDestroy(e);
  for (auto& e : h->band_gathered) cudaEventDestroy(e);
  if (h->gather_stream) { cudaStreamSynchronize(h->gather_stream); cudaStreamDestroy(h->gather_stream); }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->d_kel) cudaFree(h->d_kel);
  if (h->d_rel) cudaFree(h->d_rel);
  if (h->h_pin_in) cudaFreeHost(h->h_pin_in);
  if (h->h_pin_out) cudaFreeHost(h->h_pin_out);
  for (int k = 0; k < 8; ++k)
    if (h->ev[k]) cudaEventDestroy(h->ev[k]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int maf_nnz(maf_handle* h, int64_t* nnz) {
  MAF_API_BEGIN(h)
  if (!nnz) throw std::runtime_error("null output pointer");
  *nnz = h->M.sym.nnz;
  MAF_API_END(h)
}

int maf_pattern(maf_handle* h, int64_t* colptr, int64_t* rowval) {
  MAF_API_BEGIN(h)
  if (!colptr || !rowval) throw std::runtime_error("null output pointer");
  const HostModel& M = h->M;
  for (int64_t c = 0; c <= M.nmdf; ++c) colptr[c] = M.sym.colptr[c] + 1;
  build_rowval(M.sym, M.ID0.data(), M.cfg.rowmask, rowval);
  MAF_API_END(h)
}

int maf_colptr(maf_handle* h, int64_t* colptr) {
  MAF_API_BEGIN(h)
  if (!colptr) throw std::runtime_error("null output pointer");
  for (int64_t c = 0; c <= h->M.nmdf; ++c) colptr[c] = h->M.sym.colptr[c] + 1;
  MAF_API_END(h)
}

int maf_pattern_columns(maf_handle* h, int64_t col_first, int64_t col_last, int64_t* rowval) {
  MAF_API_BEGIN(h)
  if (!rowval) throw std::runtime_error("null output pointer");
  if (col_first < 1 || col_last > h->M.nmdf || col_first > col_last + 1)
    throw std::runtime_error("column range outside 1..nmdf");
  build_rowval_columns(h->M.sym, h->M.ID0.data(), h->M.cfg.rowmask, col_first - 1, col_last, rowval);
  MAF_API_END(h)
}

// is this host pointer page-locked (cudaMallocHost / cudaHostRegister)? then it can be copied from directly
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// host state -> the handle's device buffers (asynchronous on the handle's stream)
static void upload_state(maf_handle* h, const double* xms, const double* cps) {
  const HostModel& M = h->M;
  cudaStream_t s = h->stream;
  const size_t bx = sizeof(double) * 3 * (size_t)M.numnp, bc = sizeof(double) * (size_t)M.ndf * M.numnp;
  // inputs: page-locked caller memory is copied from directly, anything else through a pinned staging buffer
  // (keeps the copy asynchronous and at full PCIe rate)
  const double *hx = xms, *hc = cps;
  if (!is_pinned(xms) || !is_pinned(cps)) {
    if (h->pin_in_bytes < bx + bc) {
      if (h->h_pin_in) cudaFreeHost(h->h_pin_in);
      CU(cudaMallocHost(&h->h_pin_in, bx + bc));
      h->pin_in_bytes = bx + bc;
    }
    std::memcpy(h->h_pin_in, xms, bx);
    std::memcpy((char*)h->h_pin_in + bx, cps, bc);
    hx = h->h_pin_in;
    hc = (const double*)((char*)h->h_pin_in + bx);
  }
  CU(cudaMemcpyAsync(h->d_xms, hx, bx, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->d_cps, hc, bc, cudaMemcpyHostToDevice, s));
  h->state_resident = true;
}

// assembly of the resident state with the results copied into host buffers; ev[0] must have been recorded
static void assemble_to_host(maf_handle* h, double time, double dt, double bend_tm, int scatter_mode, double* r,
                             double* nzval, double* rnorm2) {
  if (!r || !nzval) throw std::runtime_error("null buffer");
  if (scatter_mode != MAF_SCATTER_ATOMIC && scatter_mode != MAF_SCATTER_DETERMINISTIC)
    throw std::runtime_error("unknown scatter mode");
  const HostModel& M = h->M;
  cudaStream_t s = h->stream;
  const size_t br = sizeof(double) * (size_t)M.nmdf, bk = sizeof(double) * (size_t)M.sym.nnz;
  // Large ranges on the atomics path are assembled in sub-strips of element rows (enqueue_assembly): everything
  // below the first slot a later strip touches is final once a strip is done, so its device-to-host copy (the
  // dominant cost: nnz * 8 bytes over PCIe) overlaps the next strips.
  const bool pipelined = scatter_mode == MAF_SCATTER_ATOMIC && h->e0 == 0 && h->e1 == M.numel && wants_strips(h);
  h->host_copy_follows = pipelined;
  try {
    do_assemble_device(h, h->d_xms, h->d_cps, time, dt, bend_tm, scatter_mode, h->d_r, h->d_nz,
                       rnorm2 ? h->d_rn : nullptr, s, true);
  } catch (...) {
    h->host_copy_follows = false;
    throw;
  }
  h->host_copy_follows = false;
  if (!pipelined) {
    // results straight into the caller's buffers (pageable destinations are staged by the driver)
    CU(cudaMemcpyAsync(r, h->d_r, br, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(nzval, h->d_nz, bk, cudaMemcpyDeviceToHost, s));
    if (rnorm2) CU(cudaMemcpyAsync(rnorm2, h->d_rn, sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(h->ev[5], s));
    CU(cudaStreamSynchronize(s));
  } else {
    const size_t ns = h->strips.size();
    for (size_t q = 0; q < ns; ++q) {
      const int64_t s_lo = q == 0 ? 0 : h->strips[q].slot_lo, s_hi = q + 1 < ns ? h->strips[q + 1].slot_lo : M.sym.nnz;
      const int64_t r_lo = q == 0 ? 0 : h->strips[q].eq_lo, r_hi = q + 1 < ns ? h->strips[q + 1].eq_lo : M.nmdf;
      CU(cudaStreamWaitEvent(h->copy_stream, h->strip_ev[q], 0));
      if (s_hi > s_lo)
        CU(cudaMemcpyAsync(nzval + s_lo, h->d_nz + s_lo, sizeof(double) * (size_t)(s_hi - s_lo),
                           cudaMemcpyDeviceToHost, h->copy_stream));
      if (r_hi > r_lo)
        CU(cudaMemcpyAsync(r + r_lo, h->d_r + r_lo, sizeof(double) * (size_t)(r_hi - r_lo), cudaMemcpyDeviceToHost,
                           h->copy_stream));
    }
    // (a copy into pageable host memory blocks the host until the stream drains: issue it last)
    if (rnorm2) CU(cudaMemcpyAsync(rnorm2, h->d_rn, sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(h->copy_stream));
    CU(cudaEventRecord(h->ev[5], s));
    CU(cudaStreamSynchronize(s));
  }
  CU(cudaEventElapsedTime(&h->ms[0], h->ev[0], h->ev[1]));
  CU(cudaEventElapsedTime(&h->ms[4], h->ev[4], h->ev[5]));
  CU(cudaEventElapsedTime(&h->ms[5], h->ev[0], h->ev[5]));
}

int maf_assemble(maf_handle* h, const double* xms, const double* cps, double time, double dt, double bend_tm,
                 int scatter_mode, double* r, double* nzval, double* rnorm2) {
  MAF_API_BEGIN(h)
  if (h->strip) throw std::runtime_error("strip handle: use maf_assemble_strip / maf_assemble_strip_host");
  if (!xms || !cps || !r || !nzval) throw std::runtime_error("null buffer");
  CU(cudaEventRecord(h->ev[0], h->stream));
  upload_state(h, xms, cps);
  assemble_to_host(h, time, dt, bend_tm, scatter_mode, r, nzval, rnorm2);
  MAF_API_END(h)
}

// ---- device-resident state: xms / cps stay on the device across the Newton iterations ----
static void require_state(maf_handle* h) {
  if (!h->state_resident) throw std::runtime_error("no resident state: call maf_state_set (or maf_assemble) first");
}

int maf_state_set(maf_handle* h, const double* xms, const double* cps) {
  MAF_API_BEGIN(h)
  if (!xms || !cps) throw std::runtime_error("null buffer");
  if (h->strip) {   // only the node rows the strip reads (the arrays are the caller's full ones)
    upload_state_rows(h, xms, cps);
    CU(cudaStreamSynchronize(h->stream));
    return 0;
  }
  upload_state(h, xms, cps);
  CU(cudaStreamSynchronize(h->stream));   // the caller may reuse its buffers
  MAF_API_END(h)
}

int maf_state_get(maf_handle* h, double* xms, double* cps) {
  MAF_API_BEGIN(h)
  require_state(h);
  const HostModel& M = h->M;
  if (xms) CU(cudaMemcpyAsync(xms, h->d_xms, sizeof(double) * 3 * (size_t)M.numnp, cudaMemcpyDeviceToHost, h->stream));
  if (cps) CU(cudaMemcpyAsync(cps, h->d_cps, sizeof(double) * (size_t)M.ndf * M.numnp, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  MAF_API_END(h)
}

int maf_state_update(maf_handle* h, const double* du, double dt) {
  MAF_API_BEGIN(h)
  require_state(h);
  if (!du) throw std::runtime_error("null buffer");
  if (!(dt == dt)) throw std::runtime_error("dt is NaN");
  const HostModel& M = h->M;
  if (!h->d_du) h->d_du = dalloc<double>(h, (size_t)M.nmdf);
  CU(cudaMemcpyAsync(h->d_du, du, sizeof(double) * (size_t)M.nmdf, cudaMemcpyHostToDevice, h->stream));
  const int grid = (int)std::min<int64_t>((M.numnp * M.ndf + 255) / 256, (int64_t)h->sm_count * 16);
  state_update_kernel<<<std::max(grid, 1), 256, 0, h->stream>>>(M.cfg, h->T, h->d_du, dt, h->d_xms, h->d_cps);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaStreamSynchronize(h->stream));   // du is the caller's
  MAF_API_END(h)
}

int maf_state_predict(maf_handle* h, double dt) {
  MAF_API_BEGIN(h)
  require_state(h);
  if (!(dt == dt)) throw std::runtime_error("dt is NaN");
  const HostModel& M = h->M;
  const int grid = (int)std::min<int64_t>((M.numnp * 3 + 255) / 256, (int64_t)h->sm_count * 16);
  state_predict_kernel<<<std::max(grid, 1), 256, 0, h->stream>>>(M.cfg, h->T, dt, h->d_xms, h->d_cps);
  CU(cudaGetLastError());
  h->launches += 1;
  MAF_API_END(h)
}

int maf_assemble_resident(maf_handle* h, double time, double dt, double bend_tm, int scatter_mode, double* r,
                          double* nzval, double* rnorm2) {
  MAF_API_BEGIN(h)
  require_state(h);
  if (h->strip) throw std::runtime_error("strip handle: use maf_assemble_strip / maf_assemble_strip_host");
  CU(cudaEventRecord(h->ev[0], h->stream));
  assemble_to_host(h, time, dt, bend_tm, scatter_mode, r, nzval, rnorm2);
  MAF_API_END(h)
}

int maf_generate_output(maf_handle* h, double* xout, double* uout) {
  MAF_API_BEGIN(h)
  require_state(h);
  if (!xout || !uout) throw std::runtime_error("null buffer");
  if (h->strip) throw std::runtime_error("generate_output needs a whole-mesh handle");
  const HostModel& M = h->M;
  {   // the sampling grid is the reference's structured patch (IX of Mesh.jl:574-593)
    const int64_t num1np = M.num1el + 2;
    bool ok = M.numnp == num1np * (M.num2el + 2);
    for (int64_t e = 0; e < M.numel && ok; ++e)
      for (int a = 0; a < 9; ++a)
        if (M.IX0[9 * e + a] != (e % M.num1el + a % 3) + num1np * (e / M.num1el + a / 3)) { ok = false; break; }
    if (!ok) throw std::runtime_error("generate_output needs the structured patch numbering of Mesh.jl:574-593");
  }
  const int64_t npt = (3 * (int64_t)M.num1el + 2) * (3 * (int64_t)M.num2el + 2);
  double *dx = nullptr, *du = nullptr;
  CU(cudaMalloc(&dx, sizeof(double) * (size_t)npt * 3));
  cudaError_t err = cudaMalloc(&du, sizeof(double) * (size_t)npt * M.ndf);
  if (err == cudaSuccess) {
    const int grid = (int)std::min<int64_t>((npt * (3 + M.ndf) + 255) / 256, (int64_t)h->sm_count * 32);
    sample_output_kernel<<<grid, 256, 0, h->stream>>>(h->T, h->BT, M.ndf, M.num2el, h->d_xms, h->d_cps, dx, du);
    h->launches += 1;
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(xout, dx, sizeof(double) * (size_t)npt * 3, cudaMemcpyDeviceToHost, h->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(uout, du, sizeof(double) * (size_t)npt * M.ndf, cudaMemcpyDeviceToHost, h->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
  cudaFree(dx);
  if (du) cudaFree(du);
  CU(err);
  MAF_API_END(h)
}

int maf_elem_v_residuals(maf_handle* h, const int64_t* el_ids, int64_t n, double* rv) {
  MAF_API_BEGIN(h)
  require_state(h);
  if (!el_ids || !rv) throw std::runtime_error("null buffer");
  if (n < 0 || n > 4096) throw std::runtime_error("element count outside 0..4096");
  const HostModel& M = h->M;
  if (M.cfg.fdof[F_V][0] < 0 || M.cfg.fdof[F_V][1] < 0 || M.cfg.fdof[F_V][2] < 0)
    throw std::runtime_error("the pull force needs a 3-D velocity");   // Analysis.jl:48
  if (n == 0) return 0;
  std::vector<int32_t> els((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    if (el_ids[k] < 1 || el_ids[k] > M.numel) throw std::runtime_error("element id outside 1..numel");
    if (h->strip && (el_ids[k] - 1 < h->e0 || el_ids[k] - 1 >= h->e1))
      throw std::runtime_error("element outside this strip (a strip handle holds the tables of its own elements)");
    els[(size_t)k] = (int32_t)(el_ids[k] - 1);
  }
  int32_t* d_els = nullptr;
  double* d_rel = nullptr;
  CU(cudaMalloc(&d_els, sizeof(int32_t) * (size_t)n));
  CU(cudaMalloc(&d_rel, sizeof(double) * 72 * (size_t)n));
  std::vector<double> rel((size_t)72 * n);
  cudaError_t err = cudaMemcpyAsync(d_els, els.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (err == cudaSuccess) {
    elres_fn kern = elem_residual_kernel_of(M.motion);
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (err == cudaSuccess) {
      kern<<<(int)n, MAF_NT, h->smem_bytes, h->stream>>>(M.cfg, h->T, h->d_xms, h->d_cps, d_els, (int)n, d_rel);
      h->launches += 1;
      err = cudaGetLastError();
    }
  }
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(rel.data(), d_rel, sizeof(double) * 72 * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
  cudaFree(d_els);
  cudaFree(d_rel);
  CU(err);
  // staged [u][a] (u = v0 v1 v2 ...) -> the reference's rv_el[comp + 3 (a - 1)] (FiniteElement.jl:293-297)
  for (int64_t k = 0; k < n; ++k)
    for (int a = 0; a < 9; ++a)
      for (int i = 0; i < 3; ++i) rv[27 * k + 3 * a + i] = rel[(size_t)72 * k + 9 * i + a];
  MAF_API_END(h)
}

int maf_assemble_device(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt,
                        double bend_tm, int scatter_mode, double* d_r, double* d_nzval, double* d_rnorm2,
                        void* stream) {
  MAF_API_BEGIN(h)
  if (!d_xms || !d_cps) throw std::runtime_error("null device buffer");
  if (h->strip && (d_r || d_nzval || d_rnorm2))
    throw std::runtime_error("strip handle: results go to the handle's own slices (maf_assemble_strip)");
  do_assemble_device(h, d_xms, d_cps, time, dt, bend_tm, scatter_mode, d_r, d_nzval, d_rnorm2,
                     stream ? (cudaStream_t)stream : h->stream, false);
  MAF_API_END(h)
}

int maf_device_buffers(maf_handle* h, double** d_xms, double** d_cps, double** d_r, double** d_nzval,
                       double** d_rnorm2) {
  MAF_API_BEGIN(h)
  if (d_xms) *d_xms = h->d_xms;
  if (d_cps) *d_cps = h->d_cps;
  if (d_r) *d_r = h->d_r;
  if (d_nzval) *d_nzval = h->d_nz;
  if (d_rnorm2) *d_rnorm2 = h->d_rn;
  MAF_API_END(h)
}

int maf_download(maf_handle* h, int64_t r_first, int64_t r_count, double* r, int64_t nz_first, int64_t nz_count,
                 double* nzval) {
  MAF_API_BEGIN(h)
  const HostModel& M = h->M;
  if (r_count < 0 || nz_count < 0) throw std::runtime_error("negative count");
  if (r_count > 0 && (!r || r_first < 1 || r_first + r_count - 1 > M.nmdf)) throw std::runtime_error("row range outside 1..nmdf");
  if (nz_count > 0 && (!nzval || nz_first < 1 || nz_first + nz_count - 1 > M.sym.nnz))
    throw std::runtime_error("entry range outside 1..nnz");
  if (h->strip && ((r_count > 0 && (r_first - 1 < h->eq_lo || r_first - 1 + r_count > h->eq_hi)) ||
                   (nz_count > 0 && (nz_first - 1 < h->slot_lo || nz_first - 1 + nz_count > h->slot_hi))))
    throw std::runtime_error("a strip handle holds only the rows / entries its elements touch (maf_strip_info)");
  if (r_count > 0)
    CU(cudaMemcpyAsync(r, h->d_r + (r_first - 1), sizeof(double) * (size_t)r_count, cudaMemcpyDeviceToHost, h->stream));
  if (nz_count > 0)
    CU(cudaMemcpyAsync(nzval, h->d_nz + (nz_first - 1), sizeof(double) * (size_t)nz_count, cudaMemcpyDeviceToHost,
                       h->stream));
  CU(cudaStreamSynchronize(h->stream));
  MAF_API_END(h)
}

int maf_stream(maf_handle* h, void** stream) {
  MAF_API_BEGIN(h)
  if (!stream) throw std::runtime_error("null output pointer");
  *stream = (void*)h->stream;
  MAF_API_END(h)
}

int maf_sync(maf_handle* h) {
  MAF_API_BEGIN(h)
  CU(cudaStreamSynchronize(h->stream));
  MAF_API_END(h)
}

int maf_timings(maf_handle* h, double* out7) {
  MAF_API_BEGIN(h)
  if (!out7) throw std::runtime_error("null output pointer");
  if (h->timed_valid) {  // events of the last assembly, whichever entry point and stream launched it
    CU(cudaEventSynchronize(h->ev[4]));
    CU(cudaEventElapsedTime(&h->ms[6], h->ev[1], h->ev[6]));
    CU(cudaEventElapsedTime(&h->ms[1], h->ev[6], h->ev[2]));
    CU(cudaEventElapsedTime(&h->ms[3], h->ev[2], h->ev[3]));
    CU(cudaEventElapsedTime(&h->ms[2], h->ev[3], h->ev[4]));
  }
  for (int k = 0; k < 7; ++k) out7[k] = h->ms[k];
  MAF_API_END(h)
}

int maf_area_kernel_times(maf_handle* h, double* out_ms, int64_t n) {
  MAF_API_BEGIN(h)
  if (!out_ms || n < 0) throw std::runtime_error("null output pointer");
  if (n > MAF_RING || n > h->ring_n) throw std::runtime_error("more launches requested than the ring holds");
  for (int64_t k = 0; k < n; ++k) {   // out[0] = oldest of the last n launches
    const int q = (int)((h->ring_n - n + k) % MAF_RING);
    CU(cudaEventSynchronize(h->ring_b[q]));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ring_a[q], h->ring_b[q]));
    out_ms[k] = ms;
  }
  MAF_API_END(h)
}

int maf_launch_count(maf_handle* h, int64_t* n) {
  MAF_API_BEGIN(h)
  if (!n) throw std::runtime_error("null output pointer");
  *n = h->launches;
  MAF_API_END(h)
}

int maf_kernel_info(maf_handle* h, int64_t* out5 /* 9 values */) {
  MAF_API_BEGIN(h)
  if (!out5) throw std::runtime_error("null output pointer");
  out5[0] = MAF_NT; out5[1] = 1; out5[2] = (int64_t)h->smem_bytes; out5[3] = h->ctas_per_sm; out5[4] = h->sm_count;
  out5[5] = h->n_slot_classes;
  out5[6] = h->graph_replays;
  out5[7] = h->d_kel ? (int64_t)(h->kel_elems * ((size_t)81 * h->nij + 72) * sizeof(double)) : 0;
  out5[8] = h->band_rows;
  MAF_API_END(h)
}

int maf_chunk_plan(maf_handle* h, char* text, int64_t cap) {
  MAF_API_BEGIN(h)
  if (!text || cap < 1) throw std::runtime_error("null output pointer");
  const Config& c = h->M.cfg;
  const int nw = c.nthreads / 32;
  std::string s;
  for (int w = 0; w < nw; ++w) {
    if (w) s += '/';
    bool first = true;
    for (int r = 0; r < c.task_rounds; ++r) {
      const int id = c.chunk_slot[r * nw + w];
      if (id < 0) continue;
      if (!first) s += ',';
      s += std::to_string(id);
      first = false;
    }
  }
  if ((int64_t)s.size() + 1 > cap) throw std::runtime_error("plan text buffer too small");
  std::memcpy(text, s.c_str(), s.size() + 1);
  MAF_API_END(h)
}

// ---- strips over several GPUs ------------------------------------------------------------------------------------
static void require_strip(maf_handle* h) {
  if (!h->strip) throw std::runtime_error("not a strip handle: create it with maf_create_strip");
}

int maf_strip_info(maf_handle* h, int64_t* out12) {
  MAF_API_BEGIN(h)
  require_strip(h);
  if (!out12) throw std::runtime_error("null output pointer");
  out12[0] = h->e0 + 1; out12[1] = h->e1;
  out12[2] = h->eq_lo + 1; out12[3] = h->eq_hi;
  out12[4] = h->slot_lo + 1; out12[5] = h->slot_hi;
  out12[6] = h->own_eq_lo + 1; out12[7] = h->own_eq_hi;
  out12[8] = h->own_slot_lo + 1; out12[9] = h->own_slot_hi;
  out12[10] = h->rank; out12[11] = h->nranks;
  MAF_API_END(h)
}

int maf_peer_export(maf_handle* h, void* handle64) {
  MAF_API_BEGIN(h)
  require_strip(h);
  if (!handle64) throw std::runtime_error("null output pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t ipc;
  CU(cudaIpcGetMemHandle(&ipc, h->strip_alloc));
  std::memcpy(handle64, &ipc, 64);
  MAF_API_END(h)
}

int maf_peer_attach(maf_handle* h, const void* lower64, const void* upper64) {
  MAF_API_BEGIN(h)
  require_strip(h);
  if (h->attached) throw std::runtime_error("neighbours are already attached");
  if ((h->rank > 0) != (lower64 != nullptr) || (h->rank + 1 < h->nranks) != (upper64 != nullptr))
    throw std::runtime_error("pass the export of exactly the neighbouring strips that exist (NULL at the ends)");
  cudaIpcMemHandle_t ipc;
  if (lower64) {
    std::memcpy(&ipc, lower64, 64);
    CU(cudaIpcOpenMemHandle(&h->lower_base, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->lower_ipc = true;
  }
  if (upper64) {
    std::memcpy(&ipc, upper64, 64);
    CU(cudaIpcOpenMemHandle(&h->upper_base, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->upper_ipc = true;
  }
  h->attached = true;
  MAF_API_END(h)
}

int maf_peer_attach_local(maf_handle* h, maf_handle* lower, maf_handle* upper) {
  MAF_API_BEGIN(h)
  require_strip(h);
  if (h->attached) throw std::runtime_error("neighbours are already attached");
  if ((h->rank > 0) != (lower != nullptr) || (h->rank + 1 < h->nranks) != (upper != nullptr))
    throw std::runtime_error("pass exactly the neighbouring strip handles that exist (NULL at the ends)");
  for (maf_handle* nb : {lower, upper}) {
    if (!nb) continue;
    if (!nb->strip || nb->nranks != h->nranks || (nb != lower ? nb->rank != h->rank + 1 : nb->rank != h->rank - 1))
      throw std::runtime_error("not the neighbouring strip of the same partition");
    if (nb->device != h->device) {
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, h->device, nb->device));
      if (!can) throw std::runtime_error("the devices of neighbouring strips cannot access each other (no NVLink / P2P)");
      cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
      cudaGetLastError();
    }
  }
  h->lower_base = lower ? lower->strip_alloc : nullptr;
  h->upper_base = upper ? upper->strip_alloc : nullptr;
  h->attached = true;
  MAF_API_END(h)
}

// the strip's share of one calc_r_K, all on the handle's stream; the state is taken from d_xms / d_cps (device
// pointers, full numnp x 3 / numnp x ndf arrays) or, if NULL, from the handle's resident state
static void assemble_strip(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt,
                           double bend_tm, int mode, double* d_rn_partial) {
  if (h->nranks > 1 && !h->attached) throw std::runtime_error("attach the neighbouring strips first (maf_peer_attach)");
  if (!d_xms || !d_cps) {
    if (!h->state_resident) throw std::runtime_error("no resident state: pass device pointers or call maf_state_set");
    d_xms = h->d_xms;
    d_cps = h->d_cps;
  }
  cudaStream_t s = h->stream;
  const long long step = ++h->step;
  const long long spins = 50000000LL;   // ~10 s
  long long* err = h->flags + 2;
  // 1. our buffer may be zeroed only once the upper strip has pulled the previous step out of it
  if (h->upper_base && step > 1) {
    flag_kernel<<<1, 1, 0, s>>>(nullptr, 0, h->flags + 1, step - 1, err, spins);
    h->launches += 1;
  }
  // 2. zero-fill + kernels of our elements (sums over our own elements; partial on the two interface node rows)
  do_assemble_device(h, d_xms, d_cps, time, dt, bend_tm, mode, nullptr, nullptr, nullptr, s, false);
  CU(cudaEventRecord(h->ev_x[0], s));
  // 3. tell the upper strip that our partial sums are final; wait for the lower strip's
  long long* upper_flags = h->upper_base ? reinterpret_cast<long long*>(h->upper_base) : nullptr;
  long long* lower_flags = h->lower_base ? reinterpret_cast<long long*>(h->lower_base) : nullptr;
  if (upper_flags || lower_flags) {
    flag_kernel<<<1, 1, 0, s>>>(upper_flags ? upper_flags + 0 : nullptr, step, lower_flags ? h->flags + 0 : nullptr,
                               step, err, spins);
    h->launches += 1;
  }
  // 4. the interface belongs to us (the upper strip of the pair): add the lower strip's partial sums, read over NVLink
  if (h->lower_base) {
    const TouchedRange& L = h->lower;
    const int64_t nr = L.eq_hi - h->eq_lo, nk = L.slot_hi - h->slot_lo;
    const double* lr = reinterpret_cast<const double*>((const char*)h->lower_base + strip_r_offset()) + (h->eq_lo - L.eq_lo);
    const double* lk = reinterpret_cast<const double*>((const char*)h->lower_base + strip_nz_offset(L)) + (h->slot_lo - L.slot_lo);
    if (nr + nk > 0)
      pull_add_kernel<<<(unsigned)std::min<int64_t>((nr + nk + 255) / 256, h->sm_count * 8), 256, 0, s>>>(
          h->d_r + h->eq_lo, lr, std::max<int64_t>(nr, 0), h->d_nz + h->slot_lo, lk, std::max<int64_t>(nk, 0));
    CU(cudaGetLastError());
    h->launches += 1;
    // 5. the lower strip may reuse its buffer
    flag_kernel<<<1, 1, 0, s>>>(lower_flags + 1, step, nullptr, 0, err, spins);
    h->launches += 1;
  }
  CU(cudaEventRecord(h->ev_x[1], s));
  // 6. our share of sum(r^2): the rows we own (every row is owned by exactly one strip)
  if (d_rn_partial) {
    rnorm2_partial<<<256, 256, 0, s>>>(h->d_r + h->own_eq_lo, h->own_eq_hi - h->own_eq_lo, h->d_part);
    rnorm2_final<<<1, 256, 0, s>>>(h->d_part, 256, d_rn_partial);
    CU(cudaGetLastError());
    h->launches += 2;
  }
}

static void check_strip_error(maf_handle* h) {
  long long e = 0;
  CU(cudaMemcpyAsync(&e, h->flags + 2, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (e) throw std::runtime_error("a neighbouring strip did not arrive at step " + std::to_string(e) + " within the time-out");
}

int maf_assemble_strip(maf_handle* h, const double* d_xms, const double* d_cps, double time, double dt, double bend_tm,
                       int scatter_mode, double* d_rnorm2_partial) {
  MAF_API_BEGIN(h)
  require_strip(h);
  assemble_strip(h, d_xms, d_cps, time, dt, bend_tm, scatter_mode, d_rnorm2_partial ? d_rnorm2_partial : h->d_rn);
  MAF_API_END(h)
}

// upload of the part of the state a strip reads: the node range [node_lo, node_hi) of every column
static void upload_state_rows(maf_handle* h, const double* xms, const double* cps) {
  const HostModel& M = h->M;
  const size_t pitch = sizeof(double) * (size_t)M.numnp, width = sizeof(double) * (size_t)(h->node_hi - h->node_lo);
  CU(cudaMemcpy2DAsync(h->d_xms + h->node_lo, pitch, xms + h->node_lo, pitch, width, 3, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpy2DAsync(h->d_cps + h->node_lo, pitch, cps + h->node_lo, pitch, width, (size_t)M.ndf,
                       cudaMemcpyHostToDevice, h->stream));
  h->state_resident = true;
}

int maf_assemble_strip_host(maf_handle* h, const double* xms, const double* cps, double time, double dt, double bend_tm,
                            int scatter_mode, double* r_own, double* nzval_own, double* rnorm2_partial) {
  MAF_API_BEGIN(h)
  require_strip(h);
  if (!r_own || !nzval_own) throw std::runtime_error("null buffer");
  CU(cudaEventRecord(h->ev[0], h->stream));
  if (xms && cps) upload_state_rows(h, xms, cps);
  assemble_strip(h, nullptr, nullptr, time, dt, bend_tm, scatter_mode, h->d_rn);
  cudaStream_t s = h->stream;
  CU(cudaMemcpyAsync(r_own, h->d_r + h->own_eq_lo, sizeof(double) * (size_t)(h->own_eq_hi - h->own_eq_lo),
                     cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(nzval_own, h->d_nz + h->own_slot_lo, sizeof(double) * (size_t)(h->own_slot_hi - h->own_slot_lo),
                     cudaMemcpyDeviceToHost, s));
  if (rnorm2_partial) CU(cudaMemcpyAsync(rnorm2_partial, h->d_rn, sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(h->ev[5], s));
  check_strip_error(h);
  CU(cudaEventElapsedTime(&h->ms[5], h->ev[0], h->ev[5]));
  CU(cudaEventElapsedTime(&h->ms_exchange, h->ev_x[0], h->ev_x[1]));
  MAF_API_END(h)
}

int maf_strip_timings(maf_handle* h, double* out2) {
  MAF_API_BEGIN(h)
  require_strip(h);
  if (!out2) throw std::runtime_error("null output pointer");
  CU(cudaEventSynchronize(h->ev_x[1]));
  CU(cudaEventElapsedTime(&h->ms_exchange, h->ev_x[0], h->ev_x[1]));
  out2[0] = h->ms_exchange;
  out2[1] = (double)h->strip_bytes;
  check_strip_error(h);
  MAF_API_END(h)
}

int maf_set_element_range(maf_handle* h, int64_t el_first, int64_t el_last) {
  MAF_API_BEGIN(h)
  if (h->strip) throw std::runtime_error("a strip handle holds one fixed element range");
  if (el_first < 1 || el_last > h->M.numel || el_first > el_last + 1)
    throw std::runtime_error("element range outside 1..numel");
  h->e0 = el_first - 1;
  h->e1 = el_last;
  compute_ranges(h);
  MAF_API_END(h)
}

int maf_range_info(maf_handle* h, int64_t* out8) {
  MAF_API_BEGIN(h)
  if (!out8) throw std::runtime_error("null output pointer");
  out8[0] = h->e0 + 1; out8[1] = h->e1;              // elements, 1-based inclusive
  out8[2] = h->node_lo + 1; out8[3] = h->node_hi;    // nodes touched
  out8[4] = h->eq_lo + 1; out8[5] = h->eq_hi;        // rows of r touched
  out8[6] = h->slot_lo + 1; out8[7] = h->slot_hi;    // entries of nzval touched
  MAF_API_END(h)
}

// profiling builds (-DMAF_PHASE_TIMING): cycles per warp and phase [warp][8] summed over the CTAs since the last
// call: wait@gather, interp, wait, gauss, wait, gather-next, residual+tangent, -; then from out[32] the cycles per
// tangent chunk. Returns 1 in regular builds.
int maf_debug_phase_cycles(unsigned long long* out, int n) {
#ifdef MAF_PHASE_TIMING
  unsigned long long hbuf[MAF_NT / 32][8];
  if (cudaMemcpyFromSymbol(hbuf, g_phase_cycles, sizeof(hbuf)) != cudaSuccess) return 2;
  for (int q = 0; q < n && q < (int)(sizeof(hbuf) / 8); ++q) out[q] = (&hbuf[0][0])[q];
  std::memset(hbuf, 0, sizeof(hbuf));
  if (cudaMemcpyToSymbol(g_phase_cycles, hbuf, sizeof(hbuf)) != cudaSuccess) return 2;
  unsigned long long cbuf[MAF_MAX_CHUNKS];   // out[32 ...]: cycles per tangent chunk
  if (cudaMemcpyFromSymbol(cbuf, maf::g_chunk_cycles, sizeof(cbuf)) != cudaSuccess) return 2;
  for (int q = 0; q < MAF_MAX_CHUNKS && 32 + q < n; ++q) out[32 + q] = cbuf[q];
  std::memset(cbuf, 0, sizeof(cbuf));
  if (cudaMemcpyToSymbol(maf::g_chunk_cycles, cbuf, sizeof(cbuf)) != cudaSuccess) return 2;
  return 0;
#else
  (void)out; (void)n;
  return 1;
#endif
}

int maf_fp64_peak(int device, double* tflops) {
  try {
    if (!tflops) return 1;
    int dev = device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    double* d = nullptr;
    CU(cudaMalloc(&d, 8));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    const int iters = 1 << 14, grid = prop.multiProcessorCount * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
      CU(cudaEventRecord(a, 0));
      dfma_peak_kernel<<<grid, 256>>>(d, iters, 0.999999, 1e-9);
      CU(cudaEventRecord(b, 0));
      CU(cudaEventSynchronize(b));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, a, b));
      const double tf = 2.0 * 8.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
      if (rep > 0) best = std::max(best, tf);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return 0;
  } catch (std::exception& e) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_create_err = e.what();
    return 2;
  }
}

int maf_host_register(void* ptr, int64_t bytes) {
  try {
    if (!ptr || bytes <= 0) throw std::runtime_error("null buffer");
    CU(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
  } catch (std::exception& e) {
    cudaGetLastError();   // a refused registration is not sticky: do not leave it for the next launch check
    std::lock_guard<std::mutex> lk(g_mu);
    g_create_err = e.what();
    return 2;
  }
}

int maf_host_unregister(void* ptr) {
  try {
    if (!ptr) throw std::runtime_error("null buffer");
    CU(cudaHostUnregister(ptr));
    return 0;
  } catch (std::exception& e) {
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_mu);
    g_create_err = e.what();
    return 2;
  }
}

}  // extern "C"
