// smem_probe.cu -- what a shared-memory load costs on sm_100a as a function of width and lane pattern, and how it
// overlaps with the FP64 pipe. Decides the register tiling of the area kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/smem_probe tools/microbench/smem_probe.cu
// Prints, per test, cycles per warp-level instruction SM-wide (1 CTA per SM, W warps).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } \
  } while (0)

__constant__ double c_tab[64];

__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds128(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ float lds32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

// lane -> offset in units of the access width
__device__ int pattern_off(int pat, int lane) {
  switch (pat) {
    case 0: return 0;                                   // uniform
    case 1: return lane % 9;                            // 9 distinct, contiguous
    case 2: return lane;                                // 32 distinct, contiguous
    case 3: return 4 * ((lane % 9) / 3) + (lane % 9) % 3;   // phi_a padding
    case 4: return (lane / 9) * 84;                     // 4 distinct rows
    case 5: return 2 * lane;                            // stride 2
    case 6: return (lane % 9) + 40 * (lane / 9);        // 9 x 4 distinct
    case 7: return lane % 4;                            // 4 distinct contiguous
    case 8: return lane % 2;                            // 2 distinct
    case 9: return (lane / 8) * 37;                     // quarter-warps uniform, rows apart
    case 10: return (lane / 16) * 37;                   // half-warps uniform
    case 11: return (lane % 16);                        // 16 distinct contiguous, both halves the same
    case 12: return (lane % 8);                         // 8 distinct contiguous, all quarters the same
    default: return lane % 3;
  }
}

template <int WIDTH>
__global__ void lds_kernel(int pat, int iters, long long* cyc, double* sink) {
  extern __shared__ double sm[];
  for (int k = threadIdx.x; k < 8192; k += blockDim.x) sm[k] = k;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(pattern_off(pat, lane) * (WIDTH / 8));
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
    const unsigned b = base + ((k & 15) << 8);   // 256-byte steps, stays inside the buffer
    if (WIDTH == 64) {
      a0 += lds64(b); a1 += lds64(b + 4096); a2 += lds64(b + 8192); a3 += lds64(b + 12288);
    } else if (WIDTH == 128) {
      double x, y;
      lds128(b, x, y); a0 += x; a1 += y;
      lds128(b + 4096, x, y); a2 += x; a3 += y;
      lds128(b + 8192, x, y); a0 += x; a1 += y;
      lds128(b + 12288, x, y); a2 += x; a3 += y;
    } else {
      a0 += lds32(b); a1 += lds32(b + 4096); a2 += lds32(b + 8192); a3 += lds32(b + 12288);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (a0 + a1 + a2 + a3 == 1.2345e-300) sink[0] = a0;
}

// FP64: 8 independent chains; MODE 0 register operands, 1 one operand from the constant bank,
// 2 interleaved with uniform LDS.64 at `ratio` DFMA per LDS, 3 with lane-distinct LDS.64, 4 with uniform LDS.128
template <int MODE, int RATIO>
__global__ void dfma_kernel(int iters, long long* cyc, double* sink, double s, double t) {
  extern __shared__ double sm[];
  for (int k = threadIdx.x; k < 8192; k += blockDim.x) sm[k] = 1e-9 * k;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (MODE == 3 ? 8u * lane : 0u);
  double x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = threadIdx.x + q;
  double l0 = 0, l1 = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
    const unsigned b = base + ((k & 15) << 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) {   // 64 DFMA per trip
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (MODE == 1) x[q] = fma(x[q], c_tab[8 * u + q], t);
        else if (MODE == 5) x[q] = fma(x[q], c_tab[8 * ((k + u) & 7) + q], t);
        else x[q] = fma(x[q], s, t);
      }
      if (MODE >= 2) {
        // RATIO DFMA per load: 8 DFMA in this inner group -> 8 / RATIO loads
#pragma unroll
        for (int v = 0; v < 8 / RATIO; ++v) {
          if (MODE == 4) { double p, q2; lds128(b + 512 * u + 64 * v, p, q2); l0 += p; l1 += q2; }
          else l0 += lds64(b + 512 * u + 64 * v);
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double acc = l0 + l1;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc += x[q];
  if (acc == 1.2345e-300) sink[0] = acc;
}

static double run_avg(long long* d_cyc, int n) {
  static long long h[1024];
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  double s = 0;
  for (int k = 0; k < n; ++k) s += (double)h[k];
  return s / n;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int nsm = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, nsm);
  long long* d_cyc;
  double* d_sink;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
  CK(cudaMalloc(&d_sink, 8));
  double tab[64];
  for (int k = 0; k < 64; ++k) tab[k] = 0.999 + 1e-6 * k;
  CK(cudaMemcpyToSymbol(c_tab, tab, sizeof(tab)));
  const int iters = 2048, smem = 65536 + 16384;
  CK(cudaFuncSetAttribute(lds_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(lds_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(lds_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* pname[] = {"uniform", "9 contiguous", "32 contiguous", "9 padded(phi_a)", "4 rows x84", "stride 2",
                         "9x4 distinct", "4 contiguous", "2 contiguous", "quarter-uniform", "half-uniform",
                         "16 contig (halves same)", "8 contig (quarters same)", "3 contiguous"};
  for (int W : {4, 8, 16}) {
    printf("--- LDS, %d warps per SM: cycles per warp-level load, SM-wide ---\n", W);
    for (int pat = 0; pat < 14; ++pat) {
      double c[3];
      for (int wi = 0; wi < 3; ++wi) {
        for (int rep = 0; rep < 2; ++rep) {
          if (wi == 0) lds_kernel<32><<<nsm, 32 * W, smem>>>(pat, iters, d_cyc, d_sink);
          if (wi == 1) lds_kernel<64><<<nsm, 32 * W, smem>>>(pat, iters, d_cyc, d_sink);
          if (wi == 2) lds_kernel<128><<<nsm, 32 * W, smem>>>(pat, iters, d_cyc, d_sink);
          CK(cudaDeviceSynchronize());
        }
        c[wi] = run_avg(d_cyc, nsm) / (4.0 * iters * W);
      }
      printf("  %-26s  LDS.32 %6.2f   LDS.64 %6.2f   LDS.128 %6.2f\n", pname[pat], c[0], c[1], c[2]);
    }
  }
  const int fsm = 65536 + 16384;
#define RUN_DFMA(MODE, RATIO, label)                                                                         \
  {                                                                                                          \
    CK(cudaFuncSetAttribute(dfma_kernel<MODE, RATIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm));    \
    for (int W : {4, 8, 12, 16}) {                                                                           \
      for (int rep = 0; rep < 2; ++rep) {                                                                    \
        dfma_kernel<MODE, RATIO><<<nsm, 32 * W, fsm>>>(512, d_cyc, d_sink, 0.999999, 1e-9);                  \
        CK(cudaDeviceSynchronize());                                                                         \
      }                                                                                                      \
      printf("  %-44s W=%2d  %6.3f cycles per warp-DFMA SM-wide\n", label, W, run_avg(d_cyc, nsm) / (64.0 * 512 * W)); \
    }                                                                                                        \
  }
  printf("--- DFMA (ideal 0.5 cycles per warp instruction SM-wide) ---\n");
  RUN_DFMA(0, 1, "registers only")
  RUN_DFMA(1, 1, "one operand from the constant bank")
  RUN_DFMA(5, 1, "constant bank, index varies per trip (LDCU in loop)")
  RUN_DFMA(2, 8, "+ uniform LDS.64, 8 DFMA per load")
  RUN_DFMA(2, 4, "+ uniform LDS.64, 4 DFMA per load")
  RUN_DFMA(2, 2, "+ uniform LDS.64, 2 DFMA per load")
  RUN_DFMA(2, 1, "+ uniform LDS.64, 1 DFMA per load")
  RUN_DFMA(3, 8, "+ lane-distinct LDS.64, 8 DFMA per load")
  RUN_DFMA(3, 4, "+ lane-distinct LDS.64, 4 DFMA per load")
  RUN_DFMA(3, 2, "+ lane-distinct LDS.64, 2 DFMA per load")
  RUN_DFMA(4, 8, "+ uniform LDS.128, 8 DFMA per load")
  RUN_DFMA(4, 4, "+ uniform LDS.128, 4 DFMA per load")
  RUN_DFMA(4, 2, "+ uniform LDS.128, 2 DFMA per load")
  return 0;
}
