// lat_probe.cu -- dependent-issue latencies of the FP64 pipe and of shared memory on sm_100a (one warp, one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/lat_probe tools/microbench/lat_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE, int ILP>
__global__ void lat_kernel(int iters, long long* cyc, double* sink, double s, double t) {
  __shared__ double sm[1024];
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) sm[k] = (double)((k * 8 + 8) % 8192);   // pointer chase table (byte offsets)
  __syncthreads();
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = 1.0 + 1e-3 * (threadIdx.x + q);
  unsigned p = 8 * (threadIdx.x & 31);
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  const long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int q = 0; q < ILP; ++q) {
        if (MODE == 0) x[q] = fma(x[q], s, t);
        else if (MODE == 1) x[q] = x[q] * s;
        else if (MODE == 2) x[q] = x[q] + t;
        else if (MODE == 3) x[q] = 1.0 / x[q] + t;
        else if (MODE == 4) x[q] = sqrt(x[q]) + t;
        else if (MODE == 5) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + (unsigned)__double2int_rn(x[q]))); x[q] = v; }
        else if (MODE == 6) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + p)); p = v & 0x1ff8; }
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double acc = p;
#pragma unroll
  for (int q = 0; q < ILP; ++q) acc += x[q];
  if (acc == 1.2345e-300) sink[0] = acc;
}

int main() {
  long long* d_cyc; double* d_sink;
  CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_sink, 8));
  const int iters = 256;
#define RUN(MODE, ILP, W, label) { lat_kernel<MODE, ILP><<<1, 32 * W>>>(iters, d_cyc, d_sink, 0.9999999, 1e-9); CK(cudaDeviceSynchronize()); \
    lat_kernel<MODE, ILP><<<1, 32 * W>>>(iters, d_cyc, d_sink, 0.9999999, 1e-9); CK(cudaDeviceSynchronize()); long long c; \
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost)); \
    printf("%-34s ILP %d  warps %2d : %7.2f cycles per dependent step (%.2f per op)\n", label, ILP, W, (double)c / (16.0 * iters), (double)c / (16.0 * iters * ILP)); }
  RUN(0, 1, 1, "DFMA") RUN(0, 2, 1, "DFMA") RUN(0, 4, 1, "DFMA") RUN(0, 8, 1, "DFMA")
  RUN(0, 1, 4, "DFMA") RUN(0, 1, 12, "DFMA") RUN(0, 2, 12, "DFMA") RUN(0, 1, 16, "DFMA") RUN(0, 1, 32, "DFMA")
  RUN(1, 1, 1, "DMUL") RUN(2, 1, 1, "DADD")
  RUN(3, 1, 1, "1/x + t (division)") RUN(3, 2, 1, "1/x + t (division)") RUN(3, 1, 12, "1/x + t (division)")
  RUN(4, 1, 1, "sqrt(x) + t") RUN(4, 1, 12, "sqrt(x) + t")
  RUN(6, 1, 1, "LDS.32 pointer chase")
  return 0;
}
