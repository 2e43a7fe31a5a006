// dmma_probe.cu -- FP64 tensor-core (DMMA, mma.sync m8n8k4 / m16n8k8 f64) throughput against the DFMA rate on sm_100a:
// whole device, NW warps per SM sub-partition, ILP independent accumulator tiles per warp. Answers "would the
// contractions run faster as small GEMMs on the tensor cores?" (DESIGN.md section 4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/dmma_probe tools/microbench/dmma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// MODE 0: DFMA   MODE 1: mma.m8n8k4.f64   MODE 2: mma.m16n8k8.f64
template <int MODE, int ILP>
__global__ void probe(int iters, double* sink, double s, double t) {
  double c[ILP][4];
#pragma unroll
  for (int q = 0; q < ILP; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) c[q][r] = 1e-3 * (threadIdx.x + q + r);
  const double a0 = s, a1 = s * 0.5, a2 = s * 0.25, a3 = s * 0.125, b0 = t, b1 = t * 0.5;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int q = 0; q < ILP; ++q) {
        if (MODE == 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r) c[q][r] = fma(c[q][r], a0, b0);
        } else if (MODE == 1) {
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a0), "d"(b0));
        } else {
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3])
                       : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        }
      }
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < ILP; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc += c[q][r];
  if (acc == 1.2345e-300) sink[0] = acc;
}

template <int MODE, int ILP>
static void run(const char* name, int sms, double fma_per_warp_instr, int instr_per_inner) {
  double* sink;
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 4096, threads = 256, blocks = sms * 4;   // 32 warps per SM
  probe<MODE, ILP><<<blocks, threads>>>(64, sink, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<MODE, ILP><<<blocks, threads>>>(iters, sink, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double warps = (double)blocks * threads / 32;
  const double winstr = warps * iters * 8.0 * ILP * instr_per_inner;
  const double fma = winstr * fma_per_warp_instr;
  printf("%-28s ILP %d: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM (at 1.965 GHz)  %.3f warp-instr/clk/SM\n", name, ILP, ms,
         2.0 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / 1.965e9 / sms, winstr / (ms * 1e-3) / 1.965e9 / sms);
  CK(cudaFree(sink));
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  run<0, 2>("DFMA", sms, 32.0, 4);
  run<0, 4>("DFMA", sms, 32.0, 4);
  run<1, 2>("DMMA m8n8k4", sms, 256.0, 1);
  run<1, 4>("DMMA m8n8k4", sms, 256.0, 1);
  run<1, 8>("DMMA m8n8k4", sms, 256.0, 1);
  run<2, 2>("DMMA m16n8k8", sms, 1024.0, 1);
  run<2, 4>("DMMA m16n8k8", sms, 1024.0, 1);
  return 0;
}
