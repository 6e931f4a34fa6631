// Microbenchmark: fp64 FMA-pipe vs fp64 tensor (mma.sync m8n8k4) throughput on this GPU, plus shared-memory atomicAdd(double).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
  double a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
  double c[8][2]; for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Do the fp64 FMA pipe and the fp64 tensor path share one datapath? Even warps issue DFMA, odd warps DMMA, in the same CTA.
// If the two were independent pipes the combined rate would approach the sum of the two single rates.
__global__ void k_mixed(double* out, int iters) {
  const int wid = threadIdx.x >> 5;
  double s = 0;
  if (wid & 1) {
    double c[8][2]; for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  } else {
    double a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    // 16 DFMA warp-instructions carry the same 256 MACs per lane-row as ONE m8n8k4: 8 x 16 FMAs per iteration keeps the two halves in step
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 16; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
    }
    for (int i = 0; i < 8; i++) s += a[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_atoms(double* out, int iters, int stride) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++) atomicAdd(&sm[(lane * stride + it * 7) & 2047], 1.0);
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32, blocks = 148 * 4;
    float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("DFMA  %2d warps/blk x4 blk/SM: %.2f TFLOP/s\n", warps, 2.0 * 8 * iters * (double)threads * blocks / ms / 1e9);
    ms = timeit([&] { k_dmma<<<blocks, threads>>>(out, iters); });
    printf("DMMA  %2d warps/blk x4 blk/SM: %.2f TFLOP/s  (%.2f clk/mma/SM @1.9GHz)\n", warps, 2.0 * 256 * 8 * iters * (double)warps * blocks / ms / 1e9,
           ms * 1e-3 * 1.9e9 / (8.0 * iters * warps * 4));
  }
  for (int warps : {8, 16}) {
    const int threads = warps * 32, blocks = 148 * 4, it2 = 4000;
    float ms = timeit([&] { k_mixed<<<blocks, threads>>>(out, it2); });
    const double fl_mma = 2.0 * 256 * 8 * it2 * (double)(warps / 2) * blocks, fl_fma = 2.0 * 8 * 16 * it2 * (double)(threads / 2) * blocks;
    printf("MIXED %2d warps/blk (half DFMA, half DMMA, equal flops): DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", warps, fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
  }
  for (int stride : {1, 0, 33}) {
    float ms = timeit([&] { k_atoms<<<148 * 2, 256>>>(out, 20000, stride); });
    printf("shared atomicAdd(double) stride %d: %.1f clk per warp-instr per SM\n", stride, ms * 1e-3 * 1.9e9 / (20000.0 * 8 * 2));
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
