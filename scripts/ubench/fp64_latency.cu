// fp64 latency on sm_100a: dependent-issue latency of DFMA and of mma.sync.m8n8k4.f64 (one warp per SM), and how many independent
// chains per warp / warps per SM sub-partition it takes to reach the datapath's throughput. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o fp64_latency fp64_latency.cu ; run on one GPU. Output feeds DESIGN.md section 6 (why k_linearize is latency bound).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_f64(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k_dfma(double* out, long long* clk, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_dmma(double* out, long long* clk, int iters, double a, double b) {
  double c[ILP][2];
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) mma_f64(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <typename K>
static void run(const char* name, K kern, int ilp, int warps, double* out, long long* clk, int nsm) {
  const int iters = 2000;
  kern<<<nsm, 32 * warps>>>(out, clk, iters, 0.999999, 1e-9);
  kern<<<nsm, 32 * warps>>>(out, clk, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_round = (double)h / (iters * 8.0);   // cycles per round of ILP independent instructions per warp
  printf("%s ILP %2d, %2d warps/SM: %7.2f clk per dependent step, %6.3f warp-instr/clk/SM\n", name, ilp, warps, per_round, ilp * warps / per_round);
}
int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out; long long* clk; cudaMalloc(&out, sizeof(double) * nsm * 1024); cudaMalloc(&clk, sizeof(long long) * nsm);
  const int ws[] = {1, 4, 8, 16, 32};
  for (int w : ws) { run("DFMA", k_dfma<1>, 1, w, out, clk, nsm); run("DFMA", k_dfma<2>, 2, w, out, clk, nsm); run("DFMA", k_dfma<4>, 4, w, out, clk, nsm); run("DFMA", k_dfma<8>, 8, w, out, clk, nsm); }
  for (int w : ws) { run("DMMA", k_dmma<1>, 1, w, out, clk, nsm); run("DMMA", k_dmma<2>, 2, w, out, clk, nsm); run("DMMA", k_dmma<4>, 4, w, out, clk, nsm); run("DMMA", k_dmma<12>, 12, w, out, clk, nsm); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
