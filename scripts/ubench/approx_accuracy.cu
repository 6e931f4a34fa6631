// Accuracy of rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64 followed by n Newton steps, against correctly rounded division / sqrt.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double rcp_n(double x, int n) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); for (int i = 0; i < n; i++) { const double e = fma(-x, r, 1.0); r = fma(r, e, r); } return r; }
__device__ double rsq_n(double x, int n) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); for (int i = 0; i < n; i++) { const double e = fma(-x * r, r, 1.0); r = fma(0.5 * r, e, r); } return r; }
__global__ void k(double* out) {
  __shared__ double red[8][256];
  unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  double e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < 4096; it++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
    const double x = exp2(u * 80.0 - 40.0) * (1.0 + u);
    for (int n = 0; n < 4; n++) {
      e[n] = fmax(e[n], fabs(rcp_n(x, n) * x - 1.0));
      e[4 + n] = fmax(e[4 + n], fabs(rsq_n(x, n) * sqrt(x) - 1.0));
    }
  }
  for (int q = 0; q < 8; q++) red[q][threadIdx.x] = e[q];
  __syncthreads();
  if (threadIdx.x < 8) { double m = 0; for (int i = 0; i < 256; i++) m = fmax(m, red[threadIdx.x][i]); out[blockIdx.x * 8 + threadIdx.x] = m; }
}
int main() {
  double* d; cudaMalloc(&d, 64 * 8 * sizeof(double)); k<<<64, 256>>>(d); double h[64 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int q = 0; q < 8; q++) { double m = 0; for (int b = 0; b < 64; b++) m = fmax(m, h[b * 8 + q]); printf("%s.approx.ftz.f64 + %d Newton steps: max relative error %.3e\n", q < 4 ? "rcp" : "rsqrt", q & 3, m); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
