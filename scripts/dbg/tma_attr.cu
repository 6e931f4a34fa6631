#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
int main() {
  int v = -1; cudaError_t e = cudaDeviceGetAttribute(&v, (cudaDeviceAttr)127 /* cudaDevAttrTensorMapAccessSupported */, 0);
  printf("cudaDevAttrTensorMapAccessSupported: %d (%s)\n", v, cudaGetErrorString(e));
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); printf("%s cc %d.%d driver", p.name, p.major, p.minor);
  int dv = 0, rv = 0; cudaDriverGetVersion(&dv); cudaRuntimeGetVersion(&rv); printf(" %d runtime %d\n", dv, rv);
  int mig = -1; cudaDeviceGetAttribute(&mig, (cudaDeviceAttr)129, 0); printf("attr129 %d\n", mig);
  return 0;
}
