// Probe: which 3-D tiled tensor-map loads does this part accept? (coordinates inside / negative / past the edge; u8 32x32 and u32 24x22 boxes)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const CUtensorMap* map, int x, int y, int z, int bytes, int elems_w, int esz, uint32_t* out) {
  __shared__ __align__(128) unsigned char tile[4096];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  __syncthreads();
  asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
  if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < bytes; i++) s += tile[i]; out[0] = s; out[1] = tile[0]; out[2] = tile[(elems_w * 1 + 1) * esz]; }
}
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)fn;
  const int W = 640, H = 480;
  unsigned char* img; cudaMalloc(&img, W * H); unsigned char* h = new unsigned char[W * H]; for (int i = 0; i < W * H; i++) h[i] = (i % W + i / W) & 0xff; cudaMemcpy(img, h, W * H, cudaMemcpyHostToDevice);
  uint32_t* der; cudaMalloc(&der, W * H * 4); cudaMemset(der, 1, W * H * 4);
  CUtensorMap m8, m32, *dm; cudaMalloc(&dm, 256);
  const cuuint64_t dims[3] = {W, H, 1}; const cuuint32_t es[3] = {1, 1, 1};
  { const cuuint64_t st[2] = {W, (cuuint64_t)W * H}; const cuuint32_t box[3] = {32, 32, 1};
    printf("encode u8: %d\n", (int)enc(&m8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, img, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  { const cuuint64_t st[2] = {W * 4, (cuuint64_t)W * H * 4}; const cuuint32_t box[3] = {24, 22, 1};
    printf("encode u32: %d\n", (int)enc(&m32, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, der, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  cudaMemcpy(dm, &m8, 128, cudaMemcpyHostToDevice); cudaMemcpy(dm + 1, &m32, 128, cudaMemcpyHostToDevice);
  uint32_t* out; cudaMalloc(&out, 16); uint32_t ho[3];
  const int cs[][3] = {{100, 100, 0}, {0, 0, 0}, {-3, 50, 0}, {50, -7, 0}, {630, 470, 0}, {-40, -40, 0}, {7, 13, 0}};
  for (auto& c : cs) {
    k<<<1, 32>>>(dm, c[0], c[1], c[2], 1024, 32, 1, out); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(ho, out, 12, cudaMemcpyDeviceToHost);
    printf("u8  box at (%4d,%4d): %s sum %u first %u (1,1) %u\n", c[0], c[1], cudaGetErrorString(e), ho[0], ho[1], ho[2]);
    if (e != cudaSuccess) return 1;
    k<<<1, 32>>>(dm + 1, c[0], c[1], c[2], 2112, 24, 4, out); e = cudaDeviceSynchronize(); cudaMemcpy(ho, out, 12, cudaMemcpyDeviceToHost);
    printf("u32 box at (%4d,%4d): %s sum %u\n", c[0], c[1], cudaGetErrorString(e), ho[0]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
