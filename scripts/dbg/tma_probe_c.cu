// 2-D variant through libcu++ (cuda::device::experimental::cp_async_bulk_tensor_2d_global_to_shared)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, uint32_t* out) {
  __shared__ alignas(128) unsigned char tile[32 * 32];
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) { cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &map, x, y, bar); token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile)); }
  else token = bar.arrive();
  bar.wait(std::move(token));
  if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < 1024; i++) s += tile[i]; out[0] = s; out[1] = tile[0]; out[2] = tile[33]; }
}
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)fn;
  const int W = 640, H = 480;
  unsigned char* img; cudaMalloc(&img, W * H); unsigned char* h = new unsigned char[W * H]; for (int i = 0; i < W * H; i++) h[i] = (i % W + i / W) & 0xff; cudaMemcpy(img, h, W * H, cudaMemcpyHostToDevice);
  CUtensorMap m8;
  const cuuint64_t dims[2] = {W, H}; const cuuint32_t es[2] = {1, 1}; const cuuint64_t st[1] = {W}; const cuuint32_t box[2] = {32, 32};
  printf("encode u8 2d: %d\n", (int)enc(&m8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  uint32_t* out; cudaMalloc(&out, 16); uint32_t ho[3];
  const int cs[][2] = {{100, 100}, {0, 0}, {-3, 50}, {50, -7}, {630, 470}, {7, 13}};
  for (auto& c : cs) {
    k<<<1, 32>>>(m8, c[0], c[1], out); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(ho, out, 12, cudaMemcpyDeviceToHost);
    printf("u8 2d box at (%4d,%4d): %s sum %u first %u (1,1) %u\n", c[0], c[1], cudaGetErrorString(e), ho[0], ho[1], ho[2]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
