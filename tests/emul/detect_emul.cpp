// TEST INFRASTRUCTURE — host emulation of the detector kernels of ground-fusion2_b200/csrc/gf2_tracker_detect.cuh, one CUDA thread at a time, so that
// the CPU-only suite can check their float32 / float64 operation order against the cv2-pinned numpy oracle without a GPU (tests/test_detect_emul.py).
// It is compiled by the test with g++ -ffp-contract=off; it is not part of the product and nothing in the package references it. Only kernels whose
// threads do not communicate are launched (k_gftt_eig, k_gftt_nms); the warp reduction of the masked maximum is redone on the host.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
#define __shared__ static
struct D3 { int x, y, z; };
static D3 threadIdx, blockIdx, blockDim;
static inline void __syncthreads() {}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int) { return v; }
static inline int __shfl_up_sync(unsigned, int v, int) { return v; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; *p = std::max(o, v); return o; }
static inline int atomicAdd(int32_t* p, int v) { int o = *p; *p += v; return o; }
using std::max; using std::min;
static inline int reflect101(int i, int n) { if (n == 1) return 0; const int p = 2 * (n - 1); i %= p; if (i < 0) i += p; return i >= n ? p - i : i; }
#include "../../ground-fusion2_b200/csrc/gf2_tracker_detect.cuh"
template <class F> static void launch(D3 g, D3 b, F f) {
  blockDim = b;
  for (int bz = 0; bz < g.z; bz++) for (int by = 0; by < g.y; by++) for (int bx = 0; bx < g.x; bx++)
    for (int tz = 0; tz < b.z; tz++) for (int ty = 0; ty < b.y; ty++) for (int tx = 0; tx < b.x; tx++) { blockIdx = {bx, by, bz}; threadIdx = {tx, ty, tz}; f(); }
}
extern "C" int detect_emul(const uint8_t* img, const uint8_t* mask, int W, int H, int S, const int32_t* want, double quality, float* eig, unsigned long long* keys, int cap, int32_t* count) {
  std::vector<unsigned> vmax(S, 0);
  launch({(W + 63) / 64, S, 1}, {64, 1, 1}, [&] { k_gftt_eig(img, (size_t)W * H, mask, W, H, eig, vmax.data()); });
  for (int s = 0; s < S; s++) { vmax[s] = 0; for (size_t i = 0; i < (size_t)W * H; i++) if (!mask || mask[(size_t)s * W * H + i]) vmax[s] = std::max(vmax[s], gftt_ordered(eig[(size_t)s * W * H + i])); }
  memset(count, 0, sizeof(int32_t) * S);
  launch({(W + 31) / 32, (H + 7) / 8, S}, {32, 8, 1}, [&] { k_gftt_nms(eig, mask, W, H, vmax.data(), quality, want, keys, cap, count); });
  return 0;
}
