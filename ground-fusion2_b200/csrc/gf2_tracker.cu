// gf2_tracker.cu — pyramidal Lucas-Kanade on sm_100a: the cv::calcOpticalFlowPyrLK calls of FeatureTracker::trackImage
// (VE/featureTracker/feature_tracker.cpp:122,132,135,141) for a batch of independent image streams.
//
//   k_pyrdown   cv::pyrDown for 8-bit images (5x5 [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8)
//   k_scharr    un-normalised 3x3 Scharr derivatives of the template pyramid, int16 (dx, dy) interleaved
//   k_lk        one warp per point: all pyramid levels coarse -> fine in one launch. Per level the warp stages three tiles in shared memory
//               with bulk asynchronous copies (cp.async.bulk global -> shared, completion on the warp's mbarrier; SASS UBLKCP; every lane
//               issues the 16-byte aligned row it owns): the 22-row footprint of the template in I, its Scharr derivatives, and a
//               32-row search tile of J around the starting position (the window may move 5 px in every direction before the
//               iteration leaves the tile). The template (I, Ix, Iy) then lives in registers (14 pixels per lane) and every iteration
//               samples J out of the shared tile. Windows that overhang the image border (BORDER_REFLECT_101 / zero derivatives),
//               iterations that leave the tile, and pyramid levels whose row pitch is not a multiple of 16 bytes take the direct global
//               gather of the same arithmetic. (2-D tensor-map copies, cp.async.bulk.tensor, raise "illegal instruction" on this pool's
//               boxes even from the CUDA programming guide's own sample — scripts/dbg/tma_probe*.cu — so the tiles are moved row by row.)
//               Window sums are exact int64 warp reductions, the 2x2 solve is float32 with the operation order of OpenCV's
//               lkpyramid.cpp (no FMA contraction)
//   k_gftt_*    cv::goodFeaturesToTrack (feature_tracker.cpp:198): Sobel -> covariance planes -> 3x3 box sums with cv's running
//               double column sum -> min eigenvalue (bit-exact with cv2.cornerMinEigenVal), masked maximum, threshold + 3x3
//               non-maximum suppression -> candidate keys; the greedy min-distance selection runs on the host inside the
//               library because the feature-id order depends on it (integer-exact)
// The integer patch extraction (14-bit fixed-point bilinear weights, CV_DESCALE) is bit-exact with OpenCV; only the window
// sums differ (OpenCV accumulates them in float32 in SIMD-path order), i.e. by float32 rounding of A and b.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <math.h>
#include "gf2_common.h"

namespace gf2 {

constexpr int kMaxLevels = 4;  // levels 0..3
constexpr int kWBits = 14;

struct Pyr {  // one image pyramid on the device
  uint8_t* img[kMaxLevels];
  short2* der[kMaxLevels];
  int w[kMaxLevels], h[kMaxLevels];
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  i %= p; if (i < 0) i += p;
  return i >= n ? p - i : i;
}

__global__ void k_pyrdown(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride, uint8_t* __restrict__ dst, int dw, int dh, size_t dstride_img,
                          size_t src_img_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, s = blockIdx.z;
  if (x >= dw || y >= dh) return;
  const uint8_t* S = src + (size_t)s * src_img_stride;
  const int k[5] = {1, 4, 6, 4, 1};
  int acc = 0;
#pragma unroll
  for (int dy = -2; dy <= 2; dy++) {
    const uint8_t* row = S + (size_t)reflect101(2 * y + dy, sh) * sstride;
    int r = 0;
#pragma unroll
    for (int dx = -2; dx <= 2; dx++) r += k[dx + 2] * row[reflect101(2 * x + dx, sw)];
    acc += k[dy + 2] * r;
  }
  dst[(size_t)s * dstride_img + (size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
}

__global__ void k_scharr(const uint8_t* __restrict__ src, int w, int h, size_t stride, size_t img_stride, short2* __restrict__ dst, size_t dst_img_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, s = blockIdx.z;
  if (x >= w || y >= h) return;
  const uint8_t* S = src + (size_t)s * img_stride;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w), ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
  const uint8_t *r0 = S + (size_t)ym * stride, *r1 = S + (size_t)y * stride, *r2 = S + (size_t)yp * stride;
  const int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10, t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
  const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
  dst[(size_t)s * dst_img_stride + (size_t)y * w + x] = make_short2((short)(t0p - t0m), (short)((t1m + t1p) * 3 + t1c * 10));
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void lk_weights(float px, float py, int ipx, int ipy, int& w00, int& w01, int& w10, int& w11) {
  const float a = __fsub_rn(px, (float)ipx), b = __fsub_rn(py, (float)ipy);
  const float s = (float)(1 << kWBits);
  w00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), __fsub_rn(1.f, b)), s));
  w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.f, b)), s));
  w10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), b), s));
  w11 = (1 << kWBits) - w00 - w01 - w10;
}

struct LkArgs {
  int n_streams, max_pts, win, max_level, max_iters, flags;
  float min_eig; double eps2;
  const int32_t* n_pts;       // [n_streams]
  const int32_t* gate;        // [n_streams] or null: streams with gate == 0 are left untouched (prediction fall-back pass)
  const float* prev_pts;      // [n_streams][max_pts][2]
  float* next_pts;            // in (initial flow) / out
  uint8_t* status; float* err;
  size_t img_stride[kMaxLevels], der_stride[kMaxLevels];
  Pyr I, J;                   // template (prev) and search (next) pyramids, stream 0 base pointers
};

// Shared memory of a warp: the search tile of J (rows copied from the 16-byte aligned address at or below the wanted first pixel, so a row
// holds up to 15 leading bytes of slack) and the template record of every window pixel.
constexpr int kTileMargin = 5;                   // the search tile of J covers the 22 x 22 footprint moved by up to 5 px each way
constexpr int kTileRows = 32, kTilePitch = 48;   // 32 = 21 + 1 + 2 * 5 pixels wanted per row, + 15 of slack, rounded to 48 bytes
constexpr int kLkWarps = 4;
constexpr int kPix = 14;                         // ceil(441 / 32) window pixels per lane
struct __align__(128) LkWarpShared {
  uint8_t J[kTileRows * kTilePitch];
  short4 tmpl[32 * kPix];                        // per window pixel: I (x32), Ix, Iy, packed offset (oy << 8 | ox)
};

__device__ __forceinline__ uint32_t lk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lk_bulk_row(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lk_smem_u32(dst)), "l"(src), "r"(bytes), "r"(lk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lk_mbar_expect(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lk_mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(lk_smem_u32(bar)), "r"(parity) : "memory");
}
// Fixed-point bilinear sample (CV_DESCALE(.., W_BITS1 - 5)) of a u8 image: interior (four neighbouring bytes at `q`, row pitch `pitch`) and
// the BORDER_REFLECT_101 version for windows that overhang the image (rare: kept out of line so that the hot loops stay small — the fully
// inlined kernel was 11 k instructions and spent most of its time waiting for instruction fetches).
__device__ __forceinline__ int lk_sample4(const uint8_t* __restrict__ q, int pitch, int w00, int w01, int w10, int w11) {
  return ((int)q[0] * w00 + (int)q[1] * w01 + (int)q[pitch] * w10 + (int)q[pitch + 1] * w11 + (1 << (kWBits - 5 - 1))) >> (kWBits - 5);
}
__device__ __noinline__ int lk_sample_border(const uint8_t* __restrict__ im, int cols, int rows, int y, int x, int w00, int w01, int w10, int w11) {
  const int y0 = reflect101(y, rows), y1 = reflect101(y + 1, rows), x0 = reflect101(x, cols), x1 = reflect101(x + 1, cols);
  const int a = im[(size_t)y0 * cols + x0], b = im[(size_t)y0 * cols + x1], c = im[(size_t)y1 * cols + x0], d = im[(size_t)y1 * cols + x1];
  return (a * w00 + b * w01 + c * w10 + d * w11 + (1 << (kWBits - 5 - 1))) >> (kWBits - 5);
}
// derivative sample with the zero border of OpenCV's derivative buffer (BORDER_CONSTANT)
__device__ __noinline__ void lk_deriv_border(const short2* __restrict__ dI, int cols, int rows, int y, int x, int w00, int w01, int w10, int w11, int& ix, int& iy) {
  const bool yi0 = y >= 0 && y < rows, yi1 = y + 1 >= 0 && y + 1 < rows, xi0 = x >= 0 && x < cols, xi1 = x + 1 >= 0 && x + 1 < cols;
  const short2 z = make_short2(0, 0);
  const short2 d00 = (yi0 && xi0) ? dI[(size_t)y * cols + x] : z, d01 = (yi0 && xi1) ? dI[(size_t)y * cols + x + 1] : z;
  const short2 d10 = (yi1 && xi0) ? dI[(size_t)(y + 1) * cols + x] : z, d11 = (yi1 && xi1) ? dI[(size_t)(y + 1) * cols + x + 1] : z;
  ix = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << (kWBits - 1))) >> kWBits;
  iy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << (kWBits - 1))) >> kWBits;
}

// sum over the window of (J(q + .) - I) * (Ix, Iy) [mode 0] or |J(q + .) - I| [mode 1] at the window origin (inx, iny) of J
template <int MODE>
__device__ __forceinline__ void lk_window_sums(const LkWarpShared& W, const uint8_t* __restrict__ J, int cols, int rows, int win, int npx, int lane,
                                               int inx, int iny, int v00, int v01, int v10, int v11, bool tJ, int jx0, int jy0, int sjx, int& q1, int& q2) {
  q1 = 0; q2 = 0;
  const int tx = inx - jx0, ty = iny - jy0;
  if (tJ && tx >= 0 && ty >= 0 && tx <= 2 * kTileMargin && ty <= 2 * kTileMargin) {   // the window lies inside the staged tile
    const uint8_t* tb = &W.J[ty * kTilePitch + sjx + tx];
#pragma unroll 2
    for (int m = 0; m < kPix; m++) {
      const int idx = lane + 32 * m;
      if (idx < npx) {
        const short4 t = W.tmpl[idx];
        const int diff = lk_sample4(tb + (t.w >> 8) * kTilePitch + (t.w & 0xff), kTilePitch, v00, v01, v10, v11) - t.x;
        if (MODE == 0) { q1 += diff * t.y; q2 += diff * t.z; } else q1 += abs(diff);
      }
    }
  } else {
    const bool in_j = inx >= 0 && iny >= 0 && inx + win < cols && iny + win < rows;
#pragma unroll 2
    for (int m = 0; m < kPix; m++) {
      const int idx = lane + 32 * m;
      if (idx < npx) {
        const short4 t = W.tmpl[idx];
        const int y = iny + (t.w >> 8), x = inx + (t.w & 0xff);
        const int diff = (in_j ? lk_sample4(J + (size_t)y * cols + x, cols, v00, v01, v10, v11) : lk_sample_border(J, cols, rows, y, x, v00, v01, v10, v11)) - t.x;
        if (MODE == 0) { q1 += diff * t.y; q2 += diff * t.z; } else q1 += abs(diff);
      }
    }
  }
}

__global__ void __launch_bounds__(32 * kLkWarps) k_lk(const __grid_constant__ LkArgs a) {
  __shared__ LkWarpShared wsh[kLkWarps];
  __shared__ unsigned long long bars[kLkWarps];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int s = warp / a.max_pts, i = warp % a.max_pts;
  if (s >= a.n_streams || i >= a.n_pts[s]) return;
  if (a.gate && !a.gate[s]) return;
  LkWarpShared& W = wsh[wib];
  unsigned long long* bar = &bars[wib];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lk_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int win = a.win, npx = win * win;
  for (int m = 0; m < kPix; m++) { const int idx = lane + 32 * m; const int oy = idx / win; W.tmpl[idx].w = (short)((oy << 8) | (idx - oy * win)); }
  __syncwarp();
  uint32_t phase = 0;
  const float half = (win - 1) * 0.5f;
  const float2 pp = reinterpret_cast<const float2*>(a.prev_pts)[(size_t)s * a.max_pts + i];
  float2 np = reinterpret_cast<float2*>(a.next_pts)[(size_t)s * a.max_pts + i];
  bool ok = true; float errv = 0.f;
  const bool tile_geom = win + 1 + 2 * kTileMargin <= kTileRows;
  for (int level = a.max_level; level >= 0; level--) {
    const int cols = a.I.w[level], rows = a.I.h[level];
    const uint8_t* I = a.I.img[level] + (size_t)s * a.img_stride[level];
    const uint8_t* J = a.J.img[level] + (size_t)s * a.img_stride[level];
    const short2* dI = a.I.der[level] + (size_t)s * a.der_stride[level];
    const float scale = 1.f / (float)(1 << level);
    const float ppx = __fmul_rn(pp.x, scale), ppy = __fmul_rn(pp.y, scale);
    float nx, ny;
    if (level == a.max_level) {
      if (a.flags & GF2_LK_USE_INITIAL_FLOW) { nx = __fmul_rn(np.x, scale); ny = __fmul_rn(np.y, scale); } else { nx = ppx; ny = ppy; }
    } else { nx = __fmul_rn(np.x, 2.f); ny = __fmul_rn(np.y, 2.f); }
    np = make_float2(nx, ny);
    const float px = __fsub_rn(ppx, half), py = __fsub_rn(ppy, half);
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) { if (level == 0) { ok = false; errv = 0.f; } continue; }
    int w00, w01, w10, w11; lk_weights(px, py, ipx, ipy, w00, w01, w10, w11);
    // ---- the search tile of J: lane r copies row r (16-byte aligned start), in flight while the template is extracted
    const int jx0 = (int)floorf(__fsub_rn(nx, half)) - kTileMargin, jy0 = (int)floorf(__fsub_rn(ny, half)) - kTileMargin;
    const int ajx = jx0 & ~15, sjx = jx0 - ajx;
    const bool tJ = tile_geom && (cols & 15) == 0 && jx0 >= 0 && jy0 >= 0 && ajx + kTilePitch <= cols && jy0 + kTileRows <= rows;
    __syncwarp();                                                 // every lane is done with the previous level's tile and template
    if (tJ) {
      if (lane == 0) lk_mbar_expect(bar, (uint32_t)(kTileRows * kTilePitch));
      __syncwarp();
      lk_bulk_row(&W.J[lane * kTilePitch], J + (size_t)(jy0 + lane) * cols + ajx, kTilePitch, bar);
    }
    // ---- template: intensity (x32), Ix, Iy as int16 per window pixel -> shared; per-lane partial sums fit 32 bits
    //      (|Ix|, |Iy| <= 16 * 255 (Scharr), |J - I| <= 255 * 32, 14 pixels per lane)
    int p11 = 0, p12 = 0, p22 = 0;
    const bool in_t = ipx >= 0 && ipy >= 0 && ipx + win < cols && ipy + win < rows;
#pragma unroll 2
    for (int m = 0; m < kPix; m++) {
      const int idx = lane + 32 * m;
      if (idx < npx) {
        const int pk = W.tmpl[idx].w, y = ipy + (pk >> 8), x = ipx + (pk & 0xff);
        int iv, ix, iy;
        if (in_t) {
          iv = lk_sample4(I + (size_t)y * cols + x, cols, w00, w01, w10, w11);
          const short2* q = dI + (size_t)y * cols + x;
          const short2 d00 = q[0], d01 = q[1], d10 = q[cols], d11 = q[cols + 1];
          ix = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << (kWBits - 1))) >> kWBits;
          iy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << (kWBits - 1))) >> kWBits;
        } else {
          iv = lk_sample_border(I, cols, rows, y, x, w00, w01, w10, w11);
          lk_deriv_border(dI, cols, rows, y, x, w00, w01, w10, w11, ix, iy);
        }
        W.tmpl[idx] = make_short4((short)iv, (short)ix, (short)iy, (short)pk);
        p11 += ix * ix; p12 += ix * iy; p22 += iy * iy;
      }
    }
    if (tJ) { lk_mbar_wait(bar, phase); phase ^= 1u; }
    __syncwarp();
    long long s11 = p11, s12 = p12, s22 = p22;
    s11 = warp_sum_ll(s11); s12 = warp_sum_ll(s12); s22 = warp_sum_ll(s22);
    const float FS = 1.f / (float)(1 << 20);
    const float A11 = __fmul_rn((float)s11, FS), A12 = __fmul_rn((float)s12, FS), A22 = __fmul_rn((float)s22, FS);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float me = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))), (float)(2 * win * win));
    if (me < a.min_eig || D < 1.1920929e-07f) { if (level == 0) ok = false; continue; }
    D = __fdiv_rn(1.f, D);
    nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < a.max_iters; j++) {
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -win || inx >= cols || iny < -win || iny >= rows) { if (level == 0) ok = false; break; }
      int v00, v01, v10, v11; lk_weights(nx, ny, inx, iny, v00, v01, v10, v11);
      int q1, q2;
      lk_window_sums<0>(W, J, cols, rows, win, npx, lane, inx, iny, v00, v01, v10, v11, tJ, jx0, jy0, sjx, q1, q2);
      long long sb1 = q1, sb2 = q2;
      sb1 = warp_sum_ll(sb1); sb2 = warp_sum_ll(sb2);
      const float b1 = __fmul_rn((float)sb1, FS), b2 = __fmul_rn((float)sb2, FS);
      const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      nx = __fadd_rn(nx, ddx); ny = __fadd_rn(ny, ddy);
      np = make_float2(__fadd_rn(nx, half), __fadd_rn(ny, half));
      if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= a.eps2) break;
      if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
        np.x = __fsub_rn(np.x, __fmul_rn(ddx, 0.5f)); np.y = __fsub_rn(np.y, __fmul_rn(ddy, 0.5f));
        break;
      }
      pdx = ddx; pdy = ddy;
    }
    if (ok && level == 0) {  // err = mean |J - I| / 32 over the window at the final position
      const float fx = __fsub_rn(np.x, half), fy = __fsub_rn(np.y, half);
      const int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -win || inx >= cols || iny < -win || iny >= rows) { ok = false; errv = 0.f; continue; }
      int v00, v01, v10, v11; lk_weights(fx, fy, inx, iny, v00, v01, v10, v11);
      int pe, unused;
      lk_window_sums<1>(W, J, cols, rows, win, npx, lane, inx, iny, v00, v01, v10, v11, tJ, jx0, jy0, sjx, pe, unused);
      long long se = pe;
      se = warp_sum_ll(se);
      errv = __fdiv_rn((float)se, (float)(32 * win * win));
    }
  }
  if (lane == 0) {
    reinterpret_cast<float2*>(a.next_pts)[(size_t)s * a.max_pts + i] = np;
    a.status[(size_t)s * a.max_pts + i] = ok ? 1 : 0;
    if (a.err) a.err[(size_t)s * a.max_pts + i] = ok ? errv : 0.f;
  }
}

// status &= reverse ok && |prev - reverse| <= 0.5 (FeatureTracker::distance in double, feature_tracker.cpp:22-28,146)
__global__ void k_fb_check(int total, const float* prev_pts, const float* rev_pts, const uint8_t* rstatus, uint8_t* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double dx = (double)prev_pts[2 * i] - (double)rev_pts[2 * i], dy = (double)prev_pts[2 * i + 1] - (double)rev_pts[2 * i + 1];
  const double d = sqrt(dx * dx + dy * dy);
  status[i] = (status[i] && rstatus[i] && d <= 0.5) ? 1 : 0;
}

// succ_num < 10 -> redo the stream at full depth without the prediction (feature_tracker.cpp:124-131)
__global__ void k_count_gate(int max_pts, const int32_t* __restrict__ n_pts, const uint8_t* __restrict__ status, int min_succ, int32_t* __restrict__ gate) {
  const int s = blockIdx.x;
  int c = 0;
  for (int i = threadIdx.x; i < n_pts[s]; i += 32) c += status[(size_t)s * max_pts + i] ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (threadIdx.x == 0) gate[s] = c < min_succ ? 1 : 0;
}

#include "gf2_tracker_detect.cuh"

}  // namespace gf2

using namespace gf2;

struct gf2_tracker {
  gf2_tracker_cfg cfg;
  cudaStream_t stream;
  Pyr pyr[2];            // ping-pong: [cur_slot] holds the pyramid of the last `cur`
  size_t img_stride[kMaxLevels], der_stride[kMaxLevels];
  int cur_slot = 0; bool have_prev = false;
  int32_t* d_npts; float *d_prev_pts, *d_next_pts, *d_rev_pts, *d_err; uint8_t *d_status, *d_rstatus;
  int32_t* d_gate;
  // detector (allocated by the first gf2_tracker_detect / gf2_tracker_min_eigen_map)
  float* d_eig = nullptr; uint8_t *d_mask = nullptr, *d_det_img = nullptr; unsigned* d_vmax = nullptr;
  unsigned long long* d_keys = nullptr; int32_t *d_count = nullptr, *d_want = nullptr; int key_cap = 0;
  std::vector<unsigned long long> h_keys; std::vector<int32_t> h_count;
  // CLAHE of every uploaded image (gf2_tracker_set_equalize); the LUT buffer is allocated on first use
  double eq_clip = 0.0; int eq_tx = 8, eq_ty = 8; uint8_t* d_lut = nullptr; int lut_tiles = 0;
  std::vector<void*> allocs;
  cudaEvent_t ev[4];
  double timing[8];
};

#define GF2T_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return gf2::fail(GF2_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)

// cv::createCLAHE(clip, Size(tx, ty))->apply in place on n_streams device images of the tracker's size
static int clahe_inplace(gf2_tracker* h, uint8_t* d_img, int n_streams, double clip, int tx, int ty) {
  const int W = h->cfg.width, H = h->cfg.height;
  if (tx < 1 || ty < 1 || W % tx || H % ty) return gf2::fail(GF2_ERR_UNSUPPORTED, "CLAHE tile grid %dx%d does not divide the image %dx%d (cv pads by reflection there; not built)", tx, ty, W, H);
  if (!h->d_lut || h->lut_tiles < tx * ty) {
    uint8_t* p = nullptr;
    if (cudaMalloc((void**)&p, (size_t)h->cfg.max_streams * tx * ty * 256) != cudaSuccess) return gf2::fail(GF2_ERR_CUDA, "CLAHE LUT allocation failed");
    h->allocs.push_back(p); h->d_lut = p; h->lut_tiles = tx * ty;
  }
  const int area = (W / tx) * (H / ty);
  int clip_i = 0;
  if (clip > 0.0) { clip_i = (int)(clip * area / 256); if (clip_i < 1) clip_i = 1; }   // clahe.cpp: static_cast<int>(clipLimit * tileSizeTotal / histSize), max 1
  k_clahe_lut<<<dim3(tx * ty, n_streams), 256, 0, h->stream>>>(d_img, W, H, (size_t)W * H, tx, ty, clip_i, h->d_lut);
  dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8, n_streams);
  const volatile float one = 1.f;   // keep the two divisions in IEEE single precision exactly as cv computes inv_tw / inv_th
  const float inv_tw = one / (float)(W / tx), inv_th = one / (float)(H / ty);
  k_clahe_apply<<<g, b, 0, h->stream>>>(d_img, W, H, (size_t)W * H, tx, ty, inv_tw, inv_th, h->d_lut);
  GF2T_CUDA(cudaGetLastError());
  return GF2_OK;
}

static int build_pyramid(gf2_tracker* h, int slot, int n_streams, const uint8_t* host_img, size_t stride, bool derivs_upto_all, int deriv_levels) {
  Pyr& P = h->pyr[slot];
  const int W = h->cfg.width, H = h->cfg.height;
  GF2T_CUDA(cudaMemcpy2DAsync(P.img[0], W, host_img, stride, W, (size_t)H * n_streams, cudaMemcpyHostToDevice, h->stream));
  if (h->eq_clip > 0.0) { int rc = clahe_inplace(h, P.img[0], n_streams, h->eq_clip, h->eq_tx, h->eq_ty); if (rc) return rc; }
  for (int l = 1; l <= h->cfg.max_level; l++) {
    dim3 b(32, 8), g((P.w[l] + 31) / 32, (P.h[l] + 7) / 8, n_streams);
    k_pyrdown<<<g, b, 0, h->stream>>>(P.img[l - 1], P.w[l - 1], P.h[l - 1], P.w[l - 1], P.img[l], P.w[l], P.h[l], h->img_stride[l], h->img_stride[l - 1]);
  }
  (void)derivs_upto_all;
  for (int l = 0; l <= deriv_levels; l++) {
    dim3 b(32, 8), g((P.w[l] + 31) / 32, (P.h[l] + 7) / 8, n_streams);
    k_scharr<<<g, b, 0, h->stream>>>(P.img[l], P.w[l], P.h[l], P.w[l], h->img_stride[l], P.der[l], h->der_stride[l]);
  }
  GF2T_CUDA(cudaGetLastError());
  return GF2_OK;
}

extern "C" {

int gf2_tracker_create(const gf2_tracker_cfg* cfg, gf2_tracker** out) {
  if (!cfg || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (cfg->max_level < 0 || cfg->max_level >= kMaxLevels) return gf2::fail(GF2_ERR_INVALID, "max_level %d out of range [0, %d]", cfg->max_level, kMaxLevels - 1);
  if (cfg->win != 21) return gf2::fail(GF2_ERR_UNSUPPORTED, "window %d: the kernel is specialised for the reference's 21x21 window", cfg->win);
  if (cfg->width < 32 || cfg->height < 32 || cfg->max_pts < 1 || cfg->max_streams < 1) return gf2::fail(GF2_ERR_INVALID, "bad tracker dimensions");
  int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) { cudaGetLastError(); return gf2::fail(GF2_ERR_CUDA, "CUDA device %d not available (no CPU fallback exists)", cfg->device); }
  // OpenCV lowers maxLevel when a level gets smaller than the window; refuse such configurations instead of guessing
  { int w = cfg->width, hh = cfg->height; for (int l = 0; l < cfg->max_level; l++) { w = (w + 1) / 2; hh = (hh + 1) / 2; } if (w <= cfg->win || hh <= cfg->win) return gf2::fail(GF2_ERR_INVALID, "coarsest level smaller than the window"); }
  GF2T_CUDA(cudaSetDevice(cfg->device));
  gf2_tracker* h = new gf2_tracker();
  h->cfg = *cfg;
  cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  const int S = cfg->max_streams;
  auto alloc = [&](void** p, size_t bytes) { if (cudaMalloc(p, bytes) != cudaSuccess) return false; h->allocs.push_back(*p); return true; };
  bool ok = true;
  for (int slot = 0; slot < 2; slot++) {
    int w = cfg->width, hh = cfg->height;
    for (int l = 0; l <= cfg->max_level; l++) {
      h->pyr[slot].w[l] = w; h->pyr[slot].h[l] = hh;
      h->img_stride[l] = (size_t)w * hh; h->der_stride[l] = (size_t)w * hh;
      ok = ok && alloc((void**)&h->pyr[slot].img[l], (size_t)w * hh * S) && alloc((void**)&h->pyr[slot].der[l], (size_t)w * hh * S * sizeof(short2));
      w = (w + 1) / 2; hh = (hh + 1) / 2;
    }
  }
  const size_t np = (size_t)S * cfg->max_pts;
  ok = ok && alloc((void**)&h->d_npts, sizeof(int32_t) * S) && alloc((void**)&h->d_prev_pts, sizeof(float) * 2 * np) && alloc((void**)&h->d_next_pts, sizeof(float) * 2 * np) &&
       alloc((void**)&h->d_rev_pts, sizeof(float) * 2 * np) && alloc((void**)&h->d_err, sizeof(float) * np) && alloc((void**)&h->d_status, np) && alloc((void**)&h->d_rstatus, np) &&
       alloc((void**)&h->d_gate, sizeof(int32_t) * S);
  if (!ok) { gf2_tracker_destroy(h); return gf2::fail(GF2_ERR_CUDA, "tracker allocation failed"); }
  for (auto& e : h->ev) cudaEventCreate(&e);
  *out = h;
  return GF2_OK;
}

void gf2_tracker_destroy(gf2_tracker* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(h->stream);
  delete h;
}

static int lk_launch(gf2_tracker* h, int n_streams, int slotI, int slotJ, const float* d_prev, float* d_next, uint8_t* d_status, float* d_err, int flags, int max_level,
                     const int32_t* d_gate = nullptr) {
  LkArgs a;
  memset(&a, 0, sizeof(a));
  a.n_streams = n_streams; a.max_pts = h->cfg.max_pts; a.win = h->cfg.win; a.max_level = max_level; a.max_iters = h->cfg.max_iters; a.flags = flags;
  a.min_eig = (float)h->cfg.min_eig; a.eps2 = h->cfg.eps * h->cfg.eps;
  a.n_pts = h->d_npts; a.gate = d_gate; a.prev_pts = d_prev; a.next_pts = d_next; a.status = d_status; a.err = d_err;
  for (int l = 0; l < kMaxLevels; l++) { a.img_stride[l] = h->img_stride[l]; a.der_stride[l] = h->der_stride[l]; }
  a.I = h->pyr[slotI]; a.J = h->pyr[slotJ];
  const int warps = n_streams * h->cfg.max_pts;
  k_lk<<<(warps + kLkWarps - 1) / kLkWarps, 32 * kLkWarps, 0, h->stream>>>(a);
  GF2T_CUDA(cudaGetLastError());
  return GF2_OK;
}

static int track_common(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride, const int32_t* n_pts, const float* prev_pts,
                        float* cur_pts, uint8_t* status, float* err, int flags, int max_level, bool fb, const float* predict_pts = nullptr) {
  if (!h || !cur || !n_pts || !prev_pts || !cur_pts || !status) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (n_streams < 1 || n_streams > h->cfg.max_streams) return gf2::fail(GF2_ERR_INVALID, "n_streams %d outside [1, %d]", n_streams, h->cfg.max_streams);
  if (max_level < 0 || max_level > h->cfg.max_level) return gf2::fail(GF2_ERR_INVALID, "max_level %d exceeds the tracker capacity %d", max_level, h->cfg.max_level);
  if (stride < (size_t)h->cfg.width) return gf2::fail(GF2_ERR_INVALID, "stride smaller than the image width");
  for (int s = 0; s < n_streams; s++) if (n_pts[s] < 0 || n_pts[s] > h->cfg.max_pts) return gf2::fail(GF2_ERR_INVALID, "n_pts[%d] = %d exceeds max_pts %d", s, n_pts[s], h->cfg.max_pts);
  if (!prev && !h->have_prev) return gf2::fail(GF2_ERR_INVALID, "prev == NULL but no pyramid is cached from an earlier call");
  cudaSetDevice(h->cfg.device);
  const size_t np = (size_t)n_streams * h->cfg.max_pts;
  cudaEventRecord(h->ev[0], h->stream);
  int slotI = h->cur_slot, slotJ = 1 - h->cur_slot;
  if (prev) { int rc = build_pyramid(h, slotI, n_streams, prev, stride, true, h->cfg.max_level); if (rc) return rc; }
  { int rc = build_pyramid(h, slotJ, n_streams, cur, stride, true, h->cfg.max_level); if (rc) return rc; }  // derivatives of cur: reverse pass now, template of the next call
  GF2T_CUDA(cudaMemcpyAsync(h->d_npts, n_pts, sizeof(int32_t) * n_streams, cudaMemcpyHostToDevice, h->stream));
  GF2T_CUDA(cudaMemcpyAsync(h->d_prev_pts, prev_pts, sizeof(float) * 2 * np, cudaMemcpyHostToDevice, h->stream));
  if (predict_pts) GF2T_CUDA(cudaMemcpyAsync(h->d_next_pts, predict_pts, sizeof(float) * 2 * np, cudaMemcpyHostToDevice, h->stream));
  else if (flags & GF2_LK_USE_INITIAL_FLOW) GF2T_CUDA(cudaMemcpyAsync(h->d_next_pts, cur_pts, sizeof(float) * 2 * np, cudaMemcpyHostToDevice, h->stream));
  cudaEventRecord(h->ev[1], h->stream);
  if (predict_pts) {
    // hasPrediction (feature_tracker.cpp:118-131): level-1 LK from the predicted positions; a stream with fewer than 10 successes
    // is redone at max_level from prev_pts (flags 0), decided on the device per stream
    const int ml1 = 1 < h->cfg.max_level ? 1 : h->cfg.max_level;
    { int rc = lk_launch(h, n_streams, slotI, slotJ, h->d_prev_pts, h->d_next_pts, h->d_status, h->d_err, GF2_LK_USE_INITIAL_FLOW, ml1); if (rc) return rc; }
    k_count_gate<<<n_streams, 32, 0, h->stream>>>(h->cfg.max_pts, h->d_npts, h->d_status, 10, h->d_gate);
    { int rc = lk_launch(h, n_streams, slotI, slotJ, h->d_prev_pts, h->d_next_pts, h->d_status, h->d_err, 0, max_level, h->d_gate); if (rc) return rc; }
  } else {
    int rc = lk_launch(h, n_streams, slotI, slotJ, h->d_prev_pts, h->d_next_pts, h->d_status, h->d_err, flags, max_level); if (rc) return rc;
  }
  if (fb) {
    // reverse: cur -> prev at maxLevel 1 with the initial flow = prev_pts (feature_tracker.cpp:139-142)
    GF2T_CUDA(cudaMemcpyAsync(h->d_rev_pts, h->d_prev_pts, sizeof(float) * 2 * np, cudaMemcpyDeviceToDevice, h->stream));
    int rc = lk_launch(h, n_streams, slotJ, slotI, h->d_next_pts, h->d_rev_pts, h->d_rstatus, nullptr, GF2_LK_USE_INITIAL_FLOW, 1 < h->cfg.max_level ? 1 : h->cfg.max_level);
    if (rc) return rc;
    k_fb_check<<<((int)np + 255) / 256, 256, 0, h->stream>>>((int)np, h->d_prev_pts, h->d_rev_pts, h->d_rstatus, h->d_status);
  }
  cudaEventRecord(h->ev[2], h->stream);
  GF2T_CUDA(cudaMemcpyAsync(cur_pts, h->d_next_pts, sizeof(float) * 2 * np, cudaMemcpyDeviceToHost, h->stream));
  GF2T_CUDA(cudaMemcpyAsync(status, h->d_status, np, cudaMemcpyDeviceToHost, h->stream));
  if (err) GF2T_CUDA(cudaMemcpyAsync(err, h->d_err, sizeof(float) * np, cudaMemcpyDeviceToHost, h->stream));
  cudaEventRecord(h->ev[3], h->stream);
  GF2T_CUDA(cudaStreamSynchronize(h->stream));
  h->cur_slot = slotJ; h->have_prev = true;  // prev_img = cur_img (feature_tracker.cpp:307)
  float ms; memset(h->timing, 0, sizeof(h->timing));
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]); h->timing[0] = ms;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->timing[1] = ms;  // upload + pyramids + derivatives
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->timing[2] = ms;  // LK kernels
  h->timing[3] = (prev ? 2.0 : 1.0) * (h->cfg.max_level + (h->cfg.max_level + 1)) + (fb ? 3.0 : 1.0) + (predict_pts ? 2.0 : 0.0);  // launches
  return GF2_OK;
}

int gf2_tracker_track(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride, const int32_t* n_pts,
                      const float* prev_pts, float* cur_pts, uint8_t* status, float* err, int flags, int max_level) {
  return track_common(h, n_streams, prev, cur, stride, n_pts, prev_pts, cur_pts, status, err, flags, max_level, false);
}
int gf2_tracker_track_fb(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride, const int32_t* n_pts,
                         const float* prev_pts, float* cur_pts, uint8_t* status, int max_level) {
  return track_common(h, n_streams, prev, cur, stride, n_pts, prev_pts, cur_pts, status, nullptr, 0, max_level, true);
}
int gf2_tracker_track_image(gf2_tracker* h, int n_streams, const uint8_t* prev, const uint8_t* cur, size_t stride, const int32_t* n_pts,
                            const float* prev_pts, const float* predict_pts, int flow_back, float* cur_pts, uint8_t* status, int max_level) {
  return track_common(h, n_streams, prev, cur, stride, n_pts, prev_pts, cur_pts, status, nullptr, 0, max_level, flow_back != 0, predict_pts);
}

// ---------------------------------------------------------------------------------------------------------------- detector
// featureselect.cpp: candidates in greaterThanPtr order (a max-heap pops them lazily: selection usually stops early), greedy
// acceptance against the corners already taken within min_distance, looked up through a grid of cell size cvRound(min_distance).
// K (n sort keys ordered(value) << 32 | y * W + x) is reordered in place. Returns the number of corners written to O.
static int select_corners(unsigned long long* K, int n, int W, int H, int want, double min_distance, float* O) {
  if (n <= 0 || want <= 0) return 0;
  const bool spaced = min_distance >= 1.0;
  const int cell = spaced ? (int)lrint(min_distance) : 1;
  const int gw = (W + cell - 1) / cell, gh = (H + cell - 1) / cell;
  const double md2 = min_distance * min_distance;
  std::vector<std::vector<int>> grid;
  if (spaced) grid.assign((size_t)gw * gh, std::vector<int>());
  int got = 0;
  std::make_heap(K, K + n);
  for (int left = n; left > 0 && got < want; left--) {
    std::pop_heap(K, K + left);
    const int idx = (int)(K[left - 1] & 0xffffffffu), y = idx / W, x = idx - y * W;
    bool good = true;
    if (spaced) {
      const int xc = x / cell, yc = y / cell;
      const int x1 = xc > 0 ? xc - 1 : 0, y1 = yc > 0 ? yc - 1 : 0, x2 = xc + 1 < gw ? xc + 1 : gw - 1, y2 = yc + 1 < gh ? yc + 1 : gh - 1;
      for (int yy = y1; yy <= y2 && good; yy++)
        for (int xx = x1; xx <= x2 && good; xx++)
          for (int q : grid[(size_t)yy * gw + xx]) {
            const int qy = q / W, qx = q - qy * W; const double dx = x - qx, dy = y - qy;
            if (dx * dx + dy * dy < md2) { good = false; break; }
          }
      if (good) grid[(size_t)yc * gw + xc].push_back(idx);
    }
    if (good) { O[2 * got] = (float)x; O[2 * got + 1] = (float)y; got++; }
  }
  return got;
}

int gf2_detect_select(const uint64_t* keys, int n, int width, int height, int max_corners, double min_distance, float* out_xy, int32_t* out_n) {
  if ((!keys && n > 0) || !out_xy || !out_n || n < 0 || width < 1 || height < 1 || max_corners < 0 || min_distance < 0.0) return gf2::fail(GF2_ERR_INVALID, "bad argument");
  std::vector<unsigned long long> K(keys, keys + n);
  *out_n = select_corners(K.data(), n, width, height, max_corners, min_distance, out_xy);
  return GF2_OK;
}

static int detect_alloc(gf2_tracker* h) {
  if (h->d_eig) return GF2_OK;
  const int S = h->cfg.max_streams; const size_t plane = (size_t)h->cfg.width * h->cfg.height;
  h->key_cap = (int)(plane / 4);   // a 3x3 local maximum excludes its 8 neighbours unless they tie; overflow is reported, never truncated silently
  auto alloc = [&](void** p, size_t bytes) { if (cudaMalloc(p, bytes) != cudaSuccess) return false; h->allocs.push_back(*p); return true; };
  const bool ok = alloc((void**)&h->d_eig, sizeof(float) * plane * S) && alloc((void**)&h->d_mask, plane * S) &&
                  alloc((void**)&h->d_det_img, plane * S) && alloc((void**)&h->d_vmax, sizeof(unsigned) * S) && alloc((void**)&h->d_keys, sizeof(unsigned long long) * (size_t)h->key_cap * S) &&
                  alloc((void**)&h->d_count, sizeof(int32_t) * S) && alloc((void**)&h->d_want, sizeof(int32_t) * S);
  if (!ok) { h->d_eig = nullptr; return gf2::fail(GF2_ERR_CUDA, "detector allocation failed"); }
  h->h_count.resize(S);
  return GF2_OK;
}

// Sobel -> covariance -> box sums -> min eigenvalue (+ masked maximum) of n_streams images on the handle's stream
static int detect_eig(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, const uint8_t* mask) {
  const int W = h->cfg.width, H = h->cfg.height; const size_t plane = (size_t)W * H;
  const uint8_t* d_img;
  if (img) {
    GF2T_CUDA(cudaMemcpy2DAsync(h->d_det_img, W, img, stride, W, (size_t)H * n_streams, cudaMemcpyHostToDevice, h->stream));
    if (h->eq_clip > 0.0) { int rc = clahe_inplace(h, h->d_det_img, n_streams, h->eq_clip, h->eq_tx, h->eq_ty); if (rc) return rc; }
    d_img = h->d_det_img;
  } else {
    if (!h->have_prev) return gf2::fail(GF2_ERR_INVALID, "img == NULL but no image is cached from an earlier gf2_tracker_track* call");
    d_img = h->pyr[h->cur_slot].img[0];   // cur_img of the last track call
  }
  if (mask) GF2T_CUDA(cudaMemcpyAsync(h->d_mask, mask, plane * n_streams, cudaMemcpyHostToDevice, h->stream));
  GF2T_CUDA(cudaMemsetAsync(h->d_vmax, 0, sizeof(unsigned) * n_streams, h->stream));
  k_gftt_eig<<<dim3((W + 63) / 64, n_streams), 64, 0, h->stream>>>(d_img, plane, mask ? h->d_mask : nullptr, W, H, h->d_eig, h->d_vmax);
  GF2T_CUDA(cudaGetLastError());
  return GF2_OK;
}

static int detect_check(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride) {
  if (!h) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (n_streams < 1 || n_streams > h->cfg.max_streams) return gf2::fail(GF2_ERR_INVALID, "n_streams %d outside [1, %d]", n_streams, h->cfg.max_streams);
  if (img && stride < (size_t)h->cfg.width) return gf2::fail(GF2_ERR_INVALID, "stride smaller than the image width");
  cudaSetDevice(h->cfg.device);
  return detect_alloc(h);
}

int gf2_tracker_min_eigen_map(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, float* eig) {
  if (!eig) return gf2::fail(GF2_ERR_INVALID, "null argument");
  { int rc = detect_check(h, n_streams, img, stride); if (rc) return rc; }
  { int rc = detect_eig(h, n_streams, img, stride, nullptr); if (rc) return rc; }
  GF2T_CUDA(cudaMemcpyAsync(eig, h->d_eig, sizeof(float) * (size_t)h->cfg.width * h->cfg.height * n_streams, cudaMemcpyDeviceToHost, h->stream));
  GF2T_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_tracker_detect(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, const uint8_t* mask, const int32_t* max_corners,
                       double quality_level, double min_distance, float* out_xy, int32_t* out_n) {
  if (!max_corners || !out_xy || !out_n) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (!(quality_level > 0.0) || min_distance < 0.0) return gf2::fail(GF2_ERR_INVALID, "quality_level must be > 0 and min_distance >= 0");
  { int rc = detect_check(h, n_streams, img, stride); if (rc) return rc; }
  for (int s = 0; s < n_streams; s++) if (max_corners[s] < 0) return gf2::fail(GF2_ERR_INVALID, "max_corners[%d] = %d is negative", s, max_corners[s]);
  const int W = h->cfg.width, H = h->cfg.height;
  cudaEventRecord(h->ev[0], h->stream);
  { int rc = detect_eig(h, n_streams, img, stride, mask); if (rc) return rc; }
  GF2T_CUDA(cudaMemcpyAsync(h->d_want, max_corners, sizeof(int32_t) * n_streams, cudaMemcpyHostToDevice, h->stream));
  GF2T_CUDA(cudaMemsetAsync(h->d_count, 0, sizeof(int32_t) * n_streams, h->stream));
  dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8, n_streams);
  k_gftt_nms<<<g, b, 0, h->stream>>>(h->d_eig, mask ? h->d_mask : nullptr, W, H, h->d_vmax, quality_level, h->d_want, h->d_keys, h->key_cap, h->d_count);
  GF2T_CUDA(cudaGetLastError());
  cudaEventRecord(h->ev[1], h->stream);
  GF2T_CUDA(cudaMemcpyAsync(h->h_count.data(), h->d_count, sizeof(int32_t) * n_streams, cudaMemcpyDeviceToHost, h->stream));
  GF2T_CUDA(cudaStreamSynchronize(h->stream));
  size_t total = 0;
  for (int s = 0; s < n_streams; s++) {
    if (h->h_count[s] > h->key_cap) return gf2::fail(GF2_ERR_INVALID, "stream %d: %d corner candidates exceed the capacity %d", s, h->h_count[s], h->key_cap);
    total += (size_t)h->h_count[s];
  }
  h->h_keys.resize(total ? total : 1);
  { size_t off = 0;
    for (int s = 0; s < n_streams; s++) {
      if (h->h_count[s]) GF2T_CUDA(cudaMemcpyAsync(h->h_keys.data() + off, h->d_keys + (size_t)s * h->key_cap, sizeof(unsigned long long) * h->h_count[s], cudaMemcpyDeviceToHost, h->stream));
      off += (size_t)h->h_count[s];
    } }
  cudaEventRecord(h->ev[2], h->stream);
  GF2T_CUDA(cudaStreamSynchronize(h->stream));
  size_t off = 0;
  for (int s = 0; s < n_streams; s++) {
    const int n = h->h_count[s];
    const int want = max_corners[s] < h->cfg.max_pts ? max_corners[s] : h->cfg.max_pts;
    out_n[s] = select_corners(h->h_keys.data() + off, n, W, H, want, min_distance, out_xy + (size_t)s * h->cfg.max_pts * 2);
    off += (size_t)n;
  }
  float ms; memset(h->timing, 0, sizeof(h->timing));
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]); h->timing[0] = ms;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->timing[4] = ms;  // upload + detector kernels
  h->timing[3] = 3.0;
  h->timing[5] = (double)total;                                       // candidates moved to the host
  return GF2_OK;
}

int gf2_tracker_set_equalize(gf2_tracker* h, double clip_limit, int tiles_x, int tiles_y) {
  if (!h) return gf2::fail(GF2_ERR_INVALID, "null handle");
  if (clip_limit > 0.0 && (tiles_x < 1 || tiles_y < 1 || h->cfg.width % tiles_x || h->cfg.height % tiles_y))
    return gf2::fail(GF2_ERR_UNSUPPORTED, "CLAHE tile grid %dx%d does not divide the image %dx%d (cv pads by reflection there; not built)", tiles_x, tiles_y, h->cfg.width, h->cfg.height);
  h->eq_clip = clip_limit > 0.0 ? clip_limit : 0.0; h->eq_tx = tiles_x; h->eq_ty = tiles_y;
  return GF2_OK;
}

int gf2_tracker_equalize(gf2_tracker* h, int n_streams, const uint8_t* img, size_t stride, double clip_limit, int tiles_x, int tiles_y, uint8_t* out) {
  if (!img || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  { int rc = detect_check(h, n_streams, img, stride); if (rc) return rc; }
  const int W = h->cfg.width, H = h->cfg.height;
  GF2T_CUDA(cudaMemcpy2DAsync(h->d_det_img, W, img, stride, W, (size_t)H * n_streams, cudaMemcpyHostToDevice, h->stream));
  { int rc = clahe_inplace(h, h->d_det_img, n_streams, clip_limit, tiles_x, tiles_y); if (rc) return rc; }
  GF2T_CUDA(cudaMemcpyAsync(out, h->d_det_img, (size_t)W * H * n_streams, cudaMemcpyDeviceToHost, h->stream));
  GF2T_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_tracker_get_image(gf2_tracker* h, int n_streams, uint8_t* out) {
  if (!h || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  if (n_streams < 1 || n_streams > h->cfg.max_streams) return gf2::fail(GF2_ERR_INVALID, "n_streams %d outside [1, %d]", n_streams, h->cfg.max_streams);
  if (!h->have_prev) return gf2::fail(GF2_ERR_INVALID, "no image is cached from an earlier gf2_tracker_track* call");
  cudaSetDevice(h->cfg.device);
  GF2T_CUDA(cudaMemcpyAsync(out, h->pyr[h->cur_slot].img[0], (size_t)h->cfg.width * h->cfg.height * n_streams, cudaMemcpyDeviceToHost, h->stream));
  GF2T_CUDA(cudaStreamSynchronize(h->stream));
  return GF2_OK;
}

int gf2_tracker_last_timing(gf2_tracker* h, double out[8]) {
  if (!h || !out) return gf2::fail(GF2_ERR_INVALID, "null argument");
  memcpy(out, h->timing, sizeof(h->timing));
  return GF2_OK;
}

}  // extern "C"
