// gf2_tracker_detect.cuh — cv::goodFeaturesToTrack(img, corners, maxCorners, 0.01, MIN_DIST, mask) as FeatureTracker::trackImage
// calls it (VE/featureTracker/feature_tracker.cpp:198; blockSize 3, Sobel aperture 3, min-eigenvalue score) for a batch of
// independent image streams. OpenCV is an un-vendored dependency; the float32 / float64 operation order below is the one that
// reproduces cv2.cornerMinEigenVal bit for bit (probed against cv2 4.13; DESIGN.md section 4 documents it):
//   k_gftt_eig   Sobel with the scale folded into the smoothing taps (fused multiply-adds exactly where cv's SIMD path has them) and the
//                covariance terms xx, xy, yy in float32, formed on the fly from the image; 3x3 box sums in double: row sums left to right,
//                the column pass as cv's RUNNING sum down the image (one thread per column; the rounding history of that sum is part of
//                the result), min eigenvalue, masked maximum
//   k_gftt_nms   threshold at float(maxVal * qualityLevel), 3x3 dilation equality test, interior pixels, mask -> sort keys
// included by gf2_tracker.cu (namespace gf2)

__device__ __forceinline__ unsigned gftt_ordered(float f) {  // monotone float -> unsigned
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float gftt_unordered(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// covariance terms (dx dx, dx dy, dy dy) of one pixel straight from the 8-bit image; the float32 operation order is cv's (header comment)
__device__ __forceinline__ void gftt_cov_px(const uint8_t* __restrict__ S, int W, int H, int y, int x, float& xx, float& xy, float& yy) {
  const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W), ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
  const uint8_t *r0 = S + (size_t)ym * W, *r1 = S + (size_t)y * W, *r2 = S + (size_t)yp * W;
  const float s1 = (float)(1.0 / (4.0 * 3.0 * 255.0)), s2 = (float)(2.0 / (4.0 * 3.0 * 255.0));
  const float a0 = r0[xm], b0 = r0[x], c0 = r0[xp], a1 = r1[xm], c1 = r1[xp], a2 = r2[xm], b2 = r2[x], c2 = r2[xp];
  // Dx: row pass [-1 0 1] (exact), column pass [1 2 1]*scale as fma(d0 + d2, s, fl(d1 * 2s))
  const float d0 = c0 - a0, d1 = c1 - a1, d2 = c2 - a2;
  const float dx = __fmaf_rn(__fadd_rn(d0, d2), s1, __fmul_rn(d1, s2));
  // Dy: row pass [1 2 1]*scale as fma(I[x+1], s, fma(I[x], 2s, fl(I[x-1] * s))), column pass [-1 0 1]
  const float q0 = __fmaf_rn(c0, s1, __fmaf_rn(b0, s2, __fmul_rn(a0, s1)));
  const float q2 = __fmaf_rn(c2, s1, __fmaf_rn(b2, s2, __fmul_rn(a2, s1)));
  const float dy = __fsub_rn(q2, q0);
  xx = __fmul_rn(dx, dx); xy = __fmul_rn(dx, dy); yy = __fmul_rn(dy, dy);
}

struct GfttRow { double xx, xy, yy; };
// row sums (left to right, in double) of the three covariance terms over the columns xm, x, xp of image row y. The covariance planes are
// not materialised (round 2: k_gftt_cov wrote 12 B per pixel that k_gftt_eig read back three times); the pixels come from L1.
__device__ __forceinline__ GfttRow gftt_row_sum(const uint8_t* __restrict__ S, int W, int H, int y, int xm, int x, int xp) {
  float a[3], b[3], c[3];
  gftt_cov_px(S, W, H, y, xm, a[0], b[0], c[0]); gftt_cov_px(S, W, H, y, x, a[1], b[1], c[1]); gftt_cov_px(S, W, H, y, xp, a[2], b[2], c[2]);
  GfttRow r;
  r.xx = __dadd_rn(__dadd_rn((double)a[0], (double)a[1]), (double)a[2]);
  r.xy = __dadd_rn(__dadd_rn((double)b[0], (double)b[1]), (double)b[2]);
  r.yy = __dadd_rn(__dadd_rn((double)c[0], (double)c[1]), (double)c[2]);
  return r;
}

// The same row sums for CONSECUTIVE image rows without re-reading pixels: per image row and position p in (xm, x, xp) the Sobel row passes
//   d = I[p+1] - I[p-1]  (exact)      q = fma(I[p+1], s, fma(I[p], 2s, fl(I[p-1] * s)))
// are formed once (9 pixel loads) and kept for the three rows y - 1, y, y + 1 that use them:
//   dx = fma(d[y-1] + d[y+1], s, fl(d[y] * 2s))      dy = q[y+1] - q[y-1]         (the operations of gftt_cov_px, in its order)
struct GfttDQ { float d[3], q[3]; };
struct GfttPix { float v[9]; };   // pixels (p-1, p, p+1) of the three positions
__device__ __forceinline__ GfttPix gftt_load_row(const uint8_t* __restrict__ S, int W, int row, const int (&cols)[9]) {
  GfttPix r; const uint8_t* p = S + (size_t)row * W;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] = p[cols[i]];
  return r;
}
__device__ __forceinline__ GfttDQ gftt_dq(const GfttPix& px) {
  const float s1 = (float)(1.0 / (4.0 * 3.0 * 255.0)), s2 = (float)(2.0 / (4.0 * 3.0 * 255.0));
  GfttDQ r;
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const float a = px.v[3 * p], b = px.v[3 * p + 1], c = px.v[3 * p + 2];
    r.d[p] = c - a;
    r.q[p] = __fmaf_rn(c, s1, __fmaf_rn(b, s2, __fmul_rn(a, s1)));
  }
  return r;
}
__device__ __forceinline__ GfttRow gftt_row_sum_dq(const GfttDQ& up, const GfttDQ& mid, const GfttDQ& dn) {
  const float s1 = (float)(1.0 / (4.0 * 3.0 * 255.0)), s2 = (float)(2.0 / (4.0 * 3.0 * 255.0));
  float xx[3], xy[3], yy[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const float dx = __fmaf_rn(__fadd_rn(up.d[p], dn.d[p]), s1, __fmul_rn(mid.d[p], s2));
    const float dy = __fsub_rn(dn.q[p], up.q[p]);
    xx[p] = __fmul_rn(dx, dx); xy[p] = __fmul_rn(dx, dy); yy[p] = __fmul_rn(dy, dy);
  }
  GfttRow r;
  r.xx = __dadd_rn(__dadd_rn((double)xx[0], (double)xx[1]), (double)xx[2]);
  r.xy = __dadd_rn(__dadd_rn((double)xy[0], (double)xy[1]), (double)xy[2]);
  r.yy = __dadd_rn(__dadd_rn((double)yy[0], (double)yy[1]), (double)yy[2]);
  return r;
}

// one thread per image column; eig [stream][H][W]; vmax [stream] ordered-float maximum over the masked pixels
constexpr int kGfttUnroll = 6;   // image rows of pixel (and mask) loads in flight per thread (the column walk is latency bound)
__global__ void __launch_bounds__(64) k_gftt_eig(const uint8_t* __restrict__ img, size_t img_stride, const uint8_t* __restrict__ mask, int W, int H, float* __restrict__ eig,
                                                  unsigned* __restrict__ vmax) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  unsigned best = 0u;  // below every ordered float
  if (x < W) {
    const size_t plane = (size_t)W * H;
    const uint8_t* C = img + (size_t)s * img_stride;
    const uint8_t* M = mask ? mask + (size_t)s * plane : nullptr;
    float* E = eig + (size_t)s * plane;
    const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    const int cols[9] = {reflect101(xm - 1, W), xm, reflect101(xm + 1, W), xm, x, xp, reflect101(xp - 1, W), xp, reflect101(xp + 1, W)};
    // SUM = r[-1] + r[0] (ColumnSum: zero, then the first ksize-1 rows); r[-1] = r[1] (BORDER_REFLECT_101)
    GfttRow prev2 = gftt_row_sum(C, W, H, reflect101(-1, H), xm, x, xp);  // r[y-1] of the running update
    GfttRow prev1 = gftt_row_sum(C, W, H, 0, xm, x, xp);                   // r[y]
    double Sxx = __dadd_rn(__dadd_rn(0.0, prev2.xx), prev1.xx), Sxy = __dadd_rn(__dadd_rn(0.0, prev2.xy), prev1.xy), Syy = __dadd_rn(__dadd_rn(0.0, prev2.yy), prev1.yy);
    // rolling Sobel row passes of the image rows (y, y + 1, y + 2) that the incoming row sum r[y + 1] needs, and the pixel rows ahead of them
    GfttDQ r0 = gftt_dq(gftt_load_row(C, W, 0, cols)), r1 = gftt_dq(gftt_load_row(C, W, reflect101(1, H), cols)), r2 = gftt_dq(gftt_load_row(C, W, reflect101(2, H), cols));
    GfttPix pix[kGfttUnroll];
    uint8_t mk[kGfttUnroll];     // mask bytes of the rows ahead: a load per row consumed straight away stalled the walk for a DRAM latency per row
#pragma unroll
    for (int k = 0; k < kGfttUnroll; k++) { pix[k] = gftt_load_row(C, W, reflect101(3 + k, H), cols); mk[k] = M ? M[(size_t)min(k, H - 1) * W + x] : 1; }
    for (int y0 = 0; y0 < H; y0 += kGfttUnroll) {
#pragma unroll
      for (int k = 0; k < kGfttUnroll; k++) {
        const int y = y0 + k;
        if (y < H) {
          // r[y+1]: rows below the image reflect (the last one is r[H-2] again: not the next consecutive row, summed from the pixels)
          const GfttRow nx = (y < H - 1) ? gftt_row_sum_dq(r0, r1, r2) : gftt_row_sum(C, W, H, reflect101(y + 1, H), xm, x, xp);
          r0 = r1; r1 = r2; r2 = gftt_dq(pix[k]);                                   // image rows (y + 1, y + 2, y + 3)
          pix[k] = gftt_load_row(C, W, reflect101(y + 3 + kGfttUnroll, H), cols);   // in flight for kGfttUnroll rows
          const uint8_t mrow = mk[k];
          mk[k] = M ? M[(size_t)min(y + kGfttUnroll, H - 1) * W + x] : 1;
          const double sxx = __dadd_rn(Sxx, nx.xx), sxy = __dadd_rn(Sxy, nx.xy), syy = __dadd_rn(Syy, nx.yy);
          Sxx = __dsub_rn(sxx, prev2.xx); Sxy = __dsub_rn(sxy, prev2.xy); Syy = __dsub_rn(syy, prev2.yy);
          prev2 = prev1; prev1 = nx;
          const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
          const float t = __fsub_rn(a, c);
          const float e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
          E[(size_t)y * W + x] = e;
          if (mrow) best = max(best, gftt_ordered(e));
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best) atomicMax(&vmax[s], best);
}

// candidates: key = ordered(value) << 32 | (y * W + x); descending key order == cv's greaterThanPtr (value, then address)
__global__ void __launch_bounds__(256) k_gftt_nms(const float* __restrict__ eig, const uint8_t* __restrict__ mask, int W, int H, const unsigned* __restrict__ vmax,
                                                   double quality, const int32_t* __restrict__ want, unsigned long long* __restrict__ keys, int cap, int32_t* __restrict__ count) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, s = blockIdx.z;
  if (x < 1 || y < 1 || x >= W - 1 || y >= H - 1 || want[s] <= 0) return;
  const size_t plane = (size_t)W * H;
  if (mask && !mask[(size_t)s * plane + (size_t)y * W + x]) return;
  const unsigned km = vmax[s];
  const double max_val = km ? (double)gftt_unordered(km) : 0.0;       // minMaxLoc over an empty mask reports 0
  const float thr = (float)(max_val * quality);                       // cv::threshold casts the double threshold to float
  const float* E = eig + (size_t)s * plane + (size_t)y * W + x;
  const float v = E[0] > thr ? E[0] : 0.f;                            // THRESH_TOZERO
  if (v == 0.f) return;
  float d = v;
#pragma unroll
  for (int dy = -1; dy <= 1; dy++)
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) { const float n = E[dy * W + dx]; d = fmaxf(d, n > thr ? n : 0.f); }
  if (v != d) return;
  const int slot = atomicAdd(&count[s], 1);
  if (slot < cap) keys[(size_t)s * cap + slot] = ((unsigned long long)gftt_ordered(v) << 32) | (unsigned)(y * W + x);
}

// ------------------------------------------------------------------------------------------------ CLAHE
// cv::createCLAHE(clipLimit, tileGridSize)->apply for 8-bit images (VE/rosNodeTest.cpp:271-276, `equalize: 1` in m3dgr.yaml):
//   k_clahe_lut    one CTA per (tile, stream): histogram, clip at max(int(clipLimit * tileArea / 256), 1), redistribution (batch + the
//                  strided residual), LUT = saturate(cvRound(float(cumsum) * (255.f / tileArea)))
//   k_clahe_apply  per pixel: bilinear blend of the four neighbouring tile LUTs in float32 with cv's operation order (no contraction)
// In place on the image buffer (the LUT kernel has read every pixel before the apply kernel starts).
__global__ void __launch_bounds__(256) k_clahe_lut(const uint8_t* __restrict__ img, int W, int H, size_t img_stride, int tiles_x, int tiles_y, int clip_limit,
                                                    uint8_t* __restrict__ lut /* [stream][tiles_y * tiles_x][256] */) {
  __shared__ int hist[256];
  __shared__ int wsum[8];
  __shared__ int s_clipped;
  const int t = threadIdx.x, tile = blockIdx.x, s = blockIdx.y;
  const int tw = W / tiles_x, th = H / tiles_y, tx = tile % tiles_x, ty = tile / tiles_x;
  hist[t] = 0;
  if (t == 0) s_clipped = 0;
  __syncthreads();
  const uint8_t* S = img + (size_t)s * img_stride + (size_t)(ty * th) * W + tx * tw;
  for (int i = t; i < tw * th; i += 256) { const int y = i / tw, x = i - y * tw; atomicAdd(&hist[S[(size_t)y * W + x]], 1); }
  __syncthreads();
  int h = hist[t];
  if (clip_limit > 0) {
    if (h > clip_limit) { atomicAdd(&s_clipped, h - clip_limit); h = clip_limit; }
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    int residual = clipped - batch * 256;
    h += batch;
    if (residual != 0) {  // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
      const int step = max(256 / residual, 1);
      if (t % step == 0 && t / step < residual) h++;
    }
  }
  // inclusive prefix sum over the 256 bins
  int v = h;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if ((t & 31) >= o) v += n; }
  if ((t & 31) == 31) wsum[t >> 5] = v;
  __syncthreads();
  int base = 0;
  for (int k = 0; k < (t >> 5); k++) base += wsum[k];
  const int sum = v + base;
  const float lut_scale = __fdiv_rn(255.f, (float)(tw * th));
  int r = __float2int_rn(__fmul_rn((float)sum, lut_scale));
  r = r < 0 ? 0 : (r > 255 ? 255 : r);
  lut[((size_t)s * tiles_x * tiles_y + tile) * 256 + t] = (uint8_t)r;
}

// inv_tw = 1.f / tile width, inv_th = 1.f / tile height: IEEE divisions done once on the host (two fp32 divisions per pixel were a third
// of the kernel's instructions)
__global__ void __launch_bounds__(256) k_clahe_apply(uint8_t* __restrict__ img, int W, int H, size_t img_stride, int tiles_x, int tiles_y, float inv_tw, float inv_th,
                                                      const uint8_t* __restrict__ lut) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, s = blockIdx.z;
  if (x >= W || y >= H) return;
  const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f), tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
  int tx2 = tx1 + 1, ty2 = ty1 + 1;
  const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.f, xa), ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.f, ya);
  tx1 = max(tx1, 0); tx2 = min(tx2, tiles_x - 1); ty1 = max(ty1, 0); ty2 = min(ty2, tiles_y - 1);
  uint8_t* px = img + (size_t)s * img_stride + (size_t)y * W + x;
  const int v = *px;
  const uint8_t* L = lut + (size_t)s * tiles_x * tiles_y * 256 + v;
  const float l11 = L[(ty1 * tiles_x + tx1) * 256], l12 = L[(ty1 * tiles_x + tx2) * 256], l21 = L[(ty2 * tiles_x + tx1) * 256], l22 = L[(ty2 * tiles_x + tx2) * 256];
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  int r = __float2int_rn(res);
  *px = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
