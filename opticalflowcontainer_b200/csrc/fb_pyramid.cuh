// fb_pyramid.cuh — Farneback pyramid stage (a2): convertTo(f32) + GaussianBlur(REFLECT_101) +
// resize(INTER_LINEAR) of the full-resolution u8 frame to one level, as two separable passes.
//
// The bilinear resize only reads the blurred image at (up to) 2 source columns and 2 source rows
// per level pixel, and blur and resize are both separable, so:
//   pass H: for every SOURCE row y and every LEVEL column x:  hb[y][x] = lerp_x( Gh * src[y] )
//   pass V: for every level pixel (x, y):                     I[y][x]  = lerp_y( Gv * hb[:, x] )
// Work per level is O(H * w_l * ksize) + O(h_l * w_l * ksize) instead of the full-resolution blur.
#pragma once
#include "fb_device.cuh"

namespace ofb {

// cv::getGaussianKernel(ksize, sigma, CV_32F) — half kernel k[0..r].
constexpr int kMaxPyrRadius = 159;
struct PyrCoef {
  int r;
  float k[kMaxPyrRadius + 1];
};

// pass H.  grid: (ceil(w_l/128), H, frames), block 128.
__global__ void __launch_bounds__(128) k_pyr_h(FrameSrc src, int W, int H, float* __restrict__ hb, int w,
                                               double sx_scale, PyrCoef pc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* row = src.frame(blockIdx.z) + (size_t)y * src.pitch;
  const int r = pc.r;
  int sx0;
  float fx;
  if (w == W) { sx0 = x; fx = 0.f; } else linear_coord(x, sx_scale, W, &sx0, &fx);
  float h0 = 0.f, h1 = 0.f;
  if (sx0 - r >= 0 && sx0 + r + 1 < W) {  // interior: no border handling
    const uint8_t* p = row + sx0;
    if (fx == 0.f) {
      h0 = pc.k[0] * (float)__ldg(p);
      for (int i = 1; i <= r; i++) h0 = fmaf(pc.k[i], (float)__ldg(p - i) + (float)__ldg(p + i), h0);
    } else {
      // window of sx0 is [-r, r], of sx0+1 is [-r+1, r+1]: one sweep over [-r, r+1]
      h0 = pc.k[r] * (float)__ldg(p - r);
      for (int i = -r + 1; i <= r; i++) {
        const float v = (float)__ldg(p + i);
        h0 = fmaf(pc.k[abs(i)], v, h0);
        h1 = fmaf(pc.k[abs(i - 1)], v, h1);
      }
      h1 = fmaf(pc.k[r], (float)__ldg(p + r + 1), h1);
    }
  } else {
    for (int i = -r; i <= r + 1; i++) {
      const float v = (float)__ldg(row + reflect101(sx0 + i, W));
      if (i <= r) h0 = fmaf(pc.k[abs(i)], v, h0);
      if (i >= -r + 1) h1 = fmaf(pc.k[abs(i - 1)], v, h1);
    }
  }
  hb[((size_t)blockIdx.z * H + y) * w + x] = fx == 0.f ? h0 : h0 * (1.f - fx) + h1 * fx;
}

// pass V.  grid: (ceil(w_l/128), ceil(h_l/2), frames), block (128, 2).
__global__ void __launch_bounds__(256) k_pyr_v(const float* __restrict__ hb, int H, float* __restrict__ out, int w,
                                               int h, double sy_scale, PyrCoef pc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const float* col = hb + (size_t)blockIdx.z * H * w + x;
  const int r = pc.r;
  int sy0;
  float fy;
  if (h == H) { sy0 = y; fy = 0.f; } else linear_coord(y, sy_scale, H, &sy0, &fy);
  float v0 = 0.f, v1 = 0.f;
  if (fy == 0.f) {
    v0 = pc.k[0] * __ldg(col + (size_t)sy0 * w);
    for (int t = 1; t <= r; t++)
      v0 = fmaf(pc.k[t], __ldg(col + (size_t)reflect101(sy0 - t, H) * w) + __ldg(col + (size_t)reflect101(sy0 + t, H) * w),
                v0);
    out[((size_t)blockIdx.z * h + y) * w + x] = v0;
    return;
  }
  for (int t = -r; t <= r + 1; t++) {
    const float v = __ldg(col + (size_t)reflect101(sy0 + t, H) * w);
    if (t <= r) v0 = fmaf(pc.k[abs(t)], v, v0);
    if (t >= -r + 1) v1 = fmaf(pc.k[abs(t - 1)], v, v1);
  }
  out[((size_t)blockIdx.z * h + y) * w + x] = v0 * (1.f - fy) + v1 * fy;
}

}  // namespace ofb
