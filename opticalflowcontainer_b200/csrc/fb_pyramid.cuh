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

// pass H.  One thread per LEVEL column, marching over PYR_RPT source rows: the resize coordinate
// (double, as cv2 computes it), the border case and the frame pointer are worked out once per thread
// instead of once per output value (the first version spent ~160 instructions per output on them).
// grid: (ceil(w_l/128), ceil(H/PYR_RPT), frames), block 128.
constexpr int PYR_RPT = 16;

// RT > 0: radius known at compile time (taps unrolled, weights read as immediate constant-bank
// operands); RT = 0: runtime radius.  The node defaults use r = 1, 4, 9 (ksize 3, 9, 19).
template <int RT>
__global__ void __launch_bounds__(128) k_pyr_h(FrameSrc src, int W, int H, float* __restrict__ hb, int w,
                                               double sx_scale, PyrCoef pc, int sy_begin, int sy_end) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  // source rows [sy_begin, sy_end) (the whole frame, or what one rank's level rows need)
  const int y_begin = sy_begin + blockIdx.y * PYR_RPT, y_end = min(y_begin + PYR_RPT, sy_end);
  if (y_begin >= y_end) return;
  const int r = RT > 0 ? RT : pc.r;
  int sx0;
  float fx;
  if (w == W) { sx0 = x; fx = 0.f; } else linear_coord(x, sx_scale, W, &sx0, &fx);
  const float gx = 1.f - fx;
  const uint8_t* col = src.frame(blockIdx.z) + sx0;
  float* out = hb + ((size_t)blockIdx.z * H + y_begin) * w + x;
  const bool interior = sx0 - r >= 0 && sx0 + r + 1 < W;
  // A thread walks its rows one after the other and every row is a first touch of its cache lines,
  // so each row would pay a full memory round trip: pull all rows of the segment into L2 up front.
  for (int y = y_begin + 1; y < y_end; y++) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(col + (size_t)y * src.pitch));
  }
  if (interior && fx == 0.f) {
    for (int y = y_begin; y < y_end; y++, out += w) {
      const uint8_t* p = col + (size_t)y * src.pitch;
      float h0 = pc.k[0] * u8f(__ldg(p));
      if (RT > 0) {
#pragma unroll
        for (int i = 1; i <= RT; i++) h0 = fmaf(pc.k[i], u8f(__ldg(p - i)) + u8f(__ldg(p + i)), h0);
      } else {
#pragma unroll 4
        for (int i = 1; i <= r; i++) h0 = fmaf(pc.k[i], u8f(__ldg(p - i)) + u8f(__ldg(p + i)), h0);
      }
      *out = h0;
    }
  } else if (interior) {
    // window of sx0 is [-r, r], of sx0+1 is [-r+1, r+1]: one sweep over [-r, r+1]
    for (int y = y_begin; y < y_end; y++, out += w) {
      const uint8_t* p = col + (size_t)y * src.pitch;
      float h0 = pc.k[r] * u8f(__ldg(p - r)), h1 = 0.f;
      if (RT > 0) {
#pragma unroll
        for (int i = -RT + 1; i <= RT; i++) {
          const float v = u8f(__ldg(p + i));
          h0 = fmaf(pc.k[i < 0 ? -i : i], v, h0);
          h1 = fmaf(pc.k[i - 1 < 0 ? 1 - i : i - 1], v, h1);
        }
      } else {
#pragma unroll 4
        for (int i = -r + 1; i <= r; i++) {
          const float v = u8f(__ldg(p + i));
          h0 = fmaf(pc.k[abs(i)], v, h0);
          h1 = fmaf(pc.k[abs(i - 1)], v, h1);
        }
      }
      h1 = fmaf(pc.k[r], u8f(__ldg(p + r + 1)), h1);
      *out = h0 * gx + h1 * fx;
    }
  } else if (RT > 0 && RT + 1 < W) {
    // image border, compile-time radius: same unrolled sweep with one branch-free reflection per tap (the
    // rolled generic loop below made the two border warps of every row the tail of the whole kernel)
    for (int y = y_begin; y < y_end; y++, out += w) {
      const uint8_t* row = col - sx0 + (size_t)y * src.pitch;
      float h0 = 0.f, h1 = 0.f;
#pragma unroll
      for (int i = -RT; i <= RT + 1; i++) {
        const float v = u8f(__ldg(row + reflect101_once(sx0 + i, W)));
        if (i <= RT) h0 = fmaf(pc.k[i < 0 ? -i : i], v, h0);
        if (i >= -RT + 1) h1 = fmaf(pc.k[i - 1 < 0 ? 1 - i : i - 1], v, h1);
      }
      *out = fx == 0.f ? h0 : h0 * gx + h1 * fx;
    }
  } else {
    for (int y = y_begin; y < y_end; y++, out += w) {
      const uint8_t* row = col - sx0 + (size_t)y * src.pitch;
      float h0 = 0.f, h1 = 0.f;
      for (int i = -r; i <= r + 1; i++) {
        const float v = u8f(__ldg(row + reflect101(sx0 + i, W)));
        if (i <= r) h0 = fmaf(pc.k[abs(i)], v, h0);
        if (i >= -r + 1) h1 = fmaf(pc.k[abs(i - 1)], v, h1);
      }
      *out = fx == 0.f ? h0 : h0 * gx + h1 * fx;
    }
  }
}

// pass V.  grid: (ceil(w_l/128), ceil(h_l/2), frames), block (128, 2).  The tap loops are unrolled by
// four so that four loads are in flight per thread (the rolled loop was latency-bound).
template <int RT>
__global__ void __launch_bounds__(256) k_pyr_v(const float* __restrict__ hb, int H, float* __restrict__ out, int w,
                                               int h, double sy_scale, PyrCoef pc, int y_begin, int y_end) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = y_begin + blockIdx.y * blockDim.y + threadIdx.y;   // level rows [y_begin, y_end)
  if (x >= w || y >= y_end) return;
  const float* col = hb + (size_t)blockIdx.z * H * w + x;
  const int r = RT > 0 ? RT : pc.r;
  int sy0;
  float fy;
  if (h == H) { sy0 = y; fy = 0.f; } else linear_coord(y, sy_scale, H, &sy0, &fy);
  const bool interior = sy0 - r >= 0 && sy0 + r + 1 < H;
  float v0 = 0.f, v1 = 0.f;
  if (fy == 0.f) {
    v0 = pc.k[0] * __ldg(col + (size_t)sy0 * w);
    if (interior) {
      const float* c = col + (size_t)sy0 * w;
#pragma unroll 4
      for (int t = 1; t <= r; t++) v0 = fmaf(pc.k[t], __ldg(c - (size_t)t * w) + __ldg(c + (size_t)t * w), v0);
    } else {
      for (int t = 1; t <= r; t++)
        v0 = fmaf(pc.k[t],
                  __ldg(col + (size_t)reflect101(sy0 - t, H) * w) + __ldg(col + (size_t)reflect101(sy0 + t, H) * w), v0);
    }
    out[((size_t)blockIdx.z * h + y) * w + x] = v0;
    return;
  }
  if (interior) {
    const float* c = col + (size_t)sy0 * w;
    v0 = pc.k[r] * __ldg(c - (size_t)r * w);
    if (RT > 0) {
#pragma unroll
      for (int t = -RT + 1; t <= RT; t++) {
        const float v = __ldg(c + (ptrdiff_t)t * w);
        v0 = fmaf(pc.k[t < 0 ? -t : t], v, v0);
        v1 = fmaf(pc.k[t - 1 < 0 ? 1 - t : t - 1], v, v1);
      }
    } else {
#pragma unroll 4
      for (int t = -r + 1; t <= r; t++) {
        const float v = __ldg(c + (ptrdiff_t)t * w);
        v0 = fmaf(pc.k[abs(t)], v, v0);
        v1 = fmaf(pc.k[abs(t - 1)], v, v1);
      }
    }
    v1 = fmaf(pc.k[r], __ldg(c + (size_t)(r + 1) * w), v1);
  } else if (RT > 0 && RT + 1 < H) {
#pragma unroll
    for (int t = -RT; t <= RT + 1; t++) {
      const float v = __ldg(col + (size_t)reflect101_once(sy0 + t, H) * w);
      if (t <= RT) v0 = fmaf(pc.k[t < 0 ? -t : t], v, v0);
      if (t >= -RT + 1) v1 = fmaf(pc.k[t - 1 < 0 ? 1 - t : t - 1], v, v1);
    }
  } else {
    for (int t = -r; t <= r + 1; t++) {
      const float v = __ldg(col + (size_t)reflect101(sy0 + t, H) * w);
      if (t <= r) v0 = fmaf(pc.k[abs(t)], v, v0);
      if (t >= -r + 1) v1 = fmaf(pc.k[abs(t - 1)], v, v1);
    }
  }
  out[((size_t)blockIdx.z * h + y) * w + x] = v0 * (1.f - fy) + v1 * fy;
}

}  // namespace ofb
