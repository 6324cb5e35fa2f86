// fb_device.cuh — device helpers shared by the Farneback kernels (borders, resize coordinates,
// per-pixel UpdateMatrices, 2x2 solve).
#pragma once
#include "common.cuh"

// Debug builds (tools/build_variant.sh dbg "-DOFB_DBG=1"; never the shipped library): in-kernel bounds checks of the
// ring / staging / output indexing.  A failing check aborts the kernel with the file and line (device-side assert);
// tools/debug_asserts.sh runs the Farneback GPU tests against such a build.  No compute-sanitizer on the GPU pool, so this
// is the memcheck stand-in for the hand-indexed shared-memory and tensor-memory structures.
#if defined(OFB_DBG) && OFB_DBG
#include <assert.h>
#define OFB_DASSERT(cond) assert(cond)
#else
#define OFB_DASSERT(cond) do { } while (0)
#endif

namespace ofb {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}
// uint8 -> float without the XU pipe: ptxas turns (float)uint8 into I2F.U16, which runs at a
// fraction of the FP32 rate on B200 (the level-3 pyramid pass spent 80 us on it).  Exact for v < 2^23.
__device__ __forceinline__ float u8f(unsigned v) { return __uint_as_float(0x4B000000u | v) - 8388608.f; }

// BORDER_REFLECT_101 for an index that overshoots by less than n (one reflection, branch-free)
__device__ __forceinline__ int reflect101_once(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv::resize INTER_LINEAR source coordinate (resize.cpp): fx=(dx+0.5)*scale-0.5, clamp at both ends.
__device__ __forceinline__ void linear_coord(int d, double scale, int src_n, int* s0, float* w1) {
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= src_n - 1) { f = 0.f; s = src_n - 1; }
  *s0 = s;
  *w1 = f;
}

struct FrameSrc {
  const uint8_t* a;  // frames [0, na)
  const uint8_t* b;  // frames [na, ...)
  int na;
  size_t pitch, image_stride;
  __device__ __forceinline__ const uint8_t* frame(int f) const {
    return f < na ? a + (size_t)f * image_stride : b + (size_t)(f - na) * image_stride;
  }
};

// =====================================================================================
// Stage a5: FarnebackUpdateMatrices (per pixel; bilinear gather of R1 at x + flow).
// =====================================================================================
__device__ __forceinline__ float border_w(int i, int n) {
  // {0.14, 0.14, 0.4472, 0.4472, 0.4472} from each side, multiplicative
  float s = 1.f;
  if (i < 5) s *= (i < 2 ? 0.14f : 0.4472f);
  if (i >= n - 5) s *= (n - 1 - i < 2 ? 0.14f : 0.4472f);
  return s;
}

struct M5 {
  float g11, g12, g22, h1, h2;
};

// UpdateMatrices is split in two so a thread can put the loads of several pixels in flight before
// consuming any of them: um_issue() starts the R0 loads and the (always in-bounds) 2x2 gather of
// R1; um_finish() does the arithmetic.  Offsets are 32-bit (a level has < 2^31 pixels).
struct UmLoads {
  float4 a0, q00, q01, q10, q11;
  float b0, s00, s01, s10, s11;
  float fx, fy, dx, dy;
  bool inside;
};

__device__ __forceinline__ void um_issue(UmLoads& L, const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                         const float4* __restrict__ RA1, const float* __restrict__ RB1, float2 fl,
                                         int x, int y, int w, int h) {
  const int o = y * w + x;
  L.a0 = __ldg(RA0 + o);
  L.b0 = __ldg(RB0 + o);
  L.dx = fl.x;
  L.dy = fl.y;
  const float fx = (float)x + fl.x, fy = (float)y + fl.y;
  const float flx = floorf(fx), fly = floorf(fy);
  L.fx = fx - flx;
  L.fy = fy - fly;
  // cv2: (unsigned)x1 < (unsigned)(w-1) && (unsigned)y1 < (unsigned)(h-1); compared in float so that
  // huge / NaN flows stay outside.  Outside pixels gather from (0,0) (in bounds, result discarded).
  L.inside = flx >= 0.f && flx < (float)(w - 1) && fly >= 0.f && fly < (float)(h - 1);
  const int x1 = L.inside ? (int)flx : 0, y1 = L.inside ? (int)fly : 0;
  // w >= 2 and h >= 2 is validated at the API (constant +1 / +w offsets keep the address math short)
  const float4* pa = RA1 + (y1 * w + x1);
  const float* pb = RB1 + (y1 * w + x1);
  L.q00 = __ldg(pa);
  L.q01 = __ldg(pa + 1);
  L.q10 = __ldg(pa + w);
  L.q11 = __ldg(pa + w + 1);
  L.s00 = __ldg(pb);
  L.s01 = __ldg(pb + 1);
  L.s10 = __ldg(pb + w);
  L.s11 = __ldg(pb + w + 1);
}

__device__ __forceinline__ M5 um_finish(const UmLoads& L, int x, int y, int w, int h) {
  const float fx = L.fx, fy = L.fy;
  const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
  float r2 = a00 * L.q00.x + a01 * L.q01.x + a10 * L.q10.x + a11 * L.q11.x;
  float r3 = a00 * L.q00.y + a01 * L.q01.y + a10 * L.q10.y + a11 * L.q11.y;
  float r4 = a00 * L.q00.z + a01 * L.q01.z + a10 * L.q10.z + a11 * L.q11.z;
  float r5 = a00 * L.q00.w + a01 * L.q01.w + a10 * L.q10.w + a11 * L.q11.w;
  float r6 = a00 * L.s00 + a01 * L.s01 + a10 * L.s10 + a11 * L.s11;
  if (L.inside) {
    r4 = (L.a0.z + r4) * 0.5f;
    r5 = (L.a0.w + r5) * 0.5f;
    r6 = (L.b0 + r6) * 0.25f;
  } else {
    r2 = r3 = 0.f;
    r4 = L.a0.z;
    r5 = L.a0.w;
    r6 = L.b0 * 0.5f;
  }
  r2 = (L.a0.x - r2) * 0.5f;
  r3 = (L.a0.y - r3) * 0.5f;
  r2 += r4 * L.dy + r6 * L.dx;
  r3 += r6 * L.dy + r5 * L.dx;
  if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
    const float s = border_w(x, w) * border_w(y, h);
    r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
  }
  M5 m;
  m.g11 = r4 * r4 + r6 * r6;
  m.g12 = (r4 + r5) * r6;
  m.g22 = r5 * r5 + r6 * r6;
  m.h1 = r4 * r2 + r6 * r3;
  m.h2 = r6 * r2 + r5 * r3;
  return m;
}

__device__ __forceinline__ M5 update_matrix_px(const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                               const float4* __restrict__ RA1, const float* __restrict__ RB1,
                                               float2 fl, int x, int y, int w, int h) {
  UmLoads L;
  um_issue(L, RA0, RB0, RA1, RB1, fl, x, y, w, h);
  return um_finish(L, x, y, w, h);
}

__device__ __forceinline__ float2 solve2x2(float g11, float g12, float g22, float h1, float h2) {
  float idet = 1.f / (g11 * g22 - g12 * g12 + 1e-3f);
  return make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
}

}  // namespace ofb
