// fb_um.cuh — lean per-pixel UpdateMatrices (stage a5) for the fused iteration kernel k_iter_v:
// 32-bit unsigned element offsets from per-pair base pointers, floor via F2I + I2FP, cv2's unsigned
// inside test, border attenuation only on border pixels; the 2x2 solve on unscaled window sums
// (winsize^-2 folded into the regulariser, approximate reciprocal, 1 ulp); L2 prefetch helper.
#pragma once
#include "fb_device.cuh"

namespace ofb {

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct UmLoads2 {
  float4 a0, q00, q01, q10, q11;
  float b0, s00, s01, s10, s11;
  float fx, fy, dx, dy;
  bool inside;
};

// Starts the loads of one pixel: R0 at (x, y) and the 2x2 neighbourhood of R1 at floor((x,y) + flow).
// yw = y * w.  Outside pixels gather from (0,0) (in bounds, result discarded).
__device__ __forceinline__ void um_issue2(UmLoads2& L, const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                          const float4* __restrict__ RA1, const float* __restrict__ RB1, float2 fl,
                                          int x, int y, unsigned yw, unsigned w, unsigned h) {
  const unsigned o = yw + (unsigned)x;
  L.a0 = __ldg(RA0 + o);
  L.b0 = __ldg(RB0 + o);
  L.dx = fl.x;
  L.dy = fl.y;
  const float fx = (float)x + fl.x, fy = (float)y + fl.y;
  const int ix = __float2int_rd(fx), iy = __float2int_rd(fy);
  L.fx = fx - (float)ix;
  L.fy = fy - (float)iy;
  // cv2: (unsigned)x1 < (unsigned)(w-1) && (unsigned)y1 < (unsigned)(h-1)  (F2I saturates, so huge
  // flows stay outside)
  L.inside = (unsigned)ix < w - 1u && (unsigned)iy < h - 1u;
  const unsigned g = L.inside ? (unsigned)iy * w + (unsigned)ix : 0u;
  const float4* pa = RA1 + g;
  const float* pb = RB1 + g;
  L.q00 = __ldg(pa);
  L.q01 = __ldg(pa + 1);
  L.q10 = __ldg(pa + w);
  L.q11 = __ldg(pa + w + 1);
  L.s00 = __ldg(pb);
  L.s01 = __ldg(pb + 1);
  L.s10 = __ldg(pb + w);
  L.s11 = __ldg(pb + w + 1);
}

// Tiled variant: R0 and the two corner rows of the R1 gather may live on different ranks.  RA0/RB0 are
// the (frame-0) bases of the rank that owns row y; the corner rows pick their owner per lane.
__device__ __forceinline__ int tile_owner(int y, const PeerTab& t) { return min(y / t.rpr, t.world - 1); }

__device__ __forceinline__ void um_issue2_tiled(UmLoads2& L, const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                                const PeerTab& t, size_t f1_elems, int my_rank, float2 fl, int x, int y,
                                                unsigned yw, unsigned w, unsigned h) {
  const unsigned o = yw + (unsigned)x;
  L.a0 = __ldg(RA0 + o);
  L.b0 = __ldg(RB0 + o);
  L.dx = fl.x;
  L.dy = fl.y;
  const float fx = (float)x + fl.x, fy = (float)y + fl.y;
  const int ix = __float2int_rd(fx), iy = __float2int_rd(fy);
  L.fx = fx - (float)ix;
  L.fy = fy - (float)iy;
  L.inside = (unsigned)ix < w - 1u && (unsigned)iy < h - 1u;
  // outside pixels gather (and discard) from the start of their OWN row: always present locally —
  // row 0, as in the untiled kernel, would be a remote NVLink load on every rank but the first, and the
  // clamped columns of the last strip are all "outside": their CTAs set the kernel time (measured 1.6x)
  const int ry = L.inside ? iy : y;
  const unsigned g = L.inside ? (unsigned)iy * w + (unsigned)ix : yw;
  // where the two corner rows live: in this rank's buffer (own rows + pulled halo) in the common case,
  // one division and a remote NVLink load otherwise (displacement beyond the halo)
  const int r_top = (ry >= t.r_lo && ry < t.r_hi) ? my_rank : tile_owner(ry, t);
  const int r_bot = (ry + 1 >= t.r_lo && ry + 1 < t.r_hi) ? my_rank : tile_owner(ry + 1, t);
  const float4* pa = t.RA[r_top] + f1_elems + g;
  const float* pb = t.RB[r_top] + f1_elems + g;
  const float4* qa = t.RA[r_bot] + f1_elems + g + w;
  const float* qb = t.RB[r_bot] + f1_elems + g + w;
  L.q00 = __ldg(pa);
  L.q01 = __ldg(pa + 1);
  L.q10 = __ldg(qa);
  L.q11 = __ldg(qa + 1);
  L.s00 = __ldg(pb);
  L.s01 = __ldg(pb + 1);
  L.s10 = __ldg(qb);
  L.s11 = __ldg(qb + 1);
}

// border: this pixel lies within 5 px of the level border (attenuation table applies)
__device__ __forceinline__ M5 um_finish2(const UmLoads2& L, bool border, int x, int y, int w, int h) {
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
  if (border) {
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

// ---- row-reuse gather (k_iter_v<..., REUSE>) -------------------------------------------------------
// A producer thread walks down one column.  Where the displacement field is smooth, the 2x2 neighbourhood
// of row y+1 sits exactly one row below that of row y (same floor(x+dx), floor(y+dy) one larger): its top
// corner row IS the previous pixel's bottom corner row, still in registers.  Only the new bottom row is
// loaded then (2 x LDG.128 + 2 x LDG.32 instead of 4 + 4: the gather was the largest share of the kernel's
// L1 wavefronts).  Any other case (flow discontinuity, floor crossing, outside pixel, clamped warm-up rows)
// loads both rows — values are bitwise the ones the full gather reads.
struct UmRow {      // two horizontally adjacent R1 pixels
  float4 q0, q1;
  float s0, s1;
};
struct UmPix {
  float4 a0;
  float b0, fx, fy, dx, dy;
  unsigned g;       // element offset of the top-left corner (0 when outside)
  bool inside;
};

// prev_g: corner offset of the previous row's pixel (or an impossible value); `top` must still hold that
// pixel's bottom row when prev_g + w == g.
__device__ __forceinline__ void um_issue_rows(UmPix& P, UmRow& top, UmRow& bot, unsigned& prev_g,
                                              const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                              const float4* __restrict__ RA1, const float* __restrict__ RB1, float2 fl,
                                              int x, int y, unsigned yw, unsigned w, unsigned h) {
  const unsigned o = yw + (unsigned)x;
  P.a0 = __ldg(RA0 + o);
  P.b0 = __ldg(RB0 + o);
  P.dx = fl.x;
  P.dy = fl.y;
  const float fx = (float)x + fl.x, fy = (float)y + fl.y;
  const int ix = __float2int_rd(fx), iy = __float2int_rd(fy);
  P.fx = fx - (float)ix;
  P.fy = fy - (float)iy;
  P.inside = (unsigned)ix < w - 1u && (unsigned)iy < h - 1u;
  const unsigned g = P.inside ? (unsigned)iy * w + (unsigned)ix : 0u;
  P.g = g;
  const float4* pa = RA1 + g;
  const float* pb = RB1 + g;
  if (g != prev_g + w) {     // (prev_g = ~0u - w after an outside pixel: never equal, g < 2^31)
    top.q0 = __ldg(pa);
    top.q1 = __ldg(pa + 1);
    top.s0 = __ldg(pb);
    top.s1 = __ldg(pb + 1);
  }
  bot.q0 = __ldg(pa + w);
  bot.q1 = __ldg(pa + w + 1);
  bot.s0 = __ldg(pb + w);
  bot.s1 = __ldg(pb + w + 1);
  prev_g = P.inside ? g : ~0u - w;
}

__device__ __forceinline__ M5 um_finish_rows(const UmPix& P, const UmRow& top, const UmRow& bot, bool border, int x,
                                             int y, int w, int h) {
  const float fx = P.fx, fy = P.fy;
  const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
  float r2 = a00 * top.q0.x + a01 * top.q1.x + a10 * bot.q0.x + a11 * bot.q1.x;
  float r3 = a00 * top.q0.y + a01 * top.q1.y + a10 * bot.q0.y + a11 * bot.q1.y;
  float r4 = a00 * top.q0.z + a01 * top.q1.z + a10 * bot.q0.z + a11 * bot.q1.z;
  float r5 = a00 * top.q0.w + a01 * top.q1.w + a10 * bot.q0.w + a11 * bot.q1.w;
  float r6 = a00 * top.s0 + a01 * top.s1 + a10 * bot.s0 + a11 * bot.s1;
  if (P.inside) {
    r4 = (P.a0.z + r4) * 0.5f;
    r5 = (P.a0.w + r5) * 0.5f;
    r6 = (P.b0 + r6) * 0.25f;
  } else {
    r2 = r3 = 0.f;
    r4 = P.a0.z;
    r5 = P.a0.w;
    r6 = P.b0 * 0.5f;
  }
  r2 = (P.a0.x - r2) * 0.5f;
  r3 = (P.a0.y - r3) * 0.5f;
  r2 += r4 * P.dy + r6 * P.dx;
  r3 += r6 * P.dy + r5 * P.dx;
  if (border) {
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

// sums are unscaled window sums; reg = 1e-3 / scale^2 (scale = winsize^-2 folded into the regulariser)
__device__ __forceinline__ float2 solve2x2_sums(float g11, float g12, float g22, float h1, float h2, float reg) {
  const float idet = rcp_approx(g11 * g22 - g12 * g12 + reg);
  return make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
}

}  // namespace ofb
