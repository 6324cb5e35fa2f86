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

// The arithmetic of UpdateMatrices for one pixel, written with explicit roundings (no FMA contraction left to the
// compiler), so that every load schedule of k_iter_v produces the same bits.  The fusion pattern is the one nvcc chose
// for the original expression-form code.
__device__ __forceinline__ M5 um_arith(const float4 a0, const float b0, const float4 q00, const float4 q01, const float4 q10,
                                       const float4 q11, const float s00, const float s01, const float s10, const float s11,
                                       const float fx, const float fy, const float dx, const float dy, const bool inside,
                                       const bool border, int x, int y, int w, int h) {
  const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
  const float a00 = __fmul_rn(gx, gy), a01 = __fmul_rn(fx, gy), a10 = __fmul_rn(gx, fy), a11 = __fmul_rn(fx, fy);
#define OFB_BIL(c00, c01, c10, c11) __fmaf_rn(a11, c11, __fmaf_rn(a10, c10, __fmaf_rn(a00, c00, __fmul_rn(a01, c01))))
  float r2 = OFB_BIL(q00.x, q01.x, q10.x, q11.x);
  float r3 = OFB_BIL(q00.y, q01.y, q10.y, q11.y);
  float r4 = OFB_BIL(q00.z, q01.z, q10.z, q11.z);
  float r5 = OFB_BIL(q00.w, q01.w, q10.w, q11.w);
  float r6 = OFB_BIL(s00, s01, s10, s11);
#undef OFB_BIL
  if (inside) {
    r4 = __fmul_rn(__fadd_rn(a0.z, r4), 0.5f);
    r5 = __fmul_rn(__fadd_rn(a0.w, r5), 0.5f);
    r6 = __fmul_rn(__fadd_rn(b0, r6), 0.25f);
  } else {
    r2 = r3 = 0.f;
    r4 = a0.z;
    r5 = a0.w;
    r6 = __fmul_rn(b0, 0.5f);
  }
  r2 = __fmaf_rn(__fsub_rn(a0.x, r2), 0.5f, __fmaf_rn(r4, dy, __fmul_rn(r6, dx)));
  r3 = __fmaf_rn(__fsub_rn(a0.y, r3), 0.5f, __fmaf_rn(r6, dy, __fmul_rn(r5, dx)));
  if (border) {
    const float s = __fmul_rn(border_w(x, w), border_w(y, h));
    r2 = __fmul_rn(r2, s); r3 = __fmul_rn(r3, s); r4 = __fmul_rn(r4, s); r5 = __fmul_rn(r5, s); r6 = __fmul_rn(r6, s);
  }
  const float r66 = __fmul_rn(r6, r6);
  M5 m;
  m.g11 = __fmaf_rn(r4, r4, r66);
  m.g12 = __fmul_rn(__fadd_rn(r4, r5), r6);
  m.g22 = __fmaf_rn(r5, r5, r66);
  m.h1 = __fmaf_rn(r4, r2, __fmul_rn(r6, r3));
  m.h2 = __fmaf_rn(r5, r3, __fmul_rn(r6, r2));
  return m;
}

// border: this pixel lies within 5 px of the level border (attenuation table applies)
__device__ __forceinline__ M5 um_finish2(const UmLoads2& L, bool border, int x, int y, int w, int h) {
  return um_arith(L.a0, L.b0, L.q00, L.q01, L.q10, L.q11, L.s00, L.s01, L.s10, L.s11, L.fx, L.fy, L.dx, L.dy, L.inside,
                  border, x, y, w, h);
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
  return um_arith(P.a0, P.b0, top.q0, top.q1, bot.q0, bot.q1, top.s0, top.s1, bot.s0, bot.s1, P.fx, P.fy, P.dx, P.dy,
                  P.inside, border, x, y, w, h);
}

// pieces of the row-reuse gather for the two-rows-in-flight schedule (k_iter_v<..., RIF = 2, REUSE>)
__device__ __forceinline__ void um_pix(UmPix& P, const float4* __restrict__ RA0, const float* __restrict__ RB0, float2 fl,
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
  P.g = P.inside ? (unsigned)iy * w + (unsigned)ix : 0u;
}
__device__ __forceinline__ void um_row_load(UmRow& r, const float4* __restrict__ RA1, const float* __restrict__ RB1,
                                            unsigned g) {
  r.q0 = __ldg(RA1 + g);
  r.q1 = __ldg(RA1 + g + 1);
  r.s0 = __ldg(RB1 + g);
  r.s1 = __ldg(RB1 + g + 1);
}

// sums are unscaled window sums; reg = 1e-3 / scale^2 (scale = winsize^-2 folded into the regulariser)
__device__ __forceinline__ float2 solve2x2_sums(float g11, float g12, float g22, float h1, float h2, float reg) {
  const float idet = rcp_approx(g11 * g22 - g12 * g12 + reg);
  return make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
}

}  // namespace ofb
