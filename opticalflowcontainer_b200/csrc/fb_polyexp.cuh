// fb_polyexp.cuh — FarnebackPolyExp (stage a4) as a marching column-strip kernel, with the
// level-0 pyramid stage (a2: convertTo f32 + 3-tap GaussianBlur, REFLECT_101) fused in.
//
// A CTA of 256 threads owns a strip of 256 level columns (8 halo columns on each side, 240 useful,
// so that every 4-pixel group is 16-byte aligned in HBM) and a segment of rows, and marches down it
// PX_ROWS = 4 rows per step:
//   V  one thread per column keeps the 2n+4 level-image values it needs in registers (sliding
//      window; each step brings in 4 new rows) and computes the three vertical moments of 4 rows
//      -> shared memory (double-buffered, one barrier per step);
//   H  thread (row, quad) reads the 4+2n moment values of its 4 adjacent pixels as float4s and
//      produces the 5 polynomial coefficients per pixel: 4 x float4 (RA) + 1 x float4 (RB, 4 px).
// SRC = 1 reads the uint8 source frame and applies the 3-tap blur on the fly (level k = 0, where the
// level image has the source size): the level image never goes through HBM.  HBM traffic per level
// pixel: 1 B (u8) or 4 B (f32 level image) in, 20 B out.
// No vertical halo recompute except the 2n warm-up rows of a segment.
#pragma once
#include "fb_device.cuh"
#include "fb_pyramid.cuh"

namespace ofb {

constexpr int PX_COLS = 256;   // strip width = threads per CTA
constexpr int PX_HALO = 8;     // halo columns per side (>= poly_n, multiple of 4)
constexpr int PX_TW = PX_COLS - 2 * PX_HALO;
constexpr int PX_ROWS = 4;     // rows per step
constexpr int PX_MAXN = 8;     // largest poly_n served by this kernel
#ifndef PX_MINB
#define PX_MINB 3        // resident CTAs per SM the register budget is held to (85 registers/thread)
#endif

// Level-image value feed of one thread (= one column xc, rows visited in increasing order).  Loading
// and consuming a row are separate steps so the kernel can issue the loads of the NEXT step right
// after it has consumed the raw values of the current one: the loads then have a whole step of
// arithmetic to land (ncu on the first version: long-scoreboard stalls on the u8 loads dominated).
template <int SRC>
struct LevelColumn;

// SRC = 0: level image in HBM (levels k >= 1, written by k_pyr_h / k_pyr_v).
template <>
struct LevelColumn<0> {
  typedef float Raw;
  const float* col;
  int w, h;
  __device__ __forceinline__ void init(const float* img, int xc, int w_, int h_) {
    col = img + xc;
    w = w_;
    h = h_;
  }
  __device__ __forceinline__ void start(int) {}
  __device__ __forceinline__ Raw load(int t) const { return __ldg(col + (size_t)clampi(t, 0, h - 1) * w); }
  __device__ __forceinline__ float consume(int, Raw r) { return r; }
  __device__ __forceinline__ bool edge() const { return false; }
  __device__ __forceinline__ Raw load_fast(int t) const { return __ldg(col + (size_t)t * w); }   // 0 <= t < h
  __device__ __forceinline__ float consume_fast(int, Raw r) { return r; }
};

// SRC = 1: uint8 source frame of the level's size; I = Gv * (Gh * float(src)) with the fixed 3-tap kernel
// [1/4, 1/2, 1/4] of the k = 0 level (sigma = 0), REFLECT_101.  Every product and sum of that blur is exact in float
// (u8 values, weights 2^-1 / 2^-2), so it is computed in INTEGERS — h = a + 2b + c per row, I = (h- + 2 h0 + h+) / 16
// — and converted once per pixel by dropping the 12-bit sum into the mantissa of 2^19 (whose last mantissa bit is worth
// 2^-4): bit-identical to cv2's float arithmetic at 6 ALU instructions per pixel instead of 15.  Horizontally summed
// rows tc-1, tc, tc+1 are kept while the row index advances by 0 or 1 per call; the only new data a row needs is the
// source row below it.
__device__ __forceinline__ unsigned ldg_u8(const uint8_t* p) {      // zero-extended into a 32-bit register: no mask
  unsigned v;
  asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <>
struct LevelColumn<1> {
  struct Raw { unsigned a, b, c; };
  const uint8_t* base;
  size_t pitch;
  int xl, xc, xr, h;
  int tc;            // row the state is centred on
  unsigned hm, h0, hp;
  float cur;
  __device__ __forceinline__ void init(const uint8_t* frame, size_t pitch_, int xc_, int w, int h_, float, float) {
    base = frame;
    pitch = pitch_;
    xc = xc_;
    xl = reflect101(xc_ - 1, w);
    xr = reflect101(xc_ + 1, w);
    h = h_;
  }
  __device__ __forceinline__ Raw raw_row(int s) const {
    const uint8_t* p = base + (size_t)s * pitch;
    Raw r;
    r.a = ldg_u8(p + xl);
    r.b = ldg_u8(p + xc);
    r.c = ldg_u8(p + xr);
    return r;
  }
  __device__ __forceinline__ unsigned hval(Raw r) const { return r.a + 2u * r.b + r.c; }
  __device__ __forceinline__ float blur() const {
    return __uint_as_float(0x49000000u | (hm + 2u * h0 + hp)) - 524288.f;     // (sum <= 4080) / 16, exact
  }
  // centre the state on the first row to be consumed (direct loads, once per segment)
  __device__ __forceinline__ void start(int t) {
    const int n = clampi(t, 0, h - 1);
    h0 = hval(raw_row(n));
    hm = hval(raw_row(reflect101(n - 1, h)));
    hp = hval(raw_row(reflect101(n + 1, h)));
    tc = n;
    cur = blur();
  }
  __device__ __forceinline__ Raw load(int t) const { return raw_row(reflect101_once(clampi(t, 0, h - 1) + 1, h)); }
  // Fast path for interior rows (1 <= t, t + 1 <= h - 1) of interior columns: no clamp, no reflection,
  // one address computation per row (the generic path spent ~60 instructions per pixel on them).
  __device__ __forceinline__ bool edge() const { return xl != xc - 1 || xr != xc + 1; }
  __device__ __forceinline__ Raw load_fast(int t) const {
    const uint8_t* p = base + (size_t)(t + 1) * pitch + xc;
    Raw r;
    r.a = ldg_u8(p - 1);
    r.b = ldg_u8(p);
    r.c = ldg_u8(p + 1);
    return r;
  }
  __device__ __forceinline__ float consume_fast(int t, Raw r) {   // t == tc + 1
    hm = h0;
    h0 = hp;
    hp = hval(r);
    tc = t;
    cur = blur();
    return cur;
  }
  __device__ __forceinline__ float consume(int t, Raw r) {
    const int n = clampi(t, 0, h - 1);
    if (n != tc) {   // n == tc + 1; uniform over the CTA (depends on t only)
      hm = h0;
      h0 = hp;
      hp = hval(r);
      tc = n;
      cur = blur();
    }
    return cur;
  }
};

template <int NT, int SRC>
__global__ void __launch_bounds__(PX_COLS, PX_MINB)
    k_polyexp_march(const float* __restrict__ I, FrameSrc src, float k0, float k1, float4* __restrict__ RA,
                    float* __restrict__ RB, int w, int h, int seg_rows, int strips, PolyCoef pc, int y_begin,
                    int y_end) {
  constexpr int NMAX = NT > 0 ? NT : PX_MAXN;
  const int n = NT > 0 ? NT : pc.n;
  constexpr int WIN = 2 * NMAX + PX_ROWS;
  // vertical moments of the step, rows interleaved in pairs: sV[buf][row pair][moment][column] = (row 2p, row 2p+1),
  // so the horizontal phase works on two rows at once with packed f32x2 arithmetic (one issue slot, two FMAs)
  __shared__ __align__(16) float2 sV[2][PX_ROWS / 2][3][PX_COLS];

  const int tid = threadIdx.x;
  const int strip = blockIdx.x % strips, seg = blockIdx.x / strips;
  const int frame = blockIdx.y;
  const int x_base = strip * PX_TW - PX_HALO;
  const int y0 = y_begin + seg * seg_rows, y1 = min(y0 + seg_rows, y_end);   // rows [y_begin, y_end) of the level
  const int xc = clampi(x_base + tid, 0, w - 1);   // replicate border of PolyExp
  const size_t fbase = (size_t)frame * w * h;

  LevelColumn<SRC> feed;
  if constexpr (SRC == 0) feed.init(I + fbase, xc, w, h);
  else feed.init(src.frame(frame), src.pitch, xc, w, h, k0, k1);

  typedef typename LevelColumn<SRC>::Raw Raw;
  // window rows: win[i] = I(row ys - NMAX + i) for the step starting at output row ys
  // (generic n < NMAX: the outer rows are loaded but never used)
  float win[WIN];
  feed.start(y0 - NMAX);
  {
    Raw rw[2 * NMAX];
#pragma unroll
    for (int i = 0; i < 2 * NMAX; i++) rw[i] = feed.load(y0 - NMAX + i);
#pragma unroll
    for (int i = 0; i < 2 * NMAX; i++) win[i] = feed.consume(y0 - NMAX + i, rw[i]);
  }
  Raw nx[PX_ROWS];   // raw values of the rows entering the window in the next step (loads in flight)
#pragma unroll
  for (int r = 0; r < PX_ROWS; r++) nx[r] = feed.load(y0 + NMAX + r);

  // H-phase role: 2 adjacent pixels of 2 adjacent rows
  const int hp = tid >> 7;            // row pair within the step
  const int q0 = (tid & 127) * 2;     // first of 2 adjacent strip columns (even)
  const int hx = x_base + q0;         // image x of that column (even)
  const bool h_valid = q0 >= PX_HALO && q0 < PX_COLS - PX_HALO && hx < w;

  int buf = 0;
  for (int ys = y0; ys < y1; ys += PX_ROWS) {
    // ---------------- V: 4 new rows, vertical moments of rows ys .. ys+3
    // rows entering now: t0 .. t0+3; rows to load for the next step: t0+4 .. t0+7
    const int t0 = ys + NMAX;
    if (t0 >= 2 && t0 + PX_ROWS <= h - 1 && !feed.edge()) {       // (t0 - 1 was consumed un-clamped: tc == t0 - 1)
#pragma unroll
      for (int r = 0; r < PX_ROWS; r++) win[2 * NMAX + r] = feed.consume_fast(t0 + r, nx[r]);
    } else {
#pragma unroll
      for (int r = 0; r < PX_ROWS; r++) win[2 * NMAX + r] = feed.consume(t0 + r, nx[r]);
    }
    if (t0 + PX_ROWS >= 1 && t0 + 2 * PX_ROWS <= h - 1 && !feed.edge()) {
#pragma unroll
      for (int r = 0; r < PX_ROWS; r++) nx[r] = feed.load_fast(t0 + PX_ROWS + r);
    } else {
#pragma unroll
      for (int r = 0; r < PX_ROWS; r++) nx[r] = feed.load(t0 + PX_ROWS + r);   // rows clamp: always valid
    }
#pragma unroll
    for (int rp = 0; rp < PX_ROWS / 2; rp++) {
      float m0[2], m1[2], m2[2];
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
        const float* c = win + NMAX + 2 * rp + rr;   // centre of row ys + 2*rp + rr
        float r0 = c[0] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int k = 1; k <= NMAX; k++) {
          if (k <= n) {
            const float a = c[-k], b = c[k];
            const float p = a + b;
            r0 = fmaf(pc.g[k], p, r0);
            r1 = fmaf(pc.xg[k], b - a, r1);
            r2 = fmaf(pc.xxg[k], p, r2);
          }
        }
        m0[rr] = r0; m1[rr] = r1; m2[rr] = r2;
      }
      sV[buf][rp][0][tid] = make_float2(m0[0], m0[1]);
      sV[buf][rp][1][tid] = make_float2(m1[0], m1[1]);
      sV[buf][rp][2][tid] = make_float2(m2[0], m2[1]);
    }
#pragma unroll
    for (int i = 0; i < 2 * NMAX; i++) win[i] = win[i + PX_ROWS];
    __syncthreads();
    // ---------------- H: horizontal moments of 2 adjacent pixels of rows ys + 2*hp, ys + 2*hp + 1 (packed pairs)
    const int y = ys + 2 * hp;
    if (h_valid && y < y1) {
      constexpr int PAD = (NMAX + 1) / 2 * 2;        // even >= n: window columns q0 - PAD .. q0 + 1 + PAD
      constexpr int NE = 2 * PAD + 2;
      const float2 neg1 = make_float2(-1.f, -1.f);
      float2 ox[2], oy[2], oz[2], ow[2], ob[2];      // per pixel: (row y, row y + 1)
      {
        float2 e[NE];   // e[i] = r0 at strip column q0 - PAD + i, rows (y, y + 1)
#pragma unroll
        for (int v = 0; v < NE / 2; v++) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[buf][hp][0][q0 - PAD + 2 * v]);
          e[2 * v] = make_float2(t.x, t.y); e[2 * v + 1] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int c = PAD + j;
          float2 b1 = __fmul2_rn(e[c], pc.g2[0]), b2 = make_float2(0.f, 0.f), b4 = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 1; k <= NMAX; k++) {
            if (k <= n) {
              const float2 tg = __fadd2_rn(e[c + k], e[c - k]);
              const float2 td = __ffma2_rn(e[c - k], neg1, e[c + k]);
              b1 = __ffma2_rn(tg, pc.g2[k], b1);
              b4 = __ffma2_rn(tg, pc.xxg2[k], b4);
              b2 = __ffma2_rn(td, pc.xg2[k], b2);
            }
          }
          oy[j] = __fmul2_rn(b2, pc.ig11_2);
          oz[j] = __fmul2_rn(b1, pc.ig03_2);          // + b5 * ig33 below
          ow[j] = __ffma2_rn(b4, pc.ig33_2, oz[j]);
        }
      }
      {
        float2 e[NE];   // r1
#pragma unroll
        for (int v = 0; v < NE / 2; v++) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[buf][hp][1][q0 - PAD + 2 * v]);
          e[2 * v] = make_float2(t.x, t.y); e[2 * v + 1] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int c = PAD + j;
          float2 b3 = __fmul2_rn(e[c], pc.g2[0]), b6 = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 1; k <= NMAX; k++) {
            if (k <= n) {
              b3 = __ffma2_rn(__fadd2_rn(e[c + k], e[c - k]), pc.g2[k], b3);
              b6 = __ffma2_rn(__ffma2_rn(e[c - k], neg1, e[c + k]), pc.xg2[k], b6);
            }
          }
          ox[j] = __fmul2_rn(b3, pc.ig11_2);
          ob[j] = __fmul2_rn(b6, pc.ig55_2);
        }
      }
      {
        float2 e[NE];   // r2
#pragma unroll
        for (int v = 0; v < NE / 2; v++) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[buf][hp][2][q0 - PAD + 2 * v]);
          e[2 * v] = make_float2(t.x, t.y); e[2 * v + 1] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int c = PAD + j;
          float2 b5 = __fmul2_rn(e[c], pc.g2[0]);
#pragma unroll
          for (int k = 1; k <= NMAX; k++) {
            if (k <= n) b5 = __ffma2_rn(__fadd2_rn(e[c + k], e[c - k]), pc.g2[k], b5);
          }
          oz[j] = __ffma2_rn(b5, pc.ig33_2, oz[j]);
        }
      }
      const size_t ob0 = fbase + (size_t)y * w + hx;
      const bool two = hx + 1 < w;
      const bool vec = two && ((w & 1) == 0) && (reinterpret_cast<uintptr_t>(RA) & 31) == 0 && (reinterpret_cast<uintptr_t>(RB) & 7) == 0;   // hx is even: pixel pairs are aligned when w is even
      // (w even, hx even: the two pixels' float4 are 32 contiguous, 32-byte aligned bytes: one 256-bit store each row)
      if (vec) {
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(RA + ob0), "f"(ox[0].x), "f"(oy[0].x),
                     "f"(oz[0].x), "f"(ow[0].x), "f"(ox[1].x), "f"(oy[1].x), "f"(oz[1].x), "f"(ow[1].x) : "memory");
      } else {
        RA[ob0] = make_float4(ox[0].x, oy[0].x, oz[0].x, ow[0].x);
        if (two) RA[ob0 + 1] = make_float4(ox[1].x, oy[1].x, oz[1].x, ow[1].x);
      }
      if (vec) *reinterpret_cast<float2*>(RB + ob0) = make_float2(ob[0].x, ob[1].x);
      else { RB[ob0] = ob[0].x; if (two) RB[ob0 + 1] = ob[1].x; }
      if (y + 1 < y1) {
        if (vec) {
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(RA + ob0 + w), "f"(ox[0].y),
                       "f"(oy[0].y), "f"(oz[0].y), "f"(ow[0].y), "f"(ox[1].y), "f"(oy[1].y), "f"(oz[1].y), "f"(ow[1].y) : "memory");
        } else {
          RA[ob0 + w] = make_float4(ox[0].y, oy[0].y, oz[0].y, ow[0].y);
          if (two) RA[ob0 + w + 1] = make_float4(ox[1].y, oy[1].y, oz[1].y, ow[1].y);
        }
        if (vec) *reinterpret_cast<float2*>(RB + ob0 + w) = make_float2(ob[0].y, ob[1].y);
        else { RB[ob0 + w] = ob[0].y; if (two) RB[ob0 + w + 1] = ob[1].y; }
      }
    }
    buf ^= 1;
  }
}

}  // namespace ofb
