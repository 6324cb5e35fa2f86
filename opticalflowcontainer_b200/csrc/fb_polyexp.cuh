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

// SRC = 1: uint8 source frame of the level's size; I = Gv * (Gh * float(src)), 3 taps, REFLECT_101.
// Horizontally blurred rows tc-1, tc, tc+1 are kept while the row index advances by 0 or 1 per call;
// the only new data a row needs is the source row below it.
template <>
struct LevelColumn<1> {
  struct Raw { unsigned a, b, c; };
  const uint8_t* base;
  size_t pitch;
  int xl, xc, xr, h;
  int tc;            // row the state is centred on
  float k0, k1, hm, h0, hp, cur;
  __device__ __forceinline__ void init(const uint8_t* frame, size_t pitch_, int xc_, int w, int h_, float k0_,
                                       float k1_) {
    base = frame;
    pitch = pitch_;
    xc = xc_;
    xl = reflect101(xc_ - 1, w);
    xr = reflect101(xc_ + 1, w);
    h = h_;
    k0 = k0_;
    k1 = k1_;
  }
  __device__ __forceinline__ Raw raw_row(int s) const {
    const uint8_t* p = base + (size_t)s * pitch;
    Raw r;
    r.a = __ldg(p + xl);
    r.b = __ldg(p + xc);
    r.c = __ldg(p + xr);
    return r;
  }
  __device__ __forceinline__ float hval(Raw r) const { return fmaf(k1, u8f(r.a) + u8f(r.c), k0 * u8f(r.b)); }
  // centre the state on the first row to be consumed (direct loads, once per segment)
  __device__ __forceinline__ void start(int t) {
    const int n = clampi(t, 0, h - 1);
    h0 = hval(raw_row(n));
    hm = hval(raw_row(reflect101(n - 1, h)));
    hp = hval(raw_row(reflect101(n + 1, h)));
    tc = n;
    cur = fmaf(k1, hm + hp, k0 * h0);
  }
  __device__ __forceinline__ Raw load(int t) const { return raw_row(reflect101_once(clampi(t, 0, h - 1) + 1, h)); }
  // Fast path for interior rows (1 <= t, t + 1 <= h - 1) of interior columns: no clamp, no reflection,
  // one address computation per row (the generic path spent ~60 instructions per pixel on them).
  __device__ __forceinline__ bool edge() const { return xl != xc - 1 || xr != xc + 1; }
  __device__ __forceinline__ Raw load_fast(int t) const {
    const uint8_t* p = base + (size_t)(t + 1) * pitch + xc;
    Raw r;
    r.a = __ldg(p - 1);
    r.b = __ldg(p);
    r.c = __ldg(p + 1);
    return r;
  }
  __device__ __forceinline__ float consume_fast(int t, Raw r) {   // t == tc + 1
    hm = h0;
    h0 = hp;
    hp = hval(r);
    tc = t;
    cur = fmaf(k1, hm + hp, k0 * h0);
    return cur;
  }
  __device__ __forceinline__ float consume(int t, Raw r) {
    const int n = clampi(t, 0, h - 1);
    if (n != tc) {   // n == tc + 1; uniform over the CTA (depends on t only)
      hm = h0;
      h0 = hp;
      hp = hval(r);
      tc = n;
      cur = fmaf(k1, hm + hp, k0 * h0);
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
  __shared__ __align__(16) float sV[2][PX_ROWS][3][PX_COLS];

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

  // H-phase role
  const int hr = tid >> 6;            // row within the step
  const int q0 = (tid & 63) * 4;      // first of 4 adjacent strip columns
  const int hx = x_base + q0;         // image x of that column (multiple of 4)
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
    for (int r = 0; r < PX_ROWS; r++) {
      const float* c = win + NMAX + r;   // centre of row ys + r
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
      sV[buf][r][0][tid] = r0;
      sV[buf][r][1][tid] = r1;
      sV[buf][r][2][tid] = r2;
    }
#pragma unroll
    for (int i = 0; i < 2 * NMAX; i++) win[i] = win[i + PX_ROWS];
    __syncthreads();
    // ---------------- H: horizontal moments of 4 adjacent pixels of row ys + hr
    const int y = ys + hr;
    if (h_valid && y < y1) {
      float4 o[4];
      float ob[4];
      {
        float e[20];   // e[i] = r0 at strip column q0 - 8 + i
#pragma unroll
        for (int v = 0; v < 5; v++) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[buf][hr][0][q0 - 8 + 4 * v]);
          e[4 * v] = t.x; e[4 * v + 1] = t.y; e[4 * v + 2] = t.z; e[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = 8 + j;
          float b1 = e[c] * pc.g[0], b2 = 0.f, b4 = 0.f;
#pragma unroll
          for (int k = 1; k <= NMAX; k++) {
            if (k <= n) {
              const float tg = e[c + k] + e[c - k];
              b1 = fmaf(tg, pc.g[k], b1);
              b4 = fmaf(tg, pc.xxg[k], b4);
              b2 = fmaf(e[c + k] - e[c - k], pc.xg[k], b2);
            }
          }
          o[j].y = b2 * pc.ig11;
          o[j].z = b1 * pc.ig03;   // + b5 * ig33 below
          o[j].w = b1 * pc.ig03 + b4 * pc.ig33;
        }
      }
      {
        float e[20];   // r1
#pragma unroll
        for (int v = 0; v < 5; v++) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[buf][hr][1][q0 - 8 + 4 * v]);
          e[4 * v] = t.x; e[4 * v + 1] = t.y; e[4 * v + 2] = t.z; e[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = 8 + j;
          float b3 = e[c] * pc.g[0], b6 = 0.f;
#pragma unroll
          for (int k = 1; k <= NMAX; k++) {
            if (k <= n) {
              b3 = fmaf(e[c + k] + e[c - k], pc.g[k], b3);
              b6 = fmaf(e[c + k] - e[c - k], pc.xg[k], b6);
            }
          }
          o[j].x = b3 * pc.ig11;
          ob[j] = b6 * pc.ig55;
        }
      }
      {
        float e[20];   // r2
#pragma unroll
        for (int v = 0; v < 5; v++) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[buf][hr][2][q0 - 8 + 4 * v]);
          e[4 * v] = t.x; e[4 * v + 1] = t.y; e[4 * v + 2] = t.z; e[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = 8 + j;
          float b5 = e[c] * pc.g[0];
#pragma unroll
          for (int k = 1; k <= NMAX; k++) {
            if (k <= n) b5 = fmaf(e[c + k] + e[c - k], pc.g[k], b5);
          }
          o[j].z += b5 * pc.ig33;
        }
      }
      const size_t ob0 = fbase + (size_t)y * w + hx;
      if (hx + 3 < w && ((w & 3) == 0)) {
        // rows are 16-byte aligned when w % 4 == 0: one float4 store for the 4 ch-4 values
#pragma unroll
        for (int j = 0; j < 4; j++) RA[ob0 + j] = o[j];
        *reinterpret_cast<float4*>(RB + ob0) = make_float4(ob[0], ob[1], ob[2], ob[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (hx + j < w) {
            RA[ob0 + j] = o[j];
            RB[ob0 + j] = ob[j];
          }
      }
    }
    buf ^= 1;
  }
}

}  // namespace ofb
