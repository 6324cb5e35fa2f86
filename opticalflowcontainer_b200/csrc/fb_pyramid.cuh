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

// =====================================================================================
// Fast pyramid stage for the regular case: W = S * w_l, H = S * h_l with S = 2, 4, 8 (pyr_scale = 0.5, frame size
// divisible by S) and W, pitch multiples of 4.  Every level pixel then samples the blurred image exactly half-way
// between two source pixels in x and in y (fx = fy = 0.5), so blur + bilinear resize collapse into ONE symmetric
// even-length filter per axis:
//     I[y][x] = sum_{j=1..R+1} c[j] * ( T[cy - (j-1)][.] + T[cy + 1 + (j-1)][.] ),   cy = S*y + S/2 - 1,
//     c[j] = 0.5 * (k[j-1] + k[j])   (k = cv2's Gaussian half kernel, k[R+1] = 0)
// and the same along x.  Vertical pass first, on the u8 source: a thread owns 4 adjacent source columns (one 32-bit
// word per row), keeps the 2R+2 rows of the window in registers split into even/odd byte lanes, adds the two rows
// of a tap pair as packed 16-bit integers (exact) and converts each sum once (PRMT into 2^23 + s, FADD) — one
// conversion and one FMA per TWO taps, where the two-pass kernels above spend a conversion and two FMAs per tap.
// The 4 floats go to a shared-memory row (skewed so that the stride-S reads of the horizontal pass are
// conflict-free); after one barrier the first OUT threads apply the same filter along x and store the level row.
// HBM traffic per frame-level: N (u8, read once; the halo columns are L2 hits) + 4 n_l.
constexpr int PF_THREADS = 256;
constexpr int PF_COLS = 4 * PF_THREADS;   // source columns per CTA
constexpr int PF_HALO = 8;                // halo columns per side (>= R + 1 - S/2 for the three instantiations)
struct PyrFastCoef { float c[10]; };      // c[j-1] for tap pair j = 1..R+1

__device__ __forceinline__ int pf_skew(int x) { return x + (x >> 5); }

template <int S, int R>
__global__ void __launch_bounds__(PF_THREADS) k_pyr_fast(FrameSrc src, int W, int H, float* __restrict__ out, int w, int h,
                                                         PyrFastCoef pc, int seg_rows) {
  constexpr int WIN = 2 * R + 2;
  constexpr int OUT = (PF_COLS - 2 * PF_HALO) / S;      // level columns per CTA (even)
  static_assert(S / 2 - 1 - R + PF_HALO >= 0 && S / 2 + R - (S - 1) <= PF_HALO, "halo too small");
  __shared__ float rowbuf[2][PF_COLS + PF_COLS / 32 + 1];
  const int tid = threadIdx.x;
  const int xo0 = blockIdx.x * OUT;                     // first level column of the chunk
  const int x = S * xo0 - PF_HALO + 4 * tid;            // first of this thread's 4 source columns (multiple of 4)
  const int y0 = blockIdx.y * seg_rows, y1 = min(y0 + seg_rows, h);
  const uint8_t* frame = src.frame(blockIdx.z);
  const bool interior = x >= 0 && x + 3 < W;
  const bool dead = x > S * (w - 1) + S / 2 + R;        // right of everything a level column of this frame reads
  // the four source columns of a thread at the image border, REFLECT_101
  const int bx0 = reflect101(x, W), bx1 = reflect101(x + 1, W), bx2 = reflect101(x + 2, W), bx3 = reflect101(x + 3, W);
  const bool fastcol = interior && !dead;
  auto load_row = [&](int ry) -> unsigned {             // source row ry (any integer: reflected)
    if (fastcol && (unsigned)ry < (unsigned)H)
      return __ldg(reinterpret_cast<const unsigned*>(frame + (size_t)ry * src.pitch + x));
    const uint8_t* p = frame + (size_t)reflect101(ry, H) * src.pitch;
    if (dead) return 0u;
    if (interior) return __ldg(reinterpret_cast<const unsigned*>(p + x));
    return (unsigned)__ldg(p + bx0) | ((unsigned)__ldg(p + bx1) << 8) | ((unsigned)__ldg(p + bx2) << 16) |
           ((unsigned)__ldg(p + bx3) << 24);
  };
  // window rows cy - R .. cy + R + 1 as even / odd byte lanes (16 bits per lane)
  unsigned we[WIN], wo[WIN];
  {
    const int top = S * y0 + S / 2 - 1 - R;
#pragma unroll
    for (int i = S; i < WIN; i++) {                     // rows the first step keeps; its S newest rows come as `nx`
      const unsigned v = load_row(top + i - S);
      we[i] = v & 0x00FF00FFu;
      wo[i] = (v >> 8) & 0x00FF00FFu;
    }
  }
  // the S rows entering the window at a step: one address computation and S word loads for interior threads on
  // interior rows (the generic path costs ~60 instructions per row: reflection loop, border bytes)
  const size_t pitch_w = src.pitch >> 2;
  auto load_group = [&](int first, unsigned (&dst)[S]) {
    if (fastcol && first >= 0 && first + S <= H) {
      const unsigned* p = reinterpret_cast<const unsigned*>(frame + (size_t)first * src.pitch + x);
#pragma unroll
      for (int i = 0; i < S; i++) dst[i] = __ldg(p + (size_t)i * pitch_w);
    } else {
#pragma unroll
      for (int i = 0; i < S; i++) dst[i] = load_row(first + i);
    }
  };
  unsigned nx[S];
  load_group(S * y0 + S / 2 - 1 - R + WIN - S, nx);

  for (int y = y0; y < y1; y++) {
    // slide the window down by S rows: the S newest rows were loaded one step ahead
#pragma unroll
    for (int i = 0; i < WIN - S; i++) { we[i] = we[i + S]; wo[i] = wo[i + S]; }
#pragma unroll
    for (int i = 0; i < S; i++) {
      we[WIN - S + i] = nx[i] & 0x00FF00FFu;
      wo[WIN - S + i] = (nx[i] >> 8) & 0x00FF00FFu;
    }
    if (y + 1 < y1) load_group(S * (y + 1) + S / 2 - 1 - R + WIN - S, nx);
    // vertical filter of the 4 columns
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 1; j <= R + 1; j++) {
      const unsigned e = we[R + 1 - j] + we[R + j];     // columns 0 and 2 (two 9-bit sums in 16-bit lanes)
      const unsigned o = wo[R + 1 - j] + wo[R + j];     // columns 1 and 3
      const float c = pc.c[j - 1];
      a0 = fmaf(c, __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7610)) - 8388608.f, a0);
      a2 = fmaf(c, __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7632)) - 8388608.f, a2);
      a1 = fmaf(c, __uint_as_float(__byte_perm(o, 0x4B000000u, 0x7610)) - 8388608.f, a1);
      a3 = fmaf(c, __uint_as_float(__byte_perm(o, 0x4B000000u, 0x7632)) - 8388608.f, a3);
    }
    float* rb = rowbuf[y & 1];
    const int q = pf_skew(4 * tid);                     // 4*tid .. 4*tid+3 share one 32-group: contiguous after the skew
    rb[q] = a0; rb[q + 1] = a1; rb[q + 2] = a2; rb[q + 3] = a3;
    __syncthreads();                                    // (double-buffered: one barrier per level row)
    // horizontal filter: level column xo0 + i needs chunk columns PF_HALO + S*i + S/2 - 1 - R .. + S/2 + R
#pragma unroll
    for (int i = tid; i < OUT; i += PF_THREADS) {
      const int xo = xo0 + i;
      if (xo < w) {
        const int cl = PF_HALO + S * i + S / 2 - 1;     // chunk column of the left centre tap
        float acc = 0.f;
#pragma unroll
        for (int j = 1; j <= R + 1; j++) acc = fmaf(pc.c[j - 1], rb[pf_skew(cl - (j - 1))] + rb[pf_skew(cl + j)], acc);
        out[((size_t)blockIdx.z * h + y) * w + xo] = acc;
      }
    }
  }
}

// =====================================================================================
// The three regular levels S = 2, 4, 8 (R = 1, 4, 9: the node defaults at pyr_scale 0.5, three levels) in ONE pass over
// the source.  Launched one after the other, each k_pyr_fast reads the whole u8 frame again and is a ~65 us launch of
// its own whatever the level's size (36 frames of 1080p: 3 x 65 us for 0.33 N of output).  The 20-row window of the
// coarsest level (source rows 8Y-6 .. 8Y+13 for its row Y) contains the windows of the two rows 2Y, 2Y+1 of the S = 4
// level (10 rows each) and of the four rows 4Y .. 4Y+3 of the S = 2 level (4 rows each), so one register window
// serves all three: per step of 8 source rows a thread (4 source columns) emits 1 + 2 + 4 vertically filtered rows
// into shared memory and the CTA then applies each level's horizontal filter.  Same arithmetic, same order of
// operations per level as k_pyr_fast: identical bits.
struct PyrFast3Coef { float c1[2], c2[5], c3[10]; };
constexpr int PF3_STRIDE = PF_COLS + PF_COLS / 32 + 1;

template <int I0, int NP>
__device__ __forceinline__ void pf3_vfilt(const unsigned (&we)[20], const unsigned (&wo)[20], const float* __restrict__ c,
                                          float* __restrict__ rb) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int j = 1; j <= NP; j++) {
    const unsigned e = we[I0 - (j - 1)] + we[I0 + j];
    const unsigned o = wo[I0 - (j - 1)] + wo[I0 + j];
    const float cj = c[j - 1];
    a0 = fmaf(cj, __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7610)) - 8388608.f, a0);
    a2 = fmaf(cj, __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7632)) - 8388608.f, a2);
    a1 = fmaf(cj, __uint_as_float(__byte_perm(o, 0x4B000000u, 0x7610)) - 8388608.f, a1);
    a3 = fmaf(cj, __uint_as_float(__byte_perm(o, 0x4B000000u, 0x7632)) - 8388608.f, a3);
  }
  rb[0] = a0; rb[1] = a1; rb[2] = a2; rb[3] = a3;
}

// horizontal filter of NROWS staged rows of one level: level column xo0 + i reads chunk columns
// PF_HALO + S*i + S/2 - 1 - R .. + S/2 + R
template <int S, int R, int NROWS>
__device__ __forceinline__ void pf3_hfilt(const float* __restrict__ rows, const float* __restrict__ c, float* __restrict__ out,
                                          int w, int h, int frame, int y_first, int xo0, int tid) {
  constexpr int OUT = (PF_COLS - 2 * PF_HALO) / S;
  for (int i = tid; i < NROWS * OUT; i += PF_THREADS) {
    const int s = i / OUT, col = i - s * OUT;
    const int xo = xo0 + col;
    if (xo < w) {
      const float* rb = rows + s * PF3_STRIDE;
      const int cl = PF_HALO + S * col + S / 2 - 1;
      float acc = 0.f;
#pragma unroll
      for (int j = 1; j <= R + 1; j++) acc = fmaf(c[j - 1], rb[pf_skew(cl - (j - 1))] + rb[pf_skew(cl + j)], acc);
      out[((size_t)frame * h + (y_first + s)) * w + xo] = acc;
    }
  }
}

// grid: (ceil(w3 / 126), ceil((y3_end - y3_begin) / seg_rows), frames); out1/2/3: the level images [frames][h_l][w_l] of S = 2, 4, 8.
__global__ void __launch_bounds__(PF_THREADS, 3) k_pyr_fast3(FrameSrc src, int W, int H, float* __restrict__ out1,
                                                          float* __restrict__ out2, float* __restrict__ out3,
                                                          PyrFast3Coef pc, int seg_rows, int y3_begin, int y3_end) {
  constexpr int S = 8, R = 9, WIN = 20;
  constexpr int OUT3 = (PF_COLS - 2 * PF_HALO) / 8;
  __shared__ float rowbuf[7 * PF3_STRIDE];              // row 0: S = 8; rows 1, 2: S = 4; rows 3..6: S = 2
  const int w3 = W >> 3, h3 = H >> 3;
  const int tid = threadIdx.x;
  const int xo0 = blockIdx.x * OUT3;                    // first S = 8 level column of the chunk
  const int x = S * xo0 - PF_HALO + 4 * tid;            // first of this thread's 4 source columns (multiple of 4)
  const int y0 = y3_begin + blockIdx.y * seg_rows, y1 = min(y0 + seg_rows, y3_end);   // rows [y3_begin, y3_end) of the S = 8 level
  const uint8_t* frame = src.frame(blockIdx.z);
  const bool interior = x >= 0 && x + 3 < W;
  const bool dead = x > W + 5;                          // right of everything a level column of this frame reads
  const int bx0 = reflect101(x, W), bx1 = reflect101(x + 1, W), bx2 = reflect101(x + 2, W), bx3 = reflect101(x + 3, W);
  const bool fastcol = interior && !dead;
  auto load_row = [&](int ry) -> unsigned {             // source row ry (any integer: reflected)
    if (fastcol && (unsigned)ry < (unsigned)H)
      return __ldg(reinterpret_cast<const unsigned*>(frame + (size_t)ry * src.pitch + x));
    const uint8_t* p = frame + (size_t)reflect101(ry, H) * src.pitch;
    if (dead) return 0u;
    if (interior) return __ldg(reinterpret_cast<const unsigned*>(p + x));
    return (unsigned)__ldg(p + bx0) | ((unsigned)__ldg(p + bx1) << 8) | ((unsigned)__ldg(p + bx2) << 16) |
           ((unsigned)__ldg(p + bx3) << 24);
  };
  unsigned we[WIN], wo[WIN];
  {
    const int top = S * y0 + S / 2 - 1 - R;
#pragma unroll
    for (int i = S; i < WIN; i++) {
      const unsigned v = load_row(top + i - S);
      we[i] = v & 0x00FF00FFu;
      wo[i] = (v >> 8) & 0x00FF00FFu;
    }
  }
  const size_t pitch_w = src.pitch >> 2;
  auto load_group = [&](int first, unsigned (&dst)[S]) {
    if (fastcol && first >= 0 && first + S <= H) {
      const unsigned* p = reinterpret_cast<const unsigned*>(frame + (size_t)first * src.pitch + x);
#pragma unroll
      for (int i = 0; i < S; i++) dst[i] = __ldg(p + (size_t)i * pitch_w);
    } else {
#pragma unroll
      for (int i = 0; i < S; i++) dst[i] = load_row(first + i);
    }
  };
  unsigned nx[S];
  load_group(S * y0 + S / 2 - 1 - R + WIN - S, nx);
  float* mine = rowbuf + pf_skew(4 * tid);              // 4*tid .. 4*tid+3 share one 32-group: contiguous after the skew

  for (int y = y0; y < y1; y++) {
#pragma unroll
    for (int i = 0; i < WIN - S; i++) { we[i] = we[i + S]; wo[i] = wo[i + S]; }
#pragma unroll
    for (int i = 0; i < S; i++) {
      we[WIN - S + i] = nx[i] & 0x00FF00FFu;
      wo[WIN - S + i] = (nx[i] >> 8) & 0x00FF00FFu;
    }
    if (y + 1 < y1) load_group(S * (y + 1) + S / 2 - 1 - R + WIN - S, nx);
    // window index i = source row 8y - 6 + i.  Left centre tap of a level row: S_l * y_l + S_l/2 - 1.
    pf3_vfilt<9, 10>(we, wo, pc.c3, mine);                              // S = 8, row y:        8y + 3
    pf3_vfilt<7, 5>(we, wo, pc.c2, mine + 1 * PF3_STRIDE);              // S = 4, row 2y:       8y + 1
    pf3_vfilt<11, 5>(we, wo, pc.c2, mine + 2 * PF3_STRIDE);             //        row 2y + 1:   8y + 5
    pf3_vfilt<6, 2>(we, wo, pc.c1, mine + 3 * PF3_STRIDE);              // S = 2, row 4y:       8y
    pf3_vfilt<8, 2>(we, wo, pc.c1, mine + 4 * PF3_STRIDE);              //        row 4y + 1:   8y + 2
    pf3_vfilt<10, 2>(we, wo, pc.c1, mine + 5 * PF3_STRIDE);             //        row 4y + 2:   8y + 4
    pf3_vfilt<12, 2>(we, wo, pc.c1, mine + 6 * PF3_STRIDE);             //        row 4y + 3:   8y + 6
    __syncthreads();
    pf3_hfilt<2, 1, 4>(rowbuf + 3 * PF3_STRIDE, pc.c1, out1, W >> 1, H >> 1, blockIdx.z, 4 * y, 4 * xo0, tid);
    pf3_hfilt<4, 4, 2>(rowbuf + 1 * PF3_STRIDE, pc.c2, out2, W >> 2, H >> 2, blockIdx.z, 2 * y, 2 * xo0, tid);
    pf3_hfilt<8, 9, 1>(rowbuf, pc.c3, out3, w3, h3, blockIdx.z, y, xo0, tid);
    __syncthreads();                                    // the rows are rewritten by the next step
  }
}

}  // namespace ofb
