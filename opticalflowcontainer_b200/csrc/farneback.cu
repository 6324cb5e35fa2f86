// farneback.cu — dense Farneback optical flow for B200 (sm_100a), hand-written CUDA.
//
// Replaces cv2.calcOpticalFlowFarneback behind the node flow call
// (ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:194).  The arithmetic follows
// OpenCV 4.x modules/video/src/optflowgf.cpp (un-vendored dependency of the reference,
// ros2_ws/src/nueflow/setup.py:29); stage names below are the upstream function names.
//
// Data layout in HBM (all per level, reused across levels):
//   I   : f32   [frame][h][w]                       level image (blur + bilinear resize of the u8 source)
//   R   : float4[frame][h][w] (ch0..3) + f32[frame][h][w] (ch4)   polynomial coefficients, 20 B/px,
//         split so that the bilinear gather of UpdateMatrices is 1 LDG.128 + 1 LDG.32 per corner
//   flow: float2[pair][h][w]
// None of the stages is a dense contraction: no tensor cores; everything is HBM/L1-bound
// stencil + gather work.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace ofb {

// =====================================================================================
// Host-side schedule and coefficient preparation
// =====================================================================================

// FarnebackOpticalFlowImpl::calc: level clamp + per-level (scale, sigma, ksize, w, h), coarse → fine.
int build_schedule(int width, int height, double pyr_scale, int levels, Level* out, int* n_out) {
  const int min_size = 32;
  double scale = 1.0;
  int k = 0;
  for (; k < levels; k++) {
    scale *= pyr_scale;
    if (width * scale < min_size || height * scale < min_size) break;
  }
  int levels_eff = k;
  if (levels_eff + 1 > kMaxLevels) return OFB_ERR_INVALID_ARG;
  int n = 0;
  for (k = levels_eff; k >= 0; k--) {
    scale = 1.0;
    for (int i = 0; i < k; i++) scale *= pyr_scale;
    Level lv;
    lv.k = k;
    lv.scale = scale;
    lv.sigma = (1.0 / scale - 1.0) * 0.5;
    int ks = cv_round(lv.sigma * 5) | 1;
    lv.ksize = std::max(ks, 3);
    lv.width = cv_round(width * scale);
    lv.height = cv_round(height * scale);
    out[n++] = lv;
  }
  *n_out = n;
  return OFB_OK;
}

// 6x6 inverse by Gauss-Jordan with partial pivoting (G is SPD, tiny).
static void invert6(double a[6][6], double inv[6][6]) {
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) inv[i][j] = (i == j);
  for (int c = 0; c < 6; c++) {
    int p = c;
    for (int r = c + 1; r < 6; r++)
      if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    for (int j = 0; j < 6; j++) {
      std::swap(a[c][j], a[p][j]);
      std::swap(inv[c][j], inv[p][j]);
    }
    double d = 1.0 / a[c][c];
    for (int j = 0; j < 6; j++) {
      a[c][j] *= d;
      inv[c][j] *= d;
    }
    for (int r = 0; r < 6; r++) {
      if (r == c) continue;
      double f = a[r][c];
      if (f == 0) continue;
      for (int j = 0; j < 6; j++) {
        a[r][j] -= f * a[c][j];
        inv[r][j] -= f * inv[c][j];
      }
    }
  }
}

// FarnebackPrepareGaussian.
void prepare_poly(int n, double sigma, PolyCoef* pc) {
  if (sigma < 1.1920929e-07) sigma = n * 0.3;
  float g[2 * kMaxPolyN + 1];
  double s = 0;
  for (int x = -n; x <= n; x++) {
    g[x + n] = (float)exp(-x * x / (2 * sigma * sigma));
    s += g[x + n];
  }
  s = 1.0 / s;
  pc->n = n;
  for (int x = -n; x <= n; x++) g[x + n] = (float)(g[x + n] * s);
  for (int x = 0; x <= n; x++) {
    pc->g[x] = g[x + n];
    pc->xg[x] = (float)(x * (double)g[x + n]);
    pc->xxg[x] = (float)(x * x * (double)g[x + n]);
  }
  double G[6][6] = {{0}};
  for (int y = -n; y <= n; y++)
    for (int x = -n; x <= n; x++) {
      double w = (double)g[y + n] * g[x + n];
      G[0][0] += w;
      G[1][1] += w * x * x;
      G[3][3] += w * x * x * x * x;
      G[5][5] += w * x * x * y * y;
    }
  G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
  G[4][4] = G[3][3];
  G[3][4] = G[4][3] = G[5][5];
  double inv[6][6];
  invert6(G, inv);
  pc->ig11 = (float)inv[1][1];
  pc->ig03 = (float)inv[0][3];
  pc->ig33 = (float)inv[3][3];
  pc->ig55 = (float)inv[5][5];
}

// Box window (FarnebackUpdateFlow_Blur) or Gaussian (FarnebackUpdateFlow_GaussianBlur) weights.
void prepare_blur(int winsize, bool gaussian, BlurCoef* bc) {
  int m = winsize / 2;
  bc->m = m;
  bc->gaussian = gaussian ? 1 : 0;
  if (!gaussian) {
    for (int i = 0; i <= m; i++) bc->k[i] = 1.f;
    bc->scale = (float)(1.0 / ((double)winsize * winsize));
    return;
  }
  double sigma = m * 0.3, s = 1;
  bc->k[0] = 1.f;
  for (int i = 1; i <= m; i++) {
    float t = (float)exp(-i * i / (2 * sigma * sigma));
    bc->k[i] = t;
    s += t * 2;
  }
  s = 1. / s;
  for (int i = 0; i <= m; i++) bc->k[i] = (float)(bc->k[i] * s);
  bc->scale = 1.f;
}

// cv::getGaussianKernel(ksize, sigma, CV_32F) — half kernel k[0..r].
constexpr int kMaxPyrRadius = 159;
struct PyrCoef {
  int r;
  float k[kMaxPyrRadius + 1];
};

static int prepare_pyr(int ksize, double sigma, PyrCoef* pc) {
  int r = ksize / 2;
  if (r > kMaxPyrRadius) return OFB_ERR_INVALID_ARG;
  pc->r = r;
  if (sigma <= 0) {  // only ksize == 3 reaches here (k = 0 level)
    pc->k[0] = 0.5f;
    pc->k[1] = 0.25f;
    return OFB_OK;
  }
  std::vector<double> t(ksize);
  double sum = 0;
  for (int i = 0; i < ksize; i++) {
    double x = i - (ksize - 1) * 0.5;
    t[i] = exp(-0.5 * x * x / (sigma * sigma));
    sum += t[i];
  }
  for (int i = 0; i <= r; i++) pc->k[i] = (float)(t[r + i] / sum);
  return OFB_OK;
}

// =====================================================================================
// Device helpers
// =====================================================================================
__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
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
// Stage a2: pyramid level = convertTo(f32) + GaussianBlur(REFLECT_101) + resize(INTER_LINEAR)
// One thread per level pixel; the blur is evaluated only at the (up to) 2x2 source samples the
// bilinear resize reads.  Horizontal pass first, then vertical, as cv2's separable filter does.
// =====================================================================================
__global__ void __launch_bounds__(256) k_pyr_level(FrameSrc src, int W, int H, float* __restrict__ out, int w,
                                                   int h, double sx_scale, double sy_scale, PyrCoef pc) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const uint8_t* img = src.frame(blockIdx.z);
  const int r = pc.r;
  int sx0, sy0;
  float fx, fy;
  if (w == W) { sx0 = x; fx = 0.f; } else linear_coord(x, sx_scale, W, &sx0, &fx);
  if (h == H) { sy0 = y; fy = 0.f; } else linear_coord(y, sy_scale, H, &sy0, &fy);
  const bool need_c1 = fx != 0.f, need_r1 = fy != 0.f;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;  // blurred(sy0, sx0), (sy0, sx0+1), (sy0+1, sx0), (sy0+1, sx0+1)
  const int t_hi = need_r1 ? r + 1 : r;
  const int i_hi = need_c1 ? r + 1 : r;
  for (int t = -r; t <= t_hi; t++) {
    const uint8_t* row = img + (size_t)reflect101(sy0 + t, H) * src.pitch;
    float h0 = 0.f, h1 = 0.f;
    for (int i = -r; i <= i_hi; i++) {
      float v = (float)__ldg(row + reflect101(sx0 + i, W));
      if (i <= r) h0 = fmaf(pc.k[abs(i)], v, h0);
      if (i >= -r + 1) h1 = fmaf(pc.k[abs(i - 1)], v, h1);
    }
    if (t <= r) {
      float kt = pc.k[abs(t)];
      a00 = fmaf(kt, h0, a00);
      a01 = fmaf(kt, h1, a01);
    }
    if (t >= -r + 1) {
      float kt = pc.k[abs(t - 1)];
      a10 = fmaf(kt, h0, a10);
      a11 = fmaf(kt, h1, a11);
    }
  }
  float top = a00 * (1.f - fx) + a01 * fx;
  float bot = a10 * (1.f - fx) + a11 * fx;
  out[((size_t)blockIdx.z * h + y) * w + x] = top * (1.f - fy) + bot * fy;
}

// =====================================================================================
// Stage a4: FarnebackPolyExp.  32x32 output tile per 256-thread CTA; level image tile with an
// n-pixel replicate halo staged in shared memory; vertical pass -> 3 moment planes in shared
// memory; horizontal pass -> 5 coefficients, written as float4 + float.
// =====================================================================================
constexpr int PE_T = 32;
template <int NT>
__global__ void __launch_bounds__(256) k_polyexp(const float* __restrict__ I, float4* __restrict__ RA,
                                                 float* __restrict__ RB, int w, int h, PolyCoef pc) {
  const int n = NT > 0 ? NT : pc.n;
  constexpr int NMAX = NT > 0 ? NT : kMaxPolyN;
  constexpr int TW = PE_T + 2 * NMAX + 1;  // padded row length
  __shared__ float sI[(PE_T + 2 * NMAX) * TW];
  __shared__ float sV[3][PE_T * TW];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * PE_T, y0 = blockIdx.y * PE_T;
  const size_t fbase = (size_t)blockIdx.z * w * h;
  const float* img = I + fbase;
  const int tw = PE_T + 2 * n, th = PE_T + 2 * n;
  for (int idx = tid; idx < tw * th; idx += 256) {
    int ty = idx / tw, tx = idx - ty * tw;
    int gx = clampi(x0 + tx - n, 0, w - 1), gy = clampi(y0 + ty - n, 0, h - 1);
    sI[ty * TW + tx] = __ldg(img + (size_t)gy * w + gx);
  }
  __syncthreads();
  // vertical pass
  for (int idx = tid; idx < tw * PE_T; idx += 256) {
    int ty = idx / tw, tx = idx - ty * tw;
    const float* c = sI + (ty + n) * TW + tx;
    float r0 = c[0] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 1; k <= NMAX; k++) {
      if (k <= n) {
        float a = c[-k * TW], b = c[k * TW];
        float p = a + b;
        r0 = fmaf(pc.g[k], p, r0);
        r1 = fmaf(pc.xg[k], b - a, r1);
        r2 = fmaf(pc.xxg[k], p, r2);
      }
    }
    sV[0][ty * TW + tx] = r0;
    sV[1][ty * TW + tx] = r1;
    sV[2][ty * TW + tx] = r2;
  }
  __syncthreads();
  // horizontal pass
  const int x = x0 + threadIdx.x;
  for (int ty = threadIdx.y; ty < PE_T; ty += 8) {
    int y = y0 + ty;
    if (x >= w || y >= h) continue;
    const float* v0 = sV[0] + ty * TW + threadIdx.x + n;
    const float* v1 = sV[1] + ty * TW + threadIdx.x + n;
    const float* v2 = sV[2] + ty * TW + threadIdx.x + n;
    float b1 = v0[0] * pc.g[0], b2 = 0.f, b3 = v1[0] * pc.g[0], b4 = 0.f, b5 = v2[0] * pc.g[0], b6 = 0.f;
#pragma unroll
    for (int k = 1; k <= NMAX; k++) {
      if (k <= n) {
        float p0 = v0[k], m0 = v0[-k], p1 = v1[k], m1 = v1[-k], p2 = v2[k], m2 = v2[-k];
        float tg = p0 + m0;
        b1 = fmaf(tg, pc.g[k], b1);
        b4 = fmaf(tg, pc.xxg[k], b4);
        b2 = fmaf(p0 - m0, pc.xg[k], b2);
        b3 = fmaf(p1 + m1, pc.g[k], b3);
        b6 = fmaf(p1 - m1, pc.xg[k], b6);
        b5 = fmaf(p2 + m2, pc.g[k], b5);
      }
    }
    size_t o = fbase + (size_t)y * w + x;
    RA[o] = make_float4(b3 * pc.ig11, b2 * pc.ig11, b1 * pc.ig03 + b5 * pc.ig33, b1 * pc.ig03 + b4 * pc.ig33);
    RB[o] = b6 * pc.ig55;
  }
}

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

__device__ __forceinline__ M5 update_matrix_px(const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                               const float4* __restrict__ RA1, const float* __restrict__ RB1,
                                               float2 fl, int x, int y, int w, int h) {
  const size_t o = (size_t)y * w + x;
  const float4 a0 = __ldg(RA0 + o);
  const float b0 = __ldg(RB0 + o);
  const float dx = fl.x, dy = fl.y;
  float fx = (float)x + dx, fy = (float)y + dy;
  const float flx = floorf(fx), fly = floorf(fy);
  fx -= flx;
  fy -= fly;
  float r2, r3, r4, r5, r6;
  // (unsigned)x1 < (unsigned)(w-1): compare in float first so huge/NaN flows stay outside
  if (flx >= 0.f && flx < (float)(w - 1) && fly >= 0.f && fly < (float)(h - 1)) {
    const int x1 = (int)flx, y1 = (int)fly;
    const size_t p = (size_t)y1 * w + x1;
    const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
    const float4 q00 = __ldg(RA1 + p), q01 = __ldg(RA1 + p + 1), q10 = __ldg(RA1 + p + w), q11 = __ldg(RA1 + p + w + 1);
    const float s00 = __ldg(RB1 + p), s01 = __ldg(RB1 + p + 1), s10 = __ldg(RB1 + p + w), s11 = __ldg(RB1 + p + w + 1);
    r2 = a00 * q00.x + a01 * q01.x + a10 * q10.x + a11 * q11.x;
    r3 = a00 * q00.y + a01 * q01.y + a10 * q10.y + a11 * q11.y;
    r4 = a00 * q00.z + a01 * q01.z + a10 * q10.z + a11 * q11.z;
    r5 = a00 * q00.w + a01 * q01.w + a10 * q10.w + a11 * q11.w;
    r6 = a00 * s00 + a01 * s01 + a10 * s10 + a11 * s11;
    r4 = (a0.z + r4) * 0.5f;
    r5 = (a0.w + r5) * 0.5f;
    r6 = (b0 + r6) * 0.25f;
  } else {
    r2 = r3 = 0.f;
    r4 = a0.z;
    r5 = a0.w;
    r6 = b0 * 0.5f;
  }
  r2 = (a0.x - r2) * 0.5f;
  r3 = (a0.y - r3) * 0.5f;
  r2 += r4 * dy + r6 * dx;
  r3 += r6 * dy + r5 * dx;
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

__global__ void __launch_bounds__(256) k_update_matrices(const float4* __restrict__ RA, const float* __restrict__ RB,
                                                         const float2* __restrict__ flow, float4* __restrict__ MA,
                                                         float* __restrict__ MB, int w, int h, int f1_offset) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t n = (size_t)w * h;
  const int pair = blockIdx.z;
  const size_t f0 = (size_t)pair * n, f1 = (size_t)(pair + f1_offset) * n;
  const size_t o = (size_t)y * w + x;
  M5 m = update_matrix_px(RA + f0, RB + f0, RA + f1, RB + f1, flow[f0 + o], x, y, w, h);
  MA[f0 + o] = make_float4(m.g11, m.g12, m.g22, m.h1);
  MB[f0 + o] = m.h2;
}

// =====================================================================================
// Stage a6/a7 (generic path): separable weighted blur of the 5-channel field (replicate border)
// as two passes through global memory + the 2x2 solve.  Handles any winsize, box or Gaussian.
// =====================================================================================
__global__ void __launch_bounds__(256) k_blur_v(const float4* __restrict__ MA, const float* __restrict__ MB,
                                                float4* __restrict__ VA, float* __restrict__ VB, int w, int h,
                                                BlurCoef bc) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t base = (size_t)blockIdx.z * w * h;
  const float4* A = MA + base;
  const float* B = MB + base;
  float4 c = __ldg(A + (size_t)y * w + x);
  float k0 = bc.k[0];
  float4 s = make_float4(c.x * k0, c.y * k0, c.z * k0, c.w * k0);
  float sb = __ldg(B + (size_t)y * w + x) * k0;
  for (int i = 1; i <= bc.m; i++) {
    size_t p0 = (size_t)max(y - i, 0) * w + x, p1 = (size_t)min(y + i, h - 1) * w + x;
    float4 a = __ldg(A + p0), b = __ldg(A + p1);
    float ki = bc.k[i];
    s.x = fmaf(a.x + b.x, ki, s.x);
    s.y = fmaf(a.y + b.y, ki, s.y);
    s.z = fmaf(a.z + b.z, ki, s.z);
    s.w = fmaf(a.w + b.w, ki, s.w);
    sb = fmaf(__ldg(B + p0) + __ldg(B + p1), ki, sb);
  }
  VA[base + (size_t)y * w + x] = s;
  VB[base + (size_t)y * w + x] = sb;
}

__device__ __forceinline__ float2 solve2x2(float g11, float g12, float g22, float h1, float h2) {
  float idet = 1.f / (g11 * g22 - g12 * g12 + 1e-3f);
  return make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
}

__global__ void __launch_bounds__(256) k_blur_h_solve(const float4* __restrict__ VA, const float* __restrict__ VB,
                                                      float2* __restrict__ flow, int w, int h, BlurCoef bc) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t base = (size_t)blockIdx.z * w * h;
  const float4* A = VA + base + (size_t)y * w;
  const float* B = VB + base + (size_t)y * w;
  float4 c = __ldg(A + x);
  float k0 = bc.k[0];
  float4 s = make_float4(c.x * k0, c.y * k0, c.z * k0, c.w * k0);
  float sb = __ldg(B + x) * k0;
  for (int i = 1; i <= bc.m; i++) {
    int xa = max(x - i, 0), xb = min(x + i, w - 1);
    float4 a = __ldg(A + xa), b = __ldg(A + xb);
    float ki = bc.k[i];
    s.x = fmaf(a.x + b.x, ki, s.x);
    s.y = fmaf(a.y + b.y, ki, s.y);
    s.z = fmaf(a.z + b.z, ki, s.z);
    s.w = fmaf(a.w + b.w, ki, s.w);
    sb = fmaf(__ldg(B + xa) + __ldg(B + xb), ki, sb);
  }
  const float sc = bc.scale;
  flow[base + (size_t)y * w + x] = solve2x2(s.x * sc, s.y * sc, s.z * sc, s.w * sc, sb * sc);
}

// =====================================================================================
// Fused iteration kernel (box window): UpdateMatrices + (2m+1)^2 box blur + 2x2 solve in ONE pass.
// HBM traffic per pixel-iteration = R0 (20 B) + R1 gather (20 B) + flow in (8 B) + flow out (8 B).
//
// A CTA (8 warps) owns a strip of FI_COLS = 256 matrix columns (2m of them halo) and a segment of
// `seg_rows` output rows, and marches down it FI_CH = 4 matrix rows at a time:
//   A1  every warp computes M for one half-row (4 px per lane, 32 px apart: coalesced R0/flow loads
//       and L1-friendly gathers) into a shared staging row;
//   A2  horizontal window sums H: each lane owns 4 adjacent columns, reads the 2m+4 staged values
//       it needs as float4s and writes H into a ring of 2m+1 rows in shared memory;
//   B   one thread per column keeps the vertical window sum as a running sum in DOUBLE (add the
//       new H row, subtract the row leaving the window — exactly cv2's vsum scheme, so there is no
//       float cancellation drift), scales, solves the 2x2 system and writes flow.
// =====================================================================================
constexpr int FI_COLS = 256;
constexpr int FI_CH = 4;
constexpr int FI_THREADS = 256;

template <int MT>
__global__ void __launch_bounds__(FI_THREADS, 2)
    k_iter_box(const float4* __restrict__ RA, const float* __restrict__ RB, const float2* __restrict__ flow_in,
               float2* __restrict__ flow_out, int w, int h, int f1_offset, int m_rt, float scale, int seg_rows,
               int strips) {
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int tw = FI_COLS - 2 * m;
  extern __shared__ float smem[];
  float* stage = smem;                          // [FI_CH][5][FI_COLS]
  float* ring = smem + FI_CH * 5 * FI_COLS;     // [R][5][FI_COLS]

  const int strip = blockIdx.x % strips;
  const int seg = blockIdx.x / strips;
  const int pair = blockIdx.y;
  const int x_base = strip * tw - m;            // image x of strip column 0
  const int y0 = seg * seg_rows;
  const int y1 = min(y0 + seg_rows, h);         // exclusive
  const int t_first = y0 - m, t_last = y1 - 1 + m;

  const size_t n = (size_t)w * h;
  const float4* RA0 = RA + (size_t)pair * n;
  const float* RB0 = RB + (size_t)pair * n;
  const float4* RA1 = RA + (size_t)(pair + f1_offset) * n;
  const float* RB1 = RB + (size_t)(pair + f1_offset) * n;
  const float2* fin = flow_in + (size_t)pair * n;
  float2* fout = flow_out + (size_t)pair * n;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int a_row = warp >> 1, a_half = warp & 1;
  const int col = tid;                          // phase-B column
  const int out_x = x_base + col;
  const bool col_valid = col >= m && col < FI_COLS - m && out_x < w;

  double vs0 = 0, vs1 = 0, vs2 = 0, vs3 = 0, vs4 = 0;

  for (int tc = t_first; tc <= t_last; tc += FI_CH) {
    // ---------------- A1: matrices of row tc + a_row -> staging
    const int t = tc + a_row;
    if (t <= t_last) {
      const int y = clampi(t, 0, h - 1);
      float* srow = stage + a_row * 5 * FI_COLS;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = a_half * 128 + lane + 32 * j;
        const int x = clampi(x_base + c, 0, w - 1);
        const float2 fl = __ldg(fin + (size_t)y * w + x);
        const M5 mm = update_matrix_px(RA0, RB0, RA1, RB1, fl, x, y, w, h);
        srow[0 * FI_COLS + c] = mm.g11;
        srow[1 * FI_COLS + c] = mm.g12;
        srow[2 * FI_COLS + c] = mm.g22;
        srow[3 * FI_COLS + c] = mm.h1;
        srow[4 * FI_COLS + c] = mm.h2;
      }
    }
    // rows leaving the window: their ring slots are overwritten in A2, so fetch them now
    float old[FI_CH][5];
#pragma unroll
    for (int rr = 0; rr < FI_CH; rr++) {
      const int tt = tc + rr;
      const bool have = (tt - t_first >= R) && tt <= t_last;
      const int slot = (tt - t_first) % R;
#pragma unroll
      for (int ch = 0; ch < 5; ch++) old[rr][ch] = have ? ring[(slot * 5 + ch) * FI_COLS + col] : 0.f;
    }
    __syncthreads();
    // ---------------- A2: horizontal window sums of the staged rows -> ring
    if (t <= t_last) {
      const float* srow = stage + a_row * 5 * FI_COLS;
      float* rrow = ring + ((t - t_first) % R) * 5 * FI_COLS;
      const int q0 = a_half * 128 + 4 * lane;   // first of this lane's 4 columns
      const int kq = (m + 3) >> 2;              // quads to each side
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        const float* s = srow + ch * FI_COLS;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (MT > 0) {
#pragma unroll
          for (int k = -((MT + 3) / 4); k <= (MT + 3) / 4; k++) {
            const int cq = min(max(q0 + 4 * k, 0), FI_COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int d = 4 * k + i;  // offset from q0
              if (d >= 0 - MT && d <= 0 + MT) s0 += e[i];
              if (d >= 1 - MT && d <= 1 + MT) s1 += e[i];
              if (d >= 2 - MT && d <= 2 + MT) s2 += e[i];
              if (d >= 3 - MT && d <= 3 + MT) s3 += e[i];
            }
          }
        } else {
          for (int k = -kq; k <= kq; k++) {
            const int cq = min(max(q0 + 4 * k, 0), FI_COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int d = 4 * k + i;
              if (d >= 0 - m && d <= 0 + m) s0 += e[i];
              if (d >= 1 - m && d <= 1 + m) s1 += e[i];
              if (d >= 2 - m && d <= 2 + m) s2 += e[i];
              if (d >= 3 - m && d <= 3 + m) s3 += e[i];
            }
          }
        }
        *reinterpret_cast<float4*>(rrow + ch * FI_COLS + q0) = make_float4(s0, s1, s2, s3);
      }
    }
    __syncthreads();
    // ---------------- B: vertical running sums (double) + solve
#pragma unroll
    for (int rr = 0; rr < FI_CH; rr++) {
      const int tt = tc + rr;
      if (tt > t_last) break;
      const float* rrow = ring + ((tt - t_first) % R) * 5 * FI_COLS + col;
      vs0 += (double)rrow[0 * FI_COLS] - (double)old[rr][0];
      vs1 += (double)rrow[1 * FI_COLS] - (double)old[rr][1];
      vs2 += (double)rrow[2 * FI_COLS] - (double)old[rr][2];
      vs3 += (double)rrow[3 * FI_COLS] - (double)old[rr][3];
      vs4 += (double)rrow[4 * FI_COLS] - (double)old[rr][4];
      const int y = tt - m;
      if (y >= y0 && col_valid) {
        fout[(size_t)y * w + out_x] = solve2x2((float)vs0 * scale, (float)vs1 * scale, (float)vs2 * scale,
                                               (float)vs3 * scale, (float)vs4 * scale);
      }
    }
  }
}

// =====================================================================================
// Stage a8: inter-level flow upsample = resize(prevFlow, INTER_LINEAR) * (1/pyr_scale)
// =====================================================================================
__global__ void __launch_bounds__(256) k_upsample_flow(const float2* __restrict__ prev, int pw, int ph,
                                                       float2* __restrict__ out, int w, int h, double sx, double sy,
                                                       float mul) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const float2* p = prev + (size_t)blockIdx.z * pw * ph;
  int x0, y0;
  float fx, fy;
  linear_coord(x, sx, pw, &x0, &fx);
  linear_coord(y, sy, ph, &y0, &fy);
  int x1 = min(x0 + 1, pw - 1), y1 = min(y0 + 1, ph - 1);
  float2 q00 = __ldg(p + (size_t)y0 * pw + x0), q01 = __ldg(p + (size_t)y0 * pw + x1);
  float2 q10 = __ldg(p + (size_t)y1 * pw + x0), q11 = __ldg(p + (size_t)y1 * pw + x1);
  float ax0 = 1.f - fx, ay0 = 1.f - fy;
  float tx = q00.x * ax0 + q01.x * fx, ty = q00.y * ax0 + q01.y * fx;
  float bx = q10.x * ax0 + q11.x * fx, by = q10.y * ax0 + q11.y * fx;
  out[(size_t)blockIdx.z * w * h + (size_t)y * w + x] = make_float2((tx * ay0 + bx * fy) * mul, (ty * ay0 + by * fy) * mul);
}

// OPTFLOW_USE_INITIAL_FLOW: resize(flow0, INTER_AREA) * scale  (computeResizeAreaTab weights).
__device__ __forceinline__ void area_range(int d, double scale, int src_n, int* s_first, int* s_last, double* w_first,
                                           double* w_mid, double* w_last) {
  double fs1 = d * scale, fs2 = fs1 + scale;
  double cell = fmin(scale, src_n - fs1);
  int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
  s2 = min(s2, src_n - 1);
  s1 = min(s1, s2);
  *w_first = (s1 - fs1 > 1e-3) ? (s1 - fs1) / cell : 0.0;
  *w_mid = 1.0 / cell;
  *w_last = (fs2 - s2 > 1e-3) ? fmin(fmin(fs2 - s2, 1.0), cell) / cell : 0.0;
  *s_first = s1;
  *s_last = s2;
}

__global__ void __launch_bounds__(256) k_init_flow_area(const float2* __restrict__ flow0, int W, int H,
                                                        float2* __restrict__ out, int w, int h, double sx, double sy,
                                                        float mul) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const float2* src = flow0 + (size_t)blockIdx.z * W * H;
  float2 r;
  if (w == W && h == H) {
    r = src[(size_t)y * W + x];
  } else {
    int xa, xb, ya, yb;
    double wxf, wxm, wxl, wyf, wym, wyl;
    area_range(x, sx, W, &xa, &xb, &wxf, &wxm, &wxl);
    area_range(y, sy, H, &ya, &yb, &wyf, &wym, &wyl);
    float ax = 0.f, ay = 0.f;
    for (int yy = ya - 1; yy <= yb; yy++) {
      float wy = yy == ya - 1 ? (float)wyf : (yy == yb ? (float)wyl : (float)wym);
      if (wy == 0.f || yy < 0 || yy >= H) continue;
      float rx = 0.f, ry = 0.f;
      for (int xx = xa - 1; xx <= xb; xx++) {
        float wx = xx == xa - 1 ? (float)wxf : (xx == xb ? (float)wxl : (float)wxm);
        if (wx == 0.f || xx < 0 || xx >= W) continue;
        float2 v = __ldg(src + (size_t)yy * W + xx);
        rx = fmaf(wx, v.x, rx);
        ry = fmaf(wx, v.y, ry);
      }
      ax = fmaf(wy, rx, ax);
      ay = fmaf(wy, ry, ay);
    }
    r = make_float2(ax, ay);
  }
  out[(size_t)blockIdx.z * w * h + (size_t)y * w + x] = make_float2(r.x * mul, r.y * mul);
}

// =====================================================================================
// Driver: the multi-level schedule on the handle's stream (no host sync inside).
// =====================================================================================
static inline dim3 grid2d(int w, int h, int z, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y, z); }

int farneback_run(ofb_handle* h, int n_pairs, bool sequence, const uint8_t* d_prev, const uint8_t* d_next,
                  int width, int height, size_t pitch, size_t image_stride, float* d_flow_out,
                  const float* d_init_flow, const ofb_farneback_params* p) {
  Level sched[kMaxLevels];
  int n_levels = 0;
  if (build_schedule(width, height, p->pyr_scale, p->levels, sched, &n_levels) != OFB_OK)
    return set_error(h, OFB_ERR_INVALID_ARG, "too many pyramid levels");
  PolyCoef pc;
  prepare_poly(p->poly_n, p->poly_sigma, &pc);
  BlurCoef bc;
  prepare_blur(p->winsize, (p->flags & OFB_OPTFLOW_FARNEBACK_GAUSSIAN) != 0, &bc);

  const int frames = sequence ? n_pairs + 1 : 2 * n_pairs;
  const int f1_offset = sequence ? 1 : n_pairs;
  FrameSrc src;
  src.a = d_prev;
  src.b = sequence ? d_prev : d_next;
  src.na = sequence ? frames : n_pairs;
  src.pitch = pitch;
  src.image_stride = image_stride;
  cudaStream_t st = h->stream;
  const dim3 blk(32, 8);
  // fused box-window iteration kernel: radius 2..19 (shared-memory ring of 2m+1 rows); the generic
  // three-kernel path covers the Gaussian window and every other radius.
  const bool use_fused = !bc.gaussian && bc.m >= 2 && bc.m <= 19 && !h->force_generic;
  if (use_fused) {
    const int smem = (FI_CH + 2 * bc.m + 1) * 5 * FI_COLS * (int)sizeof(float);
    if (bc.m == 7)
      OFB_CUDA(h, cudaFuncSetAttribute(k_iter_box<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else
      OFB_CUDA(h, cudaFuncSetAttribute(k_iter_box<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }

  float2* prev_flow = nullptr;
  int prev_w = 0, prev_h = 0;
  for (int li = 0; li < n_levels; li++) {
    const Level& lv = sched[li];
    const int w = lv.width, hh = lv.height;
    const bool last_level = li == n_levels - 1;
    // flow buffers: cur (input of this level) must differ from the buffer holding the previous
    // level's result (read by the upsample); alt may alias it (first written after the upsample).
    float2* cur = prev_flow == h->d_flow[0] ? h->d_flow[1] : h->d_flow[0];
    float2* alt = cur == h->d_flow[0] ? h->d_flow[1] : h->d_flow[0];
#define TB(stage) do { int s__ = timing_begin(h, stage); if (s__) return s__; } while (0)
#define TE() do { int s__ = timing_end(h); if (s__) return s__; } while (0)
    // --- initial flow of the level
    TB(OFB_STAGE_FLOW_INIT);
    if (prev_flow == nullptr) {
      if (p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) {
        k_init_flow_area<<<grid2d(w, hh, n_pairs, blk), blk, 0, st>>>((const float2*)d_init_flow, width, height, cur, w,
                                                                      hh, (double)width / w, (double)height / hh,
                                                                      (float)lv.scale);
        OFB_LAUNCH_CHECK(h);
      } else {
        OFB_CUDA(h, cudaMemsetAsync(cur, 0, (size_t)n_pairs * w * hh * sizeof(float2), st));
      }
    } else {
      k_upsample_flow<<<grid2d(w, hh, n_pairs, blk), blk, 0, st>>>(prev_flow, prev_w, prev_h, cur, w, hh,
                                                                   1.0 / ((double)w / prev_w), 1.0 / ((double)hh / prev_h),
                                                                   (float)(1.0 / p->pyr_scale));
      OFB_LAUNCH_CHECK(h);
    }
    TE();
    // --- pyramid level + polynomial expansion of every frame
    PyrCoef pyc;
    if (prepare_pyr(lv.ksize, lv.sigma, &pyc) != OFB_OK)
      return set_error(h, OFB_ERR_INVALID_ARG, "pyramid smoothing kernel too large (ksize=%d)", lv.ksize);
    TB(OFB_STAGE_PYRAMID);
    k_pyr_level<<<grid2d(w, hh, frames, blk), blk, 0, st>>>(src, width, height, h->d_img, w, hh,
                                                            1.0 / ((double)w / width), 1.0 / ((double)hh / height), pyc);
    OFB_LAUNCH_CHECK(h);
    TE();
    TB(OFB_STAGE_POLYEXP);
    {
      dim3 g((w + PE_T - 1) / PE_T, (hh + PE_T - 1) / PE_T, frames);
      if (pc.n == 5) k_polyexp<5><<<g, blk, 0, st>>>(h->d_img, h->d_RA, h->d_RB, w, hh, pc);
      else if (pc.n == 7) k_polyexp<7><<<g, blk, 0, st>>>(h->d_img, h->d_RA, h->d_RB, w, hh, pc);
      else k_polyexp<0><<<g, blk, 0, st>>>(h->d_img, h->d_RA, h->d_RB, w, hh, pc);
      OFB_LAUNCH_CHECK(h);
    }
    TE();
    // --- iterations
    float2* fin = cur;
    for (int it = 0; it < p->iterations; it++) {
      const bool last_it = it == p->iterations - 1;
      float2* fout = (last_level && last_it) ? (float2*)d_flow_out : (fin == cur ? alt : cur);
      TB(OFB_STAGE_ITERATION);
      if (use_fused) {
        const int tw = FI_COLS - 2 * bc.m;
        const int strips = (w + tw - 1) / tw;
        const int slots = 2 * h->num_sms;
        const int per = strips * n_pairs;
        int segs = per >= slots ? 1 : slots / per;
        int seg_rows = std::max(16, (hh + segs - 1) / segs);
        segs = (hh + seg_rows - 1) / seg_rows;
        const size_t smem = (size_t)(FI_CH + 2 * bc.m + 1) * 5 * FI_COLS * sizeof(float);
        dim3 g(strips * segs, n_pairs);
        if (bc.m == 7)
          k_iter_box<7><<<g, FI_THREADS, smem, st>>>(h->d_RA, h->d_RB, fin, fout, w, hh, f1_offset, bc.m, bc.scale,
                                                     seg_rows, strips);
        else
          k_iter_box<0><<<g, FI_THREADS, smem, st>>>(h->d_RA, h->d_RB, fin, fout, w, hh, f1_offset, bc.m, bc.scale,
                                                     seg_rows, strips);
        OFB_LAUNCH_CHECK(h);
      } else {
        k_update_matrices<<<grid2d(w, hh, n_pairs, blk), blk, 0, st>>>(h->d_RA, h->d_RB, fin, h->d_MA, h->d_MB, w, hh,
                                                                       f1_offset);
        OFB_LAUNCH_CHECK(h);
        k_blur_v<<<grid2d(w, hh, n_pairs, blk), blk, 0, st>>>(h->d_MA, h->d_MB, h->d_VA, h->d_VB, w, hh, bc);
        OFB_LAUNCH_CHECK(h);
        k_blur_h_solve<<<grid2d(w, hh, n_pairs, blk), blk, 0, st>>>(h->d_VA, h->d_VB, fout, w, hh, bc);
        OFB_LAUNCH_CHECK(h);
      }
      TE();
      fin = fout;
    }
    if (p->iterations == 0 && last_level) {
      OFB_CUDA(h, cudaMemcpyAsync(d_flow_out, fin, (size_t)n_pairs * w * hh * sizeof(float2),
                                  cudaMemcpyDeviceToDevice, st));
    }
    prev_flow = fin;
    prev_w = w;
    prev_h = hh;
  }
  h->last_flow = d_flow_out;
  h->last_n = n_pairs;
  h->last_w = width;
  h->last_h = height;
  return OFB_OK;
}

}  // namespace ofb
