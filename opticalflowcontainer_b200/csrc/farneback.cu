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

#include <vector>

#include "common.cuh"
#include "fb_device.cuh"
#include "fb_iter_v.cuh"
#include "fb_iter_launch.cuh"
#include "fb_polyexp.cuh"
#include "fb_pyramid.cuh"

// Experiment builds only (tools/build_variant.sh passes -D flags; the shipped library has one path): ring of the default
// iteration kernel in tensor memory or shared memory, depth of its staging pipeline, fused inter-level upsample.
#ifndef OFB_EXP_TMEM
#define OFB_EXP_TMEM true
#endif
#ifndef OFB_EXP_NBUF
#define OFB_EXP_NBUF 4
#endif
#ifndef OFB_EXP_FUSE_UPS
#define OFB_EXP_FUSE_UPS true
#endif
#ifndef OFB_EXP_PYR3
#define OFB_EXP_PYR3 true
#endif

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
  for (int x = 0; x <= 8; x++) {
    const float a = x <= n ? pc->g[x] : 0.f, b = x <= n ? pc->xg[x] : 0.f, c = x <= n ? pc->xxg[x] : 0.f;
    pc->g2[x] = make_float2(a, a);
    pc->xg2[x] = make_float2(b, b);
    pc->xxg2[x] = make_float2(c, c);
  }
  pc->ig11_2 = make_float2(pc->ig11, pc->ig11);
  pc->ig03_2 = make_float2(pc->ig03, pc->ig03);
  pc->ig33_2 = make_float2(pc->ig33, pc->ig33);
  pc->ig55_2 = make_float2(pc->ig55, pc->ig55);
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

// The three regular levels S = 8, 4, 2 in one pass over the source (k_pyr_fast3): applicable when the schedule ends
// ... 8, 4, 2, 1 with the default smoothing radii 9, 4, 1 and the frames are word-aligned multiples of 8 pixels.
static bool pyr3_applicable(const Level* sched, int n_levels, int width, int height, size_t pitch, const uint8_t* a,
                            const uint8_t* b, PyrFast3Coef* fc3) {
  if (n_levels < 4 || (width & 7) || (height & 7) || (pitch & 3) || (reinterpret_cast<uintptr_t>(a) & 3) ||
      (reinterpret_cast<uintptr_t>(b) & 3))
    return false;
  memset(fc3, 0, sizeof(*fc3));
  for (int q = 0; q < 3; q++) {                          // q = 0, 1, 2: S = 2, 4, 8
    const Level& lq = sched[n_levels - 2 - q];
    const int S = 2 << q, r = q == 0 ? 1 : (q == 1 ? 4 : 9);
    PyrCoef pq;
    if (!(lq.width * S == width && lq.height * S == height && prepare_pyr(lq.ksize, lq.sigma, &pq) == OFB_OK && pq.r == r))
      return false;
    float* c = q == 0 ? fc3->c1 : (q == 1 ? fc3->c2 : fc3->c3);
    for (int j = 1; j <= r + 1; j++) c[j - 1] = 0.5f * (pq.k[j - 1] + (j <= r ? pq.k[j] : 0.f));
  }
  return true;
}

// =====================================================================================
// Stage a4: FarnebackPolyExp for poly_n > 8 (the marching kernel of fb_polyexp.cuh serves 1..8).  32x32 output tile per 256-thread CTA; level image tile with an
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
// Stage a8: inter-level flow upsample = resize(prevFlow, INTER_LINEAR) * (1/pyr_scale)
// =====================================================================================
// One column, UPS_ROWS consecutive rows per thread: sixteen independent gathers in flight per thread
// (the one-pixel version was latency-bound) while a warp still reads and writes contiguous row
// segments (a four-columns-per-thread variant was LSU-bound: 3x the L1 wavefronts).  Source
// coordinates come from per-level tables (LinTab) built once on the host in double, as cv::resize
// computes them.
constexpr int UPS_ROWS = 4;
__global__ void __launch_bounds__(256) k_upsample_flow(const float2* __restrict__ prev, int pw, int ph,
                                                       float2* __restrict__ out, int w, int h,
                                                       const LinTab* __restrict__ tabx,
                                                       const LinTab* __restrict__ taby, float mul) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int yb = (blockIdx.y * blockDim.y + threadIdx.y) * UPS_ROWS;
  if (x >= w || yb >= h) return;
  const float2* p = prev + (size_t)blockIdx.z * pw * ph;
  const LinTab tx = tabx[x];
  const int x0 = tx.i0, x1 = min(x0 + 1, pw - 1);
  const float fx = tx.f;
  float2 q00[UPS_ROWS], q01[UPS_ROWS], q10[UPS_ROWS], q11[UPS_ROWS];
  float fy[UPS_ROWS];
#pragma unroll
  for (int j = 0; j < UPS_ROWS; j++) {
    const LinTab ty = taby[min(yb + j, h - 1)];
    const float2* r0 = p + (size_t)ty.i0 * pw;
    const float2* r1 = p + (size_t)min(ty.i0 + 1, ph - 1) * pw;
    fy[j] = ty.f;
    q00[j] = __ldg(r0 + x0);
    q01[j] = __ldg(r0 + x1);
    q10[j] = __ldg(r1 + x0);
    q11[j] = __ldg(r1 + x1);
  }
  float2* o = out + (size_t)blockIdx.z * w * h + (size_t)yb * w + x;
#pragma unroll
  for (int j = 0; j < UPS_ROWS; j++)
    if (yb + j < h) o[(size_t)j * w] = ups_blend(q00[j], q01[j], q10[j], q11[j], fx, fy[j], mul);
}

// cv::resize INTER_LINEAR source coordinate of destination index d (host twin of linear_coord).
static inline LinTab lin_entry(int d, double scale, int src_n) {
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= src_n - 1) { f = 0.f; s = src_n - 1; }
  LinTab t;
  t.i0 = s;
  t.f = f;
  return t;
}

// (Re)builds the upsample tables when the schedule changed; the upload is ordered on the handle's stream.
static int ensure_lintabs(ofb_handle* h, const Level* sched, int n_levels, int width, int height, double pyr_scale) {
  if (h->tab_w == width && h->tab_h == height && h->tab_levels == n_levels && h->tab_scale == pyr_scale) return OFB_OK;
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));   // the pinned table may still be in flight from the last rebuild
  size_t off = 0;
  for (int li = 1; li < n_levels; li++) {
    const int w = sched[li].width, hh = sched[li].height, pw = sched[li - 1].width, ph = sched[li - 1].height;
    if (off + (size_t)w + hh > h->lintab_cap) return set_error(h, OFB_ERR_CAPACITY, "resize table capacity exceeded");
    const double sx = 1.0 / ((double)w / pw), sy = 1.0 / ((double)hh / ph);
    h->tab_x_off[li] = off;
    for (int x = 0; x < w; x++) h->h_lintab[off + x] = lin_entry(x, sx, pw);
    off += w;
    h->tab_y_off[li] = off;
    for (int y = 0; y < hh; y++) h->h_lintab[off + y] = lin_entry(y, sy, ph);
    off += hh;
  }
  if (off) OFB_CUDA(h, cudaMemcpyAsync(h->d_lintab, h->h_lintab, off * sizeof(LinTab), cudaMemcpyHostToDevice, h->stream));
  h->tab_w = width;
  h->tab_h = height;
  h->tab_levels = n_levels;
  h->tab_scale = pyr_scale;
  return OFB_OK;
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
int farneback_levels(int width, int height, const ofb_farneback_params* p, Level* out, int* n_out) {
  return build_schedule(width, height, p->pyr_scale, p->levels, out, n_out);
}

bool farneback_stream_supported(const ofb_handle* h, const ofb_farneback_params* p) {
  (void)h;
  BlurCoef bc;
  prepare_blur(p->winsize, (p->flags & OFB_OPTFLOW_FARNEBACK_GAUSSIAN) != 0, &bc);
  const bool use_fused = !bc.gaussian && bc.m >= 2 && bc.m <= 19;
  return use_fused && p->poly_n <= PX_MAXN && !(p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW);
}

static inline dim3 grid2d(int w, int h, int z, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y, z); }

static int farneback_run_impl(ofb_handle* h, int n_pairs, bool sequence, const uint8_t* d_prev, const uint8_t* d_next,
                              int width, int height, size_t pitch, size_t image_stride, float* d_flow_out,
                              const float* d_init_flow, const ofb_farneback_params* p, const StreamCtx* sc);

// Small batches are launch-bound (a VGA pair is ~40 kernels of a few microseconds each): the launch sequence of a call
// is captured once per distinct argument set into a CUDA graph and replayed.  The key holds everything the launches
// depend on (pointers, geometry, parameters, stream-cache half); coordinate tables are brought up to date outside the
// graph.  Not used while the stage timers are on (event timing cannot be captured) or with OFB_GRAPH=0.
struct GraphKey {
  int n_pairs, sequence, width, height, cur, prime_only;
  const void *d_prev, *d_next, *d_flow_out, *d_init_flow;
  const void* sc_pool;   // base of the stream cache the launches were captured with (reallocated when it grows)
  size_t pitch, image_stride;
  ofb_farneback_params p;
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t exec;
  uint64_t launches, stamp;
};
constexpr int kGraphMaxPairs = 4, kGraphCache = 16;

void farneback_graphs_destroy(ofb_handle* h) {
  auto* v = static_cast<std::vector<GraphEntry>*>(h->graph_cache);
  if (!v) return;
  for (auto& e : *v) cudaGraphExecDestroy(e.exec);
  delete v;
  h->graph_cache = nullptr;
}

int farneback_run(ofb_handle* h, int n_pairs, bool sequence, const uint8_t* d_prev, const uint8_t* d_next,
                  int width, int height, size_t pitch, size_t image_stride, float* d_flow_out,
                  const float* d_init_flow, const ofb_farneback_params* p, const StreamCtx* sc) {
  {
    Level sched[kMaxLevels];
    int n_levels = 0;
    if (build_schedule(width, height, p->pyr_scale, p->levels, sched, &n_levels) != OFB_OK)
      return set_error(h, OFB_ERR_INVALID_ARG, "too many pyramid levels");
    int s__ = ensure_lintabs(h, sched, n_levels, width, height, p->pyr_scale);
    if (s__) return s__;
  }
  if (h->no_graph || h->timing || n_pairs > kGraphMaxPairs || h->graph_bypass)
    return farneback_run_impl(h, n_pairs, sequence, d_prev, d_next, width, height, pitch, image_stride, d_flow_out,
                              d_init_flow, p, sc);
  if (!h->graph_cache) h->graph_cache = new std::vector<GraphEntry>();
  auto& cache = *static_cast<std::vector<GraphEntry>*>(h->graph_cache);
  GraphKey key;
  memset(&key, 0, sizeof(key));
  key.n_pairs = n_pairs; key.sequence = sequence; key.width = width; key.height = height;
  key.cur = sc ? sc->cur : -1; key.prime_only = sc ? (int)sc->prime_only : 0;
  key.sc_pool = sc ? (const void*)sc->RA[0][0] : nullptr;
  key.d_prev = d_prev; key.d_next = d_next; key.d_flow_out = d_flow_out; key.d_init_flow = d_init_flow;
  key.pitch = pitch; key.image_stride = image_stride;
  key.p.pyr_scale = p->pyr_scale; key.p.levels = p->levels; key.p.winsize = p->winsize; key.p.iterations = p->iterations;
  key.p.poly_n = p->poly_n; key.p.poly_sigma = p->poly_sigma; key.p.flags = p->flags;
  for (auto& e : cache)
    if (e.key == key) {
      e.stamp = ++h->graph_clock;
      OFB_CUDA(h, cudaGraphLaunch(e.exec, h->stream));
      h->launches += e.launches;
      h->last_flow = d_flow_out; h->last_n = n_pairs; h->last_w = width; h->last_h = height;
      return OFB_OK;
    }
  const uint64_t l0 = h->launches;
  OFB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = farneback_run_impl(h, n_pairs, sequence, d_prev, d_next, width, height, pitch, image_stride, d_flow_out,
                                    d_init_flow, p, sc);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
  if (rc != OFB_OK) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc;
  }
  if (ce != cudaSuccess || !graph) return set_error(h, OFB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
  GraphEntry e;
  e.key = key;
  e.launches = h->launches - l0;
  e.stamp = ++h->graph_clock;
  const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) return set_error(h, OFB_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
  if ((int)cache.size() >= kGraphCache) {             // evict the least recently used
    size_t lru = 0;
    for (size_t i = 1; i < cache.size(); i++) if (cache[i].stamp < cache[lru].stamp) lru = i;
    cudaGraphExecDestroy(cache[lru].exec);
    cache[lru] = e;
  } else {
    cache.push_back(e);
  }
  OFB_CUDA(h, cudaGraphLaunch(e.exec, h->stream));
  return OFB_OK;
}

static int farneback_run_impl(ofb_handle* h, int n_pairs, bool sequence, const uint8_t* d_prev, const uint8_t* d_next,
                              int width, int height, size_t pitch, size_t image_stride, float* d_flow_out,
                              const float* d_init_flow, const ofb_farneback_params* p, const StreamCtx* sc) {
  Level sched[kMaxLevels];
  int n_levels = 0;
  if (build_schedule(width, height, p->pyr_scale, p->levels, sched, &n_levels) != OFB_OK)
    return set_error(h, OFB_ERR_INVALID_ARG, "too many pyramid levels");
  PolyCoef pc;
  prepare_poly(p->poly_n, p->poly_sigma, &pc);
  BlurCoef bc;
  prepare_blur(p->winsize, (p->flags & OFB_OPTFLOW_FARNEBACK_GAUSSIAN) != 0, &bc);

  // camera-stream call (sc): only the n new frames (d_next) are expanded; the previous frames' expansions of every
  // level are the ones the previous call left in the other half of the stream cache
  const int frames = sc ? n_pairs : (sequence ? n_pairs + 1 : 2 * n_pairs);
  const int f1_offset = sequence ? 1 : n_pairs;
  FrameSrc src;
  src.a = sc ? d_next : d_prev;
  src.b = sequence ? d_prev : d_next;
  src.na = (sc || sequence) ? frames : n_pairs;
  src.pitch = pitch;
  src.image_stride = image_stride;
  cudaStream_t st = h->stream;
  const dim3 blk(32, 8);
  // fused box-window iteration kernel: radius 2..19 (ring of 2m+1 rows in tensor memory or shared memory); the generic
  // three-kernel path covers the Gaussian window and every other radius.
  const bool use_fused = !bc.gaussian && bc.m >= 2 && bc.m <= 19;
  if (sc && (!use_fused || pc.n > PX_MAXN || n_levels > kMaxLevels))
    return set_error(h, OFB_ERR_INVALID_ARG, "internal: stream cache used with an unsupported configuration");
  float2* prev_flow = nullptr;
  int prev_w = 0, prev_h = 0;
  // The three regular levels S = 8, 4, 2 in one pass over the source (k_pyr_fast3): when the schedule ends
  // ... 8, 4, 2, 1 with the default smoothing radii 9, 4, 1.  Their level images then live side by side in d_img
  // (which holds frames * N floats; the three take 21/64 of it).
  int fused3_li = -1;
  float* img3[3] = {nullptr, nullptr, nullptr};          // S = 2, 4, 8
  PyrFast3Coef fc3;
  if (OFB_EXP_PYR3 && (image_stride & 3) == 0 && pc.n <= PX_MAXN) {
    const bool ok = pyr3_applicable(sched, n_levels, width, height, pitch, src.a, src.b, &fc3);
    if (ok) {
      fused3_li = n_levels - 4;
      const size_t n1 = (size_t)(width / 2) * (height / 2), n2 = n1 / 4;
      img3[0] = h->d_img;
      img3[1] = img3[0] + (size_t)frames * n1;
      img3[2] = img3[1] + (size_t)frames * n2;
    }
  }
  // Expansions beside iterations: the coarse levels' iteration launches are small and latency-bound (a few CTAs
  // marching a few rows), while the expansions of the finer levels depend on nothing but the level images.  In the
  // default schedule (four regular levels, all level images from the one k_pyr_fast3 pass) the expansions of levels
  // 1 .. 3 are launched on the handle's expansion stream right after the pyramid pass and fill the SMs the coarse
  // iterations leave idle; the compute stream waits for a level's expansions just before that level's iterations.  The
  // coarse levels then need expansion buffers of their own (d_MA / d_MB, unused by the fused path; the camera-stream
  // cache has per-level buffers anyway).  Off while the stage timers run (their events assume one stream).
  const bool overlap = !h->no_overlap && !h->timing && use_fused && fused3_li == 0 && n_levels == 4 && h->s_px != nullptr;
  size_t coarse_off[kMaxLevels] = {0};
  if (overlap && !sc) {
    size_t off = 0;
    for (int li = 0; li < n_levels - 1; li++) {
      coarse_off[li] = off;
      off += (size_t)frames * sched[li].width * sched[li].height + (size_t)kRowPad * sched[li].width;
    }
  }
  auto level_RA = [&](int li) -> float4* {
    if (sc) return sc->RA[sc->cur][li];
    return (overlap && li < n_levels - 1) ? h->d_MA + coarse_off[li] : h->d_RA;
  };
  auto level_RB = [&](int li) -> float* {
    if (sc) return sc->RB[sc->cur][li];
    return (overlap && li < n_levels - 1) ? h->d_MB + coarse_off[li] : h->d_RB;
  };
  // marching PolyExp of level li (level image -> expansions) on stream s
  auto launch_px = [&](int li, cudaStream_t s, const float* level_img, bool fused_src, const PyrCoef& pyc) -> int {
    const int w = sched[li].width, hh = sched[li].height;
    float4* const RAw = level_RA(li);
    float* const RBw = level_RB(li);
    const int strips = (w + PX_TW - 1) / PX_TW;
    const int per = strips * frames;
    const int slots = 3 * h->num_sms * kPxWaves;
    int segs = std::max(1, slots / per);
    int seg_rows = std::max(16, ((hh + segs - 1) / segs + PX_ROWS - 1) / PX_ROWS * PX_ROWS);
    segs = (hh + seg_rows - 1) / seg_rows;
    dim3 g(strips * segs, frames);
#define OFB_PX_LAUNCH(NT)                                                                                           \
  do {                                                                                                              \
    if (fused_src)                                                                                                  \
      k_polyexp_march<NT, 1><<<g, PX_COLS, 0, s>>>(nullptr, src, pyc.k[0], pyc.k[1], RAw, RBw, w, hh,               \
                                                   seg_rows, strips, pc, 0, hh);                                    \
    else                                                                                                            \
      k_polyexp_march<NT, 0><<<g, PX_COLS, 0, s>>>(level_img, src, 0.f, 0.f, RAw, RBw, w, hh, seg_rows,             \
                                                   strips, pc, 0, hh);                                              \
  } while (0)
    if (pc.n == 5) OFB_PX_LAUNCH(5); else if (pc.n == 7) OFB_PX_LAUNCH(7); else OFB_PX_LAUNCH(0);
#undef OFB_PX_LAUNCH
    OFB_LAUNCH_CHECK(h);
    return OFB_OK;
  };
  for (int li = 0; li < n_levels; li++) {
    const Level& lv = sched[li];
    const int w = lv.width, hh = lv.height;
    const bool last_level = li == n_levels - 1;
    const size_t npx = (size_t)w * hh;
    float4* const RAw = level_RA(li);                            // where this level's expansions are written
    float* const RBw = level_RB(li);
    const RSet rs = sc ? RSet{sc->RA[sc->cur ^ 1][li], sc->RB[sc->cur ^ 1][li], RAw, RBw}
                       : RSet{RAw, RBw, RAw + (size_t)f1_offset * npx, RBw + (size_t)f1_offset * npx};
    const bool prime_only = sc && sc->prime_only;
    // flow buffers: cur (input of this level) must differ from the buffer holding the previous
    // level's result (read by the upsample); alt may alias it (first written after the upsample).
    float2* cur = prev_flow == h->d_flow[0] ? h->d_flow[1] : h->d_flow[0];
    float2* alt = cur == h->d_flow[0] ? h->d_flow[1] : h->d_flow[0];
#define TB(stage) do { int s__ = timing_begin(h, stage); if (s__) return s__; } while (0)
#define TE() do { int s__ = timing_end(h); if (s__) return s__; } while (0)
    // --- initial flow of the level.  With the fused iteration kernel the x2 upsample of the coarser level's result is
    // computed by the first iteration's producers (UpsSrc): no launch, no write + read of the upsampled field.
    const bool fuse_ups = OFB_EXP_FUSE_UPS && use_fused && bc.m == 7 && prev_flow != nullptr && p->iterations > 0;   // (default window: the UPS instantiation)
    TB(OFB_STAGE_FLOW_INIT);
    if (prime_only || fuse_ups) {
      // first frame of the streams: expansions only / upsample fused into the first iteration
    } else if (prev_flow == nullptr) {
      if (p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) {
        k_init_flow_area<<<grid2d(w, hh, n_pairs, blk), blk, 0, st>>>((const float2*)d_init_flow, width, height, cur, w,
                                                                      hh, (double)width / w, (double)height / hh,
                                                                      (float)lv.scale);
        OFB_LAUNCH_CHECK(h);
      } else {
        OFB_CUDA(h, cudaMemsetAsync(cur, 0, (size_t)n_pairs * w * hh * sizeof(float2), st));
      }
    } else {
      k_upsample_flow<<<grid2d(w, (hh + UPS_ROWS - 1) / UPS_ROWS, n_pairs, blk), blk, 0, st>>>(
          prev_flow, prev_w, prev_h, cur, w, hh, h->d_lintab + h->tab_x_off[li], h->d_lintab + h->tab_y_off[li],
          (float)(1.0 / p->pyr_scale));
      OFB_LAUNCH_CHECK(h);
    }
    TE();
    // --- pyramid level + polynomial expansion of every frame
    PyrCoef pyc;
    if (prepare_pyr(lv.ksize, lv.sigma, &pyc) != OFB_OK)
      return set_error(h, OFB_ERR_INVALID_ARG, "pyramid smoothing kernel too large (ksize=%d)", lv.ksize);
    // Marching PolyExp kernel (poly_n <= 8).  When the level has the source size (k = 0: 3-tap blur,
    // identity resize) the pyramid stage is fused into it and the level image never exists in HBM.
    const bool march = pc.n <= PX_MAXN;
    const bool fused_src = march && w == width && hh == height && pyc.r == 1 && pyc.k[0] == 0.5f && pyc.k[1] == 0.25f;
    // regular power-of-two level (fb_pyramid.cuh, k_pyr_fast): one kernel, source read once
    int fastS = 0;
    if (!fused_src && (width & 3) == 0 && (pitch & 3) == 0 && (image_stride & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(src.a) & 3) == 0 && (reinterpret_cast<uintptr_t>(src.b) & 3) == 0) {
      for (int S = 2; S <= 8; S *= 2)
        if (w * S == width && hh * S == height && pyc.r == (S == 2 ? 1 : (S == 4 ? 4 : 9))) fastS = S;
    }
    const float* level_img = h->d_img;
    if (fused3_li >= 0 && li >= fused3_li && li < n_levels - 1) {
      level_img = img3[n_levels - 2 - li];
      if (li == fused3_li) {
        TB(OFB_STAGE_PYRAMID);
        constexpr int out3 = (PF_COLS - 2 * PF_HALO) / 8;
        const int w3 = width / 8, h3 = height / 8;
        const int chunks = (w3 + out3 - 1) / out3;
        // about two waves of CTAs at 3 resident per SM (the window takes ~70 registers)
        const int segs = std::max(1, 2 * 3 * h->num_sms / (chunks * frames));
        const int seg_rows = std::max(2, (h3 + segs - 1) / segs);
        dim3 g(chunks, (h3 + seg_rows - 1) / seg_rows, frames);
        k_pyr_fast3<<<g, PF_THREADS, 0, st>>>(src, width, height, img3[0], img3[1], img3[2], fc3, seg_rows, 0, h3);
        OFB_LAUNCH_CHECK(h);
        TE();
      }
    } else if (fastS) {
      TB(OFB_STAGE_PYRAMID);
      PyrFastCoef fc;
      memset(&fc, 0, sizeof(fc));
      for (int j = 1; j <= pyc.r + 1; j++) fc.c[j - 1] = 0.5f * (pyc.k[j - 1] + (j <= pyc.r ? pyc.k[j] : 0.f));
      const int out_per = (PF_COLS - 2 * PF_HALO) / fastS;
      const int chunks = (w + out_per - 1) / out_per;
      // resident CTAs per SM by register use: 8 (S=2, 32 regs), 4 (S=4), 3 (S=8); two waves of short segments hide
      // the per-row barrier better than one wave of long ones
      const int per_sm = fastS == 2 ? 8 : (fastS == 4 ? 4 : 3);
      const int segs = std::max(1, 2 * per_sm * h->num_sms / (chunks * frames));
      const int seg_rows = std::max(4, (hh + segs - 1) / segs);
      dim3 g(chunks, (hh + seg_rows - 1) / seg_rows, frames);
      if (fastS == 2) k_pyr_fast<2, 1><<<g, PF_THREADS, 0, st>>>(src, width, height, h->d_img, w, hh, fc, seg_rows);
      else if (fastS == 4) k_pyr_fast<4, 4><<<g, PF_THREADS, 0, st>>>(src, width, height, h->d_img, w, hh, fc, seg_rows);
      else k_pyr_fast<8, 9><<<g, PF_THREADS, 0, st>>>(src, width, height, h->d_img, w, hh, fc, seg_rows);
      OFB_LAUNCH_CHECK(h);
      TE();
    } else if (!fused_src) {
      TB(OFB_STAGE_PYRAMID);
      // pass H writes hb[frames][H][w] into d_MA (free here: the generic iteration path only uses it
      // after the pyramid stage of the level), pass V writes the level image.
      float* hb = reinterpret_cast<float*>(h->d_MA);
      dim3 gh((w + 127) / 128, (height + PYR_RPT - 1) / PYR_RPT, frames);
      dim3 bv(128, 2), gv((w + 127) / 128, (hh + 1) / 2, frames);
#define OFB_PYR_LAUNCH(RT)                                                                             \
  do {                                                                                                 \
    k_pyr_h<RT><<<gh, 128, 0, st>>>(src, width, height, hb, w, 1.0 / ((double)w / width), pyc, 0, height); \
    OFB_LAUNCH_CHECK(h);                                                                               \
    k_pyr_v<RT><<<gv, bv, 0, st>>>(hb, height, h->d_img, w, hh, 1.0 / ((double)hh / height), pyc, 0, hh); \
    OFB_LAUNCH_CHECK(h);                                                                               \
  } while (0)
      if (pyc.r == 1) OFB_PYR_LAUNCH(1);
      else if (pyc.r == 4) OFB_PYR_LAUNCH(4);
      else if (pyc.r == 9) OFB_PYR_LAUNCH(9);
      else OFB_PYR_LAUNCH(0);
#undef OFB_PYR_LAUNCH
      TE();
    }
    TB(OFB_STAGE_POLYEXP);
    if (march && overlap) {
      if (li == 0) {
        // fork: the finer levels' expansions on the expansion stream, behind the pyramid pass
        OFB_CUDA(h, cudaEventRecord(h->ev_fork, st));
        OFB_CUDA(h, cudaStreamWaitEvent(h->s_px, h->ev_fork, 0));
        for (int l = 1; l < n_levels; l++) {
          PyrCoef pl;
          if (prepare_pyr(sched[l].ksize, sched[l].sigma, &pl) != OFB_OK) return set_error(h, OFB_ERR_INVALID_ARG, "pyramid smoothing kernel too large");
          const bool fsrc = l == n_levels - 1;
          int s__ = launch_px(l, h->s_px, fsrc ? nullptr : img3[n_levels - 2 - l], fsrc, pl);
          if (s__) return s__;
          OFB_CUDA(h, cudaEventRecord(h->ev_px[l], h->s_px));
        }
        int s__ = launch_px(0, st, level_img, false, pyc);
        if (s__) return s__;
      } else {
        OFB_CUDA(h, cudaStreamWaitEvent(st, h->ev_px[li], 0));   // join: this level's expansions are complete
      }
    } else if (march) {
      int s__ = launch_px(li, st, level_img, fused_src, pyc);
      if (s__) return s__;
    } else {
      dim3 g((w + PE_T - 1) / PE_T, (hh + PE_T - 1) / PE_T, frames);
      if (pc.n == 5) k_polyexp<5><<<g, blk, 0, st>>>(h->d_img, h->d_RA, h->d_RB, w, hh, pc);
      else if (pc.n == 7) k_polyexp<7><<<g, blk, 0, st>>>(h->d_img, h->d_RA, h->d_RB, w, hh, pc);
      else k_polyexp<0><<<g, blk, 0, st>>>(h->d_img, h->d_RA, h->d_RB, w, hh, pc);
      OFB_LAUNCH_CHECK(h);
    }
    TE();
    // --- iterations
    // (fused upsample: the first iteration reads the coarser level's buffer — which `alt` aliases — and must not
    // write it, so it writes `cur`)
    float2* fin = fuse_ups ? alt : cur;
    for (int it = 0; it < (prime_only ? 0 : p->iterations); it++) {
      const bool last_it = it == p->iterations - 1;
      float2* fout = (last_level && last_it) ? (float2*)d_flow_out : (fin == cur ? alt : cur);
      TB(OFB_STAGE_ITERATION);
      if (use_fused) {
        // k_iter_v: float van Herk / Gil-Werman vertical sums, no FP64
        const float reg = (float)(1e-3 / ((double)bc.scale * (double)bc.scale));
        UpsSrc ups;
        const UpsSrc* up = nullptr;
        if (fuse_ups && it == 0) {
          ups.prev = prev_flow; ups.pw = prev_w; ups.ph = prev_h;
          ups.tabx = h->d_lintab + h->tab_x_off[li]; ups.taby = h->d_lintab + h->tab_y_off[li];
          ups.mul = (float)(1.0 / p->pyr_scale);
          ups.exact2y = hh == 2 * prev_h;
          up = &ups;
        }
        cudaError_t e;
        if (bc.m == 7) {
          if (up) e = launch_iter_v<7, 256, 2, 2, 0, false, true, OFB_EXP_TMEM, OFB_EXP_NBUF, true>(h, fin, fout, w, hh, n_pairs, rs, bc.m, reg, st, up);
          else e = launch_iter_v<7, 256, 2, 2, 0, false, true, OFB_EXP_TMEM, OFB_EXP_NBUF>(h, fin, fout, w, hh, n_pairs, rs, bc.m, reg, st, nullptr);
        } else {
          // other window sizes: the radius as a template argument where an instantiation exists (winsize 5..31), else
          // at run time (its generic consumer loop is slow: winsize 13 measured 6.4 ms per 18 pairs against 3.0 ms)
          bool served = false;
          e = bc.m <= 8 ? launch_iter_fixed_a(h, bc.m, fin, fout, w, hh, n_pairs, rs, reg, st, &served)
                        : launch_iter_fixed_b(h, bc.m, fin, fout, w, hh, n_pairs, rs, reg, st, &served);
          if (!served) e = launch_iter_v<0, 128, 2, 3, 0, false, true, false, 2>(h, fin, fout, w, hh, n_pairs, rs, bc.m, reg, st, nullptr);
        }
        if (e != cudaSuccess)
          return set_error(h, OFB_ERR_CUDA, "k_iter_v launch failed: %s", cudaGetErrorString(e));
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
    if (p->iterations == 0 && last_level && !prime_only) {
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

#include "tiled.cuh"
