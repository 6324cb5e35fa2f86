// ingest.cu — frame ingest on the device: the step in front of the flow call in
// every node of the reference is cv_bridge -> cv2.cvtColor(BGR2GRAY) (-> cv2.resize) on the CPU
//   ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:148-159
//   ros2_ws/src/optical_flow/optical_flow/opticalflow_node.py:43-50
// This converts an interleaved 8-bit colour frame (sensor_msgs/Image `bgr8` / `rgb8`, row `step`) to
// the single-channel uint8 frame the flow engine consumes, bit-exactly as this cv2 build does it:
//   gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15          (15-bit fixed point)
// so the 3N-byte colour frame is uploaded once and never converted on the host.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "fb_device.cuh"

namespace ofb {

// One thread = 4 consecutive pixels: three 32-bit loads (12 bytes) when aligned, one 32-bit store.
__global__ void __launch_bounds__(256) k_bgr_to_gray(const uint8_t* __restrict__ src, size_t src_pitch,
                                                     uint8_t* __restrict__ dst, size_t dst_pitch, int w, int h,
                                                     int cb, int cg, int cr) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x4 >= w || y >= h) return;
  const uint8_t* s = src + (size_t)y * src_pitch + (size_t)x4 * 3;
  uint8_t* d = dst + (size_t)y * dst_pitch + x4;
  uint8_t px[12];
  const int n = min(4, w - x4);
  if (n == 4 && ((reinterpret_cast<uintptr_t>(s) & 3) == 0)) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
    const uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);
    px[0] = a; px[1] = a >> 8; px[2] = a >> 16; px[3] = a >> 24;
    px[4] = b; px[5] = b >> 8; px[6] = b >> 16; px[7] = b >> 24;
    px[8] = c; px[9] = c >> 8; px[10] = c >> 16; px[11] = c >> 24;
  } else {
    for (int i = 0; i < 3 * n; i++) px[i] = __ldg(s + i);
  }
  uint8_t g[4];
#pragma unroll
  for (int j = 0; j < 4; j++)
    g[j] = (uint8_t)((px[3 * j] * cb + px[3 * j + 1] * cg + px[3 * j + 2] * cr + 16384) >> 15);
  if (n == 4 && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
    *reinterpret_cast<uint32_t*>(d) = g[0] | (g[1] << 8) | (g[2] << 16) | ((uint32_t)g[3] << 24);
  } else {
    for (int j = 0; j < n; j++) d[j] = g[j];
  }
}

int cvt_gray_device(ofb_handle* h, const uint8_t* d_src, size_t src_pitch, uint8_t* d_dst, size_t dst_pitch, int w,
                    int hh, int rgb_order) {
  // channel 0 weight / channel 1 / channel 2: BGR -> (3735, 19235, 9798); RGB swaps the outer two
  const int c0 = rgb_order ? 9798 : 3735, c2 = rgb_order ? 3735 : 9798;
  dim3 g(((w + 3) / 4 + 255) / 256, hh);
  k_bgr_to_gray<<<g, 256, 0, h->stream>>>(d_src, src_pitch, d_dst, dst_pitch, w, hh, c0, 19235, c2);
  OFB_LAUNCH_CHECK(h);
  return OFB_OK;
}

// ---- cv2.resize(frame, (w, h)) — INTER_LINEAR on uint8, as this cv2 build computes it -----------------------------
// (lfn3_sub_node.py:152-153, lfn3_adapt_node.py:160-161: frames that do not have the configured size are resized before
// anything else.)  OpenCV's resizeGeneric_ / HResizeLinear / VResizeLinear<uchar,int,short>: 11-bit fixed-point weights
// cvRound(w * 2048) of the float coordinate (d + 0.5) * scale - 0.5; columns clamp the coordinate, rows clip the two row
// indices; out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  Restated and pinned in
// oracle/resize_np.py.  Tables are built on the host (double/float arithmetic exactly as cv2's) and cached per geometry.
struct ResizeTab { int i0, i1; int w0, w1; };

static void build_resize_tab(int dn, int sn, bool clamp_coord, ResizeTab* t) {
  const double scale = (double)sn / dn;
  for (int d = 0; d < dn; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_coord) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    t[d].i0 = std::min(std::max(s, 0), sn - 1);
    t[d].i1 = std::min(std::max(s + 1, 0), sn - 1);
    t[d].w0 = (int)__builtin_nearbyintf((1.f - f) * 2048.f);
    t[d].w1 = (int)__builtin_nearbyintf(f * 2048.f);
  }
}

template <int CN>
__global__ void __launch_bounds__(256) k_resize_u8(const uint8_t* __restrict__ src, size_t sp, uint8_t* __restrict__ dst,
                                                   size_t dp, int dw, int dh, const ResizeTab* __restrict__ xt,
                                                   const ResizeTab* __restrict__ yt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const ResizeTab tx = xt[x], ty = yt[y];
  const uint8_t* r0 = src + (size_t)ty.i0 * sp;
  const uint8_t* r1 = src + (size_t)ty.i1 * sp;
#pragma unroll
  for (int c = 0; c < CN; c++) {
    const int s0 = (int)__ldg(r0 + tx.i0 * CN + c) * tx.w0 + (int)__ldg(r0 + tx.i1 * CN + c) * tx.w1;
    const int s1 = (int)__ldg(r1 + tx.i0 * CN + c) * tx.w0 + (int)__ldg(r1 + tx.i1 * CN + c) * tx.w1;
    dst[(size_t)y * dp + x * CN + c] = (uint8_t)((((ty.w0 * (s0 >> 4)) >> 16) + ((ty.w1 * (s1 >> 4)) >> 16) + 2) >> 2);
  }
}

// src (sw x sh x cn, device) -> dst (dw x dh x cn, device), asynchronous on the handle's stream
int resize_u8_device(ofb_handle* h, const uint8_t* d_src, size_t sp, int sw, int sh, int cn, uint8_t* d_dst, size_t dp,
                     int dw, int dh) {
  if (cn != 1 && cn != 3) return set_error(h, OFB_ERR_INVALID_ARG, "resize: 1 or 3 channels");
  ofb_handle::Ingest& g = h->ingest;
  if (g.tab_sw != sw || g.tab_sh != sh || g.tab_dw != dw || g.tab_dh != dh) {
    const size_t n = (size_t)dw + dh;
    if (n > g.tab_cap) {
      OFB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (g.d_tab) cudaFree(g.d_tab);
      if (g.h_tab) cudaFreeHost(g.h_tab);
      g.d_tab = nullptr; g.h_tab = nullptr; g.tab_cap = 0;
      OFB_CUDA(h, cudaMalloc(&g.d_tab, n * sizeof(ResizeTab)));
      OFB_CUDA(h, cudaHostAlloc(&g.h_tab, n * sizeof(ResizeTab), cudaHostAllocDefault));
      g.tab_cap = n;
    } else {
      OFB_CUDA(h, cudaStreamSynchronize(h->stream));   // the pinned table may still be in flight
    }
    ResizeTab* t = static_cast<ResizeTab*>(g.h_tab);
    build_resize_tab(dw, sw, true, t);
    build_resize_tab(dh, sh, false, t + dw);
    OFB_CUDA(h, cudaMemcpyAsync(g.d_tab, t, n * sizeof(ResizeTab), cudaMemcpyHostToDevice, h->stream));
    g.tab_sw = sw; g.tab_sh = sh; g.tab_dw = dw; g.tab_dh = dh;
  }
  const ResizeTab* xt = static_cast<const ResizeTab*>(g.d_tab);
  dim3 grid((dw + 255) / 256, dh);
  if (cn == 1) k_resize_u8<1><<<grid, 256, 0, h->stream>>>(d_src, sp, d_dst, dp, dw, dh, xt, xt + dw);
  else k_resize_u8<3><<<grid, 256, 0, h->stream>>>(d_src, sp, d_dst, dp, dw, dh, xt, xt + dw);
  OFB_LAUNCH_CHECK(h);
  return OFB_OK;
}

// ---- cv2.createCLAHE(clipLimit, tileGridSize).apply(img) on uint8 ---------------------------------------------------
// (the adapt node's contrast pre-filter, lfn3_adapt_node.py:164-182.)  OpenCV's clahe.cpp restated in
// oracle/clahe_np.py and pinned against the wheel: per-tile histogram (the image is extended to a multiple of the grid
// with REFLECT_101), integer clip limit, excess redistributed, LUT = cvRound(cumsum * 255.f / tileArea); then every
// pixel blends the LUTs of the four surrounding tiles in float with separate roundings (the cv2 build has no FMA here).
__global__ void __launch_bounds__(256) k_clahe_lut(const uint8_t* __restrict__ src, size_t sp, int w, int h, int tw, int th,
                                                   int clip, float lut_scale, uint8_t* __restrict__ luts) {
  __shared__ int hist[256];
  __shared__ int scan[256];
  __shared__ int s_clipped;
  const int t = threadIdx.x;
  hist[t] = 0;
  if (t == 0) s_clipped = 0;
  __syncthreads();
  const int x0 = blockIdx.x * tw, y0 = blockIdx.y * th;
  for (int i = t; i < tw * th; i += 256) {
    const int ey = y0 + i / tw, ex = x0 + i % tw;
    const int sy = ey < h ? ey : 2 * (h - 1) - ey, sx = ex < w ? ex : 2 * (w - 1) - ex;   // REFLECT_101 of the extension
    atomicAdd(&hist[__ldg(src + (size_t)sy * sp + sx)], 1);
  }
  __syncthreads();
  int v = hist[t];
  if (clip > 0) {
    if (v > clip) { atomicAdd(&s_clipped, v - clip); v = clip; }
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    const int residual = clipped - batch * 256;
    v += batch;
    if (residual != 0) {
      const int step = max(256 / residual, 1);
      if (t % step == 0 && t / step < residual) v++;
    }
  }
  scan[t] = v;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {            // inclusive scan
    const int a = t >= d ? scan[t - d] : 0;
    __syncthreads();
    scan[t] += a;
    __syncthreads();
  }
  const int r = __float2int_rn(__fmul_rn((float)scan[t], lut_scale));
  luts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + t] = (uint8_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) k_clahe_interp(const uint8_t* __restrict__ src, size_t sp, uint8_t* __restrict__ dst,
                                                      size_t dp, int w, int h, int tiles_x, int tiles_y, float inv_tw,
                                                      float inv_th, const uint8_t* __restrict__ luts) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
  const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
  const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
  const float xa1 = __fsub_rn(1.f, xa), ya1 = __fsub_rn(1.f, ya);
  const int tx2 = min(tx1 + 1, tiles_x - 1), ty2 = min(ty1 + 1, tiles_y - 1);
  tx1 = max(tx1, 0);
  ty1 = max(ty1, 0);
  const int v = src[(size_t)y * sp + x];
  const uint8_t* p1 = luts + (size_t)ty1 * tiles_x * 256 + v;
  const uint8_t* p2 = luts + (size_t)ty2 * tiles_x * 256 + v;
  const float l11 = (float)__ldg(p1 + tx1 * 256), l12 = (float)__ldg(p1 + tx2 * 256);
  const float l21 = (float)__ldg(p2 + tx1 * 256), l22 = (float)__ldg(p2 + tx2 * 256);
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
  const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const int r = __float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
  dst[(size_t)y * dp + x] = (uint8_t)min(max(r, 0), 255);
}

// ---- the adapt node's colour pre-filter: BGR -> HSV, CLAHE on V, HSV -> RGB (lfn3_adapt_node.py:164-184) -----------
// cv2.cvtColor(BGR2HSV) on uint8 is 12-bit fixed point with two division tables; cv2.cvtColor(HSV2RGB) is float32 with
// 1 - s*f fused, truncated in the 32-pixel vector steps of a row and rounded in the scalar tail (oracle/prefilter_np.py,
// pinned against the wheel).
__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

__global__ void __launch_bounds__(256) k_bgr2hsv_planes(const uint8_t* __restrict__ src, size_t sp, int w, int h,
                                                        uint8_t* __restrict__ Hp, uint8_t* __restrict__ Sp,
                                                        uint8_t* __restrict__ Vp, size_t pp,
                                                        unsigned long long* __restrict__ sums) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  unsigned long long sv = 0, sv2 = 0;
  if (x < w && y < h) {
    const uint8_t* p = src + (size_t)y * sp + (size_t)x * 3;
    const int b = p[0], g = p[1], r = p[2];
    const int v = max(max(b, g), r), vmin = min(min(b, g), r);
    const int diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    const int sat = (diff * c_sdiv[v] + (1 << 11)) >> 12;
    int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    hh = (hh * c_hdiv[diff] + (1 << 11)) >> 12;
    hh += hh < 0 ? 180 : 0;
    const size_t o = (size_t)y * pp + x;
    Hp[o] = (uint8_t)hh; Sp[o] = (uint8_t)sat; Vp[o] = (uint8_t)v;
    sv = (unsigned long long)v; sv2 = (unsigned long long)(v * v);
  }
  // sum(v), sum(v^2) of the frame for the adaptive clip limit (exact integers)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sv += __shfl_down_sync(0xffffffffu, sv, o);
    sv2 += __shfl_down_sync(0xffffffffu, sv2, o);
  }
  if ((threadIdx.x & 31) == 0 && sv2) { atomicAdd(sums, sv); atomicAdd(sums + 1, sv2); }
}

__global__ void __launch_bounds__(256) k_hsv2rgb_planes(const uint8_t* __restrict__ Hp, const uint8_t* __restrict__ Sp,
                                                        const uint8_t* __restrict__ Vp, size_t pp, int w, int h,
                                                        uint8_t* __restrict__ dst, size_t dp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const size_t o = (size_t)y * pp + x;
  const float s = __fmul_rn((float)Sp[o], 1.f / 255.f), v = __fmul_rn((float)Vp[o], 1.f / 255.f);
  float hf = __fmul_rn((float)Hp[o], 6.f / 180.f);
  int sec = (int)floorf(hf);
  const float f = __fsub_rn(hf, (float)sec);
  sec %= 6;
  const float t1 = __fmul_rn(v, __fsub_rn(1.f, s));
  const float t2 = __fmul_rn(v, __fmaf_rn(-s, f, 1.f));
  const float t3 = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.f, f), 1.f));
  float b, g, r;
  switch (sec) {            // sector table {b, g, r} <- {v, t1, t2, t3}
    case 0: b = t1; g = t3; r = v; break;
    case 1: b = t1; g = v; r = t2; break;
    case 2: b = t3; g = v; r = t1; break;
    case 3: b = v; g = t2; r = t1; break;
    case 4: b = v; g = t1; r = t3; break;
    default: b = t2; g = t1; r = v; break;
  }
  const bool body = x < (w / 32) * 32;       // the wheel's vector loop truncates, its scalar tail rounds
  const float fr = __fmul_rn(r, 255.f), fg = __fmul_rn(g, 255.f), fb = __fmul_rn(b, 255.f);
  const int ir = body ? (int)fr : __float2int_rn(fr), ig = body ? (int)fg : __float2int_rn(fg),
            ib = body ? (int)fb : __float2int_rn(fb);
  uint8_t* d = dst + (size_t)y * dp + (size_t)x * 3;
  d[0] = (uint8_t)min(max(ir, 0), 255); d[1] = (uint8_t)min(max(ig, 0), 255); d[2] = (uint8_t)min(max(ib, 0), 255);
}

int ingest_reserve(ofb_handle* h, size_t bytes_a, size_t bytes_b) {
  ofb_handle::Ingest& g = h->ingest;
  if (bytes_a > g.a_bytes) {
    OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (g.d_a) cudaFree(g.d_a);
    g.d_a = nullptr; g.a_bytes = 0;
    OFB_CUDA(h, cudaMalloc(&g.d_a, bytes_a));
    g.a_bytes = bytes_a;
  }
  if (bytes_b > g.b_bytes) {
    OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (g.d_b) cudaFree(g.d_b);
    g.d_b = nullptr; g.b_bytes = 0;
    OFB_CUDA(h, cudaMalloc(&g.d_b, bytes_b));
    g.b_bytes = bytes_b;
  }
  return OFB_OK;
}

// ---- bilateral filter of the adapt node (lfn3_adapt_node.py:186-190: cv2.bilateralFilter(rgb, d, sigmaColor, sigmaSpace)) ----
// OpenCV's own algorithm for 8UC3 (modules/imgproc/src/bilateral_filter.dispatch.cpp / .simd.hpp, restated in
// oracle/prefilter_np.py::bilateral_u8c3): a circular support of radius d/2 (REFLECT_101 border), per tap
//   w = space_weight[k] * color_weight[|db| + |dg| + |dr|]        (float product of two float tables)
//   wsum += w ; sum_c = fma(neighbour_c, w, sum_c)                 (taps in row-major order of the support)
// and out_c = round(sum_c * (1 / wsum)).  Bit-exact against the restatement; against the installed wheel the result
// differs at rounding ties only (a few values per 100 000, by one): the wheel routes 8-bit bilateralFilter through Intel
// IPP, whose arithmetic is not published (tests/test_node_gpu.py::test_bilateral states both facts).
constexpr int kBilMaxTaps = 1024;                     // (2 * 15 + 1)^2 = 961 >= the support of radius 15
__constant__ float c_bil_cw[768];
__constant__ float c_bil_sw[kBilMaxTaps];
__constant__ short2 c_bil_off[kBilMaxTaps];           // (dy, dx)

__global__ void __launch_bounds__(256) k_bilateral_u8c3(const uint8_t* __restrict__ src, size_t sp, int w, int h, int n_taps,
                                                        uint8_t* __restrict__ dst, size_t dp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* c = src + (size_t)y * sp + (size_t)x * 3;
  const int c0 = c[0], c1 = c[1], c2 = c[2];
  float wsum = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int k = 0; k < n_taps; k++) {
    const short2 o = c_bil_off[k];
    const int yy = reflect101(y + o.x, h), xx = reflect101(x + o.y, w);
    const uint8_t* q = src + (size_t)yy * sp + (size_t)xx * 3;
    const int b = q[0], g = q[1], r = q[2];
    const float wgt = __fmul_rn(c_bil_sw[k], c_bil_cw[abs(b - c0) + abs(g - c1) + abs(r - c2)]);
    wsum = __fadd_rn(wsum, wgt);
    s0 = __fmaf_rn((float)b, wgt, s0);
    s1 = __fmaf_rn((float)g, wgt, s1);
    s2 = __fmaf_rn((float)r, wgt, s2);
  }
  const float inv = __fdiv_rn(1.f, wsum);
  uint8_t* o = dst + (size_t)y * dp + (size_t)x * 3;
  o[0] = (uint8_t)min(max(__float2int_rn(__fmul_rn(s0, inv)), 0), 255);
  o[1] = (uint8_t)min(max(__float2int_rn(__fmul_rn(s1, inv)), 0), 255);
  o[2] = (uint8_t)min(max(__float2int_rn(__fmul_rn(s2, inv)), 0), 255);
}

// the filter on a device frame (3 bytes per pixel): tables to constant memory, one launch
static int bilateral_device(ofb_handle* h, const uint8_t* d_src, size_t sp, int width, int height, int d, double sigma_color,
                            double sigma_space, uint8_t* d_dst, size_t dp) {
  if (sigma_color <= 0) sigma_color = 1.0;
  if (sigma_space <= 0) sigma_space = 1.0;
  int radius = d <= 0 ? (int)__builtin_nearbyint(sigma_space * 1.5) : d / 2;
  radius = std::max(radius, 1);
  if (radius > 15) return set_error(h, OFB_ERR_INVALID_ARG, "bilateral filter: radius %d > 15", radius);
  if (radius >= width || radius >= height) return set_error(h, OFB_ERR_INVALID_ARG, "bilateral filter: image smaller than the support");
  const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
  static thread_local float cw[768], sw[kBilMaxTaps];
  static thread_local short2 off[kBilMaxTaps];
  for (int i = 0; i < 768; i++) cw[i] = (float)std::exp((double)i * i * gc);
  int n = 0;
  for (int i = -radius; i <= radius; i++)
    for (int j = -radius; j <= radius; j++) {
      const double r = std::sqrt((double)i * i + (double)j * j);
      if (r > radius) continue;
      sw[n] = (float)std::exp(r * r * gs);
      off[n] = make_short2((short)i, (short)j);
      n++;
    }
  cudaStream_t sm = h->stream;
  OFB_CUDA(h, cudaMemcpyToSymbolAsync(c_bil_cw, cw, sizeof(cw), 0, cudaMemcpyHostToDevice, sm));
  OFB_CUDA(h, cudaMemcpyToSymbolAsync(c_bil_sw, sw, n * sizeof(float), 0, cudaMemcpyHostToDevice, sm));
  OFB_CUDA(h, cudaMemcpyToSymbolAsync(c_bil_off, off, n * sizeof(short2), 0, cudaMemcpyHostToDevice, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));               // (the host tables are reused by the next call)
  k_bilateral_u8c3<<<dim3((width + 255) / 256, height), 256, 0, sm>>>(d_src, sp, width, height, n, d_dst, dp);
  OFB_LAUNCH_CHECK(h);
  return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_resize_u8(ofb_handle* h, const uint8_t* src, int src_width, int src_height, size_t src_stride_bytes, int channels,
                  uint8_t* dst, int dst_width, int dst_height, size_t dst_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!src || !dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (channels != 1 && channels != 3) return set_error(h, OFB_ERR_INVALID_ARG, "resize: 1 or 3 channels");
  if (src_width < 1 || src_height < 1 || dst_width < 1 || dst_height < 1) return set_error(h, OFB_ERR_INVALID_ARG, "bad size");
  const size_t srow = (size_t)src_width * channels, drow = (size_t)dst_width * channels;
  if (src_stride_bytes == 0) src_stride_bytes = srow;
  if (dst_stride_bytes == 0) dst_stride_bytes = drow;
  if (src_stride_bytes < srow || dst_stride_bytes < drow) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  OFB_CUDA(h, cudaSetDevice(h->device));
  int st = ingest_reserve(h, srow * src_height, drow * dst_height);
  if (st) return st;
  ofb_handle::Ingest& g = h->ingest;
  OFB_CUDA(h, cudaMemcpy2DAsync(g.d_a, srow, src, src_stride_bytes, srow, src_height, cudaMemcpyHostToDevice, h->stream));
  if ((st = resize_u8_device(h, g.d_a, srow, src_width, src_height, channels, g.d_b, drow, dst_width, dst_height))) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(dst, dst_stride_bytes, g.d_b, drow, drow, dst_height, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_ingest_gray(ofb_handle* h, const uint8_t* src, int src_width, int src_height, size_t src_stride_bytes, int rgb_order,
                    uint8_t* dst, int dst_width, int dst_height, size_t dst_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!src || !dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (src_width < 1 || src_height < 1 || dst_width < 1 || dst_height < 1) return set_error(h, OFB_ERR_INVALID_ARG, "bad size");
  const size_t srow = (size_t)src_width * 3, drow = (size_t)dst_width * 3;
  if (src_stride_bytes == 0) src_stride_bytes = srow;
  if (dst_stride_bytes == 0) dst_stride_bytes = (size_t)dst_width;
  if (src_stride_bytes < srow || dst_stride_bytes < (size_t)dst_width)
    return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  OFB_CUDA(h, cudaSetDevice(h->device));
  const bool same = src_width == dst_width && src_height == dst_height;
  const size_t gpitch = ((size_t)dst_width + 15) & ~(size_t)15;
  int st = ingest_reserve(h, srow * src_height, (same ? 0 : drow * dst_height) + gpitch * dst_height);
  if (st) return st;
  ofb_handle::Ingest& g = h->ingest;
  OFB_CUDA(h, cudaMemcpy2DAsync(g.d_a, srow, src, src_stride_bytes, srow, src_height, cudaMemcpyHostToDevice, h->stream));
  const uint8_t* col = g.d_a;
  size_t col_pitch = srow;
  uint8_t* gray = g.d_b;
  if (!same) {
    // cv2.resize on the colour frame first, as the nodes do, then the gray conversion
    uint8_t* small = g.d_b + gpitch * dst_height;
    if ((st = resize_u8_device(h, g.d_a, srow, src_width, src_height, 3, small, drow, dst_width, dst_height))) return st;
    col = small;
    col_pitch = drow;
  }
  if ((st = cvt_gray_device(h, col, col_pitch, gray, gpitch, dst_width, dst_height, rgb_order))) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(dst, dst_stride_bytes, gray, gpitch, dst_width, dst_height, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_cvt_gray_device(ofb_handle* h, const uint8_t* d_src, int width, int height, size_t src_pitch_bytes,
                        int rgb_order, uint8_t* d_dst, size_t dst_pitch_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!d_src || !d_dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL device pointer");
  if (width < 1 || height < 1 || src_pitch_bytes < (size_t)width * 3 || dst_pitch_bytes < (size_t)width)
    return set_error(h, OFB_ERR_INVALID_ARG, "bad size or pitch");
  OFB_CUDA(h, cudaSetDevice(h->device));
  return cvt_gray_device(h, d_src, src_pitch_bytes, d_dst, dst_pitch_bytes, width, height, rgb_order);
}

// Host-buffer variant (synchronous): colour frame in, gray frame out; the gray frame also stays in the
// handle's source staging slot `slot` (0 or 1) so a following ofb_farneback_staged call can use it
// without another upload.
int ofb_cvt_gray(ofb_handle* h, const uint8_t* src, int width, int height, size_t stride_bytes, int rgb_order,
                 uint8_t* dst, size_t dst_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!src || !dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (width < 1 || height < 1) return set_error(h, OFB_ERR_INVALID_ARG, "bad size");
  if (stride_bytes == 0) stride_bytes = (size_t)width * 3;
  if (dst_stride_bytes == 0) dst_stride_bytes = (size_t)width;
  if (stride_bytes < (size_t)width * 3 || dst_stride_bytes < (size_t)width)
    return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  if (width > h->max_w || height > h->max_h || (size_t)width * height > (size_t)h->max_w * h->max_h)
    return set_error(h, OFB_ERR_CAPACITY, "frame %dx%d exceeds handle capacity %dx%d", width, height, h->max_w, h->max_h);
  OFB_CUDA(h, cudaSetDevice(h->device));
  // colour staging: the level-image buffer (4 bytes per pixel per frame >= 3 bytes per pixel) is free
  // outside a flow call
  uint8_t* d_col = reinterpret_cast<uint8_t*>(h->d_img);
  const size_t col_pitch = (size_t)width * 3;
  OFB_CUDA(h, cudaMemcpy2DAsync(d_col, col_pitch, src, stride_bytes, col_pitch, height, cudaMemcpyHostToDevice, h->stream));
  int st = cvt_gray_device(h, d_col, col_pitch, h->d_src, h->src_pitch, width, height, rgb_order);
  if (st) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(dst, dst_stride_bytes, h->d_src, h->src_pitch, width, height, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_clahe(ofb_handle* h, const uint8_t* src, int width, int height, size_t src_stride_bytes, double clip_limit,
              int tiles_x, int tiles_y, uint8_t* dst, size_t dst_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!src || !dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (width < 1 || height < 1 || tiles_x < 1 || tiles_y < 1 || tiles_x > 256 || tiles_y > 256)
    return set_error(h, OFB_ERR_INVALID_ARG, "bad size or tile grid");
  if (src_stride_bytes == 0) src_stride_bytes = (size_t)width;
  if (dst_stride_bytes == 0) dst_stride_bytes = (size_t)width;
  if (src_stride_bytes < (size_t)width || dst_stride_bytes < (size_t)width)
    return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  // the extension to a multiple of the grid (REFLECT_101) must stay inside one reflection
  const bool exact = width % tiles_x == 0 && height % tiles_y == 0;
  const int ew = exact ? width : width + tiles_x - width % tiles_x, eh = exact ? height : height + tiles_y - height % tiles_y;
  if (ew - width >= width || eh - height >= height)
    return set_error(h, OFB_ERR_INVALID_ARG, "image too small for a %dx%d tile grid", tiles_x, tiles_y);
  const int tw = ew / tiles_x, th = eh / tiles_y;
  const int area = tw * th;
  int clip = 0;
  if (clip_limit > 0.0) clip = std::max((int)(clip_limit * area / 256), 1);
  const float lut_scale = 255.f / (float)area;
  OFB_CUDA(h, cudaSetDevice(h->device));
  const size_t pitch = ((size_t)width + 15) & ~(size_t)15;
  const size_t lut_bytes = (size_t)tiles_x * tiles_y * 256;
  int st = ingest_reserve(h, pitch * height, pitch * height + lut_bytes);
  if (st) return st;
  ofb_handle::Ingest& g = h->ingest;
  uint8_t* luts = g.d_b + pitch * height;
  OFB_CUDA(h, cudaMemcpy2DAsync(g.d_a, pitch, src, src_stride_bytes, width, height, cudaMemcpyHostToDevice, h->stream));
  k_clahe_lut<<<dim3(tiles_x, tiles_y), 256, 0, h->stream>>>(g.d_a, pitch, width, height, tw, th, clip, lut_scale, luts);
  OFB_LAUNCH_CHECK(h);
  k_clahe_interp<<<dim3((width + 255) / 256, height), 256, 0, h->stream>>>(g.d_a, pitch, g.d_b, pitch, width, height, tiles_x,
                                                                          tiles_y, 1.f / (float)tw, 1.f / (float)th, luts);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemcpy2DAsync(dst, dst_stride_bytes, g.d_b, pitch, width, height, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_bilateral_u8c3(ofb_handle* h, const uint8_t* src, int width, int height, size_t src_stride_bytes, int d,
                       double sigma_color, double sigma_space, uint8_t* dst, size_t dst_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!src || !dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (width < 2 || height < 2) return set_error(h, OFB_ERR_INVALID_ARG, "bad size");
  const size_t row3 = (size_t)width * 3;
  if (src_stride_bytes == 0) src_stride_bytes = row3;
  if (dst_stride_bytes == 0) dst_stride_bytes = row3;
  if (src_stride_bytes < row3 || dst_stride_bytes < row3) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  OFB_CUDA(h, cudaSetDevice(h->device));
  int st = ingest_reserve(h, row3 * height, row3 * height);
  if (st) return st;
  ofb_handle::Ingest& g = h->ingest;
  OFB_CUDA(h, cudaMemcpy2DAsync(g.d_a, row3, src, src_stride_bytes, row3, height, cudaMemcpyHostToDevice, h->stream));
  if ((st = timing_begin(h, OFB_STAGE_OTHER))) return st;
  if ((st = bilateral_device(h, g.d_a, row3, width, height, d, sigma_color, sigma_space, g.d_b, row3))) return st;
  if ((st = timing_end(h))) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(dst, dst_stride_bytes, g.d_b, row3, row3, height, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_adapt_prefilter(ofb_handle* h, const uint8_t* bgr, int width, int height, size_t stride_bytes,
                        const ofb_clahe_params* p, uint8_t* rgb, size_t rgb_stride_bytes, double* clip_used) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!bgr || !rgb || !p) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (width < 1 || height < 1 || p->tiles_x < 1 || p->tiles_y < 1 || p->tiles_x > 256 || p->tiles_y > 256)
    return set_error(h, OFB_ERR_INVALID_ARG, "bad size or tile grid");
  const size_t row3 = (size_t)width * 3;
  if (stride_bytes == 0) stride_bytes = row3;
  if (rgb_stride_bytes == 0) rgb_stride_bytes = row3;
  if (stride_bytes < row3 || rgb_stride_bytes < row3) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  const int tiles_x = p->tiles_x, tiles_y = p->tiles_y;
  const bool exact = width % tiles_x == 0 && height % tiles_y == 0;
  const int ew = exact ? width : width + tiles_x - width % tiles_x, eh = exact ? height : height + tiles_y - height % tiles_y;
  if (ew - width >= width || eh - height >= height)
    return set_error(h, OFB_ERR_INVALID_ARG, "image too small for a %dx%d tile grid", tiles_x, tiles_y);
  OFB_CUDA(h, cudaSetDevice(h->device));
  static bool tables[64] = {false};
  if (!tables[h->device & 63]) {
    int sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
      sdiv[i] = (int)__builtin_nearbyint((255 << 12) / (1.0 * i));
      hdiv[i] = (int)__builtin_nearbyint((180 << 12) / (6.0 * i));
    }
    OFB_CUDA(h, cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv)));
    OFB_CUDA(h, cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv)));
    tables[h->device & 63] = true;
  }
  const size_t pp = ((size_t)width + 15) & ~(size_t)15, plane = pp * height;
  const size_t lut_bytes = (size_t)tiles_x * tiles_y * 256;
  int st = ingest_reserve(h, row3 * height, 4 * plane + row3 * height + lut_bytes + 64);
  if (st) return st;
  ofb_handle::Ingest& g = h->ingest;
  unsigned long long* d_sums = reinterpret_cast<unsigned long long*>(g.d_b);   // (first: cudaMalloc alignment)
  uint8_t *Hp = g.d_b + 64, *Sp = Hp + plane, *Vp = Sp + plane, *V2 = Vp + plane, *luts = V2 + plane, *out = luts + lut_bytes;
  cudaStream_t sm = h->stream;
  OFB_CUDA(h, cudaMemcpy2DAsync(g.d_a, row3, bgr, stride_bytes, row3, height, cudaMemcpyHostToDevice, sm));
  OFB_CUDA(h, cudaMemsetAsync(d_sums, 0, 16, sm));
  const dim3 grid((width + 255) / 256, height);
  k_bgr2hsv_planes<<<grid, 256, 0, sm>>>(g.d_a, row3, width, height, Hp, Sp, Vp, pp, d_sums);
  OFB_LAUNCH_CHECK(h);
  double clip_limit = p->clip_limit;
  if (p->adaptive) {
    // contrast = std(v) / (mean(v) + 1e-3), mapped linearly to [clip_min, clip_max] (lfn3_adapt_node.py:170-175)
    unsigned long long hs[2];
    OFB_CUDA(h, cudaMemcpyAsync(hs, d_sums, 16, cudaMemcpyDeviceToHost, sm));
    OFB_CUDA(h, cudaStreamSynchronize(sm));
    const double n = (double)width * height, mean = (double)hs[0] / n;
    const double var = std::max((double)hs[1] / n - mean * mean, 0.0);
    const double contrast = std::sqrt(var) / (mean + 1e-3);
    clip_limit = p->clip_min + (contrast - p->c_min) / (p->c_max - p->c_min) * (p->clip_max - p->clip_min);
    clip_limit = std::min(std::max(clip_limit, p->clip_min), p->clip_max);
  }
  if (clip_used) *clip_used = clip_limit;
  const int tw = ew / tiles_x, th = eh / tiles_y, area = tw * th;
  const int clip = clip_limit > 0.0 ? std::max((int)(clip_limit * area / 256), 1) : 0;
  k_clahe_lut<<<dim3(tiles_x, tiles_y), 256, 0, sm>>>(Vp, pp, width, height, tw, th, clip, 255.f / (float)area, luts);
  OFB_LAUNCH_CHECK(h);
  k_clahe_interp<<<grid, 256, 0, sm>>>(Vp, pp, V2, pp, width, height, tiles_x, tiles_y, 1.f / (float)tw, 1.f / (float)th, luts);
  OFB_LAUNCH_CHECK(h);
  k_hsv2rgb_planes<<<grid, 256, 0, sm>>>(Hp, Sp, V2, pp, width, height, out, row3);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemcpy2DAsync(rgb, rgb_stride_bytes, out, row3, row3, height, cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  return OFB_OK;
}

}  // extern "C"
