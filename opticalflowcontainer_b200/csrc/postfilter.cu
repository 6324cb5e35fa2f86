// postfilter.cu — the "adapt" node's flow post-processing on the device (SURVEY.md 8f rank 3 / rank 1 masks).
//
// ros2_ws/src/liteflownet3/liteflownet3/lfn3_adapt_node.py:
//   :236-238  flow[c] = cv2.medianBlur(flow[c], k)              k = 3 | 5, float32, BORDER_REPLICATE, exact selection
//   :241-244  m = sqrt(u*u + v*v) >= threshold ; u *= m ; v *= m   (float32, every operation rounded: no FMA)
//   :247-251  m = gray < intensity_threshold   ; u *= m ; v *= m
// followed by np.mean(u) (:254), which ofb_flow_u_stats computes on the result.  One kernel, one pass: 8 B in,
// 8 B out per pixel (+1 B gray), the field never leaves HBM.  Bit-exact against cv2 / NumPy (signed zeros aside).
#include <algorithm>

#include "common.cuh"

namespace ofb {

// moves the minimum of a[0..M) to a[0] and the maximum to a[M-1]
template <int M, int CAP>
__device__ __forceinline__ void extract_minmax(float (&a)[CAP]) {
#pragma unroll
  for (int i = 0; i < M / 2; i++) {
    const float lo = fminf(a[i], a[M - 1 - i]), hi = fmaxf(a[i], a[M - 1 - i]);
    a[i] = lo; a[M - 1 - i] = hi;
  }
#pragma unroll
  for (int i = 1; i < (M + 1) / 2; i++) {
    const float lo = fminf(a[0], a[i]), hi = fmaxf(a[0], a[i]);
    a[0] = lo; a[i] = hi;
  }
#pragma unroll
  for (int i = M / 2; i < M - 1; i++) {
    const float lo = fminf(a[i], a[M - 1]), hi = fmaxf(a[i], a[M - 1]);
    a[i] = lo; a[M - 1] = hi;
  }
}

// Forgetful selection of the median of N = 2*W0 - 3 values: keep W0 = N/2 + 2 candidates, discard their minimum and
// maximum (neither can be the median), take the next value in, until three are left.
template <int M, int W0, int N>
__device__ __forceinline__ float forget(float (&a)[W0], const float (&v)[N]) {
  extract_minmax<M, W0>(a);
  if constexpr (M == 3) {
    return a[1];
  } else {
    a[0] = v[2 * W0 - M];          // replaces the minimum; the maximum a[M-1] drops out with the smaller M
    return forget<M - 1, W0, N>(a, v);
  }
}

template <int K>
__device__ __forceinline__ float median_kk(const float (&v)[K * K]) {
  constexpr int N = K * K, W0 = N / 2 + 2;
  float a[W0];
#pragma unroll
  for (int i = 0; i < W0; i++) a[i] = v[i];
  return forget<W0, W0, N>(a, v);
}

constexpr int PFX = 32, PFY = 8;

// K = 0: masks only.  gray (optional): n images [h][gray_pitch] on the device.
template <int K>
__global__ void __launch_bounds__(PFX* PFY) k_flow_postfilter(const float2* __restrict__ in, float2* __restrict__ out, int w,
                                                              int h, bool use_mag, float mag_thr,
                                                              const uint8_t* __restrict__ gray, size_t gray_pitch,
                                                              int intensity_thr) {
  constexpr int R = K / 2;
  __shared__ float2 tile[PFY + 2 * R][PFX + 2 * R];
  const size_t n = (size_t)w * h;
  const float2* f = in + (size_t)blockIdx.z * n;
  const int bx = blockIdx.x * PFX, by = blockIdx.y * PFY;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (K > 0) {
    for (int i = ty * PFX + tx; i < (PFY + 2 * R) * (PFX + 2 * R); i += PFX * PFY) {
      const int ry = i / (PFX + 2 * R), rx = i - ry * (PFX + 2 * R);
      const int sy = min(max(by + ry - R, 0), h - 1), sx = min(max(bx + rx - R, 0), w - 1);   // BORDER_REPLICATE
      tile[ry][rx] = __ldg(f + (size_t)sy * w + sx);
    }
    __syncthreads();
  }
  const int x = bx + tx, y = by + ty;
  if (x >= w || y >= h) return;
  float u, v;
  if constexpr (K > 0) {
    float a[K * K > 0 ? K * K : 1];
#pragma unroll
    for (int j = 0; j < K; j++)
#pragma unroll
      for (int i = 0; i < K; i++) a[j * K + i] = tile[ty + j][tx + i].x;
    u = median_kk<K>(a);
#pragma unroll
    for (int j = 0; j < K; j++)
#pragma unroll
      for (int i = 0; i < K; i++) a[j * K + i] = tile[ty + j][tx + i].y;
    v = median_kk<K>(a);
  } else {
    const float2 t = __ldg(f + (size_t)y * w + x);
    u = t.x; v = t.y;
  }
  if (use_mag) {
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
    const float m = mag >= mag_thr ? 1.f : 0.f;
    u = __fmul_rn(u, m); v = __fmul_rn(v, m);
  }
  if (gray) {
    const float m = (int)gray[(size_t)blockIdx.z * gray_pitch * h + (size_t)y * gray_pitch + x] < intensity_thr ? 1.f : 0.f;
    u = __fmul_rn(u, m); v = __fmul_rn(v, m);
  }
  out[(size_t)blockIdx.z * n + (size_t)y * w + x] = make_float2(u, v);
}

int flow_postfilter(ofb_handle* h, int n, int median_ksize, float magnitude_threshold, const uint8_t* const* gray,
                    size_t gray_stride, int intensity_threshold) {
  if (!h->last_flow || n < 1 || n > h->last_n)
    return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_postfilter: no flow field of %d pair(s) on the device", n);
  if (median_ksize != 0 && median_ksize != 3 && median_ksize != 5)
    return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_postfilter: median_ksize must be 0, 3 or 5 (cv2.medianBlur, float32)");
  OFB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int w = h->last_w, hh = h->last_h;
  const size_t npix = (size_t)w * hh;
  const uint8_t* dgray = nullptr;
  const size_t gpitch = ((size_t)w + 15) & ~(size_t)15;
  if (gray) {
    if (gray_stride == 0) gray_stride = (size_t)w;
    if (gray_stride < (size_t)w) return set_error(h, OFB_ERR_INVALID_ARG, "gray stride smaller than width");
    for (int i = 0; i < n; i++)
      if (!gray[i]) return set_error(h, OFB_ERR_INVALID_ARG, "NULL gray image");
    const size_t need = gpitch * hh * (size_t)n;
    if (need > h->gray_bytes) {
      OFB_CUDA(h, cudaStreamSynchronize(st));
      if (h->d_gray) cudaFree(h->d_gray);
      h->d_gray = nullptr; h->gray_bytes = 0;
      OFB_CUDA(h, cudaMalloc(&h->d_gray, need));
      h->gray_bytes = need;
    }
    for (int i = 0; i < n; i++)
      OFB_CUDA(h, cudaMemcpy2DAsync(h->d_gray + (size_t)i * gpitch * hh, gpitch, gray[i], gray_stride, w, hh,
                                    cudaMemcpyHostToDevice, st));
    dgray = h->d_gray;
  }
  // the result goes to one of the level ping-pong buffers (free between calls) and becomes the handle's field
  float2* dst = reinterpret_cast<const float2*>(h->last_flow) == h->d_flow[0] ? h->d_flow[1] : h->d_flow[0];
  const float2* srcf = reinterpret_cast<const float2*>(h->last_flow);
  dim3 blk(PFX, PFY), g((w + PFX - 1) / PFX, (hh + PFY - 1) / PFY, n);
  const bool use_mag = magnitude_threshold >= 0.f;
  int s;
  if ((s = timing_begin(h, OFB_STAGE_OTHER))) return s;
  if (median_ksize == 3)
    k_flow_postfilter<3><<<g, blk, 0, st>>>(srcf, dst, w, hh, use_mag, magnitude_threshold, dgray, gpitch, intensity_threshold);
  else if (median_ksize == 5)
    k_flow_postfilter<5><<<g, blk, 0, st>>>(srcf, dst, w, hh, use_mag, magnitude_threshold, dgray, gpitch, intensity_threshold);
  else
    k_flow_postfilter<0><<<g, blk, 0, st>>>(srcf, dst, w, hh, use_mag, magnitude_threshold, dgray, gpitch, intensity_threshold);
  OFB_LAUNCH_CHECK(h);
  if ((s = timing_end(h))) return s;
  h->last_flow = reinterpret_cast<const float*>(dst);
  (void)npix;
  return OFB_OK;
}

// flow at integer pixel positions (the junction node's lookup, lfn3_junction_node.py:207-214); outside -> NaN
__global__ void k_flow_sample(const float2* __restrict__ f, int w, int h, const int2* __restrict__ pts, int n,
                              float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 p = pts[i];
  const float nanv = __int_as_float(0x7fc00000);
  out[i] = (p.x >= 0 && p.x < w && p.y >= 0 && p.y < h) ? f[(size_t)p.y * w + p.x] : make_float2(nanv, nanv);
}

int flow_sample(ofb_handle* h, int pair, int n_points, const int* xy, float* out_dxdy) {
  if (!h->last_flow || pair < 0 || pair >= h->last_n)
    return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_sample: no flow field %d on the device", pair);
  if (n_points < 0 || (n_points > 0 && (!xy || !out_dxdy))) return set_error(h, OFB_ERR_INVALID_ARG, "bad point arrays");
  if (n_points == 0) return OFB_OK;
  OFB_CUDA(h, cudaSetDevice(h->device));
  // the points and their results share the mask buffer (>= 16 B per point available: npix bytes)
  const size_t npix = (size_t)h->max_w * h->max_h;
  if ((size_t)n_points * 16 > npix) return set_error(h, OFB_ERR_INVALID_ARG, "too many points (%d)", n_points);
  int2* dp = reinterpret_cast<int2*>(h->d_mask);
  float2* dout = reinterpret_cast<float2*>(h->d_mask + (size_t)n_points * 8);
  OFB_CUDA(h, cudaMemcpyAsync(dp, xy, (size_t)n_points * 8, cudaMemcpyHostToDevice, h->stream));
  k_flow_sample<<<(n_points + 127) / 128, 128, 0, h->stream>>>(
      reinterpret_cast<const float2*>(h->last_flow) + (size_t)pair * h->last_w * h->last_h, h->last_w, h->last_h, dp,
      n_points, dout);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemcpyAsync(out_dxdy, dout, (size_t)n_points * 8, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

// ---- flow visualisation: the nodes' flow_to_color (sub_n_pub_lfn3_node.py:132-140) on the device ------------------
//   mag, ang = cv2.cartToPolar(u, v); H = uint8(ang * 180 / pi / 2); S = 255;
//   V = uint8(cv2.normalize(mag, None, 0, 255, NORM_MINMAX)); cv2.cvtColor(hsv, COLOR_HSV2BGR)
// with the wheel's arithmetic (oracle/visual_np.py, pinned bit for bit): magnitude sqrt(fma(u, u, v*v)); fastAtan's
// 7th-order polynomial evaluated with FMAs in degrees, times pi/180; normalisation fma(mag, scale, shift) with scale and
// shift from the field's min / max in double; HSV2BGR as in the colour pre-filter (truncating in the 32-pixel vector body
// of a row, rounding in its tail).
__device__ __forceinline__ float viz_mag(float2 f) { return __fsqrt_rn(__fmaf_rn(f.x, f.x, __fmul_rn(f.y, f.y))); }

__global__ void __launch_bounds__(256) k_viz_minmax(const float2* __restrict__ f, size_t n, unsigned int* __restrict__ mm) {
  float lo = __int_as_float(0x7f800000), hi = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float m = viz_mag(f[i]);
    lo = fminf(lo, m);
    hi = fmaxf(hi, m);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {      // magnitudes are >= 0: float order = unsigned bit order
    atomicMin(mm, __float_as_uint(lo));
    atomicMax(mm + 1, __float_as_uint(hi));
  }
}

// mode 0: flow_to_color (sub_n_pub_lfn3_node.py:132-140): H = angle / 2 in degrees, V = min-max normalised magnitude.
// mode 1: the sub node's dense view (lfn3_sub_node.py:244-262): H = uint8(ang * 90 / pi), V = uint8(clip(mag / dt *
// pixel_to_meter / max_speed, 0, 1) * 255), every operation a float32 operation as NumPy evaluates the expression.
__global__ void __launch_bounds__(256) k_viz_color(const float2* __restrict__ f, int w, int h, const unsigned int* __restrict__ mm,
                                                   uint8_t* __restrict__ dst, size_t dp, int mode, float dt, float p2m, float vmax) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const float2 v = f[(size_t)y * w + x];
  // cartToPolar angle (v_atan_f32): degrees by the polynomial, then * pi/180
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k);
  const float p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
  const float ax = fabsf(v.x), ay = fabsf(v.y);
  const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), 2.220446049250313e-16f));
  const float cc = __fmul_rn(c, c);
  float a = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(cc, p7, p5), cc, p3), cc, p1), c);
  a = ax >= ay ? a : __fsub_rn(90.f, a);
  a = v.x < 0.f ? __fsub_rn(180.f, a) : a;
  a = v.y < 0.f ? __fsub_rn(360.f, a) : a;
  const float ang = __fmul_rn(a, (float)(3.14159265358979323846 / 180.0));
  int H, V;
  if (mode == 0) {
    H = (int)__fdiv_rn(__fdiv_rn(__fmul_rn(ang, 180.f), (float)3.14159265358979323846), 2.f) & 255;
    // normalize(NORM_MINMAX, 0..255): scale / shift in double from the field's extrema, applied as one float FMA
    const double smin = (double)__uint_as_float(mm[0]), smax = (double)__uint_as_float(mm[1]);
    const double scale = 255.0 * (smax - smin > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    const double shift = 0.0 - smin * scale;
    V = (int)__fmaf_rn(viz_mag(v), (float)scale, (float)shift) & 255;
  } else {
    H = (int)__fdiv_rn(__fmul_rn(ang, 90.f), (float)3.14159265358979323846) & 255;
    const float t = __fdiv_rn(__fmul_rn(__fdiv_rn(viz_mag(v), dt), p2m), vmax);
    V = (int)__fmul_rn(fminf(fmaxf(t, 0.f), 1.f), 255.f) & 255;
  }
  // HSV -> BGR, S = 255
  const float s = __fmul_rn(255.f, 1.f / 255.f), vv = __fmul_rn((float)V, 1.f / 255.f);
  const float hf = __fmul_rn((float)H, 6.f / 180.f);
  int sec = (int)floorf(hf);
  const float fr = __fsub_rn(hf, (float)sec);
  sec %= 6;
  const float t1 = __fmul_rn(vv, __fsub_rn(1.f, s));
  const float t2 = __fmul_rn(vv, __fmaf_rn(-s, fr, 1.f));
  const float t3 = __fmul_rn(vv, __fmaf_rn(-s, __fsub_rn(1.f, fr), 1.f));
  float b, g, r;
  switch (sec) {
    case 0: b = t1; g = t3; r = vv; break;
    case 1: b = t1; g = vv; r = t2; break;
    case 2: b = t3; g = vv; r = t1; break;
    case 3: b = vv; g = t2; r = t1; break;
    case 4: b = vv; g = t1; r = t3; break;
    default: b = t2; g = t1; r = vv; break;
  }
  const bool body = x < (w / 32) * 32;
  const float fR = __fmul_rn(r, 255.f), fG = __fmul_rn(g, 255.f), fB = __fmul_rn(b, 255.f);
  uint8_t* o = dst + (size_t)y * dp + (size_t)x * 3;
  o[0] = (uint8_t)min(max(body ? (int)fB : __float2int_rn(fB), 0), 255);
  o[1] = (uint8_t)min(max(body ? (int)fG : __float2int_rn(fG), 0), 255);
  o[2] = (uint8_t)min(max(body ? (int)fR : __float2int_rn(fR), 0), 255);
}

int flow_to_bgr(ofb_handle* h, int pair, uint8_t* bgr_out, size_t stride_bytes, int mode, float dt, float p2m, float vmax) {
  if (!h->last_flow || pair < 0 || pair >= h->last_n)
    return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_to_bgr: no flow field %d on the device", pair);
  if (!bgr_out) return set_error(h, OFB_ERR_INVALID_ARG, "NULL output pointer");
  const int w = h->last_w, hh = h->last_h;
  if (stride_bytes == 0) stride_bytes = (size_t)w * 3;
  if (stride_bytes < (size_t)w * 3) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  OFB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // the image and the two extrema live in the scratch the reductions use (3 B per pixel <= the 8 B per pixel of d_flow[x])
  const float2* f = reinterpret_cast<const float2*>(h->last_flow) + (size_t)pair * w * hh;
  float2* freebuf = reinterpret_cast<const float2*>(h->last_flow) == h->d_flow[0] ? h->d_flow[1] : h->d_flow[0];
  uint8_t* img = reinterpret_cast<uint8_t*>(freebuf) + 256;
  unsigned int* mm = reinterpret_cast<unsigned int*>(freebuf);
  const unsigned int init[2] = {0x7f800000u, 0u};
  OFB_CUDA(h, cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
  int s;
  if ((s = timing_begin(h, OFB_STAGE_OTHER))) return s;
  if (mode == 0) {
    k_viz_minmax<<<2 * h->num_sms, 256, 0, st>>>(f, (size_t)w * hh, mm);
    OFB_LAUNCH_CHECK(h);
  }
  k_viz_color<<<dim3((w + 255) / 256, hh), 256, 0, st>>>(f, w, hh, mm, img, (size_t)w * 3, mode, dt, p2m, vmax);
  OFB_LAUNCH_CHECK(h);
  if ((s = timing_end(h))) return s;
  OFB_CUDA(h, cudaMemcpy2DAsync(bgr_out, stride_bytes, img, (size_t)w * 3, (size_t)w * 3, hh, cudaMemcpyDeviceToHost, st));
  OFB_CUDA(h, cudaStreamSynchronize(st));
  return OFB_OK;
}

int flow_download(ofb_handle* h, int n, float* const* flow, size_t flow_stride_bytes) {
  if (!h->last_flow || n < 1 || n > h->last_n)
    return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_download: no flow field of %d pair(s) on the device", n);
  if (!flow) return set_error(h, OFB_ERR_INVALID_ARG, "NULL array pointer");
  OFB_CUDA(h, cudaSetDevice(h->device));
  const int w = h->last_w, hh = h->last_h;
  const size_t row = (size_t)w * 2 * sizeof(float);
  if (flow_stride_bytes == 0) flow_stride_bytes = row;
  if (flow_stride_bytes < row) return set_error(h, OFB_ERR_INVALID_ARG, "flow stride smaller than a row");
  for (int i = 0; i < n; i++) {
    if (!flow[i]) return set_error(h, OFB_ERR_INVALID_ARG, "NULL flow pointer");
    OFB_CUDA(h, cudaMemcpy2DAsync(flow[i], flow_stride_bytes, h->last_flow + (size_t)i * w * hh * 2, row, row, hh,
                                  cudaMemcpyDeviceToHost, h->stream));
  }
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

}  // namespace ofb
