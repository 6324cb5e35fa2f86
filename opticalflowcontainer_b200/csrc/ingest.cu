// ingest.cu — frame ingest on the device: the step in front of the flow call in
// every node of the reference is cv_bridge -> cv2.cvtColor(BGR2GRAY) (-> cv2.resize) on the CPU
//   ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:148-159
//   ros2_ws/src/optical_flow/optical_flow/opticalflow_node.py:43-50
// This converts an interleaved 8-bit colour frame (sensor_msgs/Image `bgr8` / `rgb8`, row `step`) to
// the single-channel uint8 frame the flow engine consumes, bit-exactly as this cv2 build does it:
//   gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15          (15-bit fixed point)
// so the 3N-byte colour frame is uploaded once and never converted on the host.
#include "common.cuh"

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

}  // namespace ofb

using namespace ofb;

extern "C" {

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

}  // extern "C"
