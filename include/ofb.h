/* ofb.h — C ABI of libofb.so, the B200 (sm_100a) dense/sparse optical-flow engine.
 *
 * Drop-in boundary for the per-frame-pair flow call of the ROS 2 image-subscriber
 * nodes of Hagestregen/OpticalFlowContainer.  The reference has no FFI for this
 * path: the boundary is the expression at the node's flow call,
 *   ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:194   flow = self.net(t1,t2)
 *   ros2_ws/src/optical_flow/optical_flow/opticalflow_node.py:87
 *   ros2_ws/src/pwc_net/pwc_net/pwc_sub_node.py:183
 *   ros2_ws/src/liteflownet3/liteflownet3/lfn3_node.py:184
 * which a Farneback/LK node fills with cv2.calcOpticalFlowFarneback /
 * cv2.goodFeaturesToTrack / cv2.calcOpticalFlowPyrLK (OpenCV: un-vendored
 * dependency, ros2_ws/src/nueflow/setup.py:29).  Each entry point below names the
 * cv2 call (and the reference line it sits behind) that it replaces.  The
 * reference's only native-binding precedent is the pybind module
 * ros2_ws/src/liteflownet3/correlation_package/correlation_cuda.cc:169-172
 * (free functions, caller-allocated outputs, int status) — the same conventions
 * are kept here, minus the torch types.
 *
 * Conventions
 *   - plain C, no C++ exceptions cross the ABI, never aborts; every function
 *     returns an ofb_status (0 = OK) and records a message readable through
 *     ofb_last_error().
 *   - the caller owns every host/device array it passes; the library owns the
 *     device work buffers and pinned staging buffers inside the handle.
 *   - a handle is bound to one CUDA device and one stream; it is NOT re-entrant
 *     (one call at a time per handle, from any thread); different handles are
 *     independent (one per GPU / per camera stream).
 *   - there is no CPU fallback: if no CUDA device is usable ofb_create fails.
 */
#ifndef OFB_H_
#define OFB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* libofb.so is built with -fvisibility=hidden */
#endif

#define OFB_VERSION 100 /* 0.1.0 */

typedef enum ofb_status {
  OFB_OK = 0,
  OFB_ERR_INVALID_ARG = 1, /* cv2 would raise cv::Exception (bad size/type/flags) */
  OFB_ERR_CUDA = 2,        /* a CUDA runtime call or kernel launch failed */
  OFB_ERR_NO_DEVICE = 3,   /* no usable CUDA device: there is no CPU fallback */
  OFB_ERR_CAPACITY = 4,    /* frame or batch larger than the handle was created for */
  OFB_ERR_ALLOC = 5,
  OFB_ERR_UNSUPPORTED = 6  /* a stream the device path does not decode (e.g. progressive JPEG): decode it on the host */
} ofb_status;

/* cv2 flag values (cv2.OPTFLOW_*), identical numbers */
#define OFB_OPTFLOW_USE_INITIAL_FLOW 4
#define OFB_OPTFLOW_LK_GET_MIN_EIGENVALS 8
#define OFB_OPTFLOW_FARNEBACK_GAUSSIAN 256

typedef struct ofb_handle ofb_handle;

/* cv2.calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize,
 *                              iterations, poly_n, poly_sigma, flags) — same meaning, same order. */
typedef struct ofb_farneback_params {
  double pyr_scale; /* < 1 */
  int levels;       /* cv2 semantics: levels+1 scales, clamped so the coarsest is >= 32 px */
  int winsize;
  int iterations;
  int poly_n;
  double poly_sigma;
  int flags; /* 0 | OFB_OPTFLOW_USE_INITIAL_FLOW | OFB_OPTFLOW_FARNEBACK_GAUSSIAN */
} ofb_farneback_params;

/* cv2.calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status, err,
 *                          winSize, maxLevel, criteria, flags, minEigThreshold) */
typedef struct ofb_lk_params {
  int win_w, win_h;      /* winSize, default 21x21 */
  int max_level;         /* default 3 */
  int max_count;         /* criteria.maxCount, clamped to [0,100] as cv2 does */
  double epsilon;        /* criteria.epsilon, clamped to [0,10], squared internally */
  int flags;             /* 0 | OFB_OPTFLOW_USE_INITIAL_FLOW | OFB_OPTFLOW_LK_GET_MIN_EIGENVALS */
  double min_eig_threshold; /* default 1e-4 */
} ofb_lk_params;

/* cv2.goodFeaturesToTrack(image, maxCorners, qualityLevel, minDistance, mask,
 *                         blockSize, useHarrisDetector=False, k) — Shi-Tomasi (the Harris response of the
 * wheel is not reproduced: useHarrisDetector=True is refused by the Python mirror). */
typedef struct ofb_gftt_params {
  int max_corners;
  double quality_level;
  double min_distance;
  int block_size; /* 3 */
  int use_harris_detector; /* 0: Shi-Tomasi (cornerMinEigenVal), 1: cv2.cornerHarris(image, blockSize, 3, harris_k) */
  double harris_k;         /* 0.04 */
} ofb_gftt_params;

/* ---- lifetime ---------------------------------------------------------------- */
int ofb_version(void);
const char* ofb_status_string(int status);

/* Creates an engine on CUDA device `device` able to process up to `max_batch`
 * frame pairs of up to max_width x max_height per call.  Allocates all device
 * work buffers and pinned staging buffers up front (nothing is allocated on the
 * hot path). */
int ofb_create(int device, int max_width, int max_height, int max_batch, ofb_handle** out);
int ofb_destroy(ofb_handle* h);
/* Last error message of this handle (or of the failed ofb_create when h == NULL). */
const char* ofb_last_error(const ofb_handle* h);
/* The handle's CUDA stream (a cudaStream_t) so callers can order their own work. */
void* ofb_stream(ofb_handle* h);
int ofb_synchronize(ofb_handle* h);

/* ---- Farneback: replaces cv2.calcOpticalFlowFarneback at lfn3_sub_node.py:194 -- */

/* Host-buffer call, synchronous: uploads prev/next (uint8, `stride_bytes` per row)
 * through pinned staging on the handle's stream, runs all levels, downloads
 * flow (float32 [height][width][2] = (dx,dy), `flow_stride_bytes` per row, 0 =
 * packed).  With OFB_OPTFLOW_USE_INITIAL_FLOW `flow` is read as the initial
 * estimate first (as cv2 does with its in/out `flow` argument). */
int ofb_farneback(ofb_handle* h, const uint8_t* prev, const uint8_t* next, int width, int height,
                  size_t stride_bytes, float* flow, size_t flow_stride_bytes,
                  const ofb_farneback_params* params);

/* Batched host-buffer call: n independent pairs (n <= max_batch); prev[i], next[i],
 * flow[i] as above.  One upload, one batched launch sequence, one download. */
int ofb_farneback_batch(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next,
                        int width, int height, size_t stride_bytes, float* const* flow,
                        size_t flow_stride_bytes, const ofb_farneback_params* params);

/* Asynchronous variant for throughput: enqueues the uploads, kernels and downloads of the batch and
 * returns.  Buffers must be page-locked (else it degrades to the synchronous path) and must stay
 * untouched until ofb_wait.  Successive calls pipeline ACROSS calls: the uploads and kernels of call
 * i+1 overlap the downloads of call i (use a second set of flow buffers for it). */
int ofb_farneback_batch_async(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next,
                              int width, int height, size_t stride_bytes, float* const* flow,
                              size_t flow_stride_bytes, const ofb_farneback_params* params);
/* The node contract in one call (lfn3_sub_node.py:194-212: flow -> np.median / np.mean of u): uploads
 * the n pairs (pipelined with the kernels), computes the flow and reduces it on the device; only the
 * n scalars come back — the 8N-byte field never crosses PCIe.  mask/out_mean/out_median as in
 * ofb_flow_u_stats.  Synchronous. */
int ofb_farneback_batch_stats(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next,
                              int width, int height, size_t stride_bytes, const ofb_farneback_params* params,
                              const uint8_t* mask, double* out_mean, float* out_median);
/* Asynchronous form of ofb_farneback_batch_stats for page-locked frames: returns once the work is enqueued, so
 * the uploads and kernels of the next call run behind this one's (a camera node keeps a call in flight while it
 * receives the next frames).  out_mean / out_median are written by ofb_wait (or by a later call, once more than
 * a few reductions are pending) and must stay valid until then; the frames must stay untouched as for
 * ofb_farneback_batch_async.  Pageable frames are served synchronously.
 * Replaces the same node lines as ofb_farneback_batch_stats (lfn3_sub_node.py:194-212). */
int ofb_farneback_batch_stats_async(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next,
                                    int width, int height, size_t stride_bytes, const ofb_farneback_params* params,
                                    const uint8_t* mask, double* out_mean, float* out_median);
/* Blocks until everything enqueued on the handle (all three streams) has finished and hands the results of
 * pending asynchronous reductions to their callers' arrays. */
int ofb_wait(ofb_handle* h);

/* Device-resident call, asynchronous on the handle's stream: d_prev/d_next are n
 * uint8 images (row pitch `pitch_bytes`, image i at + i*image_stride_bytes);
 * d_flow is n packed float32 [height][width][2] fields.  Returns after enqueueing. */
int ofb_farneback_device(ofb_handle* h, int n, const uint8_t* d_prev, const uint8_t* d_next, int width,
                         int height, size_t pitch_bytes, size_t image_stride_bytes, float* d_flow,
                         const ofb_farneback_params* params);

/* Camera-stream call (consecutive frames of ONE stream: pair i = (frame i, frame i+1)):
 * d_frames holds n+1 frames; the polynomial expansion of each frame is computed
 * once and shared by the two pairs it belongs to.  Results are identical to n
 * independent ofb_farneback_device pairs. */
int ofb_farneback_sequence_device(ofb_handle* h, int n_pairs, const uint8_t* d_frames, int width,
                                  int height, size_t pitch_bytes, size_t image_stride_bytes,
                                  float* d_flow, const ofb_farneback_params* params);

/* ---- camera streams: the temporal state stays on the GPU ---------------------------------------
 * A node calls once per received frame (lfn3_sub_node.py:141-222 keeps `self.prev_tensor` and runs the flow on
 * (previous, current)).  The handle keeps, per stream, what the next call needs from the current frame: the
 * polynomial expansion at every pyramid level (default window; the previous uint8 frame otherwise), so a call
 * uploads ONE frame per stream and expands only that frame.  Results are identical to ofb_farneback on
 * (previous frame, frame).
 *   frames[i]: the new frame of stream i, i < n_streams <= max_batch.  The first call — and the first after a change
 *   of n_streams, size or parameters, or after ofb_stream_reset — only primes the state: *produced = 0 and no flow is
 *   written (the node's "first frame only primes" branch, lfn3_sub_node.py:164-167).  Later calls write flow[i]
 *   (float32 [height][width][2]; flow == NULL keeps the fields on the device for ofb_flow_u_stats /
 *   ofb_flow_postfilter) and set *produced = n_streams.  Synchronous when flow != NULL; with flow == NULL the call
 *   returns once the work is enqueued (page-locked frames: keep them untouched until ofb_wait or a synchronous call). */
int ofb_farneback_stream(ofb_handle* h, int n_streams, const uint8_t* const* frames, int width, int height,
                         size_t stride_bytes, float* const* flow, size_t flow_stride_bytes,
                         const ofb_farneback_params* params, int* produced);
/* The same on device-resident frames (n_streams images at d_frames, row pitch / image stride in bytes), asynchronous
 * on the handle's stream; d_flow receives n_streams packed fields when *produced != 0. */
int ofb_farneback_stream_device(ofb_handle* h, int n_streams, const uint8_t* d_frames, int width, int height,
                                size_t pitch_bytes, size_t image_stride_bytes, float* d_flow,
                                const ofb_farneback_params* params, int* produced);
/* Forgets the streams' state: the next stream call primes again. */
int ofb_stream_reset(ofb_handle* h);

/* ---- spatially tiled mode: ONE frame pair split into row strips over the GPUs of a node -------
 * (BASELINE.json config 5: 7680x4320 over 8 B200.)  One handle per GPU ("rank"), all created with the
 * same max_width/max_height.  Every rank holds the two source frames and computes the rows it owns at
 * every pyramid level; rows owned by neighbours are read through NVLink peer pointers inside the
 * kernels, and a flag barrier in peer memory orders the stages.  Set-up:
 *   ofb_tiled_init(h, rank, world)                        on every rank
 *   ofb_tiled_export(h, blob)  -> all-gather the blobs -> ofb_tiled_import(h, all_blobs)   (one process per GPU,
 *       CUDA IPC), or ofb_tiled_import_local(h, handles) when all handles live in one process. */
#define OFB_TILED_EXPORT_BYTES 320 /* 5 CUDA IPC memory handles */
int ofb_tiled_init(ofb_handle* h, int rank, int world);
int ofb_tiled_export(ofb_handle* h, void* blob /* OFB_TILED_EXPORT_BYTES */);
int ofb_tiled_import(ofb_handle* h, const void* all_blobs /* world * OFB_TILED_EXPORT_BYTES, rank order */);
int ofb_tiled_import_local(ofb_handle* h, ofb_handle* const* handles /* world handles, rank order */);
/* Asynchronous on the handle's stream.  d_prev/d_next: the full uint8 frames on this rank's device;
 * d_flow: a full-size float32 [height][width][2] buffer of which this rank writes rows
 * [*row_begin, *row_end) only.  flags must be 0, winsize in [4, 39], poly_n <= 8, iterations >= 1.
 * All ranks must make the same call; a rank that does not arrive makes the others time out (~2 s),
 * reported by ofb_tiled_status. */
int ofb_farneback_tiled_device(ofb_handle* h, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                               size_t pitch_bytes, float* d_flow, const ofb_farneback_params* params,
                               int* row_begin, int* row_end);
/* Enqueues one cross-GPU flag barrier on the handle's stream (every rank must call it; real multi-GPU
 * set-ups only — never with several ranks on one device). */
int ofb_tiled_barrier(ofb_handle* h);
/* Synchronises the stream; *timed_out = 1 if a cross-GPU barrier gave up waiting since the last call. */
int ofb_tiled_status(ofb_handle* h, int* timed_out);
/* Test path for fewer GPUs than ranks: all `world` handles live in this process on ONE device (set up
 * with ofb_tiled_init + ofb_tiled_import_local); runs the stages of all ranks in order, synchronously,
 * without the barrier kernel, and fills the whole of d_flow. */
int ofb_farneback_tiled_emulated(ofb_handle* const* handles, int world, const uint8_t* d_prev, const uint8_t* d_next,
                                 int width, int height, size_t pitch_bytes, float* d_flow,
                                 const ofb_farneback_params* params);

/* Number of kernel launches this handle has enqueued since creation (bench evidence). */
uint64_t ofb_launch_count(const ofb_handle* h);

/* Per-stage device timing with CUDA events recorded on the handle's stream around every
 * kernel launch of a stage (what bench.py's roofline block reads).  Stages: */
enum {
  OFB_STAGE_PYRAMID = 0,   /* convert + GaussianBlur + resize (a2) */
  OFB_STAGE_POLYEXP = 1,   /* FarnebackPolyExp (a4) */
  OFB_STAGE_ITERATION = 2, /* UpdateMatrices + blur + solve (a5-a7): the dominant kernel */
  OFB_STAGE_FLOW_INIT = 3, /* zero / initial-flow / inter-level upsample (a8, a9) */
  OFB_STAGE_OTHER = 4,
  OFB_NUM_STAGES = 5
};
/* enable != 0 starts a fresh recording (drops earlier samples); 0 stops recording. */
int ofb_timing_enable(ofb_handle* h, int enable);
/* Synchronises the stream and returns, per stage, the summed event time in milliseconds and the
 * number of launches recorded since ofb_timing_enable(h, 1).  Arrays of OFB_NUM_STAGES. */
int ofb_timing_read(ofb_handle* h, double* ms_out, uint64_t* launches_out);
/* The individual samples of one stage, in launch order: ms_out receives up to `capacity` event
 * times (milliseconds), *n_out the number of samples recorded for that stage. */
int ofb_timing_read_samples(ofb_handle* h, int stage, double* ms_out, int capacity, int* n_out);

/* ---- on-device reduction of the flow field (the node contract) ----------------
 * Every reference node collapses the field to one scalar right after the flow
 * call (lfn3_sub_node.py:205-212: np.median(flow_np[0]); opticalflow_node.py:98:
 * np.mean(flow_np[0]); masked: sub_n_pub_lfn3_node.py:195-210).  These reduce the
 * most recent flow field(s) held on the device, so the 8N-byte D2H is avoided.
 * mask: optional uint8 [height][width] host array (non-zero = use), NULL = all.
 * out_mean / out_median: n values (u component); either may be NULL. */
int ofb_flow_u_stats(ofb_handle* h, int n, const uint8_t* mask, double* out_mean, float* out_median);
/* The same, returning once the reduction is enqueued; out_mean / out_median are written by ofb_wait (or by a later
 * asynchronous reduction once a few are pending) and must stay valid until then.  With ofb_farneback_stream(flow =
 * NULL) this keeps a camera node's frame-by-frame loop fully asynchronous: the next frames upload behind the kernels. */
int ofb_flow_u_stats_async(ofb_handle* h, int n, const uint8_t* mask, double* out_mean, float* out_median);

/* The "adapt" node's flow post-processing, on the device, applied to the field(s) of the last flow call on this
 * handle (replaces ros2_ws/src/liteflownet3/liteflownet3/lfn3_adapt_node.py:236-251):
 *   median_ksize 3 | 5 : u = cv2.medianBlur(u, k), v = cv2.medianBlur(v, k)   (0 = off; float32 fields allow 3 and 5)
 *   magnitude_threshold >= 0 : u, v *= (sqrt(u*u + v*v) >= threshold)         (negative = off)
 *   gray != NULL : u, v *= (gray < intensity_threshold); gray = n host uint8 [height][gray_stride_bytes] images
 * in that order, bit-exact against cv2 / NumPy float32.  The filtered field replaces the handle's current field:
 * ofb_flow_u_stats (np.mean(u), :254) and ofb_flow_download see it.  Asynchronous on the handle's stream. */
int ofb_flow_postfilter(ofb_handle* h, int n, int median_ksize, float magnitude_threshold, const uint8_t* const* gray,
                        size_t gray_stride_bytes, int intensity_threshold);
/* Copies the handle's current field(s) (the last flow result, post-filtered or not) to n host float32
 * [height][width][2] arrays (flow_stride_bytes = 0: packed rows).  Synchronous. */
int ofb_flow_download(ofb_handle* h, int n, float* const* flow, size_t flow_stride_bytes);
/* The nodes' flow visualisation `flow_to_color` (ros2_ws/src/liteflownet3/liteflownet3/sub_n_pub_lfn3_node.py:132-140:
 * cartToPolar -> hue = angle / 2, value = magnitude normalised to 0..255 over the frame, saturation 255 -> HSV2BGR) of
 * field `pair` of the handle's current result, computed on the device, bit-exact against the cv2 recipe.
 * bgr_out: host uint8 [height][width][3] (stride_bytes = 0: packed rows).  Synchronous. */
int ofb_flow_to_bgr(ofb_handle* h, int pair, uint8_t* bgr_out, size_t stride_bytes);
/* The sub node's dense view (ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:244-262): hue = uint8(ang * 90 / pi),
 * saturation 255, value = uint8(clip(mag / dt * pixel_to_meter / max_speed, 0, 1) * 255) -> HSV2BGR; float32 arithmetic
 * in NumPy's evaluation order, bit-exact against the recipe (oracle/visual_np.py::flow_to_color_speed).  Synchronous. */
int ofb_flow_to_bgr_speed(ofb_handle* h, int pair, double dt, double pixel_to_meter, double max_speed, uint8_t* bgr_out,
                          size_t stride_bytes);
/* Flow of pair `pair` of the handle's current field at n integer pixel positions xy = [n][2] (x, y): out_dxdy =
 * [n][2] float32 (dx, dy), NaN for positions outside the frame.  The junction node's lookup of the predicted
 * junction positions (ros2_ws/src/liteflownet3/liteflownet3/lfn3_junction_node.py:207-214) without downloading the
 * field.  Synchronous. */
int ofb_flow_sample(ofb_handle* h, int pair, int n_points, const int* xy, float* out_dxdy);

/* ---- frame ingest (the step in front of the flow call: lfn3_sub_node.py:148-159) --------------
 * cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) / COLOR_RGB2GRAY on an interleaved 8-bit 3-channel frame
 * (sensor_msgs/Image bgr8 / rgb8 with row `step`), bit-exact with cv2's 15-bit fixed-point formula.
 * rgb_order: 0 = BGR (cv2 / bgr8), 1 = RGB (rgb8).  Host variant is synchronous; stride 0 = packed. */
int ofb_cvt_gray(ofb_handle* h, const uint8_t* src, int width, int height, size_t stride_bytes, int rgb_order,
                 uint8_t* dst, size_t dst_stride_bytes);
/* Device-resident, asynchronous on the handle's stream. */
int ofb_cvt_gray_device(ofb_handle* h, const uint8_t* d_src, int width, int height, size_t src_pitch_bytes,
                        int rgb_order, uint8_t* d_dst, size_t dst_pitch_bytes);

/* cv2.resize(src, (dst_width, dst_height)) — INTER_LINEAR on uint8, 1 or 3 interleaved channels, bit-exact with this
 * cv2 build (11-bit fixed-point weights; oracle/resize_np.py).  The nodes resize every frame that does not have the
 * configured size before anything else: lfn3_sub_node.py:152-153, lfn3_adapt_node.py:160-161.  Frames of any size
 * (not limited by the handle's max_width/max_height).  Synchronous; stride 0 = packed. */
int ofb_resize_u8(ofb_handle* h, const uint8_t* src, int src_width, int src_height, size_t src_stride_bytes, int channels,
                  uint8_t* dst, int dst_width, int dst_height, size_t dst_stride_bytes);
/* The whole ingest of a node in one call (lfn3_sub_node.py:148-159): a bgr8 / rgb8 frame of any size is uploaded once,
 * resized to dst_width x dst_height if it has another size (cv2.resize, on the colour frame, as the nodes do) and
 * converted to gray (cv2.cvtColor); dst receives the uint8 [dst_height][dst_width] frame the flow calls take. */
int ofb_ingest_gray(ofb_handle* h, const uint8_t* src, int src_width, int src_height, size_t src_stride_bytes,
                    int rgb_order, uint8_t* dst, int dst_width, int dst_height, size_t dst_stride_bytes);

/* ---- JPEG frames (sensor_msgs/CompressedImage): cv2.imdecode(np.frombuffer(msg.data, np.uint8), cv2.IMREAD_COLOR) of the
 * compressed-image node, ros2_ws/src/optical_flow/optical_flow/opticalflow_comprerssed_node.py:43-46.  Baseline JPEG
 * (SOF0/SOF1, 8 bit, Huffman, one interleaved scan; gray-scale or YCbCr 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0, restart intervals),
 * bit-exact with the libjpeg-turbo 3.1.2 build inside the cv2 wheel (islow IDCT, fancy up-sampling, BGR output;
 * oracle/jpeg_np.py).  Entropy decoding (self-synchronising parallel Huffman decode), dequantisation, IDCT, chroma
 * up-sampling, colour conversion and the gray conversion run on the device; the host parses the markers and removes the
 * byte stuffing.  Anything else (progressive,
 * arithmetic, 12 bit, CMYK) returns OFB_ERR_UNSUPPORTED — never a wrong frame, never a CPU decode. */
/* Frame header only (no handle, no device work): size and component count for sizing the output arrays. */
int ofb_jpeg_info(const uint8_t* jpeg, size_t n_bytes, int* width, int* height, int* components);
/* The host half on its own (no handle, no device work): the quantised DCT coefficients of the scan, int16, component
 * after component (Y, Cb, Cr), each a row-major grid of whole blocks (padded to whole MCUs) of 64 coefficients in
 * natural (row-major) order.  coef may be NULL to query *n_coef. */
int ofb_jpeg_entropy_decode(const uint8_t* jpeg, size_t n_bytes, int16_t* coef, size_t coef_capacity, size_t* n_coef);
/* Where the Huffman stream is walked.  Default (0): on the device (self-synchronising parallel decode, restart intervals
 * as independent chains: only the compressed bytes cross PCIe).  1: on the host (coefficient blocks cross PCIe).  Same
 * coefficients either way. */
int ofb_jpeg_set_host_entropy(ofb_handle* h, int on);
/* bgr: host uint8 [height][width][3] = cv2.imdecode(buf, IMREAD_COLOR), and/or gray: host uint8 [height][width] =
 * cv2.cvtColor(that, COLOR_BGR2GRAY); either may be NULL, strides 0 = packed.  Synchronous. */
int ofb_jpeg_decode(ofb_handle* h, const uint8_t* jpeg, size_t n_bytes, uint8_t* bgr, size_t bgr_stride_bytes, uint8_t* gray,
                    size_t gray_stride_bytes);
/* gray: host uint8 [height][width] = cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE): the luma plane as libjpeg delivers it
 * (no colour conversion; differs from cvtColor(BGR2GRAY) of the colour decode by rounding).  Synchronous. */
int ofb_jpeg_decode_luma(ofb_handle* h, const uint8_t* jpeg, size_t n_bytes, uint8_t* gray, size_t gray_stride_bytes);
/* ofb_ingest_gray for a compressed frame: decode, cv2.resize of the colour frame to dst_width x dst_height if it has
 * another size, cv2.cvtColor(BGR2GRAY); only the gray frame crosses PCIe back.  Synchronous. */
int ofb_ingest_jpeg_gray(ofb_handle* h, const uint8_t* jpeg, size_t n_bytes, uint8_t* dst, int dst_width, int dst_height,
                         size_t dst_stride_bytes);

/* cv2.createCLAHE(clip_limit, (tiles_x, tiles_y)).apply(src) on a uint8 single-channel image, bit-exact with this
 * cv2 build (oracle/clahe_np.py).  The adapt node's contrast pre-filter: lfn3_adapt_node.py:164-182
 * (`self.clahe.setClipLimit(clip); v_enhanced = self.clahe.apply(v)`).  Synchronous; stride 0 = packed. */
int ofb_clahe(ofb_handle* h, const uint8_t* src, int width, int height, size_t src_stride_bytes, double clip_limit,
              int tiles_x, int tiles_y, uint8_t* dst, size_t dst_stride_bytes);

/* The adapt node's colour pre-filter in one call (lfn3_adapt_node.py:164-184): cv2.cvtColor(bgr, BGR2HSV), CLAHE on the V
 * channel — with the node's adaptive clip limit (contrast = std(v) / (mean(v) + 1e-3) mapped linearly from
 * [c_min, c_max] to [clip_min, clip_max]) when `adaptive` is set, else `clip_limit` — and cv2.cvtColor(..., HSV2RGB):
 * a bgr8 frame in, the rgb8 frame the node continues with out, every step bit-exact with this cv2 build
 * (oracle/prefilter_np.py).  *clip_used receives the clip limit applied.  The bilateral filter that may follow
 * (:186-190) is ofb_bilateral_u8c3.  Synchronous. */
typedef struct ofb_clahe_params {
  int adaptive;                              /* 1: clip limit from the frame's contrast, 0: clip_limit */
  double clip_limit;
  double clip_min, clip_max, c_min, c_max;   /* adaptive mapping */
  int tiles_x, tiles_y;                      /* tileGridSize (cv2 default 8 x 8) */
} ofb_clahe_params;
int ofb_adapt_prefilter(ofb_handle* h, const uint8_t* bgr, int width, int height, size_t stride_bytes,
                        const ofb_clahe_params* params, uint8_t* rgb, size_t rgb_stride_bytes, double* clip_used);

/* cv2.bilateralFilter(src, d, sigmaColor, sigmaSpace) on a uint8 3-channel frame: the optional last step of the adapt
 * node's pre-filter chain (lfn3_adapt_node.py:186-190, applied to the rgb frame ofb_adapt_prefilter returns).  OpenCV's
 * own 8UC3 algorithm (circular support of radius d/2 — round(1.5 sigmaSpace) for d <= 0 —, REFLECT_101 border, float
 * space / colour weight tables, out = round(sum * (1 / wsum))), bit-exact with its restatement
 * oracle/prefilter_np.py::bilateral_u8c3.  The installed wheel sends 8-bit images through Intel IPP instead and
 * differs from both at rounding ties (a few values per 100 000, by one).  Radius <= 15.  Synchronous. */
int ofb_bilateral_u8c3(ofb_handle* h, const uint8_t* src, int width, int height, size_t src_stride_bytes, int d,
                       double sigma_color, double sigma_space, uint8_t* dst, size_t dst_stride_bytes);

/* ---- junction detector: the reference's own sparse point source --------------------------------------------------
 * find_junctions_not_rotated(img, grid_area, grid_area_threshold, false, eps) of
 * ros2_ws/src/junction_point_detector/src/junction_detector.cpp:31-214, optionally preceded by dampenIntensity(img,
 * dampen_min, dampen_max) (:3-28; the ROS wrapper fishnet_detector_ros.cpp:49-58 calls dampenIntensity(img, -20, 15) and
 * find_junctions_not_rotated(img, 200, 2.0, false, 6)).  cvtColor -> GaussianBlur 3x3 -> adaptiveThreshold (Gaussian 11, 2)
 * -> findContours(RETR_TREE) -> contourArea / boundingRect tests -> box corners -> nanoflann radius clusters.  Pixel
 * stages and contour measurements run on the device, bit-exact with the cv2 4.13.0 wheel; the contours come from two
 * connected-component labellings instead of border following (same areas, boxes and order as cv2.findContours); the
 * final ordering of the passing contours and the KD-tree clustering run on the host exactly as nanoflann does them
 * (approximate search with eps 10 included).  Output: junction centres (x, y) float32, in the reference's order — the
 * points the junction node tracks (lfn3_junction_node.py:203-231) and a point source for ofb_pyrlk. */
typedef struct ofb_junction_params {
  int grid_area;              /* expected cell area in pixels (header default 250; the ROS node passes 200) */
  float grid_area_threshold;  /* accepted area window: grid_area / (2 t) < area < grid_area * 2 t (default 2) */
  int eps;                    /* cluster radius in pixels (header default 4; the ROS node passes 6) */
  int dampen;                 /* 1: dampenIntensity(img, dampen_min, dampen_max) first (3-channel frames only) */
  double dampen_min, dampen_max;
} ofb_junction_params;
/* img: host uint8, 1 channel (gray) or 3 (bgr8).  junctions_xy: capacity x 2 floats, *n_out receives the count.
 * candidates_xy (may be NULL): the box corners before clustering, candidate_capacity x 2 floats, *n_candidates their
 * count.  Synchronous.  OFB_ERR_CAPACITY if an output is too small or contours nest deeper than 48 levels. */
int ofb_find_junctions(ofb_handle* h, const uint8_t* img, int width, int height, size_t stride_bytes, int channels,
                       const ofb_junction_params* params, float* junctions_xy, int capacity, int* n_out,
                       float* candidates_xy, int candidate_capacity, int* n_candidates);
/* The detector's binary image (adaptiveThreshold output, 0 / 255) for inspection and stage-level tests. */
int ofb_junction_threshold(ofb_handle* h, const uint8_t* img, int width, int height, size_t stride_bytes, int channels,
                           const ofb_junction_params* params, uint8_t* thresh, size_t thresh_stride_bytes);
/* The clustering step alone (host only, no handle): candidates -> cluster centres, junction_detector.cpp:123-185. */
int ofb_cluster_junctions(const float* candidates_xy, int n_candidates, int eps, float* junctions_xy, int capacity,
                          int* n_out);

/* ---- sparse path: replaces cv2.goodFeaturesToTrack + cv2.calcOpticalFlowPyrLK -- */

/* Shi-Tomasi corners of a uint8 image (host buffer).  corners_xy: capacity
 * max_corners*2 floats, receives (x,y) pairs in cv2's order; *n_out their count. */
int ofb_good_features(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                      const ofb_gftt_params* params, float* corners_xy, int* n_out);

/* The same with cv2's `mask` argument (uint8 [height][width], 0 = excluded, `mask_stride_bytes` per row, 0 = packed;
 * mask == NULL: no mask).  As in cv2 the quality threshold is relative to the strongest response INSIDE the mask. */
int ofb_good_features_masked(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                             const uint8_t* mask, size_t mask_stride_bytes, const ofb_gftt_params* params,
                             float* corners_xy, int* n_out);

/* cv2.cornerMinEigenVal(image, blockSize=3, ksize=3) — exposed for bit-exact parity tests. */
int ofb_corner_min_eigenval(ofb_handle* h, const uint8_t* image, int width, int height,
                            size_t stride_bytes, int block_size, float* eig_out);

/* cv2.buildOpticalFlowPyramid levels as a cv2.pyrDown chain (uint8, bit-exact) and
 * the Scharr derivative images (int16 [h][w][2] = (dx,dy), bit-exact).
 * level_out[l] (may be NULL) receives packed uint8 [h_l][w_l]; deriv_out[l] (may be
 * NULL) packed int16 [h_l][w_l][2]; *n_levels_out the number of levels built. */
int ofb_lk_pyramid(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                   int win_w, int win_h, int max_level, uint8_t* const* level_out,
                   int16_t* const* deriv_out, int* n_levels_out);

/* Pyramidal Lucas-Kanade.  prev_pts/next_pts: n_points (x,y) float pairs; next_pts is
 * read as the initial guess when OFB_OPTFLOW_USE_INITIAL_FLOW is set; status: n bytes;
 * err: n floats (may be NULL). */
int ofb_pyrlk(ofb_handle* h, const uint8_t* prev, const uint8_t* next, int width, int height,
              size_t stride_bytes, const float* prev_pts, int n_points, float* next_pts,
              uint8_t* status, float* err, const ofb_lk_params* params);

/* Camera-stream form of the sparse path — the worker loop of a direct-camera node
 * (ros2_ws/src/liteflownet3/liteflownet3/lfn3_node.py:145-210: keep the previous frame, process the new one) with
 * goodFeaturesToTrack + calcOpticalFlowPyrLK as the flow call.  One new frame per call; the previous frame's pyramid,
 * Scharr derivatives and corner list stay on the GPU, so a frame is uploaded ONCE and nothing but the results comes
 * back.  Call t tracks the corners detected in frame t-1 into frame t and detects the corners of frame t for call t+1:
 *   prev_pts / next_pts / status / err : the tracked points of (frame t-1 -> frame t), *n_tracked of them
 *                                        (capacity params->max_corners each; any may be NULL); *n_tracked = -1 on a
 *                                        priming call (first frame, or the stream was re-primed)
 *   new_corners / n_new                : the corners of frame t (optional)
 * Bit-identical to ofb_good_features(frame t-1) + ofb_pyrlk(frame t-1, frame t, those corners).  Needs
 * max_corners > 0 (fixed-size result block) and no OFB_OPTFLOW_USE_INITIAL_FLOW.  A change of size or parameters, or
 * any other sparse call on the handle, re-primes the stream. */
int ofb_lk_stream(ofb_handle* h, const uint8_t* frame, int width, int height, size_t stride_bytes,
                  const ofb_gftt_params* gftt, const ofb_lk_params* lk, float* prev_pts, float* next_pts,
                  uint8_t* status, float* err, int* n_tracked, float* new_corners, int* n_new);
int ofb_lk_stream_reset(ofb_handle* h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OFB_H_ */
