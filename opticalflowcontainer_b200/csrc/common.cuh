// common.cuh — shared declarations of libofb (B200 / sm_100a optical-flow engine).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ofb.h"

namespace ofb {

constexpr int kMaxLevels = 16;     // scales per call (cv2 clamps by the 32-px rule long before)
constexpr int kMaxPolyN = 10;      // poly_n <= 10 (cv2 users: 5 or 7)
constexpr int kMaxBlurRadius = 64; // winsize <= 129
constexpr int kRowPad = 10;        // spare rows at the end of the R buffers (the prefetching schedule reads up to 4 rows past a pixel)
constexpr int kPxWaves = 4;        // target CTA waves of the marching PolyExp kernel

// One pyramid scale of the Farneback schedule (FarnebackOpticalFlowImpl::calc).
struct Level {
  int k;
  double scale;
  double sigma;
  int ksize;
  int width, height;
};

// Polynomial-expansion constants (FarnebackPrepareGaussian).
struct PolyCoef {
  int n;
  float g[kMaxPolyN + 1], xg[kMaxPolyN + 1], xxg[kMaxPolyN + 1];
  float ig11, ig03, ig33, ig55;
  // the same weights duplicated into pairs (taps 0..8): 64-bit constant operands of the packed f32x2 arithmetic in
  // the horizontal phase of k_polyexp_march
  float2 g2[9], xg2[9], xxg2[9];
  float2 ig11_2, ig03_2, ig33_2, ig55_2;
};

struct BlurCoef {
  int m;               // radius = winsize / 2
  float scale;         // box: 1 / winsize^2 ; gaussian: 1
  int gaussian;
  float k[kMaxBlurRadius + 1];  // box: all ones
};

// One entry of a cv::resize INTER_LINEAR coordinate table: source index and weight of source index + 1.
struct LinTab {
  int i0;
  float f;
};

// Spatially tiled mode (one frame pair split into row strips over the GPUs of a node): every rank
// holds full-size level buffers but fills only the rows it owns; rows owned by another rank are read
// from that rank's buffer (same offsets) through NVLink peer pointers.  Level rows are split evenly:
// rank r owns [r * rpr, min((r+1) * rpr, h)), rpr = ceil(h / world).
constexpr int kMaxTileRanks = 8;
struct PeerTab {
  const float4* RA[kMaxTileRanks];
  const float* RB[kMaxTileRanks];
  int rpr;                             // rows per rank at the level of the launch
  int world;
  // rows [r_lo, r_hi) of R are valid in THIS rank's buffers (the band it computed); rows outside are read from the
  // rank that owns them.  The flow a launch reads is always local (independent bands, tiled.cuh).
  int r_lo, r_hi;
};

// Polynomial coefficients of the two frames of every pair: frame 0 of pair i at A0/B0 + i*n, frame 1 at A1/B1 + i*n
// (n = level pixels).  Independent pairs: A1 = A0 + n_pairs*n; consecutive frames of a stream: A1 = A0 + n; the
// camera-stream call: A0 = the expansions kept from the previous call, A1 = the new frames' (different arrays).
struct RSet {
  const float4* A0;
  const float* B0;
  const float4* A1;
  const float* B1;
};

// Camera-stream cache (ofb_farneback_stream*): the polynomial expansions of the streams' latest frames at every pyramid
// level, double-buffered — a call expands only the new frames into half `cur` and reads the previous frames' from the
// other half.
struct StreamCtx {
  float4* RA[2][kMaxLevels];
  float* RB[2][kMaxLevels];
  int cur;
  bool prime_only;     // first frame of the streams: expansions only, no flow
};

inline int cv_round(double v) { return (int)__builtin_nearbyint(v); }  // round-half-even like cvRound

}  // namespace ofb

struct ofb_handle {
  int device = 0;
  int max_w = 0, max_h = 0, max_batch = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  bool no_graph = false;       // OFB_GRAPH=0: no CUDA-graph replay of the launch sequence of small batches
  bool graph_bypass = false;   // set by the host-buffer batch call while a batch is split into more chunks than the cache holds
  void* graph_cache = nullptr; // farneback.cu: captured launch sequences (std::vector<GraphEntry>)
  uint64_t graph_clock = 0;    // LRU stamp of the graph cache
  // host-buffer pipeline: copy-in / copy-out streams and their events (api.cu)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  // expansion stream (farneback.cu): the polynomial expansions of the finer levels run beside the coarse levels' iterations
  cudaStream_t s_px = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_px[ofb::kMaxLevels] = {nullptr};
  bool no_overlap = false;     // OFB_OVERLAP=0: everything on the one stream
  std::vector<cudaEvent_t> pipe_ev;
  int pipe_chunk = 0;          // OFB_PIPE_CHUNK: pairs per pipeline chunk of the host-buffer batch call (0 = auto)
  int pipe_parity = 0;         // source staging set of the last whole-batch (reduction) call
  int pipe_n = 0, pipe_c = 0, pipe_w = 0, pipe_h = 0;   // staging layout of the previous pipelined call
  bool no_pipeline = false;    // OFB_NO_PIPELINE=1: serial upload -> compute -> download
  uint64_t launches = 0;
  std::string err;

  // device buffers (allocated once in ofb_create)
  uint8_t* d_src = nullptr;       // staging for host API: [2*max_batch (+1)] images, pitch src_pitch
  size_t src_pitch = 0, src_image_stride = 0;
  float* d_img = nullptr;         // level images I: [frames][h][w] f32
  float4* d_RA = nullptr;         // polynomial coefficients ch0..3: [frames][h][w] float4
  float* d_RB = nullptr;          // polynomial coefficient ch4:     [frames][h][w] float
  float4* d_MA = nullptr;         // generic path: matrix field ch0..3 / blurred
  float* d_MB = nullptr;
  float4* d_VA = nullptr;         // generic path: vertical-pass intermediate
  float* d_VB = nullptr;
  float2* d_flow[3] = {nullptr, nullptr, nullptr};  // ping/pong/previous-level, [batch][h][w] float2
  float* d_flow_out = nullptr;    // staging for host API output [batch][H][W][2]
  float* d_init_flow = nullptr;   // USE_INITIAL_FLOW input [batch][H][W][2]
  // INTER_LINEAR coordinate tables of the inter-level flow upsample, one (x, y) pair per level; built
  // on the host (double arithmetic, as cv::resize does) and cached for the current schedule
  ofb::LinTab* d_lintab = nullptr;
  ofb::LinTab* h_lintab = nullptr;   // pinned
  size_t lintab_cap = 0;             // entries
  int tab_w = 0, tab_h = 0, tab_levels = -1;
  double tab_scale = 0;
  size_t tab_x_off[ofb::kMaxLevels] = {0}, tab_y_off[ofb::kMaxLevels] = {0};
  // spatially tiled mode (tiled.cuh / ofb_tiled_*)
  struct Tile {
    int rank = 0, world = 0;
    bool imported = false, same_process = false;
    void* peer_RA[ofb::kMaxTileRanks] = {nullptr};
    void* peer_RB[ofb::kMaxTileRanks] = {nullptr};
    void* peer_MA[ofb::kMaxTileRanks] = {nullptr};   // expansions of the coarser levels (the generic path's buffers)
    void* peer_MB[ofb::kMaxTileRanks] = {nullptr};
    unsigned* peer_flags[ofb::kMaxTileRanks] = {nullptr};
    unsigned* d_flags = nullptr;     // [world] written by the peers
    int* d_err = nullptr;            // barrier timeout flag
    unsigned epoch = 0;
  } tile;
  // frame ingest (ingest.cu): staging for frames of any size and the cached cv2.resize coordinate tables
  struct Ingest {
    uint8_t* d_a = nullptr;      // source frame
    uint8_t* d_b = nullptr;      // result(s)
    size_t a_bytes = 0, b_bytes = 0;
    void* d_tab = nullptr;
    void* h_tab = nullptr;
    size_t tab_cap = 0;
    int tab_sw = 0, tab_sh = 0, tab_dw = 0, tab_dh = 0;
  } ingest;
  // camera-stream state (api.cu: ofb_farneback_stream*)
  struct Stream {
    ofb::StreamCtx ctx = {};
    void* pool = nullptr;        // one allocation behind ctx.RA / ctx.RB
    uint8_t* d_prev = nullptr;   // the streams' previous frames (u8), for configurations the cache does not serve
    size_t pool_bytes = 0, prev_bytes = 0;
    int n = 0, w = 0, h = 0;     // streams and frame size of the cached state (n == 0: not primed)
    ofb_farneback_params params = {};
    bool cached = false;         // expansions cached (fused path) vs previous frames only
    int up_parity = 0;           // staging half of the last host-buffer stream call
  } stream_state;
  // last result bookkeeping for ofb_flow_u_stats
  const float* last_flow = nullptr;
  int last_n = 0, last_w = 0, last_h = 0;
  double* d_stats = nullptr;
  uint32_t* d_sel = nullptr;   // radix-select state of ofb_flow_u_stats (reduce.cu)
  uint8_t* d_mask = nullptr;
  uint8_t* d_gray = nullptr;   // gray frames of the intensity mask (ofb_flow_postfilter), allocated on first use
  size_t gray_bytes = 0;
  float* d_scratch = nullptr;  // median selection scratch
  // asynchronous reductions (ofb_farneback_batch_stats_async): results land in pinned slots and are handed
  // to the caller's arrays by ofb_wait
  struct PendingStats { double* out_mean; float* out_median; int n; int slot; };
  cudaEvent_t stats_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // "results of slot i are in the pinned buffer"
  std::vector<PendingStats> pending_stats;
  char* h_stats = nullptr;     // pinned, kStatSlots x max_batch x 16 B
  int stats_slot = 0;
  size_t scratch_bytes = 0;

  // pinned host staging
  uint8_t* h_src = nullptr;
  float* h_flow = nullptr;
  size_t h_src_bytes = 0, h_flow_bytes = 0;

  // sparse path buffers are owned by lk.cu / features.cu state
  void* sparse = nullptr;
  void* junction = nullptr;     // junction.cu: planes, labels and contour records of the junction detector
  bool jpeg_host_entropy = false;   // ofb_jpeg_set_host_entropy: Huffman decoding on the host even without restart markers
  void* jpeg = nullptr;         // jpeg.cu: coefficient staging and sample planes

  // per-stage event timing (ofb_timing_enable / ofb_timing_read)
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;     // pairs (start, stop)
  std::vector<int> ev_stage;            // stage of pair i
  size_t ev_used = 0;                   // pairs in use
};

namespace ofb {

extern thread_local std::string g_create_error;

int set_error(ofb_handle* h, int status, const char* fmt, ...);

#define OFB_CUDA(h, call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ofb::set_error((h), OFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,              \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                   \
  } while (0)

#define OFB_LAUNCH_CHECK(h)                                                                 \
  do {                                                                                      \
    (h)->launches++;                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return ofb::set_error((h), OFB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",          \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                   \
  } while (0)

// RAII-free stage timer: call begin before the launch(es) of a stage and end after.
int timing_begin(ofb_handle* h, int stage);
int timing_end(ofb_handle* h);

// ---- farneback.cu --------------------------------------------------------------------
int build_schedule(int width, int height, double pyr_scale, int levels, Level* out, int* n_out);
void prepare_poly(int n, double sigma, PolyCoef* pc);
void prepare_blur(int winsize, bool gaussian, BlurCoef* bc);

// Runs the whole multi-level schedule for n_pairs pairs on the handle's stream.
//   sequence == false: frames are [prev_0..prev_{n-1}] at d_prev and [next_0..] at d_next
//   sequence == true : d_prev holds n_pairs+1 consecutive frames (d_next ignored)
int farneback_run(ofb_handle* h, int n_pairs, bool sequence, const uint8_t* d_prev, const uint8_t* d_next,
                  int width, int height, size_t pitch, size_t image_stride, float* d_flow_out,
                  const float* d_init_flow, const ofb_farneback_params* p, const StreamCtx* sc = nullptr);
// k_iter_v with the window radius as a template argument for the common window sizes other than the default
// (iter_fixed_a.cu, iter_fixed_b.cu: separate translation units so that they compile in parallel).  *served = false if
// there is no instantiation for m (the caller then takes the run-time-radius kernel).
cudaError_t launch_iter_fixed_a(ofb_handle* h, int m, const float2* fin, float2* fout, int w, int hh, int n_pairs,
                                const RSet& rs, float reg, cudaStream_t st, bool* served);
cudaError_t launch_iter_fixed_b(ofb_handle* h, int m, const float2* fin, float2* fout, int w, int hh, int n_pairs,
                                const RSet& rs, float reg, cudaStream_t st, bool* served);
void farneback_graphs_destroy(ofb_handle* h);
// the schedule farneback_run would use (level sizes), for sizing the stream cache
int farneback_levels(int width, int height, const ofb_farneback_params* p, Level* out, int* n_out);
// true if farneback_run can serve this configuration from the stream cache (fused box-window path, marching PolyExp)
bool farneback_stream_supported(const ofb_handle* h, const ofb_farneback_params* p);

// Spatially tiled mode (tiled.cuh)
int farneback_run_tiled(ofb_handle* h, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                        size_t pitch, float* d_flow_out, const ofb_farneback_params* p, int* row_begin, int* row_end);
int tiled_barrier_public(ofb_handle* h);
int farneback_run_tiled_emulated(ofb_handle* const* hs, int world, const uint8_t* d_prev, const uint8_t* d_next,
                                 int width, int height, size_t pitch, float* d_flow_out,
                                 const ofb_farneback_params* p);

constexpr int kStatSlots = 4;   // reductions in flight before ofb_farneback_batch_stats_async drains them itself
int flow_u_stats(ofb_handle* h, int n, const uint8_t* host_mask, double* out_mean, float* out_median, bool async = false);
int finish_pending_stats(ofb_handle* h);
// ---- postfilter.cu
int flow_postfilter(ofb_handle* h, int n, int median_ksize, float magnitude_threshold, const uint8_t* const* gray,
                    size_t gray_stride, int intensity_threshold);
int flow_download(ofb_handle* h, int n, float* const* flow, size_t flow_stride_bytes);
int flow_to_bgr(ofb_handle* h, int pair, uint8_t* bgr_out, size_t stride_bytes, int mode = 0, float dt = 1.f, float p2m = 1.f,
                float vmax = 1.f);
int flow_sample(ofb_handle* h, int pair, int n_points, const int* xy, float* out_dxdy);   // synchronises the stream, copies staged scalars to the callers' arrays

// ---- ingest.cu / jpeg.cu
int ingest_reserve(ofb_handle* h, size_t bytes_a, size_t bytes_b);
int resize_u8_device(ofb_handle* h, const uint8_t* d_src, size_t sp, int sw, int sh, int cn, uint8_t* d_dst, size_t dp,
                     int dw, int dh);
int cvt_gray_device(ofb_handle* h, const uint8_t* d_src, size_t src_pitch, uint8_t* d_dst, size_t dst_pitch, int w,
                    int hh, int rgb_order);
void jpeg_destroy(ofb_handle* h);
void junction_destroy(ofb_handle* h);

// ---- sparse ---------------------------------------------------------------------------
void sparse_destroy(ofb_handle* h);

}  // namespace ofb
