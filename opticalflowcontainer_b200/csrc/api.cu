// api.cu — the C ABI of libofb.so (include/ofb.h): handle lifetime, argument validation with
// cv2's error behaviour, pinned-host staging and the host/device entry points of the dense path.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>

#include <new>

#include "common.cuh"

namespace ofb {

thread_local std::string g_create_error;

int set_error(ofb_handle* h, int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return status;
}

int timing_begin(ofb_handle* h, int stage) {
  if (!h->timing) return OFB_OK;
  if (h->ev_used * 2 + 2 > h->ev_pool.size()) {
    for (int i = 0; i < 2; i++) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return set_error(h, OFB_ERR_CUDA, "cudaEventCreate failed");
      h->ev_pool.push_back(e);
    }
    h->ev_stage.push_back(stage);
  }
  h->ev_stage[h->ev_used] = stage;
  OFB_CUDA(h, cudaEventRecord(h->ev_pool[h->ev_used * 2], h->stream));
  return OFB_OK;
}

int timing_end(ofb_handle* h) {
  if (!h->timing) return OFB_OK;
  OFB_CUDA(h, cudaEventRecord(h->ev_pool[h->ev_used * 2 + 1], h->stream));
  h->ev_used++;
  return OFB_OK;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// cv2 argument checks of calcOpticalFlowFarneback (optflowgf.cpp: CV_Assert lines) — INVALID_ARG
// where cv2 raises cv::error.
static int validate_farneback(ofb_handle* h, int n, int width, int height, const ofb_farneback_params* p) {
  if (!p) return set_error(h, OFB_ERR_INVALID_ARG, "params is NULL");
  if (n < 1) return set_error(h, OFB_ERR_INVALID_ARG, "need at least one frame pair");
  if (width < 2 || height < 2) return set_error(h, OFB_ERR_INVALID_ARG, "image must be at least 2x2");
  if (n > h->max_batch || width > h->max_w || height > h->max_h || (size_t)width * height > (size_t)h->max_w * h->max_h)
    return set_error(h, OFB_ERR_CAPACITY, "request %d x %dx%d exceeds handle capacity %d x %dx%d", n, width, height,
                     h->max_batch, h->max_w, h->max_h);
  if (!(p->pyr_scale > 0.0 && p->pyr_scale < 1.0))
    return set_error(h, OFB_ERR_INVALID_ARG, "pyr_scale must be in (0,1)");
  if (p->levels < 0) return set_error(h, OFB_ERR_INVALID_ARG, "levels must be >= 0");
  if (p->iterations < 0) return set_error(h, OFB_ERR_INVALID_ARG, "iterations must be >= 0");
  if (p->winsize < 1 || p->winsize / 2 > kMaxBlurRadius)
    return set_error(h, OFB_ERR_INVALID_ARG, "winsize must be in [1,%d]", 2 * kMaxBlurRadius + 1);
  if (p->poly_n < 1 || p->poly_n > kMaxPolyN)
    return set_error(h, OFB_ERR_INVALID_ARG, "poly_n must be in [1,%d]", kMaxPolyN);
  if (p->flags & ~(OFB_OPTFLOW_USE_INITIAL_FLOW | OFB_OPTFLOW_FARNEBACK_GAUSSIAN))
    return set_error(h, OFB_ERR_INVALID_ARG, "unsupported flags 0x%x", p->flags);
  return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_version(void) { return OFB_VERSION; }

const char* ofb_status_string(int s) {
  switch (s) {
    case OFB_OK: return "OK";
    case OFB_ERR_INVALID_ARG: return "invalid argument";
    case OFB_ERR_CUDA: return "CUDA error";
    case OFB_ERR_NO_DEVICE: return "no usable CUDA device (libofb has no CPU fallback)";
    case OFB_ERR_CAPACITY: return "request exceeds handle capacity";
    case OFB_ERR_ALLOC: return "allocation failed";
    case OFB_ERR_UNSUPPORTED: return "stream not decoded on the device";
    default: return "unknown status";
  }
}

const char* ofb_last_error(const ofb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void* ofb_stream(ofb_handle* h) { return h ? (void*)h->stream : nullptr; }

uint64_t ofb_launch_count(const ofb_handle* h) { return h ? h->launches : 0; }

int ofb_timing_enable(ofb_handle* h, int enable) {
  if (!h) return OFB_ERR_INVALID_ARG;
  OFB_CUDA(h, cudaSetDevice(h->device));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->timing = enable != 0;
  h->ev_used = 0;
  return OFB_OK;
}

int ofb_timing_read(ofb_handle* h, double* ms_out, uint64_t* launches_out) {
  if (!h || !ms_out || !launches_out) return OFB_ERR_INVALID_ARG;
  OFB_CUDA(h, cudaSetDevice(h->device));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < OFB_NUM_STAGES; i++) { ms_out[i] = 0; launches_out[i] = 0; }
  for (size_t i = 0; i < h->ev_used; i++) {
    float ms = 0.f;
    OFB_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
    ms_out[h->ev_stage[i]] += ms;
    launches_out[h->ev_stage[i]]++;
  }
  return OFB_OK;
}

int ofb_timing_read_samples(ofb_handle* h, int stage, double* ms_out, int capacity, int* n_out) {
  if (!h || !ms_out || !n_out || stage < 0 || stage >= OFB_NUM_STAGES || capacity < 0) return OFB_ERR_INVALID_ARG;
  OFB_CUDA(h, cudaSetDevice(h->device));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  int n = 0;
  for (size_t i = 0; i < h->ev_used; i++) {
    if (h->ev_stage[i] != stage) continue;
    if (n < capacity) {
      float ms = 0.f;
      OFB_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
      ms_out[n] = ms;
    }
    n++;
  }
  *n_out = n;
  return OFB_OK;
}

int ofb_synchronize(ofb_handle* h) {
  if (!h) return OFB_ERR_INVALID_ARG;
  OFB_CUDA(h, cudaSetDevice(h->device));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->tile.imported && h->tile.d_err) {
    // tiled mode: a flag barrier that timed out let later stages run on halo rows that were not final
    int v = 0;
    OFB_CUDA(h, cudaMemcpy(&v, h->tile.d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (v) {
      OFB_CUDA(h, cudaMemset(h->tile.d_err, 0, sizeof(int)));
      return set_error(h, OFB_ERR_CUDA, "tiled mode: a peer did not reach the stage barrier in time; the field of this call is invalid");
    }
  }
  return OFB_OK;
}

int ofb_destroy(ofb_handle* h) {
  if (!h) return OFB_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  sparse_destroy(h);
  jpeg_destroy(h);
  junction_destroy(h);
  cudaFree(h->d_src); cudaFree(h->d_img); cudaFree(h->d_RA); cudaFree(h->d_RB);
  cudaFree(h->d_MA); cudaFree(h->d_MB); cudaFree(h->d_VA); cudaFree(h->d_VB);
  for (int i = 0; i < 3; i++) cudaFree(h->d_flow[i]);
  cudaFree(h->d_flow_out); cudaFree(h->d_init_flow); cudaFree(h->d_stats); cudaFree(h->d_mask);
  cudaFree(h->d_scratch); cudaFree(h->d_sel); cudaFree(h->d_lintab);
  if (h->h_lintab) cudaFreeHost(h->h_lintab);
  if (h->h_src) cudaFreeHost(h->h_src);
  if (h->h_flow) cudaFreeHost(h->h_flow);
  farneback_graphs_destroy(h);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  for (auto& e : h->stats_ev) if (e) cudaEventDestroy(e);
  if (h->d_gray) cudaFree(h->d_gray);
  if (h->ingest.d_a) cudaFree(h->ingest.d_a);
  if (h->ingest.d_b) cudaFree(h->ingest.d_b);
  if (h->ingest.d_tab) cudaFree(h->ingest.d_tab);
  if (h->ingest.h_tab) cudaFreeHost(h->ingest.h_tab);
  if (h->stream_state.pool) cudaFree(h->stream_state.pool);
  if (h->stream_state.d_prev) cudaFree(h->stream_state.d_prev);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->tile.imported && !h->tile.same_process) {
    for (int r = 0; r < h->tile.world; r++) {
      if (r == h->tile.rank) continue;
      cudaIpcCloseMemHandle(h->tile.peer_RA[r]);
      cudaIpcCloseMemHandle(h->tile.peer_RB[r]);
      cudaIpcCloseMemHandle(h->tile.peer_MA[r]);
      cudaIpcCloseMemHandle(h->tile.peer_MB[r]);
      cudaIpcCloseMemHandle(h->tile.peer_flags[r]);
    }
  }
  cudaFree(h->tile.d_flags);
  cudaFree(h->tile.d_err);
  for (cudaEvent_t e : h->pipe_ev) cudaEventDestroy(e);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->s_px) cudaStreamDestroy(h->s_px);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int i = 0; i < ofb::kMaxLevels; i++) if (h->ev_px[i]) cudaEventDestroy(h->ev_px[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return OFB_OK;
}

int ofb_create(int device, int max_width, int max_height, int max_batch, ofb_handle** out) {
  if (!out) return set_error(nullptr, OFB_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (max_width < 1 || max_height < 1 || max_batch < 1)
    return set_error(nullptr, OFB_ERR_INVALID_ARG, "max_width, max_height, max_batch must be >= 1");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(nullptr, OFB_ERR_NO_DEVICE, "no CUDA device: %s (libofb has no CPU fallback)",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count)
    return set_error(nullptr, OFB_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, count);
  ofb_handle* h = new (std::nothrow) ofb_handle();
  if (!h) return set_error(nullptr, OFB_ERR_ALLOC, "out of host memory");
  h->device = device;
  h->max_w = max_width;
  h->max_h = max_height;
  h->max_batch = max_batch;
#define CREATE_CUDA(call)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      int st__ = set_error(nullptr, e__ == cudaErrorMemoryAllocation ? OFB_ERR_ALLOC : OFB_ERR_CUDA, \
                           "%s failed: %s", #call, cudaGetErrorString(e__));                   \
      ofb_destroy(h);                                                                          \
      return st__;                                                                             \
    }                                                                                          \
  } while (0)
  CREATE_CUDA(cudaSetDevice(device));
  {
    // the compute stream outranks the expansion stream: when both have CTAs waiting, the critical path (the iterations)
    // gets the SM first and the expansions of the finer levels fill what is left
    int least = 0, greatest = 0;
    CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CREATE_CUDA(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
    CREATE_CUDA(cudaStreamCreateWithPriority(&h->s_px, cudaStreamNonBlocking, least));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < ofb::kMaxLevels; i++) CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_px[i], cudaEventDisableTiming));
  }
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  CREATE_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  {
    // host-side debugging knobs only (kernel experiments are compile-time: tools/build_variant.sh)
    const char* gr = getenv("OFB_GRAPH");
    h->no_graph = gr && gr[0] == '0';
    const char* pch = getenv("OFB_PIPE_CHUNK");
    if (pch) h->pipe_chunk = std::max(0, atoi(pch));
    const char* ov = getenv("OFB_OVERLAP");
    h->no_overlap = ov && ov[0] == '0';
    const char* np = getenv("OFB_NO_PIPELINE");
    h->no_pipeline = np && np[0] == '1';
  }
  const size_t N = (size_t)max_width * max_height;
  const size_t frames = 2 * (size_t)max_batch;            // prev + next of a full batch
  const size_t staged_frames = 2 * frames;                // two such sets: the reduction calls alternate between them
  h->src_pitch = align_up((size_t)max_width, 256);
  h->src_image_stride = h->src_pitch * max_height;
  CREATE_CUDA(cudaMalloc(&h->d_src, h->src_image_stride * staged_frames));
  CREATE_CUDA(cudaMalloc(&h->d_img, frames * N * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&h->d_RA, (frames * N + (size_t)kRowPad * max_width) * sizeof(float4)));
  CREATE_CUDA(cudaMalloc(&h->d_RB, (frames * N + (size_t)kRowPad * max_width) * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&h->d_MA, max_batch * N * sizeof(float4)));
  CREATE_CUDA(cudaMalloc(&h->d_MB, max_batch * N * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&h->d_VA, max_batch * N * sizeof(float4)));
  CREATE_CUDA(cudaMalloc(&h->d_VB, max_batch * N * sizeof(float)));
  for (int i = 0; i < 2; i++) CREATE_CUDA(cudaMalloc(&h->d_flow[i], max_batch * N * sizeof(float2)));
  CREATE_CUDA(cudaMalloc(&h->d_flow_out, max_batch * N * sizeof(float2)));
  CREATE_CUDA(cudaMalloc(&h->d_init_flow, max_batch * N * sizeof(float2)));
  CREATE_CUDA(cudaMalloc(&h->d_stats, 64 * sizeof(double) * (size_t)max_batch));
  CREATE_CUDA(cudaMalloc(&h->d_mask, N));
  CREATE_CUDA(cudaMalloc(&h->d_sel, (size_t)max_batch * 4104 * sizeof(uint32_t)));
  h->lintab_cap = (size_t)kMaxLevels * ((size_t)max_width + max_height);
  CREATE_CUDA(cudaMalloc(&h->d_lintab, h->lintab_cap * sizeof(LinTab)));
  CREATE_CUDA(cudaHostAlloc(&h->h_lintab, h->lintab_cap * sizeof(LinTab), cudaHostAllocDefault));
  h->h_src_bytes = h->src_image_stride * frames;
  h->h_flow_bytes = max_batch * N * sizeof(float2);
  CREATE_CUDA(cudaHostAlloc(&h->h_src, h->h_src_bytes, cudaHostAllocDefault));
  CREATE_CUDA(cudaHostAlloc(&h->h_flow, h->h_flow_bytes, cudaHostAllocDefault));
  CREATE_CUDA(cudaHostAlloc(&h->h_stats, (size_t)kStatSlots * max_batch * 16, cudaHostAllocDefault));
#undef CREATE_CUDA
  *out = h;
  return OFB_OK;
}

// ------------------------------------------------------------------------------------------
// Dense path entry points
// ------------------------------------------------------------------------------------------
int ofb_farneback_device(ofb_handle* h, int n, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                         size_t pitch_bytes, size_t image_stride_bytes, float* d_flow,
                         const ofb_farneback_params* params) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!d_prev || !d_next || !d_flow) return set_error(h, OFB_ERR_INVALID_ARG, "NULL device pointer");
  int st = validate_farneback(h, n, width, height, params);
  if (st) return st;
  if (pitch_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "pitch smaller than width");
  OFB_CUDA(h, cudaSetDevice(h->device));
  const float* init = nullptr;
  if (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) {
    // cv2 semantics: `flow` is in/out.  Copy the initial estimate aside (d_flow is overwritten last).
    OFB_CUDA(h, cudaMemcpyAsync(h->d_init_flow, d_flow, (size_t)n * width * height * sizeof(float2),
                                cudaMemcpyDeviceToDevice, h->stream));
    init = h->d_init_flow;
  }
  return farneback_run(h, n, false, d_prev, d_next, width, height, pitch_bytes, image_stride_bytes, d_flow, init,
                       params);
}

// ------------------------------------------------------------------------------------------
// Spatially tiled mode
// ------------------------------------------------------------------------------------------
int ofb_tiled_init(ofb_handle* h, int rank, int world) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (world < 1 || world > kMaxTileRanks || rank < 0 || rank >= world)
    return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode: need 0 <= rank < world <= %d", kMaxTileRanks);
  OFB_CUDA(h, cudaSetDevice(h->device));
  if (!h->tile.d_flags) {
    OFB_CUDA(h, cudaMalloc(&h->tile.d_flags, kMaxTileRanks * sizeof(unsigned)));
    OFB_CUDA(h, cudaMalloc(&h->tile.d_err, sizeof(int)));
  }
  OFB_CUDA(h, cudaMemset(h->tile.d_flags, 0, kMaxTileRanks * sizeof(unsigned)));
  OFB_CUDA(h, cudaMemset(h->tile.d_err, 0, sizeof(int)));
  h->tile.rank = rank;
  h->tile.world = world;
  h->tile.epoch = 0;
  h->tile.imported = false;
  return OFB_OK;
}

int ofb_tiled_export(ofb_handle* h, void* blob) {
  if (!h || !blob) return OFB_ERR_INVALID_ARG;
  if (!h->tile.d_flags) return set_error(h, OFB_ERR_INVALID_ARG, "ofb_tiled_init first");
  OFB_CUDA(h, cudaSetDevice(h->device));
  static_assert(5 * sizeof(cudaIpcMemHandle_t) == OFB_TILED_EXPORT_BYTES, "export blob size");
  cudaIpcMemHandle_t* out = reinterpret_cast<cudaIpcMemHandle_t*>(blob);
  OFB_CUDA(h, cudaIpcGetMemHandle(&out[0], h->d_RA));
  OFB_CUDA(h, cudaIpcGetMemHandle(&out[1], h->d_RB));
  OFB_CUDA(h, cudaIpcGetMemHandle(&out[2], h->d_MA));
  OFB_CUDA(h, cudaIpcGetMemHandle(&out[3], h->d_MB));
  OFB_CUDA(h, cudaIpcGetMemHandle(&out[4], h->tile.d_flags));
  return OFB_OK;
}

static void tiled_set_self(ofb_handle* h) {
  const int r = h->tile.rank;
  h->tile.peer_RA[r] = h->d_RA;
  h->tile.peer_RB[r] = h->d_RB;
  h->tile.peer_MA[r] = h->d_MA;
  h->tile.peer_MB[r] = h->d_MB;
  h->tile.peer_flags[r] = h->tile.d_flags;
}

int ofb_tiled_import(ofb_handle* h, const void* all_blobs) {
  if (!h || !all_blobs) return OFB_ERR_INVALID_ARG;
  if (!h->tile.d_flags) return set_error(h, OFB_ERR_INVALID_ARG, "ofb_tiled_init first");
  OFB_CUDA(h, cudaSetDevice(h->device));
  const cudaIpcMemHandle_t* in = reinterpret_cast<const cudaIpcMemHandle_t*>(all_blobs);
  for (int r = 0; r < h->tile.world; r++) {
    if (r == h->tile.rank) continue;
    void* p[5];
    for (int k = 0; k < 5; k++)
      OFB_CUDA(h, cudaIpcOpenMemHandle(&p[k], in[r * 5 + k], cudaIpcMemLazyEnablePeerAccess));
    h->tile.peer_RA[r] = p[0];
    h->tile.peer_RB[r] = p[1];
    h->tile.peer_MA[r] = p[2];
    h->tile.peer_MB[r] = p[3];
    h->tile.peer_flags[r] = reinterpret_cast<unsigned*>(p[4]);
  }
  tiled_set_self(h);
  h->tile.imported = true;
  h->tile.same_process = false;
  return OFB_OK;
}

int ofb_tiled_import_local(ofb_handle* h, ofb_handle* const* handles) {
  if (!h || !handles) return OFB_ERR_INVALID_ARG;
  if (!h->tile.d_flags) return set_error(h, OFB_ERR_INVALID_ARG, "ofb_tiled_init first");
  OFB_CUDA(h, cudaSetDevice(h->device));
  for (int r = 0; r < h->tile.world; r++) {
    ofb_handle* q = handles[r];
    if (!q || !q->tile.d_flags) return set_error(h, OFB_ERR_INVALID_ARG, "peer %d is not initialised for tiled mode", r);
    if (q->max_w != h->max_w || q->max_h != h->max_h)
      return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode needs handles of identical capacity");
    if (q->device != h->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return set_error(h, OFB_ERR_CUDA, "no peer access %d -> %d: %s", h->device, q->device, cudaGetErrorString(e));
    }
    h->tile.peer_RA[r] = q->d_RA;
    h->tile.peer_RB[r] = q->d_RB;
    h->tile.peer_MA[r] = q->d_MA;
    h->tile.peer_MB[r] = q->d_MB;
    h->tile.peer_flags[r] = q->tile.d_flags;
  }
  h->tile.imported = true;
  h->tile.same_process = true;
  return OFB_OK;
}

static int validate_tiled(ofb_handle* h, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                          size_t pitch_bytes, float* d_flow, const ofb_farneback_params* params) {
  if (!d_prev || !d_next || !d_flow) return set_error(h, OFB_ERR_INVALID_ARG, "NULL device pointer");
  int st = validate_farneback(h, 1, width, height, params);
  if (st) return st;
  if (pitch_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "pitch smaller than width");
  if (!h->tile.imported) return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode is not set up (init/export/import)");
  return OFB_OK;
}

int ofb_farneback_tiled_device(ofb_handle* h, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                               size_t pitch_bytes, float* d_flow, const ofb_farneback_params* params,
                               int* row_begin, int* row_end) {
  if (!h || !row_begin || !row_end) return OFB_ERR_INVALID_ARG;
  int st = validate_tiled(h, d_prev, d_next, width, height, pitch_bytes, d_flow, params);
  if (st) return st;
  OFB_CUDA(h, cudaSetDevice(h->device));
  return farneback_run_tiled(h, d_prev, d_next, width, height, pitch_bytes, d_flow, params, row_begin, row_end);
}

int ofb_tiled_barrier(ofb_handle* h) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!h->tile.imported) return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode is not set up (init/export/import)");
  OFB_CUDA(h, cudaSetDevice(h->device));
  return tiled_barrier_public(h);
}

int ofb_tiled_status(ofb_handle* h, int* timed_out) {
  if (!h || !timed_out) return OFB_ERR_INVALID_ARG;
  if (!h->tile.d_err) return set_error(h, OFB_ERR_INVALID_ARG, "ofb_tiled_init first");
  OFB_CUDA(h, cudaSetDevice(h->device));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  int v = 0;
  OFB_CUDA(h, cudaMemcpy(&v, h->tile.d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (v) OFB_CUDA(h, cudaMemset(h->tile.d_err, 0, sizeof(int)));
  *timed_out = v;
  return OFB_OK;
}

int ofb_farneback_tiled_emulated(ofb_handle* const* handles, int world, const uint8_t* d_prev, const uint8_t* d_next,
                                 int width, int height, size_t pitch_bytes, float* d_flow,
                                 const ofb_farneback_params* params) {
  if (!handles || world < 1 || world > kMaxTileRanks || !handles[0]) return OFB_ERR_INVALID_ARG;
  for (int r = 0; r < world; r++) {
    if (!handles[r]) return OFB_ERR_INVALID_ARG;
    int st = validate_tiled(handles[r], d_prev, d_next, width, height, pitch_bytes, d_flow, params);
    if (st) return st;
    if (handles[r]->device != handles[0]->device || handles[r]->tile.world != world || handles[r]->tile.rank != r)
      return set_error(handles[r], OFB_ERR_INVALID_ARG, "emulated tiled run needs ranks 0..world-1 on one device");
  }
  OFB_CUDA(handles[0], cudaSetDevice(handles[0]->device));
  return farneback_run_tiled_emulated(handles, world, d_prev, d_next, width, height, pitch_bytes, d_flow, params);
}

int ofb_farneback_sequence_device(ofb_handle* h, int n_pairs, const uint8_t* d_frames, int width, int height,
                                  size_t pitch_bytes, size_t image_stride_bytes, float* d_flow,
                                  const ofb_farneback_params* params) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!d_frames || !d_flow) return set_error(h, OFB_ERR_INVALID_ARG, "NULL device pointer");
  int st = validate_farneback(h, n_pairs, width, height, params);
  if (st) return st;
  if (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)
    return set_error(h, OFB_ERR_INVALID_ARG, "USE_INITIAL_FLOW is not supported by the sequence call");
  if (pitch_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "pitch smaller than width");
  OFB_CUDA(h, cudaSetDevice(h->device));
  return farneback_run(h, n_pairs, true, d_frames, d_frames, width, height, pitch_bytes, image_stride_bytes, d_flow,
                       nullptr, params);
}

// ---- camera streams: temporal state on the device --------------------------------------------------------------
static bool same_params(const ofb_farneback_params& a, const ofb_farneback_params& b) {
  return a.pyr_scale == b.pyr_scale && a.levels == b.levels && a.winsize == b.winsize && a.iterations == b.iterations &&
         a.poly_n == b.poly_n && a.poly_sigma == b.poly_sigma && a.flags == b.flags;
}

// (re)build the stream cache for n streams of width x height; returns OFB_OK with st.n == n
static int stream_prepare(ofb_handle* h, int n, int width, int height, const ofb_farneback_params* p) {
  ofb_handle::Stream& st = h->stream_state;
  const bool cached = farneback_stream_supported(h, p);
  const size_t N = (size_t)width * height;
  if (cached) {
    Level sched[kMaxLevels];
    int nl = 0;
    if (farneback_levels(width, height, p, sched, &nl) != OFB_OK)
      return set_error(h, OFB_ERR_INVALID_ARG, "too many pyramid levels");
    size_t need = 0;
    for (int l = 0; l < nl; l++)   // per level and half: n frames + spare rows (prefetching schedules read ahead)
      need += 2 * (((size_t)n * sched[l].width * sched[l].height + (size_t)kRowPad * sched[l].width) * 20 + 512);
    if (need > st.pool_bytes) {
      OFB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (st.pool) cudaFree(st.pool);
      st.pool = nullptr; st.pool_bytes = 0;
      if (cudaMalloc(&st.pool, need) != cudaSuccess) {
        cudaGetLastError();
        return set_error(h, OFB_ERR_ALLOC, "stream cache: cannot allocate %zu MB", need >> 20);
      }
      st.pool_bytes = need;
    }
    char* q = static_cast<char*>(st.pool);
    for (int half = 0; half < 2; half++)
      for (int l = 0; l < nl; l++) {
        const size_t e = (size_t)n * sched[l].width * sched[l].height + (size_t)kRowPad * sched[l].width;
        st.ctx.RA[half][l] = reinterpret_cast<float4*>(q); q += align_up(e * 16, 256);
        st.ctx.RB[half][l] = reinterpret_cast<float*>(q);  q += align_up(e * 4, 256);
      }
  } else if (2 * n * N > st.prev_bytes) {      // two packed frame sets: previous and current, alternating
    OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (st.d_prev) cudaFree(st.d_prev);
    st.d_prev = nullptr; st.prev_bytes = 0;
    OFB_CUDA(h, cudaMalloc(&st.d_prev, 2 * n * N));
    st.prev_bytes = 2 * n * N;
  }
  st.ctx.cur = 0;
  st.cached = cached;
  st.n = n; st.w = width; st.h = height; st.params = *p;
  return OFB_OK;
}

int ofb_stream_reset(ofb_handle* h) {
  if (!h) return OFB_ERR_INVALID_ARG;
  h->stream_state.n = 0;
  return OFB_OK;
}

int ofb_farneback_stream_device(ofb_handle* h, int n_streams, const uint8_t* d_frames, int width, int height,
                                size_t pitch_bytes, size_t image_stride_bytes, float* d_flow,
                                const ofb_farneback_params* params, int* produced) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (produced) *produced = 0;
  if (!d_frames || !d_flow) return set_error(h, OFB_ERR_INVALID_ARG, "NULL device pointer");
  int s = validate_farneback(h, n_streams, width, height, params);
  if (s) return s;
  if (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)
    return set_error(h, OFB_ERR_INVALID_ARG, "USE_INITIAL_FLOW is not supported by the stream call");
  if (pitch_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "pitch smaller than width");
  OFB_CUDA(h, cudaSetDevice(h->device));
  ofb_handle::Stream& st = h->stream_state;
  const bool prime = st.n != n_streams || st.w != width || st.h != height || !same_params(st.params, *params);
  if (prime && (s = stream_prepare(h, n_streams, width, height, params))) return s;
  const size_t N = (size_t)width * height;
  if (st.cached) {
    st.ctx.prime_only = prime;
    st.ctx.cur ^= 1;                              // the new frames' expansions go to the other half
    s = farneback_run(h, n_streams, false, nullptr, d_frames, width, height, pitch_bytes, image_stride_bytes, d_flow,
                      nullptr, params, &st.ctx);
    if (s) { st.n = 0; return s; }
  } else {
    // configurations the expansion cache does not serve (Gaussian window, unusual radii, ...): the previous frames are
    // kept instead (packed, two alternating sets) and the pair path runs on (previous, current)
    st.ctx.cur ^= 1;
    uint8_t* cur = st.d_prev + (size_t)st.ctx.cur * n_streams * N;
    const uint8_t* prev = st.d_prev + (size_t)(st.ctx.cur ^ 1) * n_streams * N;
    for (int i = 0; i < n_streams; i++)
      OFB_CUDA(h, cudaMemcpy2DAsync(cur + (size_t)i * N, (size_t)width, d_frames + (size_t)i * image_stride_bytes, pitch_bytes,
                                    width, height, cudaMemcpyDeviceToDevice, h->stream));
    if (!prime) {
      s = farneback_run(h, n_streams, false, prev, cur, width, height, (size_t)width, N, d_flow, nullptr, params);
      if (s) { st.n = 0; return s; }
    }
  }
  if (produced) *produced = prime ? 0 : n_streams;
  return OFB_OK;
}

int ofb_farneback_stream(ofb_handle* h, int n_streams, const uint8_t* const* frames, int width, int height,
                         size_t stride_bytes, float* const* flow, size_t flow_stride_bytes,
                         const ofb_farneback_params* params, int* produced) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (produced) *produced = 0;
  if (!frames) return set_error(h, OFB_ERR_INVALID_ARG, "NULL array pointer");
  int s = validate_farneback(h, n_streams, width, height, params);
  if (s) return s;
  if (stride_bytes == 0) stride_bytes = (size_t)width;
  if (stride_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than width");
  const size_t row_flow = (size_t)width * 2 * sizeof(float);
  if (flow_stride_bytes == 0) flow_stride_bytes = row_flow;
  if (flow_stride_bytes < row_flow) return set_error(h, OFB_ERR_INVALID_ARG, "flow stride smaller than a row");
  for (int i = 0; i < n_streams; i++)
    if (!frames[i]) return set_error(h, OFB_ERR_INVALID_ARG, "NULL frame pointer");
  OFB_CUDA(h, cudaSetDevice(h->device));
  // the stream call shares the batch calls' staging and events: their pipeline restarts after it
  if (h->pipe_n != 0) {
    OFB_CUDA(h, cudaStreamSynchronize(h->s_in));
    OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    OFB_CUDA(h, cudaStreamSynchronize(h->s_out));
    h->pipe_n = 0;
  }
  while ((int)h->pipe_ev.size() < 6) {
    cudaEvent_t e;
    OFB_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->pipe_ev.push_back(e);
  }
  // Uploads go through the copy-in stream into two alternating staging halves, so that with flow == NULL (fields stay
  // on the device, nothing to wait for) the next call's upload runs behind this call's kernels.
  const size_t pitch = align_up((size_t)width, 16), istride = pitch * height;
  h->stream_state.up_parity ^= 1;
  const int par = h->stream_state.up_parity;
  uint8_t* stage = h->d_src + (size_t)par * n_streams * istride;
  cudaEvent_t ev_in = h->pipe_ev[3 * par], ev_comp = h->pipe_ev[3 * par + 1];
  OFB_CUDA(h, cudaStreamWaitEvent(h->s_in, ev_comp, 0));      // the kernels that read this half two calls ago are done
  for (int i = 0; i < n_streams; i++)   // (pageable frames: the runtime stages them; pinned ones are DMA'd directly)
    OFB_CUDA(h, cudaMemcpy2DAsync(stage + (size_t)i * istride, pitch, frames[i], stride_bytes, width, height,
                                  cudaMemcpyHostToDevice, h->s_in));
  OFB_CUDA(h, cudaEventRecord(ev_in, h->s_in));
  OFB_CUDA(h, cudaStreamWaitEvent(h->stream, ev_in, 0));
  int got = 0;
  s = ofb_farneback_stream_device(h, n_streams, stage, width, height, pitch, istride, h->d_flow_out, params, &got);
  if (s) return s;
  OFB_CUDA(h, cudaEventRecord(ev_comp, h->stream));
  if (produced) *produced = got;
  if (!flow) return OFB_OK;                                   // asynchronous: ofb_wait / ofb_flow_u_stats synchronise
  if (got && flow) {
    const size_t fl_img = (size_t)width * height * 2;
    for (int i = 0; i < n_streams; i++) {
      if (!flow[i]) return set_error(h, OFB_ERR_INVALID_ARG, "NULL flow pointer");
      OFB_CUDA(h, cudaMemcpy2DAsync(flow[i], flow_stride_bytes, h->d_flow_out + i * fl_img, row_flow, row_flow, height,
                                    cudaMemcpyDeviceToHost, h->stream));
    }
  }
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (produced) *produced = got;
  return OFB_OK;
}

static int farneback_batch_impl(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next,
                                int width, int height, size_t stride_bytes, float* const* flow,
                                size_t flow_stride_bytes, const ofb_farneback_params* params, bool wait) {
  if (!h) return OFB_ERR_INVALID_ARG;
  const bool download = flow != nullptr;   // nullptr: the field stays on the device (ofb_farneback_batch_stats)
  if (!prev || !next) return set_error(h, OFB_ERR_INVALID_ARG, "NULL array pointer");
  if (!download && (params && (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)))
    return set_error(h, OFB_ERR_INVALID_ARG, "USE_INITIAL_FLOW needs the flow arrays");
  int st = validate_farneback(h, n, width, height, params);
  if (st) return st;
  if (stride_bytes == 0) stride_bytes = (size_t)width;
  if (stride_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than width");
  const size_t row_flow = (size_t)width * 2 * sizeof(float);
  if (flow_stride_bytes == 0) flow_stride_bytes = row_flow;
  if (flow_stride_bytes < row_flow) return set_error(h, OFB_ERR_INVALID_ARG, "flow stride smaller than a row");
  for (int i = 0; i < n; i++)
    if (!prev[i] || !next[i] || (download && !flow[i])) return set_error(h, OFB_ERR_INVALID_ARG, "NULL image/flow pointer");
  OFB_CUDA(h, cudaSetDevice(h->device));
  // Pinned (page-locked / registered) caller buffers are DMA'd directly; pageable ones are staged
  // through the handle's pinned buffers (one memcpy each way, as cudaMemcpy would do internally).
  bool pinned = true;
  for (int i = 0; i < n && pinned; i++)
    pinned = is_pinned(prev[i]) && is_pinned(next[i]) && (!download || is_pinned(flow[i]));
  const size_t pitch = align_up((size_t)width, 16);
  const size_t istride = pitch * height;
  const size_t fl_img = (size_t)width * height * 2;
  const float* init = nullptr;
  if (pinned && n > 1 && !h->no_pipeline) {
    // Pipelined path: chunks of `c` pairs; H2D of chunk i+1 (copy-in stream) and D2H of chunk i-1
    // (copy-out stream) overlap the kernels of chunk i (handle stream).  PCIe is full duplex, so the
    // steady state is bounded by max(compute, D2H of the 8N-byte field).
    // With the asynchronous entry point the pipeline also runs ACROSS calls (the next call's uploads and
    // kernels overlap this call's downloads): every chunk slot keeps three events (upload done, kernels
    // done, download done) that the next call's work on the same staging regions waits for.
    // Without a download to hide (the reduction call) the whole batch runs as ONE chunk — full-size launches — and
    // successive calls alternate between two source staging sets, so the next call's upload runs behind this call's
    // kernels.
    const bool whole = !download && h->pipe_chunk == 0;
    const int c = h->pipe_chunk > 0 ? std::min(h->pipe_chunk, n) : whole ? n : (n >= 16 ? 4 : (n >= 8 ? 2 : 1));
    const int chunks = (n + c - 1) / c;
    while ((int)h->pipe_ev.size() < 3 * std::max(chunks, 2)) {
      cudaEvent_t e;
      OFB_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->pipe_ev.push_back(e);
    }
    if (h->pipe_n != n || h->pipe_c != c || h->pipe_w != width || h->pipe_h != height) {
      // different staging layout than the call before: drain it first
      OFB_CUDA(h, cudaStreamSynchronize(h->s_in));
      OFB_CUDA(h, cudaStreamSynchronize(h->stream));
      OFB_CUDA(h, cudaStreamSynchronize(h->s_out));
      h->pipe_n = n; h->pipe_c = c; h->pipe_w = width; h->pipe_h = height;
    }
    const bool use_init = (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) != 0;
    if (whole) h->pipe_parity ^= 1;
    // every chunk has its own staging and result pointers, i.e. its own graph key: a batch split into more chunks than
    // half the graph cache would evict its own entries on every call (capture + instantiate per chunk) — and such a
    // batch is not launch-bound anyway
    h->graph_bypass = chunks > 8;
    for (int ci = 0; ci < chunks; ci++) {
      const int i0 = ci * c, cn = std::min(c, n - i0);
      const int slot = whole ? h->pipe_parity : ci;             // staging region and its three events
      uint8_t* base = h->d_src + (whole ? (size_t)slot * 2 * n * istride : (size_t)2 * i0 * istride);   // [prev x cn][next x cn]
      cudaEvent_t ev_in = h->pipe_ev[3 * slot], ev_comp = h->pipe_ev[3 * slot + 1], ev_out = h->pipe_ev[3 * slot + 2];
      OFB_CUDA(h, cudaStreamWaitEvent(h->s_in, ev_comp, 0));     // previous call's kernels have read this source slot
      for (int i = 0; i < cn; i++) {
        OFB_CUDA(h, cudaMemcpy2DAsync(base + (size_t)i * istride, pitch, prev[i0 + i], stride_bytes, width, height,
                                      cudaMemcpyHostToDevice, h->s_in));
        OFB_CUDA(h, cudaMemcpy2DAsync(base + (size_t)(cn + i) * istride, pitch, next[i0 + i], stride_bytes, width,
                                      height, cudaMemcpyHostToDevice, h->s_in));
        if (use_init)
          OFB_CUDA(h, cudaMemcpy2DAsync(h->d_init_flow + (i0 + i) * fl_img, row_flow, flow[i0 + i], flow_stride_bytes,
                                        row_flow, height, cudaMemcpyHostToDevice, h->s_in));
      }
      OFB_CUDA(h, cudaEventRecord(ev_in, h->s_in));
      OFB_CUDA(h, cudaStreamWaitEvent(h->stream, ev_in, 0));
      OFB_CUDA(h, cudaStreamWaitEvent(h->stream, ev_out, 0));    // previous call's download of this result slot is done
      st = farneback_run(h, cn, false, base, base + (size_t)cn * istride, width, height, pitch, istride,
                         h->d_flow_out + i0 * fl_img, use_init ? h->d_init_flow + i0 * fl_img : nullptr, params);
      if (st) { h->graph_bypass = false; return st; }
      OFB_CUDA(h, cudaEventRecord(ev_comp, h->stream));
      OFB_CUDA(h, cudaStreamWaitEvent(h->s_out, ev_comp, 0));
      for (int i = 0; i < cn && download; i++)
        OFB_CUDA(h, cudaMemcpy2DAsync(flow[i0 + i], flow_stride_bytes, h->d_flow_out + (i0 + i) * fl_img, row_flow,
                                      row_flow, height, cudaMemcpyDeviceToHost, h->s_out));
      OFB_CUDA(h, cudaEventRecord(ev_out, h->s_out));
    }
    h->graph_bypass = false;
    h->last_flow = h->d_flow_out;
    h->last_n = n;
    if (wait) {
      OFB_CUDA(h, cudaStreamSynchronize(h->s_out));
      OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return OFB_OK;
  }
  // serial paths (single pair, pageable buffers, OFB_NO_PIPELINE): always synchronous
  OFB_CUDA(h, cudaStreamSynchronize(h->s_in));
  OFB_CUDA(h, cudaStreamSynchronize(h->s_out));
  h->pipe_n = 0;
  if (pinned) {
    for (int i = 0; i < n; i++) {
      OFB_CUDA(h, cudaMemcpy2DAsync(h->d_src + (size_t)i * istride, pitch, prev[i], stride_bytes, width, height,
                                    cudaMemcpyHostToDevice, h->stream));
      OFB_CUDA(h, cudaMemcpy2DAsync(h->d_src + (size_t)(n + i) * istride, pitch, next[i], stride_bytes, width, height,
                                    cudaMemcpyHostToDevice, h->stream));
    }
    if (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) {
      for (int i = 0; i < n; i++)
        OFB_CUDA(h, cudaMemcpy2DAsync(h->d_init_flow + i * fl_img, row_flow, flow[i], flow_stride_bytes, row_flow,
                                      height, cudaMemcpyHostToDevice, h->stream));
      init = h->d_init_flow;
    }
  } else {
    for (int i = 0; i < n; i++)
      for (int img = 0; img < 2; img++) {
        const uint8_t* s = img ? next[i] : prev[i];
        uint8_t* d = h->h_src + (size_t)(img * n + i) * istride;
        if (stride_bytes == pitch) memcpy(d, s, istride);
        else for (int y = 0; y < height; y++) memcpy(d + (size_t)y * pitch, s + (size_t)y * stride_bytes, width);
      }
    OFB_CUDA(h, cudaMemcpyAsync(h->d_src, h->h_src, istride * 2 * n, cudaMemcpyHostToDevice, h->stream));
    if (params->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) {
      for (int i = 0; i < n; i++)
        for (int y = 0; y < height; y++)
          memcpy(h->h_flow + i * fl_img + (size_t)y * width * 2, (const char*)flow[i] + (size_t)y * flow_stride_bytes,
                 row_flow);
      OFB_CUDA(h, cudaMemcpyAsync(h->d_init_flow, h->h_flow, n * fl_img * sizeof(float), cudaMemcpyHostToDevice,
                                  h->stream));
      init = h->d_init_flow;
    }
  }
  st = farneback_run(h, n, false, h->d_src, h->d_src + (size_t)n * istride, width, height, pitch, istride,
                     h->d_flow_out, init, params);
  if (st) return st;
  if (!download) {
    OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return OFB_OK;
  }
  if (pinned) {
    for (int i = 0; i < n; i++)
      OFB_CUDA(h, cudaMemcpy2DAsync(flow[i], flow_stride_bytes, h->d_flow_out + i * fl_img, row_flow, row_flow, height,
                                    cudaMemcpyDeviceToHost, h->stream));
    OFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return OFB_OK;
  }
  OFB_CUDA(h, cudaMemcpyAsync(h->h_flow, h->d_flow_out, n * fl_img * sizeof(float), cudaMemcpyDeviceToHost,
                              h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; i++) {
    if (flow_stride_bytes == row_flow) memcpy(flow[i], h->h_flow + i * fl_img, fl_img * sizeof(float));
    else
      for (int y = 0; y < height; y++)
        memcpy((char*)flow[i] + (size_t)y * flow_stride_bytes, h->h_flow + i * fl_img + (size_t)y * width * 2, row_flow);
  }
  return OFB_OK;
}

int ofb_farneback_batch(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next, int width,
                        int height, size_t stride_bytes, float* const* flow, size_t flow_stride_bytes,
                        const ofb_farneback_params* params) {
  return farneback_batch_impl(h, n, prev, next, width, height, stride_bytes, flow, flow_stride_bytes, params, true);
}

int ofb_farneback_batch_async(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next, int width,
                              int height, size_t stride_bytes, float* const* flow, size_t flow_stride_bytes,
                              const ofb_farneback_params* params) {
  return farneback_batch_impl(h, n, prev, next, width, height, stride_bytes, flow, flow_stride_bytes, params, false);
}

int ofb_farneback_batch_stats(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next, int width,
                              int height, size_t stride_bytes, const ofb_farneback_params* params, const uint8_t* mask,
                              double* out_mean, float* out_median) {
  int st = farneback_batch_impl(h, n, prev, next, width, height, stride_bytes, nullptr, 0, params, false);
  if (st) return st;
  return flow_u_stats(h, n, mask, out_mean, out_median);   // synchronises the stream and copies n scalars back
}

int ofb_farneback_batch_stats_async(ofb_handle* h, int n, const uint8_t* const* prev, const uint8_t* const* next,
                                    int width, int height, size_t stride_bytes, const ofb_farneback_params* params,
                                    const uint8_t* mask, double* out_mean, float* out_median) {
  int st = farneback_batch_impl(h, n, prev, next, width, height, stride_bytes, nullptr, 0, params, false);
  if (st) return st;
  if (h->pipe_n == 0) return flow_u_stats(h, n, mask, out_mean, out_median);   // pageable frames: served synchronously
  return flow_u_stats(h, n, mask, out_mean, out_median, true);
}

int ofb_wait(ofb_handle* h) {
  if (!h) return OFB_ERR_INVALID_ARG;
  OFB_CUDA(h, cudaSetDevice(h->device));
  OFB_CUDA(h, cudaStreamSynchronize(h->s_in));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->s_out));
  return finish_pending_stats(h);
}

int ofb_farneback(ofb_handle* h, const uint8_t* prev, const uint8_t* next, int width, int height, size_t stride_bytes,
                  float* flow, size_t flow_stride_bytes, const ofb_farneback_params* params) {
  const uint8_t* pp[1] = {prev};
  const uint8_t* nn[1] = {next};
  float* ff[1] = {flow};
  return ofb_farneback_batch(h, 1, pp, nn, width, height, stride_bytes, ff, flow_stride_bytes, params);
}

int ofb_flow_postfilter(ofb_handle* h, int n, int median_ksize, float magnitude_threshold, const uint8_t* const* gray,
                        size_t gray_stride_bytes, int intensity_threshold) {
  if (!h) return OFB_ERR_INVALID_ARG;
  return flow_postfilter(h, n, median_ksize, magnitude_threshold, gray, gray_stride_bytes, intensity_threshold);
}

int ofb_flow_sample(ofb_handle* h, int pair, int n_points, const int* xy, float* out_dxdy) {
  if (!h) return OFB_ERR_INVALID_ARG;
  return flow_sample(h, pair, n_points, xy, out_dxdy);
}

int ofb_flow_to_bgr(ofb_handle* h, int pair, uint8_t* bgr_out, size_t stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  return flow_to_bgr(h, pair, bgr_out, stride_bytes);
}

int ofb_flow_to_bgr_speed(ofb_handle* h, int pair, double dt, double pixel_to_meter, double max_speed, uint8_t* bgr_out,
                          size_t stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!(dt > 0) || !(max_speed > 0)) return set_error(h, OFB_ERR_INVALID_ARG, "dt and max_speed must be positive");
  return flow_to_bgr(h, pair, bgr_out, stride_bytes, 1, (float)dt, (float)pixel_to_meter, (float)max_speed);
}

int ofb_flow_download(ofb_handle* h, int n, float* const* flow, size_t flow_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  return flow_download(h, n, flow, flow_stride_bytes);
}

int ofb_flow_u_stats_async(ofb_handle* h, int n, const uint8_t* mask, double* out_mean, float* out_median) {
  if (!h) return OFB_ERR_INVALID_ARG;
  return flow_u_stats(h, n, mask, out_mean, out_median, true);
}

int ofb_flow_u_stats(ofb_handle* h, int n, const uint8_t* mask, double* out_mean, float* out_median) {
  if (!h) return OFB_ERR_INVALID_ARG;
  return flow_u_stats(h, n, mask, out_mean, out_median);
}

}  // extern "C"
