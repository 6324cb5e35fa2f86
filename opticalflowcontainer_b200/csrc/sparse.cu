// sparse.cu — sparse path for B200 (sm_100a): uint8 pyrDown pyramid, Scharr derivatives,
// Shi-Tomasi corners (cornerMinEigenVal + goodFeaturesToTrack selection) and the pyramidal
// Lucas-Kanade tracker.  Replaces cv2.goodFeaturesToTrack + cv2.calcOpticalFlowPyrLK behind the
// node flow call (ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:194).  Arithmetic follows
// OpenCV 4.x modules/imgproc/src/{pyramids,corner,featureselect,deriv}.cpp and
// modules/video/src/lkpyramid.cpp (un-vendored dependency, ros2_ws/src/nueflow/setup.py:29).
//
// Integer stages (pyrDown, Scharr, the LK fixed-point patches) are bit-exact; the eigenvalue map
// reproduces the wheel's optimized (FMA) float recipe bit for bit so that the corner ranking, and
// therefore the corner list, is identical.
//
// Everything between the upload of a frame and the download of its results runs on the handle's stream
// without a host round trip: candidate and corner counts stay on the device (kernels read them there,
// grids are sized by capacity), and the candidates are ordered by a bucketed partial sort of our own
// (k_candidates / k_bucket_scan / k_bucket_scatter + an in-CTA bitonic sort per selection round) — the
// greedy selection stops at maxCorners, so only the strongest few thousand candidates are ever sorted.
// ofb_lk_stream keeps a camera's temporal state (previous frame's pyramid, Scharr derivatives and corner
// list) on the GPU: one upload per frame; behind the upload the tracker and the new frame's corner detection run as two
// chains on two streams, captured once per frame slot into a CUDA graph.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "fb_device.cuh"

namespace ofb {

constexpr int kMaxLkLevels = 8;
constexpr int kGridSlots = 4;

constexpr int kLogBuckets = 16, kBuckets = 1 << kLogBuckets;   // value buckets of the candidate ordering

// One frame of a camera: level 0 + the pyrDown chain + the Scharr derivatives of every level.
struct FramePyr {
  int n_levels = 0;
  const uint8_t* lv[kMaxLkLevels] = {nullptr};
  const short2* D[kMaxLkLevels] = {nullptr};
  int w[kMaxLkLevels] = {0}, h[kMaxLkLevels] = {0};
};

struct SparseState {
  int cap_w = 0, cap_h = 0;
  uint8_t* img[2] = {nullptr, nullptr};        // level-0 images of the two frame slots, packed pitch = width
  uint8_t* pyr[2] = {nullptr, nullptr};        // levels 1.. of both pyramids, packed one after another
  short2* deriv[2] = {nullptr, nullptr};       // Scharr (dx,dy) of every level of a slot
  float* eig = nullptr;
  unsigned long long* keys = nullptr;          // 2 x cand_cap: candidates as found | bucket order
  unsigned long long* sortbuf = nullptr;       // power-of-two scratch of the oversized-bucket path
  size_t cand_cap = 0, sort_cap = 0;
  unsigned int* hist = nullptr;                // kBuckets: candidates per value bucket, then the scatter cursors
  unsigned int* bstart = nullptr;              // kBuckets + 1: first position of every bucket, strongest bucket first
  unsigned int* bpart = nullptr;               // kBuckets / 1024 partial sums of the bucket scan
  unsigned int* counters = nullptr;            // [0] candidate count, [1] max(eig) bits, [4 + slot] corner count of a slot
  unsigned int* grid_cnt = nullptr;            // per cell
  ushort2* grid_pts = nullptr;                 // per cell x kGridSlots
  float2* corners[2] = {nullptr, nullptr};     // accepted corners of a slot (device)
  uint8_t* mask = nullptr;                     // goodFeaturesToTrack mask (device), allocated on first use
  float2* pts_prev = nullptr;                  // LK inputs / outputs (device)
  // LK results of a call over n points, one block [next pts n x 8 | err n x 4 | status n] so that one copy brings
  // them back (lk_out_layout sets the three pointers for the call's n)
  char* lk_out = nullptr;
  float2* pts_next = nullptr;
  uint8_t* lk_status = nullptr;
  float* lk_err = nullptr;
  int pts_cap = 0;
  // pinned host staging
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  // second stream of the camera-stream call: the corner detection of the new frame runs beside the tracker
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_up = nullptr, ev_aux = nullptr;
  // camera-stream state (ofb_lk_stream)
  struct Stream {
    bool primed = false;
    int cur = 0, w = 0, h = 0;
    ofb_gftt_params gp = {};
    ofb_lk_params lp = {};
    FramePyr fp[2];
    // the launch sequence of a tracking call with the new frame in slot 0 / 1, captured once (CUDA graph): ~30 launches,
    // copies and event calls per frame become one launch — the driver's launch path is what 8 camera threads contend for
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    uint64_t graph_launches[2] = {0, 0};
    const void* graph_out = nullptr;           // lk_out / bound the graphs were captured with
    int graph_bound = 0;
    int n_host[2] = {0, 0};                    // corner count of a slot as last downloaded
    std::vector<float> host_corners[2];        // ... and the list itself (what the next call reports as prev_pts)
  } st;
};

static void stream_graphs_drop(SparseState* s) {
  for (int i = 0; i < 2; i++) {
    if (s->st.graph[i]) cudaGraphExecDestroy(s->st.graph[i]);
    s->st.graph[i] = nullptr;
  }
}

static void sparse_free(SparseState* s) {
  if (!s) return;
  stream_graphs_drop(s);
  for (int i = 0; i < 2; i++) { cudaFree(s->img[i]); cudaFree(s->pyr[i]); cudaFree(s->deriv[i]); cudaFree(s->corners[i]); }
  cudaFree(s->eig); cudaFree(s->keys); cudaFree(s->sortbuf); cudaFree(s->hist); cudaFree(s->bstart);
  cudaFree(s->counters); cudaFree(s->bpart); cudaFree(s->grid_cnt); cudaFree(s->grid_pts); cudaFree(s->mask);
  cudaFree(s->pts_prev); cudaFree(s->lk_out);
  if (s->h_stage) cudaFreeHost(s->h_stage);
  if (s->aux) cudaStreamDestroy(s->aux);
  if (s->ev_up) cudaEventDestroy(s->ev_up);
  if (s->ev_aux) cudaEventDestroy(s->ev_aux);
  delete s;
}

void sparse_destroy(ofb_handle* h) {
  sparse_free(reinterpret_cast<SparseState*>(h->sparse));
  h->sparse = nullptr;
}

#define SP_CUDA(h, call)                                                                                  \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      return set_error((h), e__ == cudaErrorMemoryAllocation ? OFB_ERR_ALLOC : OFB_ERR_CUDA, "%s failed: %s", \
                       #call, cudaGetErrorString(e__));                                                   \
  } while (0)

static int sparse_get(ofb_handle* h, SparseState** out) {
  if (h->sparse) { *out = reinterpret_cast<SparseState*>(h->sparse); return OFB_OK; }
  SparseState* s = new SparseState();
  h->sparse = s;
  s->cap_w = h->max_w;
  s->cap_h = h->max_h;
  const size_t N = (size_t)h->max_w * h->max_h;
  for (int i = 0; i < 2; i++) {
    SP_CUDA(h, cudaMalloc(&s->img[i], N));
    SP_CUDA(h, cudaMalloc(&s->pyr[i], N));     // sum of levels >= 1 is < N/2 (+ rounding)
  }
  for (int i = 0; i < 2; i++) SP_CUDA(h, cudaMalloc(&s->deriv[i], 2 * N * sizeof(short2)));
  SP_CUDA(h, cudaMalloc(&s->eig, N * sizeof(float)));
  s->cand_cap = N / 2 + 1024;
  SP_CUDA(h, cudaMalloc(&s->keys, 2 * s->cand_cap * sizeof(unsigned long long)));
  s->sort_cap = 1024;
  while (s->sort_cap < s->cand_cap) s->sort_cap *= 2;
  SP_CUDA(h, cudaMalloc(&s->sortbuf, s->sort_cap * sizeof(unsigned long long)));
  SP_CUDA(h, cudaMalloc(&s->hist, kBuckets * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->bstart, (kBuckets + 1) * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->counters, 16 * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->bpart, (kBuckets / 1024) * sizeof(unsigned int)));
  SP_CUDA(h, cudaMemset(s->counters, 0, 16 * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->grid_cnt, N * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->grid_pts, N * kGridSlots * sizeof(ushort2)));
  // the selection can keep every candidate (minDistance < 1, maxCorners <= 0): corner buffers sized like the candidate list
  for (int i = 0; i < 2; i++) SP_CUDA(h, cudaMalloc(&s->corners[i], s->cand_cap * sizeof(float2)));
  s->h_stage_bytes = std::max<size_t>(2 * N, s->cand_cap * sizeof(float2) + 64);
  SP_CUDA(h, cudaHostAlloc(&s->h_stage, s->h_stage_bytes, cudaHostAllocDefault));
  SP_CUDA(h, cudaStreamCreateWithFlags(&s->aux, cudaStreamNonBlocking));
  SP_CUDA(h, cudaEventCreateWithFlags(&s->ev_up, cudaEventDisableTiming));
  SP_CUDA(h, cudaEventCreateWithFlags(&s->ev_aux, cudaEventDisableTiming));
  *out = s;
  return OFB_OK;
}

static int sparse_points(ofb_handle* h, SparseState* s, int n) {
  if (n <= s->pts_cap) return OFB_OK;
  cudaFree(s->pts_prev); cudaFree(s->lk_out);
  s->pts_prev = s->pts_next = nullptr; s->lk_out = nullptr; s->lk_status = nullptr; s->lk_err = nullptr;
  s->pts_cap = 0;
  const int cap = std::max(n, 4096);
  SP_CUDA(h, cudaMalloc(&s->pts_prev, cap * sizeof(float2)));
  SP_CUDA(h, cudaMalloc(&s->lk_out, (size_t)cap * 13));
  s->pts_cap = cap;
  return OFB_OK;
}

static void lk_out_layout(SparseState* s, int n) {
  s->pts_next = reinterpret_cast<float2*>(s->lk_out);
  s->lk_err = reinterpret_cast<float*>(s->lk_out + (size_t)n * 8);
  s->lk_status = reinterpret_cast<uint8_t*>(s->lk_out + (size_t)n * 12);
}

// ======================================================================================
// a10: pyrDown (uint8): [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8, size ((w+1)/2,(h+1)/2)
// ======================================================================================
__global__ void __launch_bounds__(256) k_pyrdown_u8(const uint8_t* __restrict__ src, int w, int h, size_t spitch,
                                                    uint8_t* __restrict__ dst, int ow, int oh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  const int kk[5] = {1, 4, 6, 4, 1};
  int xi[5];
#pragma unroll
  for (int i = 0; i < 5; i++) xi[i] = reflect101(2 * x + i - 2, w);
  int acc = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const uint8_t* row = src + (size_t)reflect101(2 * y + j - 2, h) * spitch;
    int r = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) r += kk[i] * (int)__ldg(row + xi[i]);
    acc += kk[j] * r;
  }
  dst[(size_t)y * ow + x] = (uint8_t)((acc + 128) >> 8);
}

// ======================================================================================
// a11: calcScharrDeriv: int16 (dx, dy) interleaved, BORDER_REFLECT_101
// ======================================================================================
__global__ void __launch_bounds__(256) k_scharr(const uint8_t* __restrict__ src, int w, int h, size_t spitch,
                                                short2* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * spitch;
  const uint8_t* r1 = src + (size_t)y * spitch;
  const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * spitch;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
  const int a0 = r0[xm], a1 = r0[x], a2 = r0[xp];
  const int b0 = r1[xm], b2 = r1[xp];
  const int c0 = r2[xm], c1 = r2[x], c2 = r2[xp];
  // t0 = (above + below)*3 + centre*10 ; t1 = below - above
  const int t0m = (a0 + c0) * 3 + b0 * 10, t0p = (a2 + c2) * 3 + b2 * 10;
  const int t1m = c0 - a0, t1c = c1 - a1, t1p = c2 - a2;
  dst[(size_t)y * w + x] = make_short2((short)(t0p - t0m), (short)((t1p + t1m) * 3 + t1c * 10));
}

// ======================================================================================
// a13: cornerMinEigenVal(blockSize, ksize=3) — the wheel's optimized float recipe, bit for bit
// ---- four pixels per thread: rows read as aligned 32-bit words, results stored as one vector ---------------------------
// (the byte-per-thread forms above issue 9 / 25 byte loads per output pixel; these need w % 4 == 0, a pitch that is a
// multiple of 4 and aligned bases — every level of a 1080p / VGA pyramid — and are the same integer arithmetic.)
__device__ __forceinline__ unsigned int ldg_u32(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ int byte_of(unsigned int w, int i) { return (int)((w >> (8 * i)) & 0xffu); }

__global__ void __launch_bounds__(256) k_scharr4(const uint8_t* __restrict__ src, int w, int h, size_t spitch,
                                                 short2* __restrict__ dst) {
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= w || y >= h) return;
  const uint8_t* rows[3] = {src + (size_t)reflect101(y - 1, h) * spitch, src + (size_t)y * spitch,
                            src + (size_t)reflect101(y + 1, h) * spitch};
  const int xl = x0 == 0 ? 1 : x0 - 1, xr = x0 + 4 >= w ? w - 2 : x0 + 4;     // REFLECT_101 of x0 - 1 and x0 + 4
  int p[3][6];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const unsigned int wd = ldg_u32(rows[j] + x0);
    p[j][0] = rows[j][xl];
#pragma unroll
    for (int i = 0; i < 4; i++) p[j][1 + i] = byte_of(wd, i);
    p[j][5] = rows[j][xr];
  }
  // t0 = (above + below)*3 + centre*10 ; t1 = below - above   (per column), then the horizontal difference / smoothing
  int t0[6], t1[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    t0[i] = (p[0][i] + p[2][i]) * 3 + p[1][i] * 10;
    t1[i] = p[2][i] - p[0][i];
  }
  uint4 out;
  unsigned int* o = &out.x;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int dx = t0[i + 2] - t0[i], dy = (t1[i + 2] + t1[i]) * 3 + t1[i + 1] * 10;
    o[i] = ((unsigned int)dx & 0xffffu) | ((unsigned int)dy << 16);
  }
  *reinterpret_cast<uint4*>(dst + (size_t)y * w + x0) = out;
}

__global__ void __launch_bounds__(256) k_pyrdown_u8x4(const uint8_t* __restrict__ src, int w, int h, size_t spitch,
                                                      uint8_t* __restrict__ dst, int ow, int oh) {
  // needs w == 2 * ow, w % 4 == 0, ow % 4 == 0
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);       // first of the thread's four output columns
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= ow || y >= oh) return;
  // source columns 2 x0 - 2 .. 2 x0 + 8.  2 x0 .. 2 x0 + 7 are two aligned words inside the row; the two columns to the
  // left come from the word before — or, in the first group, are REFLECT_101 of -2 and -1: columns 2 and 1 of the first
  // word; the column to the right is byte 0 of the word behind — or, in the last group, column w reflected to w - 2.
  const bool first = x0 == 0, last = 2 * x0 + 8 >= w;
  int acc[4] = {0, 0, 0, 0};
  const int kk[5] = {1, 4, 6, 4, 1};
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const uint8_t* row = src + (size_t)reflect101_once(2 * y + j - 2, h) * spitch + 2 * x0;
    const unsigned int w1 = ldg_u32(row), w2 = ldg_u32(row + 4);
    const unsigned int w0 = ldg_u32(first ? row : row - 4), w3 = ldg_u32(last ? row : row + 8);
    int v[11];
    v[0] = first ? byte_of(w1, 2) : byte_of(w0, 2);
    v[1] = first ? byte_of(w1, 1) : byte_of(w0, 3);
#pragma unroll
    for (int i = 0; i < 4; i++) { v[2 + i] = byte_of(w1, i); v[6 + i] = byte_of(w2, i); }
    v[10] = last ? byte_of(w2, 2) : byte_of(w3, 0);
#pragma unroll
    for (int o = 0; o < 4; o++)
      acc[o] += kk[j] * (v[2 * o] + 4 * v[2 * o + 1] + 6 * v[2 * o + 2] + 4 * v[2 * o + 3] + v[2 * o + 4]);
  }
  unsigned int out = 0;
#pragma unroll
  for (int o = 0; o < 4; o++) out |= (unsigned int)((acc[o] + 128) >> 8) << (8 * o);
  *reinterpret_cast<unsigned int*>(dst + (size_t)y * ow + x0) = out;
}

// ======================================================================================
// Sobel with the scale folded into the smoothing kernel k = f32([1,2,1]/(4*blockSize*255)):
//   dx = fma(r[y-1] + r[y+1], k0, r[y]*k1),  r = p[x+1] - p[x-1]
//   dy = rw[y+1] - rw[y-1],  rw = fma(k2, p[x+1], fma(k1, p[x], k0*p[x-1]))   (vector body)
//        rw = (p[x-1]*k0 + p[x]*k1) + p[x+1]*k2   for columns past the last full block of 32 (SIMD tail)
// Both stages in one kernel: a CTA computes the three derivative products for its 32 x 16 output tile plus the
// box-filter apron into shared memory (the apron positions are the REFLECT_101 images of interior pixels, so each is the
// product AT the reflected position, as a filter over a stored covariance image would read it), then sums the windows.
// Nothing but the eigenvalue map is written: the 12 B/px covariance image of the two-kernel form (written once, read
// blockSize^2 times through L1) is gone.
//
// unnormalised blockSize^2 box sum in double (exact for these magnitudes), then
// eig = (a + c) - sqrt((a - c)^2 + b^2), a = Sxx/2, c = Syy/2 — plain mul/add, no FMA.
// harris != 0: cv2.cornerHarris instead, as the wheel computes it over the image as ONE continuous row of w * h pixels —
// (a c - b b) - k ((a + c)(a + c)) in float in the 8-wide body, (a c - b b) - (k (a + c)) (a + c) in the 4-wide step
// behind it, and the last (w * h) % 4 pixels in double with the caller's double k (oracle/features_np.py::corner_harris).
// Maximum of a 256-thread CTA into a global word with at most ONE atomic per CTA, and none when the word already holds
// a larger value (it only grows, so a stale read can only cost a redundant atomic): one atomicMax per warp was 32 000
// same-address atomics per 1080p frame, which L2 serialises at ~1 ns each — the whole duration of the kernel.
__device__ __forceinline__ void block_max_to_global(float m, int tid, unsigned int* max_bits) {
  __shared__ float wmax[8];
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) wmax[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int k = 1; k < 8; k++) m = fmaxf(m, wmax[k]);
    if (m > 0.f && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(max_bits)) atomicMax(max_bits, __float_as_uint(m));
  }
}

constexpr int kEigTileH = 16;
// derivative products at image position (x, y) [already inside the image]; BORDER: the 3x3 neighbourhood may leave it
template <bool BORDER>
__device__ __forceinline__ void sobel_products(const uint8_t* __restrict__ src, int w, int h, size_t spitch, int x, int y,
                                               float k0, float k1, float k2, int wb, double* pxx, double* pxy, double* pyy) {
  int xm = x - 1, xp = x + 1, ym = y - 1, yp = y + 1;
  if (BORDER) { xm = reflect101(xm, w); xp = reflect101(xp, w); ym = reflect101(ym, h); yp = reflect101(yp, h); }
  const uint8_t* rows[3] = {src + (size_t)ym * spitch, src + (size_t)y * spitch, src + (size_t)yp * spitch};
  float r[3], rw[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const float pm = u8f(rows[j][xm]), pc = u8f(rows[j][x]), pp = u8f(rows[j][xp]);   // (exact, off the XU pipe)
    r[j] = pp - pm;
    if (x < wb) rw[j] = fmaf(k2, pp, fmaf(k1, pc, __fmul_rn(k0, pm)));
    else rw[j] = __fadd_rn(__fadd_rn(__fmul_rn(pm, k0), __fmul_rn(pc, k1)), __fmul_rn(pp, k2));
  }
  const float dx = fmaf(r[0] + r[2], k0, __fmul_rn(r[1], k1));
  const float dy = rw[2] - rw[0];
  *pxx = (double)__fmul_rn(dx, dx);
  *pxy = (double)__fmul_rn(dx, dy);
  *pyy = (double)__fmul_rn(dy, dy);
}

__global__ void __launch_bounds__(256) k_sobel_min_eig(const uint8_t* __restrict__ src, int w, int h, size_t spitch, float k0,
                                                       float k1, float k2, int wb, int block_size,
                                                       float* __restrict__ eig, unsigned int* __restrict__ max_bits,
                                                       const uint8_t* __restrict__ mask, int harris, float kf, double kd) {
  // the products as doubles: one conversion per product instead of one per window tap (ncu on the float tile: the XU
  // pipe — I2F of the pixels and F2F of the taps — was the busiest unit at 57 %)
  extern __shared__ double cov_tile[];               // [3][th][tw]
  const int rb = block_size / 2, tw = 32 + 2 * rb, th = kEigTileH + 2 * rb, ta = tw * th;
  double* cxx = cov_tile;
  double* cxy = cov_tile + ta;
  double* cyy = cov_tile + 2 * ta;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int bx = blockIdx.x * 32, by = blockIdx.y * kEigTileH;
  // tile + apron + the Sobel neighbourhood inside the image: no border arithmetic in the whole CTA
  const bool interior = bx - rb - 1 >= 0 && by - rb - 1 >= 0 && bx + 32 + rb + 1 <= w && by + kEigTileH + rb + 1 <= h;
  if (interior) {
    for (int i = tid; i < ta; i += 256) {
      const int ty = i / tw, tx = i - ty * tw;
      sobel_products<false>(src, w, h, spitch, bx - rb + tx, by - rb + ty, k0, k1, k2, wb, cxx + i, cxy + i, cyy + i);
    }
  } else {
    for (int i = tid; i < ta; i += 256) {
      const int ty = i / tw, tx = i - ty * tw;
      // (tile positions beyond the apron of the image's last row / column are never read: keep them in range)
      const int x = reflect101(min(bx - rb + tx, w - 1 + rb), w), y = reflect101(min(by - rb + ty, h - 1 + rb), h);
      sobel_products<true>(src, w, h, spitch, x, y, k0, k1, k2, wb, cxx + i, cxy + i, cyy + i);
    }
  }
  __syncthreads();
  const size_t n = (size_t)w * h;
  float m = 0.f;
#pragma unroll
  for (int half = 0; half < kEigTileH / 8; half++) {
    const int lx = threadIdx.x, ly = threadIdx.y + 8 * half;
    const int x = bx + lx, y = by + ly;
    if (x >= w || y >= h) continue;
    double sxx = 0, sxy = 0, syy = 0;
    for (int j = 0; j <= 2 * rb; j++) {
      const int ro = (ly + j) * tw + lx;
      for (int i = 0; i <= 2 * rb; i++) {
        sxx += cxx[ro + i];
        sxy += cxy[ro + i];
        syy += cyy[ro + i];
      }
    }
    float e;
    const size_t o = (size_t)y * w + x;
    if (!harris) {
      const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
      const float d = __fsub_rn(a, c);
      e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
    } else {
      const float a = (float)sxx, b = (float)sxy, c = (float)syy;
      const float t1 = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, b)), sm = __fadd_rn(a, c);
      if (o < n - (n & 7)) e = __fsub_rn(t1, __fmul_rn(kf, __fmul_rn(sm, sm)));
      else if (o < n - (n & 3)) e = __fsub_rn(t1, __fmul_rn(__fmul_rn(kf, sm), sm));
      else e = (float)__dsub_rn((double)t1, __dmul_rn(__dmul_rn(kd, (double)sm), (double)sm));
    }
    eig[o] = e;
    // max over the image — over the masked pixels with a mask — (minMaxLoc); the map is >= 0 up to rounding, a negative
    // max means "no corners"
    if (mask && mask[o] == 0) e = 0.f;
    m = fmaxf(m, e);
  }
  block_max_to_global(m, tid, max_bits);
}

// blockSize 3 (the default): the same map with the tile pass in groups of four columns — aligned word loads, the border
// decision per CTA — storing the HORIZONTAL 3-sums of the products (doubles, exact), so that the window pass of a thread
// is a vertical 3-sum over one column, shared by its four vertically adjacent outputs.  ncu on the generic kernel at
// blockSize 3: 388 instructions per pixel, issue slots 84 % busy; a first form of this kernel that stored the products
// themselves: 225 instructions per pixel, the shared-memory pipe 72 % busy (36 LDS.64 per two outputs, strided stores
// with 1.5 M bank conflicts).  Here: 4.5 LDS.64 per output, vector stores.
constexpr int kE3T = 32;                            // outputs per CTA: 32 x 32
constexpr int kE3H = kE3T + 2;                      // rows of row sums: by - 1 .. by + 32
__device__ __forceinline__ void sts_d2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

__global__ void __launch_bounds__(256) k_sobel_min_eig3(const uint8_t* __restrict__ src, int w, int h, size_t spitch, float k0,
                                                        float k1, float k2, int wb, float* __restrict__ eig,
                                                        unsigned int* __restrict__ max_bits, const uint8_t* __restrict__ mask,
                                                        int harris, float kf, double kd) {
  __shared__ __align__(16) double hxx[kE3H][kE3T], hxy[kE3H][kE3T], hyy[kE3H][kE3T];   // 3-sums centred on column bx + c
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int bx = blockIdx.x * kE3T, by = blockIdx.y * kE3T;
  // products are needed at columns bx - 1 .. bx + 32, rows by - 1 .. by + 32; their Sobel neighbourhood one further
  const bool interior = bx - 4 >= 0 && by - 2 >= 0 && bx + kE3T + 4 <= w && by + kE3T + 2 <= h && (spitch & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  if (interior) {
    for (int t = tid; t < (kE3T / 4) * kE3H; t += 256) {
      const int g = t % (kE3T / 4), r = t / (kE3T / 4);
      const int x0 = bx + 4 * g, y = by - 1 + r;          // row sums at columns x0 .. x0 + 3: products at x0 - 1 .. x0 + 4
      float p[3][8];                                      // pixels x0 - 2 .. x0 + 5 of rows y - 1 .. y + 1
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const uint8_t* row = src + (size_t)(y - 1 + j) * spitch + x0;
        const unsigned int wl = ldg_u32(row - 4), wc = ldg_u32(row), wr = ldg_u32(row + 4);
        p[j][0] = u8f((wl >> 16) & 0xffu); p[j][1] = u8f(wl >> 24);
#pragma unroll
        for (int i = 0; i < 4; i++) p[j][2 + i] = u8f((wc >> (8 * i)) & 0xffu);
        p[j][6] = u8f(wr & 0xffu); p[j][7] = u8f((wr >> 8) & 0xffu);
      }
      double pxx[6], pxy[6], pyy[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {                       // product at column x0 - 1 + i
        float rr[3], rw[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float pm = p[j][i], pc = p[j][i + 1], pp = p[j][i + 2];
          rr[j] = pp - pm;
          if (x0 - 1 + i < wb) rw[j] = fmaf(k2, pp, fmaf(k1, pc, __fmul_rn(k0, pm)));
          else rw[j] = __fadd_rn(__fadd_rn(__fmul_rn(pm, k0), __fmul_rn(pc, k1)), __fmul_rn(pp, k2));
        }
        const float dx = fmaf(rr[0] + rr[2], k0, __fmul_rn(rr[1], k1));
        const float dy = rw[2] - rw[0];
        pxx[i] = (double)__fmul_rn(dx, dx);
        pxy[i] = (double)__fmul_rn(dx, dy);
        pyy[i] = (double)__fmul_rn(dy, dy);
      }
      sts_d2(&hxx[r][4 * g], (pxx[0] + pxx[1]) + pxx[2], (pxx[1] + pxx[2]) + pxx[3]);
      sts_d2(&hxx[r][4 * g + 2], (pxx[2] + pxx[3]) + pxx[4], (pxx[3] + pxx[4]) + pxx[5]);
      sts_d2(&hxy[r][4 * g], (pxy[0] + pxy[1]) + pxy[2], (pxy[1] + pxy[2]) + pxy[3]);
      sts_d2(&hxy[r][4 * g + 2], (pxy[2] + pxy[3]) + pxy[4], (pxy[3] + pxy[4]) + pxy[5]);
      sts_d2(&hyy[r][4 * g], (pyy[0] + pyy[1]) + pyy[2], (pyy[1] + pyy[2]) + pyy[3]);
      sts_d2(&hyy[r][4 * g + 2], (pyy[2] + pyy[3]) + pyy[4], (pyy[3] + pyy[4]) + pyy[5]);
    }
  } else {
    for (int i = tid; i < kE3T * kE3H; i += 256) {
      const int ty = i / kE3T, tx = i - ty * kE3T;
      // (rows / columns beyond the image's last ones are never read: keep them in range)
      const int xc = min(bx + tx, w - 1), y = reflect101(min(by - 1 + ty, h), h);
      double sxx = 0, sxy = 0, syy = 0;
#pragma unroll
      for (int d = -1; d <= 1; d++) {
        double a, b, c;
        sobel_products<true>(src, w, h, spitch, reflect101(xc + d, w), y, k0, k1, k2, wb, &a, &b, &c);
        sxx += a; sxy += b; syy += c;
      }
      hxx[ty][tx] = sxx; hxy[ty][tx] = sxy; hyy[ty][tx] = syy;
    }
  }
  __syncthreads();
  const size_t n = (size_t)w * h;
  float m = 0.f;
  const int lx = threadIdx.x, x = bx + lx;
  double vxx[6], vxy[6], vyy[6];
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const int tr = 4 * threadIdx.y + r;
    vxx[r] = hxx[tr][lx]; vxy[r] = hxy[tr][lx]; vyy[r] = hyy[tr][lx];
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int y = by + 4 * threadIdx.y + q;
    if (x >= w || y >= h) continue;
    const double sxx = (vxx[q] + vxx[q + 1]) + vxx[q + 2], sxy = (vxy[q] + vxy[q + 1]) + vxy[q + 2],
                 syy = (vyy[q] + vyy[q + 1]) + vyy[q + 2];
    float e;
    const size_t o = (size_t)y * w + x;
    if (!harris) {
      const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
      const float d = __fsub_rn(a, c);
      e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
    } else {
      const float a = (float)sxx, b = (float)sxy, c = (float)syy;
      const float t1 = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, b)), sm = __fadd_rn(a, c);
      if (o < n - (n & 7)) e = __fsub_rn(t1, __fmul_rn(kf, __fmul_rn(sm, sm)));
      else if (o < n - (n & 3)) e = __fsub_rn(t1, __fmul_rn(__fmul_rn(kf, sm), sm));
      else e = (float)__dsub_rn((double)t1, __dmul_rn(__dmul_rn(kd, (double)sm), (double)sm));
    }
    eig[o] = e;
    if (mask && mask[o] == 0) e = 0.f;
    m = fmaxf(m, e);
  }
  block_max_to_global(m, tid, max_bits);
}

// Candidate ordering.  cv2 sorts all candidates (std::sort, value descending, ties by DESCENDING address) and walks the
// sorted list until maxCorners are accepted.  Here a candidate is the 64-bit key (value bits << 32 | linear index) —
// descending key order IS cv2's order — and only the part of the order the walk consumes is ever established:
//   * k_candidates drops every candidate into one of 65536 VALUE buckets (order-preserving: positive floats compare
//     like their bit patterns; bucket = (bits - threshold bits) >> shift) and counts the buckets;
//   * k_bucket_scan turns the counts into bucket positions, strongest bucket first; k_bucket_scatter moves the keys
//     there (unordered inside a bucket);
//   * the selection kernel takes whole buckets, <= 1024 keys per round, and sorts a round in shared memory (bitonic).
// No candidate count crosses to the host; a bucket larger than a round (many equal values: checkerboards) is sorted
// by the CTA in global memory first.
struct CandRange {
  unsigned int thr_bits;
  int shift;
};
__device__ __forceinline__ CandRange cand_range(unsigned int max_bits, double quality, float* thr_out) {
  const float thr = (float)((double)__uint_as_float(max_bits) * quality);
  *thr_out = thr;
  CandRange r;
  r.thr_bits = __float_as_uint(fmaxf(thr, 0.f));
  const unsigned int range = max_bits > r.thr_bits ? max_bits - r.thr_bits : 1u;
  r.shift = max(0, 32 - __clz(range) - kLogBuckets);
  return r;
}

// threshold (THRESH_TOZERO at max*quality, strict >) + 3x3 local maximum, 1-px frame skipped, optional mask.
// One CTA per 32 x 32 tile: the tile's candidates are collected in shared memory and the global list is extended once per
// CTA (every candidate — then every warp — adding to the one counter itself was ~5e4 same-address atomics per 1080p frame,
// which L2 serialises: the kernel took 39 us for 8 MB of input).
constexpr int kCandTile = 32;
__global__ void __launch_bounds__(256) k_candidates(const float* __restrict__ eig, int w, int h, double quality,
                                                    const unsigned int* __restrict__ max_bits,
                                                    unsigned long long* __restrict__ keys, unsigned int* __restrict__ count,
                                                    unsigned int cap, unsigned int* __restrict__ hist,
                                                    const uint8_t* __restrict__ mask) {
  __shared__ unsigned long long s_keys[kCandTile * kCandTile];
  __shared__ unsigned int s_n, s_base;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid == 0) s_n = 0;
  __syncthreads();
  float thr;
  const CandRange cr = cand_range(*max_bits, quality, &thr);
  const int x = blockIdx.x * kCandTile + threadIdx.x + 1;
#pragma unroll
  for (int r = 0; r < kCandTile / 8; r++) {
    const int y = blockIdx.y * kCandTile + threadIdx.y + 8 * r + 1;
    if (x >= w - 1 || y >= h - 1) continue;
    const float v = eig[(size_t)y * w + x];
    if (!(v > thr)) continue;
    float mx = v;
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
      for (int i = -1; i <= 1; i++) mx = fmaxf(mx, eig[(size_t)(y + j) * w + x + i]);
    if (v != mx) continue;
    if (mask && mask[(size_t)y * w + x] == 0) continue;
    s_keys[atomicAdd(&s_n, 1u)] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(y * w + x);
  }
  __syncthreads();
  const unsigned int n = s_n;
  if (n == 0u) return;
  if (tid == 0) s_base = atomicAdd(count, n);
  __syncthreads();
  for (unsigned int i = tid; i < n; i += 256) {
    const unsigned int slot = s_base + i;
    if (slot < cap) {
      const unsigned long long key = s_keys[i];
      keys[slot] = key;
      atomicAdd(&hist[min(((unsigned int)(key >> 32) - cr.thr_bits) >> cr.shift, (unsigned int)kBuckets - 1u)], 1u);
    }
  }
}

// bstart[r] = first position of the r-th strongest bucket (r = kBuckets - 1 - bucket), bstart[kBuckets] = total;
// the counts are cleared: the scatter uses them as cursors.
// Two launches of kBuckets / 1024 CTAs, every access coalesced (one CTA reading 64 buckets per thread was latency-bound:
// 156 us per frame): k_bucket_sums adds up each CTA's 1024 buckets, k_bucket_scan places them behind the CTAs before it.
// Position r of the order = bucket kBuckets - 1 - r (strongest bucket first).
__device__ __forceinline__ unsigned int block_scan_1024(unsigned int v, unsigned int* wsum, unsigned int* total) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  unsigned int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    unsigned int w = wsum[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  *total = wsum[31];
  return inc - v + (wid ? wsum[wid - 1] : 0u);          // exclusive
}

__global__ void __launch_bounds__(1024) k_bucket_sums(const unsigned int* __restrict__ hist, unsigned int* __restrict__ part) {
  __shared__ unsigned int wsum[32];
  unsigned int total;
  block_scan_1024(hist[kBuckets - 1 - (blockIdx.x * 1024 + threadIdx.x)], wsum, &total);
  if (threadIdx.x == 0) part[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_bucket_scan(unsigned int* __restrict__ hist, unsigned int* __restrict__ bstart,
                                                      const unsigned int* __restrict__ part) {
  __shared__ unsigned int wsum[32];
  __shared__ unsigned int base;
  if (threadIdx.x < 32) {                                // candidates in the CTAs before this one (<= 64 partial sums)
    unsigned int v = 0;
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += 32) v += part[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) base = v;
  }
  const int r = blockIdx.x * 1024 + threadIdx.x, b = kBuckets - 1 - r;
  const unsigned int c = hist[b];
  unsigned int total;
  const unsigned int excl = block_scan_1024(c, wsum, &total);   // (its barriers also publish `base`)
  bstart[r] = base + excl;
  hist[b] = 0;
  if (r == kBuckets - 1) bstart[kBuckets] = base + total;
}

__global__ void __launch_bounds__(256) k_bucket_scatter(const unsigned long long* __restrict__ keys,
                                                        const unsigned int* __restrict__ count, unsigned int cap,
                                                        double quality, const unsigned int* __restrict__ max_bits,
                                                        unsigned int* __restrict__ cursor, const unsigned int* __restrict__ bstart,
                                                        unsigned long long* __restrict__ out) {
  const unsigned int n = min(*count, cap);
  float thr;
  const CandRange cr = cand_range(*max_bits, quality, &thr);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const unsigned int b = min(((unsigned int)(k >> 32) - cr.thr_bits) >> cr.shift, (unsigned int)kBuckets - 1u);
    out[bstart[kBuckets - 1 - b] + atomicAdd(&cursor[b], 1u)] = k;
  }
}

// Greedy minimum-distance selection in descending key order (lexicographically-first maximal independent set), exactly
// cv2's sequential rule, in rounds of <= 1024 candidates by one CTA:
//   fetch   : the next run of whole buckets (<= 1024 keys), sorted in shared memory;
//   phase 1 : every candidate of the round is tested against the corners accepted in EARLIER rounds
//             (grid of cell = cvRound(minDistance), +-1 cell); the survivors are compacted in rank order;
//   phase 2 : the survivors decide among themselves by the parallel form of the same rule (see below)
//             and the accepted ones are committed in rank order, up to the corner limit.
constexpr int GS_THREADS = 1024;
constexpr int GS_BUCKETS = 4096;

__device__ __forceinline__ bool gs_far_from_accepted(int x, int y, int cell, int gw, int gh, float md2,
                                                     const unsigned int* grid_cnt, const ushort2* grid_pts) {
  const int xc = x / cell, yc = y / cell;
  // the nine cell counts first (independent loads), then the cells' points: the walk is a chain of L2 round trips otherwise
  unsigned int cnt[9];
#pragma unroll
  for (int q = 0; q < 9; q++) {
    const int yy = yc + q / 3 - 1, xx = xc + q % 3 - 1;
    const bool in = (unsigned)yy < (unsigned)gh && (unsigned)xx < (unsigned)gw;
    cnt[q] = in ? min(grid_cnt[yy * gw + xx], (unsigned int)kGridSlots) : 0u;   // (one CTA: barrier/fence-ordered, L1-coherent)
  }
  bool far = true;
#pragma unroll
  for (int q = 0; q < 9; q++) {
    const int c = (yc + q / 3 - 1) * gw + xc + q % 3 - 1;
    for (unsigned int k = 0; k < cnt[q]; k++) {
      const ushort2 p = grid_pts[c * kGridSlots + k];
      const float dx = (float)(x - (int)p.x), dy = (float)(y - (int)p.y);
      if (dx * dx + dy * dy < md2) far = false;
    }
  }
  return far;
}

// descending bitonic sort of n = 2^k keys by the whole CTA; keys live in shared or global memory
__device__ __forceinline__ void gs_bitonic_desc(unsigned long long* a, unsigned int n) {
  for (unsigned int k = 2; k <= n; k <<= 1)
    for (unsigned int j = k >> 1; j > 0; j >>= 1) {
      for (unsigned int i = threadIdx.x; i < n; i += GS_THREADS) {
        const unsigned int p = i ^ j;
        if (p > i) {
          const unsigned long long u = a[i], v = a[p];
          if (((i & k) == 0) ? (u < v) : (u > v)) { a[i] = v; a[p] = u; }
        }
      }
      __syncthreads();
    }
}

// The same order for exactly GS_THREADS keys, one per thread: exchanges at distances < 32 are warp shuffles on the key
// in a register (40 of the 55 steps, no barrier), the 15 longer ones go through two alternating shared-memory buffers
// (one barrier per step).  Leaves the sorted keys in `a`.
__device__ __forceinline__ void gs_bitonic_desc_round(unsigned long long* a, unsigned long long* b) {
  const unsigned int i = threadIdx.x;
  unsigned long long u = a[i];
  unsigned long long* wr = b;                        // (a is still being read by other warps: first exchange goes to b)
  for (unsigned int k = 2; k <= (unsigned int)GS_THREADS; k <<= 1)
    for (unsigned int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long v;
      if (j >= 32u) {
        wr[i] = u;
        __syncthreads();
        v = wr[i ^ j];
        wr = wr == a ? b : a;
      } else {
        v = __shfl_xor_sync(0xffffffffu, u, (int)j);
      }
      // descending block ((i & k) == 0): the lower index keeps the larger key
      const bool keep_max = ((i & j) == 0u) == ((i & k) == 0u);
      u = keep_max ? (u > v ? u : v) : (u < v ? u : v);
    }
  __syncthreads();                                   // every read of the last exchange is done
  a[i] = u;
  __syncthreads();
}

__global__ void __launch_bounds__(GS_THREADS) k_greedy_select(const unsigned long long* __restrict__ bkeys,
                                                              const unsigned int* __restrict__ bstart,
                                                              unsigned long long* __restrict__ sortbuf,
                                                              int w, int h, float min_dist, int max_corners,
                                                              unsigned int corner_cap, unsigned int* grid_cnt,
                                                              ushort2* grid_pts, float2* __restrict__ corners,
                                                              unsigned int* __restrict__ n_out) {
  __shared__ unsigned long long rk[GS_THREADS];      // the round's keys, descending
  __shared__ unsigned long long rk2[GS_THREADS];     // exchange buffer of the round sort
  __shared__ unsigned int surv[GS_THREADS];          // x | y << 16, rank order
  __shared__ unsigned int warp_cnt[GS_THREADS / 32];
  __shared__ unsigned int s_nsurv, s_accepted;
  __shared__ unsigned int bucket[GS_BUCKETS];        // head of the chain of in-round survivors per hashed grid cell
  __shared__ unsigned short chain[GS_THREADS];
  __shared__ unsigned char state[GS_THREADS];        // 0 undecided, 1 accepted, 2 rejected
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned int total = bstart[kBuckets];
  const unsigned int limit = min(max_corners > 0 ? (unsigned int)max_corners : 0xffffffffu, corner_cap);
  const bool use_dist = min_dist >= 1.f;
  const int cell = use_dist ? __float2int_rn(min_dist) : 1;
  const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
  const float md2 = min_dist * min_dist;
  if (tid == 0) s_accepted = 0;
  __syncthreads();
  // walk state (uniform over the CTA): next bucket rank, next key position, keys left of an oversized bucket
  unsigned int rb = 0, pos = 0, ov_left = 0, ov_pos = 0;
  for (;;) {
    if (s_accepted >= limit) break;                  // (uniform: read after the barrier that ends a round)
    // ---- fetch the next round
    unsigned int nr = 0;
    if (ov_left > 0) {
      nr = min(ov_left, (unsigned int)GS_THREADS);
      rk[tid] = (unsigned int)tid < nr ? sortbuf[ov_pos + tid] : 0ull;
      ov_pos += nr;
      ov_left -= nr;
      __syncthreads();
    } else {
      if (pos >= total) break;
      // how many whole buckets fit into one round: bstart is non-decreasing, so "fits" is a prefix property and the CTA
      // finds its end in two probes — every 64th bucket behind rb, then the 64 buckets behind the last fitting one
      // (walking 1024 buckets per step took up to 64 dependent L2 round trips in the sparse strong end of the range)
      static_assert(GS_THREADS * 64 >= kBuckets, "coarse probes cover every bucket");
      unsigned int rb_end;
      {
        const unsigned int p1 = rb + ((unsigned int)tid + 1u) * 64u;
        const bool ok1 = p1 <= (unsigned int)kBuckets && bstart[p1] - pos <= (unsigned int)GS_THREADS;
        const unsigned int base = rb + 64u * (unsigned int)__syncthreads_count(ok1);
        const unsigned int p2 = base + 1u + (unsigned int)tid;
        const bool ok2 = tid < 64 && p2 <= (unsigned int)kBuckets && bstart[p2] - pos <= (unsigned int)GS_THREADS;
        rb_end = base + (unsigned int)__syncthreads_count(ok2);
      }
      if (rb_end == rb) {
        // the strongest remaining bucket alone is larger than a round: order it in global memory first
        const unsigned int m_b = bstart[rb + 1] - pos;
        unsigned int p2 = GS_THREADS;
        while (p2 < m_b) p2 <<= 1;
        for (unsigned int i = tid; i < p2; i += GS_THREADS) sortbuf[i] = i < m_b ? bkeys[pos + i] : 0ull;
        __syncthreads();
        gs_bitonic_desc(sortbuf, p2);
        ov_left = m_b;
        ov_pos = 0;
        pos += m_b;
        rb += 1;
        continue;
      }
      nr = bstart[rb_end] - pos;
      rk[tid] = (unsigned int)tid < nr ? bkeys[pos + tid] : 0ull;
      pos += nr;
      rb = rb_end;
      __syncthreads();
      if (nr > 1) gs_bitonic_desc_round(rk, rk2);
    }
    if (nr == 0) continue;
    if (!use_dist) {
      // minDistance < 1: every candidate is a corner, in order
      const unsigned int o = s_accepted + tid;
      if ((unsigned int)tid < nr && o < limit) {
        const unsigned int idx = (unsigned int)(rk[tid] & 0xffffffffu);
        corners[o] = make_float2((float)(idx % w), (float)(idx / w));
      }
      __syncthreads();
      if (tid == 0) s_accepted = min(s_accepted + nr, limit);
      __syncthreads();
      continue;
    }
    // ---- phase 1
    bool alive = (unsigned int)tid < nr;
    int x = 0, y = 0;
    if (alive) {
      const unsigned int idx = (unsigned int)(rk[tid] & 0xffffffffu);
      x = idx % w;
      y = idx / w;
      alive = gs_far_from_accepted(x, y, cell, gw, gh, md2, grid_cnt, grid_pts);
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    unsigned int off = 0;
    for (int k = 0; k < wid; k++) off += warp_cnt[k];
    if (alive) surv[off + __popc(bal & ((1u << lane) - 1u))] = (unsigned int)x | ((unsigned int)y << 16);
    if (tid == GS_THREADS - 1) s_nsurv = off + __popc(bal);
    __syncthreads();
    // ---- phase 2: the survivors of the round decide among themselves, all in parallel.  Survivor i (rank order) is
    // rejected as soon as an EARLIER survivor within minDistance is accepted, and accepted once all of those are
    // rejected; the lowest-ranked undecided survivor can always decide, so the loop ends, and the result is the
    // sequential rule's (the lexicographically first maximal independent set is unique).  Neighbours are found through
    // a hash of the grid cell into 4096 chains in shared memory.
    {
      const unsigned int ns = s_nsurv;
      for (int b = tid; b < GS_BUCKETS; b += GS_THREADS) bucket[b] = 0xffffu;
      __syncthreads();
      int sx = 0, sy = 0;
      if ((unsigned int)tid < ns) {
        const unsigned int v = surv[tid];
        sx = (int)(v & 0xffffu);
        sy = (int)(v >> 16);
        state[tid] = 0;
      }
      __syncthreads();
      if ((unsigned int)tid < ns) {
        const unsigned int bk = ((unsigned int)(sy / cell) * 73u + (unsigned int)(sx / cell)) & (GS_BUCKETS - 1);
        chain[tid] = (unsigned short)atomicExch(&bucket[bk], (unsigned int)tid);
      }
      __syncthreads();
      // the earlier survivors within minDistance of this one, found ONCE (the decision loop below re-reads only their
      // states; walking the nine cells' chains in every iteration was a third of the kernel's instructions, and chains
      // of mutually close candidates take ten and more iterations).  Up to kNb of them in the sort's exchange buffer;
      // a survivor with more walks the chains as before.
      constexpr int kNb = 4;
      unsigned short* nbl = reinterpret_cast<unsigned short*>(rk2);     // [GS_THREADS][kNb]
      int ncnt = 0;
      auto walk = [&](auto&& visit) {
        const int xc = sx / cell, yc = sy / cell;
        for (int yy = max(0, yc - 1); yy <= min(gh - 1, yc + 1); yy++)
          for (int xx = max(0, xc - 1); xx <= min(gw - 1, xc + 1); xx++) {
            unsigned int k = bucket[((unsigned int)yy * 73u + (unsigned int)xx) & (GS_BUCKETS - 1)];
            while (k != 0xffffu) {
              if (k < (unsigned int)tid) {
                const unsigned int p = surv[k];
                const float dx = (float)(sx - (int)(p & 0xffffu)), dy = (float)(sy - (int)(p >> 16));
                if (dx * dx + dy * dy < md2 && visit(k)) return;
              }
              k = chain[k];
            }
          }
      };
      if ((unsigned int)tid < ns)
        walk([&](unsigned int k) {
          if (ncnt < kNb) nbl[tid * kNb + ncnt] = (unsigned short)k;
          ncnt++;
          return false;
        });
      for (;;) {
        bool undecided = false;
        if ((unsigned int)tid < ns && state[tid] == 0) {
          bool blocked = false, rejected = false;
          auto look = [&](unsigned int k) {
            const unsigned char sk = *(volatile unsigned char*)&state[k];
            if (sk == 1) { rejected = true; return true; }
            if (sk == 0) blocked = true;
            return false;
          };
          if (ncnt <= kNb) {
            for (int q = 0; q < ncnt; q++)
              if (look(nbl[tid * kNb + q])) break;
          } else {
            walk(look);
          }
          if (rejected) state[tid] = 2;
          else if (!blocked) state[tid] = 1;
          else undecided = true;
        }
        if (!__syncthreads_or(undecided)) break;
      }
      // accepted survivors in rank order -> positions; stop at the corner limit
      const bool acc = (unsigned int)tid < ns && state[tid] == 1;
      const unsigned int bal2 = __ballot_sync(0xffffffffu, acc);
      if (lane == 0) warp_cnt[wid] = __popc(bal2);
      __syncthreads();
      unsigned int before = 0, total_acc = 0;
      for (int k = 0; k < GS_THREADS / 32; k++) {
        if (k < wid) before += warp_cnt[k];
        total_acc += warp_cnt[k];
      }
      const unsigned int o = s_accepted + before + __popc(bal2 & ((1u << lane) - 1u));
      if (acc && o < limit) {
        corners[o] = make_float2((float)sx, (float)sy);
        const int c = (sy / cell) * gw + (sx / cell);
        const unsigned int sl = atomicAdd(&grid_cnt[c], 1u);
        if (sl < (unsigned int)kGridSlots) grid_pts[c * kGridSlots + sl] = make_ushort2((unsigned short)sx, (unsigned short)sy);
      }
      __syncthreads();
      if (tid == 0) s_accepted = min(s_accepted + total_acc, limit);
    }
    __threadfence_block();
    __syncthreads();
  }
  if (tid == 0) *n_out = s_accepted;
}

// ======================================================================================
// a12: pyramidal Lucas-Kanade tracker — one warp per point, all levels in one launch
// ======================================================================================
struct LkLevels {
  int n_levels;                      // levels 0..n_levels-1
  const uint8_t* I[kMaxLkLevels];
  const uint8_t* J[kMaxLkLevels];
  const short2* D[kMaxLkLevels];
  int w[kMaxLkLevels], h[kMaxLkLevels];
  size_t pitch0;                     // pitch of level 0 (others are packed, pitch = w)
};

__device__ __forceinline__ void lk_weights(float a, float b, int* iw00, int* iw01, int* iw10, int* iw11) {
  const float s = 16384.f;  // 1 << W_BITS
  *iw00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), s));
  *iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), s));
  *iw10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), s));
  *iw11 = 16384 - *iw00 - *iw01 - *iw10;
}

__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// fixed-point bilinear sample of a u8 image with the REFLECT_101 border cv2's pyramid carries
__device__ __forceinline__ int lk_sample_u8(const uint8_t* __restrict__ img, size_t pitch, int w, int h, int x, int y,
                                            int iw00, int iw01, int iw10, int iw11, bool interior) {
  int x0 = x, x1 = x + 1, y0 = y, y1 = y + 1;
  if (!interior) { x0 = reflect101(x0, w); x1 = reflect101(x1, w); y0 = reflect101(y0, h); y1 = reflect101(y1, h); }
  const uint8_t* r0 = img + (size_t)y0 * pitch;
  const uint8_t* r1 = img + (size_t)y1 * pitch;
  return descale((int)r0[x0] * iw00 + (int)r0[x1] * iw01 + (int)r1[x0] * iw10 + (int)r1[x1] * iw11, 9);
}

__device__ __forceinline__ short2 lk_deriv_at(const short2* __restrict__ d, int w, int h, int x, int y) {
  if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return make_short2(0, 0);  // constant-0 border
  return d[(size_t)y * w + x];
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int LK_WARPS = 4;
#ifndef OFB_LK_WPP
#define OFB_LK_WPP 2          // warps per point of k_lk_track_cta at the 21 x 21 window
#endif

// Unroll factor of the interior window loops (experiment builds: -DOFB_LK_UNROLL=n; default: complete).
#define OFB_PRAGMA_(x) _Pragma(#x)
#define OFB_PRAGMA(x) OFB_PRAGMA_(x)
#ifdef OFB_LK_UNROLL
#define OFB_LK_PRAGMA_UNROLL OFB_PRAGMA(unroll OFB_LK_UNROLL)
#else
#define OFB_LK_PRAGMA_UNROLL OFB_PRAGMA(unroll)
#endif

// The window passes of the tracker, with the border decision as a template argument: whether the window (plus the
// bilinear neighbour) lies inside the level is uniform over the warp and decided ONCE per pass — inside the loops it
// was a branch (with reconvergence bookkeeping) per load: ncu counted 120-170 SASS instructions per window pixel,
// a tenth of them branches.  Same arithmetic in the same order.
//   MODE 0: window inside the level.   MODE 1: window crosses the border of a level at least as large as the window (every
//   pyramid level cv2 builds): REFLECT_101 is a single branch-free reflection, the derivative's constant border a
//   predicated load.  The launch lasts as long as its slowest warp, and the slowest warps were the points within half a
//   window of the border of the COARSEST level (88 px at level 0: a quarter of the corners of a 1080p frame) on the
//   generic path.   MODE 2: generic (multiple reflections; level 0 smaller than the window).
template <int MODE>
__device__ __forceinline__ int lk_sample_j(const uint8_t* __restrict__ J, size_t pitch, int cols, int rows, int x, int y,
                                           int iw00, int iw01, int iw10, int iw11) {
  if (MODE == 0) {
    const uint8_t* r0 = J + (size_t)y * pitch + x;
    const uint8_t* r1 = r0 + pitch;
    return descale((int)r0[0] * iw00 + (int)r0[1] * iw01 + (int)r1[0] * iw10 + (int)r1[1] * iw11, 9);
  }
  if (MODE == 1) {      // every coordinate overshoots by less than the level's size: one reflection, no branches
    const int x0 = reflect101_once(x, cols), x1 = reflect101_once(x + 1, cols);
    const uint8_t* r0 = J + (size_t)reflect101_once(y, rows) * pitch;
    const uint8_t* r1 = J + (size_t)reflect101_once(y + 1, rows) * pitch;
    return descale((int)r0[x0] * iw00 + (int)r0[x1] * iw01 + (int)r1[x0] * iw10 + (int)r1[x1] * iw11, 9);
  }
  return lk_sample_u8(J, pitch, cols, rows, x, y, iw00, iw01, iw10, iw11, false);
}

template <int MODE, int TW, int TH>
__device__ __forceinline__ void lk_window_build(const uint8_t* __restrict__ I, const short2* __restrict__ D, size_t pitch,
                                                int cols, int rows, int ipx, int ipy, int iw00, int iw01, int iw10, int iw11,
                                                int ww, int area, int lane, short* Iwin, short* dIx, short* dIy,
                                                float* pa11, float* pa12, float* pa22) {
  float a11 = 0.f, a12 = 0.f, a22 = 0.f;
  auto step = [&](int e) {
    const int wy = e / ww, wx = e - wy * ww;
    const int x = ipx + wx, y = ipy + wy;
    int ival;
    short2 d00, d01, d10, d11;
    if (MODE == 0) {
      const uint8_t* r0 = I + (size_t)y * pitch + x;
      const uint8_t* r1 = r0 + pitch;
      ival = descale((int)r0[0] * iw00 + (int)r0[1] * iw01 + (int)r1[0] * iw10 + (int)r1[1] * iw11, 9);
      const short2* q0 = D + (size_t)y * cols + x;
      const short2* q1 = q0 + cols;
      d00 = q0[0]; d01 = q0[1]; d10 = q1[0]; d11 = q1[1];
    } else if (MODE == 1) {
      ival = lk_sample_j<1>(I, pitch, cols, rows, x, y, iw00, iw01, iw10, iw11);
      const bool xa = (unsigned)x < (unsigned)cols, xb = (unsigned)(x + 1) < (unsigned)cols;
      const bool ya = (unsigned)y < (unsigned)rows, yb = (unsigned)(y + 1) < (unsigned)rows;
      const short2* q0 = D + (ptrdiff_t)y * cols + x;    // (only dereferenced where inside)
      const short2* q1 = q0 + cols;
      const short2 z = make_short2(0, 0);                // constant-0 border of the derivative image
      d00 = xa && ya ? q0[0] : z; d01 = xb && ya ? q0[1] : z; d10 = xa && yb ? q1[0] : z; d11 = xb && yb ? q1[1] : z;
    } else {
      ival = lk_sample_u8(I, pitch, cols, rows, x, y, iw00, iw01, iw10, iw11, false);
      d00 = lk_deriv_at(D, cols, rows, x, y); d01 = lk_deriv_at(D, cols, rows, x + 1, y);
      d10 = lk_deriv_at(D, cols, rows, x, y + 1); d11 = lk_deriv_at(D, cols, rows, x + 1, y + 1);
    }
    const int ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
    const int iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
    Iwin[e] = (short)ival;
    dIx[e] = (short)ix;
    dIy[e] = (short)iy;
    a11 += (float)(ix * ix);
    a12 += (float)(ix * iy);
    a22 += (float)(iy * iy);
  };
  if (MODE != 2) {
    OFB_LK_PRAGMA_UNROLL
    for (int e = lane; e < area; e += 32) step(e);
  } else {           // (windows larger than the level: rare, keep it small)
#pragma unroll 1
    for (int e = lane; e < area; e += 32) step(e);
  }
  *pa11 = a11; *pa12 = a12; *pa22 = a22;
}

template <int MODE>
__device__ __forceinline__ void lk_window_mismatch(const uint8_t* __restrict__ J, size_t pitch, int cols, int rows, int jx,
                                                   int jy, int iw00, int iw01, int iw10, int iw11, int ww, int area, int lane,
                                                   const short* Iwin, const short* dIx, const short* dIy, float* pb1,
                                                   float* pb2) {
  float b1 = 0.f, b2 = 0.f;
  auto step = [&](int e) {
    const int wy = e / ww, wx = e - wy * ww;
    const int diff = lk_sample_j<MODE>(J, pitch, cols, rows, jx + wx, jy + wy, iw00, iw01, iw10, iw11) - (int)Iwin[e];
    b1 += (float)(diff * (int)dIx[e]);
    b2 += (float)(diff * (int)dIy[e]);
  };
  if (MODE != 2) {
    OFB_LK_PRAGMA_UNROLL
    for (int e = lane; e < area; e += 32) step(e);
  } else {           // (windows larger than the level: rare, keep it small)
#pragma unroll 1
    for (int e = lane; e < area; e += 32) step(e);
  }
  *pb1 = b1; *pb2 = b2;
}

template <int MODE>
__device__ __forceinline__ float lk_window_abs_error(const uint8_t* __restrict__ J, size_t pitch, int cols, int rows, int jx,
                                                     int jy, int iw00, int iw01, int iw10, int iw11, int ww, int area,
                                                     int lane, const short* Iwin) {
  float ev = 0.f;
  auto step = [&](int e) {
    const int wy = e / ww, wx = e - wy * ww;
    const int diff = lk_sample_j<MODE>(J, pitch, cols, rows, jx + wx, jy + wy, iw00, iw01, iw10, iw11) - (int)Iwin[e];
    ev += fabsf((float)diff);
  };
  if (MODE != 2) {
    OFB_LK_PRAGMA_UNROLL
    for (int e = lane; e < area; e += 32) step(e);
  } else {           // (windows larger than the level: rare, keep it small)
#pragma unroll 1
    for (int e = lane; e < area; e += 32) step(e);
  }
  return ev;
}

// TW x TH: the window as a compile-time constant (0 = run time).  The window loops then have a constant trip count and
// unroll completely: the 4 byte loads of all ceil(area / 32) window pixels of a lane are in flight together instead of
// one dependent round trip per step, and e / ww is a multiply.  The arithmetic and its order are the same (same bits).
template <int TW, int TH>
__global__ void __launch_bounds__(LK_WARPS * 32) k_lk_track(LkLevels lv, const float2* __restrict__ prev_pts,
                                                            float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                            float* __restrict__ err, int n_points,
                                                            const unsigned int* __restrict__ n_dev, int ww_rt, int wh_rt,
                                                            int max_count, double eps2, int flags, double min_eig_thr) {
  extern __shared__ short lk_smem[];  // per warp: Iwin[area], dIx[area], dIy[area]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = blockIdx.x * LK_WARPS + warp;
  // n_dev: the point count is a device-side value (corner list of the previous frame); the grid covers n_points = its bound
  if (pid >= (n_dev ? (int)min(*n_dev, (unsigned int)n_points) : n_points)) return;
  const int ww = TW ? TW : ww_rt, wh = TH ? TH : wh_rt;
  const int area = ww * wh;
  short* Iwin = lk_smem + (size_t)warp * 3 * area;
  short* dIx = Iwin + area;
  short* dIy = dIx + area;

  const float2 p0 = prev_pts[pid];
  const float halfx = (ww - 1) * 0.5f, halfy = (wh - 1) * 0.5f;
  float2 np = make_float2(0.f, 0.f);
  bool st = true;
  float er = 0.f;
  const bool use_init = (flags & OFB_OPTFLOW_USE_INITIAL_FLOW) != 0;
  const bool get_min_eig = (flags & OFB_OPTFLOW_LK_GET_MIN_EIGENVALS) != 0;
  const float2 init = use_init ? next_pts[pid] : p0;
  const float FLT_SCALE = 1.f / (1 << 20);

  for (int level = lv.n_levels - 1; level >= 0; level--) {
    const float sc = (float)(1. / (1 << level));
    const uint8_t* I = lv.I[level];
    const uint8_t* J = lv.J[level];
    const short2* D = lv.D[level];
    const int cols = lv.w[level], rows = lv.h[level];
    const size_t pitch = level == 0 ? lv.pitch0 : (size_t)cols;
    float2 pp = make_float2(__fmul_rn(p0.x, sc), __fmul_rn(p0.y, sc));
    if (level == lv.n_levels - 1) np = use_init ? make_float2(__fmul_rn(init.x, sc), __fmul_rn(init.y, sc)) : pp;
    else np = make_float2(__fmul_rn(np.x, 2.f), __fmul_rn(np.y, 2.f));

    pp.x -= halfx;
    pp.y -= halfy;
    const int ipx = (int)floorf(pp.x), ipy = (int)floorf(pp.y);
    if (ipx < -ww || ipx >= cols || ipy < -wh || ipy >= rows) {
      if (level == 0) { st = false; er = 0.f; }
      continue;
    }
    int iw00, iw01, iw10, iw11;
    lk_weights(pp.x - (float)ipx, pp.y - (float)ipy, &iw00, &iw01, &iw10, &iw11);
    const bool in_i = ipx >= 0 && ipy >= 0 && ipx + ww + 1 <= cols && ipy + wh + 1 <= rows;
    float a11, a12, a22;
    __syncwarp();
    const bool small = cols <= ww || rows <= wh;       // a coordinate may need more than one reflection
    if (in_i) lk_window_build<0, TW, TH>(I, D, pitch, cols, rows, ipx, ipy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin, dIx, dIy, &a11, &a12, &a22);
    else if (!small) lk_window_build<1, TW, TH>(I, D, pitch, cols, rows, ipx, ipy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin, dIx, dIy, &a11, &a12, &a22);
    else lk_window_build<2, TW, TH>(I, D, pitch, cols, rows, ipx, ipy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin, dIx, dIy, &a11, &a12, &a22);
    __syncwarp();
    const float A11 = __fmul_rn(warp_sum(a11), FLT_SCALE), A12 = __fmul_rn(warp_sum(a12), FLT_SCALE),
                A22 = __fmul_rn(warp_sum(a22), FLT_SCALE);
    float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float min_eig =
        __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11),
                            __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                  (float)(2 * ww * wh));
    if (get_min_eig) er = min_eig;
    if ((double)min_eig < min_eig_thr || Dt < 1.1920929e-07f) {
      if (level == 0) st = false;
      continue;
    }
    Dt = __fdiv_rn(1.f, Dt);
    float2 q = make_float2(np.x - halfx, np.y - halfy);
    float2 prev_delta = make_float2(0.f, 0.f);
    for (int j = 0; j < max_count; j++) {
      const int jx = (int)floorf(q.x), jy = (int)floorf(q.y);
      if (jx < -ww || jx >= cols || jy < -wh || jy >= rows) {
        if (level == 0) st = false;
        break;
      }
      lk_weights(q.x - (float)jx, q.y - (float)jy, &iw00, &iw01, &iw10, &iw11);
      const bool in_j = jx >= 0 && jy >= 0 && jx + ww + 1 <= cols && jy + wh + 1 <= rows;
      float b1, b2;
      if (in_j) lk_window_mismatch<0>(J, pitch, cols, rows, jx, jy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin, dIx, dIy, &b1, &b2);
      else if (!small) lk_window_mismatch<1>(J, pitch, cols, rows, jx, jy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin, dIx, dIy, &b1, &b2);
      else lk_window_mismatch<2>(J, pitch, cols, rows, jx, jy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin, dIx, dIy, &b1, &b2);
      const float B1 = __fmul_rn(warp_sum(b1), FLT_SCALE), B2 = __fmul_rn(warp_sum(b2), FLT_SCALE);
      const float2 delta = make_float2(__fmul_rn(__fsub_rn(__fmul_rn(A12, B2), __fmul_rn(A22, B1)), Dt),
                                       __fmul_rn(__fsub_rn(__fmul_rn(A12, B1), __fmul_rn(A11, B2)), Dt));
      q.x += delta.x;
      q.y += delta.y;
      np = make_float2(q.x + halfx, q.y + halfy);
      if ((double)delta.x * delta.x + (double)delta.y * delta.y <= eps2) break;
      if (j > 0 && fabs((double)(delta.x + prev_delta.x)) < 0.01 && fabs((double)(delta.y + prev_delta.y)) < 0.01) {
        np.x -= __fmul_rn(delta.x, 0.5f);
        np.y -= __fmul_rn(delta.y, 0.5f);
        break;
      }
      prev_delta = delta;
    }
    if (st && level == 0 && !get_min_eig) {
      const float2 r = make_float2(np.x - halfx, np.y - halfy);
      const int jx = (int)floorf(r.x), jy = (int)floorf(r.y);
      if (jx < -ww || jx >= cols || jy < -wh || jy >= rows) {
        st = false;
        continue;
      }
      lk_weights(r.x - (float)jx, r.y - (float)jy, &iw00, &iw01, &iw10, &iw11);
      const bool in_j = jx >= 0 && jy >= 0 && jx + ww + 1 <= cols && jy + wh + 1 <= rows;
      const float ev = in_j ? lk_window_abs_error<0>(J, pitch, cols, rows, jx, jy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin)
                       : !small ? lk_window_abs_error<1>(J, pitch, cols, rows, jx, jy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin)
                                : lk_window_abs_error<2>(J, pitch, cols, rows, jx, jy, iw00, iw01, iw10, iw11, ww, area, lane, Iwin);
      er = __fdiv_rn(warp_sum(ev), (float)(32 * ww * wh));
    }
  }
  if (lane == 0) {
    next_pts[pid] = np;
    status[pid] = st ? 1 : 0;
    err[pid] = er;
  }
}

// ---- the tracker with ONE POINT PER CTA of WPP warps (fixed window TW x TH) -------------------------------------------------
// k_lk_track above lasts as long as its slowest warp: all points are resident at once (2000 warps on 148 SMs), every warp
// walks its 4 levels x (1 build + up to 30 mismatch passes) alone, and a pass over a 21 x 21 window is 14 pixels per lane
// (ncu: 28 M warp-instructions, 14 k per warp, 6.6 cycles per instruction with 2.4 warps per scheduler: 103 us, the
// average warp done after half of that).  Here the window of a point is spread over WPP warps: TW * TH / (32 WPP) pixels
// per thread, whose patch values stay in REGISTERS (a thread only ever re-reads the pixels it built: no shared-memory
// patch), partial sums meet through one double-buffered shared-memory slot and one barrier per pass, and every thread
// carries the point's scalar state redundantly (uniform control flow, no broadcast).  The float sums are taken in a
// different order than in the warp kernel (and than cv2's): positions agree within the contract tolerance, not bit for bit.
template <int MODE>
__device__ __forceinline__ void lk_build_px(const uint8_t* __restrict__ I, const short2* __restrict__ D, size_t pitch,
                                            int cols, int rows, int x, int y, int iw00, int iw01, int iw10, int iw11,
                                            int* ival, int* ix, int* iy) {
  short2 d00, d01, d10, d11;
  *ival = lk_sample_j<MODE>(I, pitch, cols, rows, x, y, iw00, iw01, iw10, iw11);
  if (MODE == 0) {
    const short2* q0 = D + (size_t)y * cols + x;
    const short2* q1 = q0 + cols;
    d00 = q0[0]; d01 = q0[1]; d10 = q1[0]; d11 = q1[1];
  } else {
    d00 = lk_deriv_at(D, cols, rows, x, y); d01 = lk_deriv_at(D, cols, rows, x + 1, y);
    d10 = lk_deriv_at(D, cols, rows, x, y + 1); d11 = lk_deriv_at(D, cols, rows, x + 1, y + 1);
  }
  *ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
  *iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
}

template <int TW, int TH, int WPP>
__global__ void __launch_bounds__(WPP * 32) k_lk_track_cta(LkLevels lv, const float2* __restrict__ prev_pts,
                                                           float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                           float* __restrict__ err, int n_points,
                                                           const unsigned int* __restrict__ n_dev, int max_count, double eps2,
                                                           int flags, double min_eig_thr) {
  constexpr int NT = WPP * 32, AREA = TW * TH, STEPS = (AREA + NT - 1) / NT;
  constexpr int ww = TW, wh = TH;
  __shared__ float part[2][WPP][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pid = blockIdx.x;
  if (pid >= (n_dev ? (int)min(*n_dev, (unsigned int)n_points) : n_points)) return;   // (uniform over the CTA)
  int par = 0;
  // sums over the CTA: butterfly inside the warps, the WPP partial sums through shared memory in a fixed order.  The slot
  // alternates, so one barrier per reduction is enough: a thread reaches the next use of a slot only through the barrier
  // of the reduction in between, which every thread passes after its reads of this one.
  auto cta_sum3 = [&](float& a, float& b, float& c) {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { part[par][warp][0] = a; part[par][warp][1] = b; part[par][warp][2] = c; }
    __syncthreads();
    a = part[par][0][0]; b = part[par][0][1]; c = part[par][0][2];
#pragma unroll
    for (int k = 1; k < WPP; k++) { a += part[par][k][0]; b += part[par][k][1]; c += part[par][k][2]; }
    par ^= 1;
  };
  // window coordinates of this thread's pixels
  int wxs[STEPS], wys[STEPS];
#pragma unroll
  for (int s2 = 0; s2 < STEPS; s2++) {
    const int e = tid + s2 * NT;
    wys[s2] = e / TW;
    wxs[s2] = e - wys[s2] * TW;
  }

  const float2 p0 = prev_pts[pid];
  const float halfx = (ww - 1) * 0.5f, halfy = (wh - 1) * 0.5f;
  float2 np = make_float2(0.f, 0.f);
  bool st = true;
  float er = 0.f;
  const bool use_init = (flags & OFB_OPTFLOW_USE_INITIAL_FLOW) != 0;
  const bool get_min_eig = (flags & OFB_OPTFLOW_LK_GET_MIN_EIGENVALS) != 0;
  const float2 init = use_init ? next_pts[pid] : p0;
  const float FLT_SCALE = 1.f / (1 << 20);

  for (int level = lv.n_levels - 1; level >= 0; level--) {
    const float sc = (float)(1. / (1 << level));
    const uint8_t* I = lv.I[level];
    const uint8_t* J = lv.J[level];
    const short2* D = lv.D[level];
    const int cols = lv.w[level], rows = lv.h[level];
    const size_t pitch = level == 0 ? lv.pitch0 : (size_t)cols;
    float2 pp = make_float2(__fmul_rn(p0.x, sc), __fmul_rn(p0.y, sc));
    if (level == lv.n_levels - 1) np = use_init ? make_float2(__fmul_rn(init.x, sc), __fmul_rn(init.y, sc)) : pp;
    else np = make_float2(__fmul_rn(np.x, 2.f), __fmul_rn(np.y, 2.f));

    pp.x -= halfx;
    pp.y -= halfy;
    const int ipx = (int)floorf(pp.x), ipy = (int)floorf(pp.y);
    if (ipx < -ww || ipx >= cols || ipy < -wh || ipy >= rows) {
      if (level == 0) { st = false; er = 0.f; }
      continue;
    }
    int iw00, iw01, iw10, iw11;
    lk_weights(pp.x - (float)ipx, pp.y - (float)ipy, &iw00, &iw01, &iw10, &iw11);
    const bool in_i = ipx >= 0 && ipy >= 0 && ipx + ww + 1 <= cols && ipy + wh + 1 <= rows;
    const bool small = cols <= ww || rows <= wh;       // a coordinate may need more than one reflection
    int Iw[STEPS], dX[STEPS], dY[STEPS];               // this thread's part of the patch
    float a11 = 0.f, a12 = 0.f, a22 = 0.f;
    auto build = [&](auto mode) {
      constexpr int MODE = decltype(mode)::value;
#pragma unroll
      for (int s2 = 0; s2 < STEPS; s2++) {
        Iw[s2] = 0; dX[s2] = 0; dY[s2] = 0;
        if (s2 < STEPS - 1 || tid + s2 * NT < AREA) {
          lk_build_px<MODE>(I, D, pitch, cols, rows, ipx + wxs[s2], ipy + wys[s2], iw00, iw01, iw10, iw11, &Iw[s2], &dX[s2],
                            &dY[s2]);
          a11 += (float)(dX[s2] * dX[s2]);
          a12 += (float)(dX[s2] * dY[s2]);
          a22 += (float)(dY[s2] * dY[s2]);
        }
      }
    };
    if (in_i) build(std::integral_constant<int, 0>());
    else if (!small) build(std::integral_constant<int, 1>());
    else build(std::integral_constant<int, 2>());
    cta_sum3(a11, a12, a22);
    const float A11 = __fmul_rn(a11, FLT_SCALE), A12 = __fmul_rn(a12, FLT_SCALE), A22 = __fmul_rn(a22, FLT_SCALE);
    float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dd = __fsub_rn(A11, A22);
    const float min_eig =
        __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11),
                            __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                  (float)(2 * ww * wh));
    if (get_min_eig) er = min_eig;
    if ((double)min_eig < min_eig_thr || Dt < 1.1920929e-07f) {
      if (level == 0) st = false;
      continue;
    }
    Dt = __fdiv_rn(1.f, Dt);
    float2 q = make_float2(np.x - halfx, np.y - halfy);
    float2 prev_delta = make_float2(0.f, 0.f);
    for (int j = 0; j < max_count; j++) {
      const int jx = (int)floorf(q.x), jy = (int)floorf(q.y);
      if (jx < -ww || jx >= cols || jy < -wh || jy >= rows) {
        if (level == 0) st = false;
        break;
      }
      lk_weights(q.x - (float)jx, q.y - (float)jy, &iw00, &iw01, &iw10, &iw11);
      const bool in_j = jx >= 0 && jy >= 0 && jx + ww + 1 <= cols && jy + wh + 1 <= rows;
      float b1 = 0.f, b2 = 0.f, unused = 0.f;
      auto mismatch = [&](auto mode) {
        constexpr int MODE = decltype(mode)::value;
#pragma unroll
        for (int s2 = 0; s2 < STEPS; s2++)
          if (s2 < STEPS - 1 || tid + s2 * NT < AREA) {
            const int diff = lk_sample_j<MODE>(J, pitch, cols, rows, jx + wxs[s2], jy + wys[s2], iw00, iw01, iw10, iw11) - Iw[s2];
            b1 += (float)(diff * dX[s2]);
            b2 += (float)(diff * dY[s2]);
          }
      };
      if (in_j) mismatch(std::integral_constant<int, 0>());
      else if (!small) mismatch(std::integral_constant<int, 1>());
      else mismatch(std::integral_constant<int, 2>());
      cta_sum3(b1, b2, unused);
      const float B1 = __fmul_rn(b1, FLT_SCALE), B2 = __fmul_rn(b2, FLT_SCALE);
      const float2 delta = make_float2(__fmul_rn(__fsub_rn(__fmul_rn(A12, B2), __fmul_rn(A22, B1)), Dt),
                                       __fmul_rn(__fsub_rn(__fmul_rn(A12, B1), __fmul_rn(A11, B2)), Dt));
      q.x += delta.x;
      q.y += delta.y;
      np = make_float2(q.x + halfx, q.y + halfy);
      if ((double)delta.x * delta.x + (double)delta.y * delta.y <= eps2) break;
      if (j > 0 && fabs((double)(delta.x + prev_delta.x)) < 0.01 && fabs((double)(delta.y + prev_delta.y)) < 0.01) {
        np.x -= __fmul_rn(delta.x, 0.5f);
        np.y -= __fmul_rn(delta.y, 0.5f);
        break;
      }
      prev_delta = delta;
    }
    if (st && level == 0 && !get_min_eig) {
      const float2 r = make_float2(np.x - halfx, np.y - halfy);
      const int jx = (int)floorf(r.x), jy = (int)floorf(r.y);
      if (jx < -ww || jx >= cols || jy < -wh || jy >= rows) {
        st = false;
        continue;
      }
      lk_weights(r.x - (float)jx, r.y - (float)jy, &iw00, &iw01, &iw10, &iw11);
      const bool in_j = jx >= 0 && jy >= 0 && jx + ww + 1 <= cols && jy + wh + 1 <= rows;
      float ev = 0.f, u1 = 0.f, u2 = 0.f;
      auto abs_err = [&](auto mode) {
        constexpr int MODE = decltype(mode)::value;
#pragma unroll
        for (int s2 = 0; s2 < STEPS; s2++)
          if (s2 < STEPS - 1 || tid + s2 * NT < AREA) {
            const int diff = lk_sample_j<MODE>(J, pitch, cols, rows, jx + wxs[s2], jy + wys[s2], iw00, iw01, iw10, iw11) - Iw[s2];
            ev += fabsf((float)diff);
          }
      };
      if (in_j) abs_err(std::integral_constant<int, 0>());
      else if (!small) abs_err(std::integral_constant<int, 1>());
      else abs_err(std::integral_constant<int, 2>());
      cta_sum3(ev, u1, u2);
      er = __fdiv_rn(ev, (float)(32 * ww * wh));
    }
  }
  if (tid == 0) {
    next_pts[pid] = np;
    status[pid] = st ? 1 : 0;
    err[pid] = er;
  }
}

// ======================================================================================
// host-side drivers
// ======================================================================================
static inline dim3 g2(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

static int upload_image(ofb_handle* h, SparseState* s, int which, const uint8_t* host, int width, int height,
                        size_t stride) {
  if (width < 2 || height < 2) return set_error(h, OFB_ERR_INVALID_ARG, "image must be at least 2x2");
  if (width > s->cap_w || height > s->cap_h || (size_t)width * height > (size_t)s->cap_w * s->cap_h)
    return set_error(h, OFB_ERR_CAPACITY, "image %dx%d exceeds handle capacity %dx%d", width, height, s->cap_w, s->cap_h);
  if (stride == 0) stride = width;
  if (stride < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than width");
  cudaPointerAttributes a;
  bool pinned = cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const uint8_t* from = host;
  size_t from_stride = stride;
  if (!pinned) {
    SP_CUDA(h, cudaStreamSynchronize(h->stream));   // the staging half may still be read by the previous call's copy
    uint8_t* stg = reinterpret_cast<uint8_t*>(s->h_stage) + (size_t)which * width * height;
    // row bands: the DMA of a band runs while the host copies the next one into the staging buffer
    const int bands = height >= 256 ? 4 : 1;
    for (int b = 0; b < bands; b++) {
      const int y0 = (int)((long long)height * b / bands), y1 = (int)((long long)height * (b + 1) / bands);
      if (stride == (size_t)width) memcpy(stg + (size_t)y0 * width, host + (size_t)y0 * width, (size_t)(y1 - y0) * width);
      else for (int y = y0; y < y1; y++) memcpy(stg + (size_t)y * width, host + (size_t)y * stride, width);
      SP_CUDA(h, cudaMemcpyAsync(s->img[which] + (size_t)y0 * width, stg + (size_t)y0 * width, (size_t)(y1 - y0) * width,
                                 cudaMemcpyHostToDevice, h->stream));
    }
    return OFB_OK;
  }
  SP_CUDA(h, cudaMemcpy2DAsync(s->img[which], width, from, from_stride, width, height, cudaMemcpyHostToDevice, h->stream));
  return OFB_OK;
}

// Scharr derivatives of every level of the pyramid `fp` of slot `which`.
static int build_derivs(ofb_handle* h, SparseState* s, int which, FramePyr* fp) {
  dim3 b(32, 8);
  short2* d = s->deriv[which];
  for (int l = 0; l < fp->n_levels; l++) {
    const int lw = fp->w[l], lh = fp->h[l];
    if (lw >= 8 && (lw & 3) == 0 && (reinterpret_cast<uintptr_t>(fp->lv[l]) & 3) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0)
      k_scharr4<<<g2(lw / 4, lh, b), b, 0, h->stream>>>(fp->lv[l], lw, lh, (size_t)lw, d);
    else
      k_scharr<<<g2(lw, lh, b), b, 0, h->stream>>>(fp->lv[l], lw, lh, (size_t)lw, d);
    OFB_LAUNCH_CHECK(h);
    fp->D[l] = d;
    d += (size_t)fp->w[l] * fp->h[l];
  }
  return OFB_OK;
}

// Builds levels 1.. of the pyramid of slot `which` (level 0 = s->img[which]) and, with `derivs`, the Scharr derivatives
// of every level.
static int build_pyr(ofb_handle* h, SparseState* s, int which, int width, int height, int win_w, int win_h,
                     int max_level, bool derivs, FramePyr* fp) {
  fp->lv[0] = s->img[which];
  fp->w[0] = width;
  fp->h[0] = height;
  int n = 1;
  uint8_t* next = s->pyr[which];
  dim3 b(32, 8);
  for (int l = 1; l <= max_level && l < kMaxLkLevels; l++) {
    const int ow = (fp->w[l - 1] + 1) / 2, oh = (fp->h[l - 1] + 1) / 2;
    if (ow <= win_w || oh <= win_h) break;
    const int sw = fp->w[l - 1];
    if (sw >= 16 && fp->h[l - 1] >= 4 && (sw & 3) == 0 && (ow & 3) == 0 && (reinterpret_cast<uintptr_t>(fp->lv[l - 1]) & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(next) & 3) == 0)
      k_pyrdown_u8x4<<<g2(ow / 4, oh, b), b, 0, h->stream>>>(fp->lv[l - 1], sw, fp->h[l - 1], (size_t)sw, next, ow, oh);
    else
      k_pyrdown_u8<<<g2(ow, oh, b), b, 0, h->stream>>>(fp->lv[l - 1], sw, fp->h[l - 1], (size_t)sw, next, ow, oh);
    OFB_LAUNCH_CHECK(h);
    fp->lv[l] = next;
    fp->w[l] = ow;
    fp->h[l] = oh;
    next += (size_t)ow * oh;
    n++;
  }
  fp->n_levels = n;
  for (int l = 0; l < n; l++) fp->D[l] = nullptr;
  return derivs ? build_derivs(h, s, which, fp) : OFB_OK;
}

static int eigen_map(ofb_handle* h, SparseState* s, int which, int width, int height, int block_size,
                     const uint8_t* d_mask, int harris = 0, double harris_k = 0.04, cudaStream_t sm = nullptr) {
  if (!sm) sm = h->stream;
  const double scale = 1.0 / (4.0 * block_size * 255.0);
  const float k0 = (float)(1.0 * scale), k1 = (float)(2.0 * scale), k2 = (float)(1.0 * scale);
  dim3 b(32, 8);
  SP_CUDA(h, cudaMemsetAsync(s->counters, 0, 4 * sizeof(unsigned int), sm));
  // columns past the last full block of 32 take the row filter's scalar tail (no FMA): the same on the AVX2 and the
  // AVX-512 dispatch of the wheel (tests/test_oracle_sparse.py probes both with OPENCV_CPU_DISABLE)
  if (block_size == 3) {
    k_sobel_min_eig3<<<dim3((width + kE3T - 1) / kE3T, (height + kE3T - 1) / kE3T), b, 0, sm>>>(
        s->img[which], width, height, (size_t)width, k0, k1, k2, (width / 32) * 32, s->eig, s->counters + 1, d_mask, harris,
        (float)harris_k, harris_k);
    OFB_LAUNCH_CHECK(h);
    return OFB_OK;
  }
  const int rb = block_size / 2;
  const size_t smem = (size_t)3 * (32 + 2 * rb) * (kEigTileH + 2 * rb) * sizeof(double);   // 14.7 KB at blockSize 3, 68 KB at 31
  if (smem > 48 * 1024)
    SP_CUDA(h, cudaFuncSetAttribute(k_sobel_min_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_sobel_min_eig<<<dim3((width + 31) / 32, (height + kEigTileH - 1) / kEigTileH), b, smem, sm>>>(
      s->img[which], width, height, (size_t)width, k0, k1, k2, (width / 32) * 32, block_size, s->eig, s->counters + 1,
      d_mask, harris, (float)harris_k, harris_k);
  OFB_LAUNCH_CHECK(h);
  return OFB_OK;
}

static int validate_gftt(ofb_handle* h, const ofb_gftt_params* p, int width, int height) {
  if (!(p->quality_level > 0) || p->min_distance < 0)
    return set_error(h, OFB_ERR_INVALID_ARG, "qualityLevel must be > 0 and minDistance >= 0");
  if (p->block_size < 1 || p->block_size % 2 == 0 || p->block_size > 31)
    return set_error(h, OFB_ERR_INVALID_ARG, "blockSize must be odd and in [1,31]");
  if (width > 65535 || height > 65535) return set_error(h, OFB_ERR_INVALID_ARG, "image larger than 65535 px");
  return OFB_OK;
}

// goodFeaturesToTrack of the image in slot `which`; the list lands in s->corners[which], its length in
// s->counters[4 + which] — both on the device, nothing is synchronised.
static int detect_corners(ofb_handle* h, SparseState* s, int which, int width, int height, const ofb_gftt_params* p,
                          const uint8_t* d_mask, cudaStream_t sm = nullptr) {
  if (!sm) sm = h->stream;
  int st = eigen_map(h, s, which, width, height, p->block_size, d_mask, p->use_harris_detector, p->harris_k, sm);
  if (st) return st;
  SP_CUDA(h, cudaMemsetAsync(s->hist, 0, kBuckets * sizeof(unsigned int), sm));
  dim3 b(32, 8);
  k_candidates<<<dim3((width - 2 + kCandTile - 1) / kCandTile, (height - 2 + kCandTile - 1) / kCandTile), b, 0, sm>>>(s->eig, width, height, p->quality_level, s->counters + 1,
                                                          s->keys, s->counters, (unsigned int)s->cand_cap, s->hist, d_mask);
  OFB_LAUNCH_CHECK(h);
  k_bucket_sums<<<kBuckets / 1024, 1024, 0, sm>>>(s->hist, s->bpart);
  OFB_LAUNCH_CHECK(h);
  k_bucket_scan<<<kBuckets / 1024, 1024, 0, sm>>>(s->hist, s->bstart, s->bpart);
  OFB_LAUNCH_CHECK(h);
  k_bucket_scatter<<<2 * h->num_sms, 256, 0, sm>>>(s->keys, s->counters, (unsigned int)s->cand_cap, p->quality_level,
                                                   s->counters + 1, s->hist, s->bstart, s->keys + s->cand_cap);
  OFB_LAUNCH_CHECK(h);
  const int cell = p->min_distance >= 1 ? (int)__builtin_nearbyint(p->min_distance) : 1;
  const size_t cells = (size_t)((width + cell - 1) / cell) * ((height + cell - 1) / cell);
  if (p->min_distance >= 1) SP_CUDA(h, cudaMemsetAsync(s->grid_cnt, 0, cells * sizeof(unsigned int), sm));
  k_greedy_select<<<1, GS_THREADS, 0, sm>>>(s->keys + s->cand_cap, s->bstart, s->sortbuf, width, height,
                                            (float)p->min_distance, p->max_corners, (unsigned int)s->cand_cap, s->grid_cnt,
                                            s->grid_pts, s->corners[which], s->counters + 4 + which);
  OFB_LAUNCH_CHECK(h);
  return OFB_OK;
}

static int upload_mask(ofb_handle* h, SparseState* s, const uint8_t* mask, int width, int height, size_t stride,
                       const uint8_t** d_mask) {
  *d_mask = nullptr;
  if (!mask) return OFB_OK;
  if (stride == 0) stride = width;
  if (stride < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "mask stride smaller than width");
  if (!s->mask) SP_CUDA(h, cudaMalloc(&s->mask, (size_t)s->cap_w * s->cap_h));
  SP_CUDA(h, cudaMemcpy2DAsync(s->mask, width, mask, stride, width, height, cudaMemcpyHostToDevice, h->stream));
  *d_mask = s->mask;
  return OFB_OK;
}

static int lk_launch(ofb_handle* h, SparseState* s, const FramePyr& I, const FramePyr& J, int width, const float2* d_prev,
                     int n_bound, const unsigned int* n_dev, const ofb_lk_params* p) {
  LkLevels lv;
  lv.n_levels = I.n_levels;
  lv.pitch0 = (size_t)width;
  for (int l = 0; l < I.n_levels; l++) {
    lv.I[l] = I.lv[l]; lv.J[l] = J.lv[l]; lv.D[l] = I.D[l]; lv.w[l] = I.w[l]; lv.h[l] = I.h[l];
  }
  const int max_count = std::min(std::max(p->max_count, 0), 100);
  double eps = std::min(std::max(p->epsilon, 0.0), 10.0);
  eps *= eps;
  const size_t smem = (size_t)p->win_w * p->win_h * 3 * sizeof(short) * LK_WARPS;
  if (n_bound <= 0) return OFB_OK;
  const dim3 grid((n_bound + LK_WARPS - 1) / LK_WARPS);
  if (p->win_w == 21 && p->win_h == 21) {        // cv2's default window: one point per CTA
    k_lk_track_cta<21, 21, OFB_LK_WPP><<<n_bound, OFB_LK_WPP * 32, 0, h->stream>>>(
        lv, d_prev, s->pts_next, s->lk_status, s->lk_err, n_bound, n_dev, max_count, eps, p->flags, p->min_eig_threshold);
  } else if (p->win_w == 15 && p->win_h == 15) {
    k_lk_track_cta<15, 15, 2><<<n_bound, 64, 0, h->stream>>>(lv, d_prev, s->pts_next, s->lk_status, s->lk_err, n_bound, n_dev,
                                                              max_count, eps, p->flags, p->min_eig_threshold);
  } else {
    if (smem > 48 * 1024)
      OFB_CUDA(h, cudaFuncSetAttribute(k_lk_track<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_lk_track<0, 0><<<grid, LK_WARPS * 32, smem, h->stream>>>(lv, d_prev, s->pts_next, s->lk_status, s->lk_err, n_bound,
                                                               n_dev, p->win_w, p->win_h, max_count, eps, p->flags,
                                                               p->min_eig_threshold);
  }
  OFB_LAUNCH_CHECK(h);
  return OFB_OK;
}

static int validate_lk(ofb_handle* h, const ofb_lk_params* p) {
  if (p->win_w <= 2 || p->win_h <= 2) return set_error(h, OFB_ERR_INVALID_ARG, "winSize must be > 2x2");
  if (p->max_level < 0) return set_error(h, OFB_ERR_INVALID_ARG, "maxLevel must be >= 0");
  if ((size_t)p->win_w * p->win_h * 3 * sizeof(short) * LK_WARPS > 200 * 1024)
    return set_error(h, OFB_ERR_INVALID_ARG, "winSize too large");
  return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_corner_min_eigenval(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                            int block_size, float* eig_out) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!image || !eig_out) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (block_size < 1 || block_size % 2 == 0 || block_size > 31)
    return set_error(h, OFB_ERR_INVALID_ARG, "blockSize must be odd and in [1,31]");
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  int st = sparse_get(h, &s);
  if (st) return st;
  s->st.primed = false;
  if ((st = upload_image(h, s, 0, image, width, height, stride_bytes))) return st;
  if ((st = eigen_map(h, s, 0, width, height, block_size, nullptr))) return st;
  OFB_CUDA(h, cudaMemcpyAsync(eig_out, s->eig, (size_t)width * height * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_good_features_masked(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                             const uint8_t* mask, size_t mask_stride_bytes, const ofb_gftt_params* p,
                             float* corners_xy, int* n_out) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!image || !p || !corners_xy || !n_out) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  int st = validate_gftt(h, p, width, height);
  if (st) return st;
  if (width < 3 || height < 3) {     // no interior pixel: cv2 returns an empty list
    *n_out = 0;
    return OFB_OK;
  }
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  if ((st = sparse_get(h, &s))) return st;
  s->st.primed = false;
  if ((st = upload_image(h, s, 0, image, width, height, stride_bytes))) return st;
  const uint8_t* d_mask;
  if ((st = upload_mask(h, s, mask, width, height, mask_stride_bytes, &d_mask))) return st;
  if ((st = detect_corners(h, s, 0, width, height, p, d_mask))) return st;
  cudaStream_t sm = h->stream;
  unsigned int* hc = reinterpret_cast<unsigned int*>(s->h_stage);
  float* hp = reinterpret_cast<float*>(reinterpret_cast<char*>(s->h_stage) + 64);
  OFB_CUDA(h, cudaMemcpyAsync(hc, s->counters + 4, sizeof(unsigned int), cudaMemcpyDeviceToHost, sm));
  // a bounded list comes back with its count in ONE synchronisation; an unbounded one (maxCorners <= 0) needs the count first
  const size_t bound = p->max_corners > 0 ? std::min<size_t>((size_t)p->max_corners, s->cand_cap) : 0;
  if (bound) OFB_CUDA(h, cudaMemcpyAsync(hp, s->corners[0], bound * sizeof(float2), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  const unsigned int n = hc[0];
  if (n && !bound) {
    OFB_CUDA(h, cudaMemcpyAsync(hp, s->corners[0], (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, sm));
    OFB_CUDA(h, cudaStreamSynchronize(sm));
  }
  if (n) memcpy(corners_xy, hp, (size_t)n * sizeof(float2));
  *n_out = (int)n;
  return OFB_OK;
}

int ofb_good_features(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                      const ofb_gftt_params* p, float* corners_xy, int* n_out) {
  return ofb_good_features_masked(h, image, width, height, stride_bytes, nullptr, 0, p, corners_xy, n_out);
}

int ofb_lk_pyramid(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes, int win_w,
                   int win_h, int max_level, uint8_t* const* level_out, int16_t* const* deriv_out, int* n_levels_out) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!image || !n_levels_out) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (win_w < 3 || win_h < 3 || max_level < 0) return set_error(h, OFB_ERR_INVALID_ARG, "bad winSize / maxLevel");
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  int st = sparse_get(h, &s);
  if (st) return st;
  s->st.primed = false;
  if ((st = upload_image(h, s, 0, image, width, height, stride_bytes))) return st;
  FramePyr fp;
  if ((st = build_pyr(h, s, 0, width, height, win_w, win_h, max_level, deriv_out != nullptr, &fp))) return st;
  for (int l = 0; l < fp.n_levels; l++) {
    if (level_out && level_out[l])
      OFB_CUDA(h, cudaMemcpyAsync(level_out[l], fp.lv[l], (size_t)fp.w[l] * fp.h[l], cudaMemcpyDeviceToHost, h->stream));
    if (deriv_out && deriv_out[l])
      OFB_CUDA(h, cudaMemcpyAsync(deriv_out[l], fp.D[l], (size_t)fp.w[l] * fp.h[l] * sizeof(short2), cudaMemcpyDeviceToHost,
                                  h->stream));
  }
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_levels_out = fp.n_levels;
  return OFB_OK;
}

int ofb_pyrlk(ofb_handle* h, const uint8_t* prev, const uint8_t* next, int width, int height, size_t stride_bytes,
              const float* prev_pts, int n_points, float* next_pts, uint8_t* status, float* err,
              const ofb_lk_params* p) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!prev || !next || !p || !status || !next_pts || (!prev_pts && n_points > 0))
    return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (n_points < 0) return set_error(h, OFB_ERR_INVALID_ARG, "negative point count");
  int st = validate_lk(h, p);
  if (st) return st;
  if (n_points == 0) return OFB_OK;
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  if ((st = sparse_get(h, &s))) return st;
  s->st.primed = false;
  if ((st = sparse_points(h, s, n_points))) return st;
  lk_out_layout(s, n_points);
  if ((st = upload_image(h, s, 0, prev, width, height, stride_bytes))) return st;
  if ((st = upload_image(h, s, 1, next, width, height, stride_bytes))) return st;
  cudaStream_t sm = h->stream;
  FramePyr fi, fj;
  if ((st = build_pyr(h, s, 0, width, height, p->win_w, p->win_h, p->max_level, true, &fi))) return st;
  if ((st = build_pyr(h, s, 1, width, height, p->win_w, p->win_h, p->max_level, false, &fj))) return st;
  OFB_CUDA(h, cudaMemcpyAsync(s->pts_prev, prev_pts, (size_t)n_points * sizeof(float2), cudaMemcpyHostToDevice, sm));
  if (p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)
    OFB_CUDA(h, cudaMemcpyAsync(s->pts_next, next_pts, (size_t)n_points * sizeof(float2), cudaMemcpyHostToDevice, sm));
  if ((st = lk_launch(h, s, fi, fj, width, s->pts_prev, n_points, nullptr, p))) return st;
  OFB_CUDA(h, cudaMemcpyAsync(next_pts, s->pts_next, (size_t)n_points * sizeof(float2), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaMemcpyAsync(status, s->lk_status, (size_t)n_points, cudaMemcpyDeviceToHost, sm));
  if (err) OFB_CUDA(h, cudaMemcpyAsync(err, s->lk_err, (size_t)n_points * sizeof(float), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  return OFB_OK;
}

// Camera-stream form of the sparse path: one new frame per call, the temporal state stays on the GPU.
//   call t:  upload frame t (ONE copy) -> pyramid of t -> track the corners of frame t-1 (their list, the pyramid and the
//            Scharr derivatives of t-1 are on the device) into t -> Scharr derivatives of t -> corners of t for the next
//            call -> ONE download (counts, tracked points, status, error, the new corner list) -> ONE synchronisation.
// Same kernels and arguments as ofb_good_features(t-1) followed by ofb_pyrlk(t-1, t, corners): identical bits.
int ofb_lk_stream(ofb_handle* h, const uint8_t* frame, int width, int height, size_t stride_bytes,
                  const ofb_gftt_params* gp, const ofb_lk_params* lp, float* prev_pts, float* next_pts, uint8_t* status,
                  float* err, int* n_tracked, float* new_corners, int* n_new) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!frame || !gp || !lp || !n_tracked) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  *n_tracked = -1;
  if (n_new) *n_new = 0;
  int st = validate_gftt(h, gp, width, height);
  if (st) return st;
  if ((st = validate_lk(h, lp))) return st;
  if (gp->max_corners <= 0) return set_error(h, OFB_ERR_INVALID_ARG, "the stream call needs maxCorners > 0 (fixed-size result buffers)");
  if (lp->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)
    return set_error(h, OFB_ERR_INVALID_ARG, "USE_INITIAL_FLOW is not supported by the stream call");
  if (width < 3 || height < 3) return set_error(h, OFB_ERR_INVALID_ARG, "image must be at least 3x3");
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  if ((st = sparse_get(h, &s))) return st;
  const int bound = (int)std::min<size_t>((size_t)gp->max_corners, s->cand_cap);
  if ((st = sparse_points(h, s, bound))) return st;
  SparseState::Stream& S = s->st;
  const bool same = S.primed && S.w == width && S.h == height && memcmp(&S.gp, gp, sizeof(*gp)) == 0 &&
                    memcmp(&S.lp, lp, sizeof(*lp)) == 0;
  if (!same) { S.primed = false; stream_graphs_drop(s); }
  const int cur = S.primed ? (S.cur ^ 1) : 0, prv = cur ^ 1;
  cudaStream_t sm = h->stream;
  if ((st = upload_image(h, s, cur, frame, width, height, stride_bytes))) { S.primed = false; return st; }
  // result block in pinned memory: [count prev, count cur | next pts | err | status | new corners]; the tracker's three
  // arrays have the layout of the device block (lk_out_layout) and come back in one copy
  char* hb = reinterpret_cast<char*>(s->h_stage);
  const size_t lk_bytes = (size_t)bound * 13, new_off = 64 + ((lk_bytes + 15) & ~(size_t)15);
  const size_t need = new_off + (size_t)bound * 8;
  if (need > s->h_stage_bytes) { S.primed = false; return set_error(h, OFB_ERR_CAPACITY, "maxCorners exceeds the staging capacity"); }
  unsigned int* hc = reinterpret_cast<unsigned int*>(hb);
  float* h_next = reinterpret_cast<float*>(hb + 64);
  float* h_err = h_next + 2 * (size_t)bound;
  uint8_t* h_status = reinterpret_cast<uint8_t*>(h_err + bound);
  float* h_new = reinterpret_cast<float*>(hb + new_off);
  lk_out_layout(s, bound);
  const bool track = S.primed;
  // Two chains behind the upload: tracker (pyramid of the new frame -> LK from the previous frame's corners) on the
  // handle's stream, corner detection of the new frame (needs level 0 only) on the second stream.  They share no buffer:
  // the tracker reads slot prv's corners / derivatives and slot cur's levels, the detector writes slot cur's corners and
  // the detection scratch.  The new frame's Scharr derivatives — needed by the NEXT call, in which it is the previous
  // frame — go behind the tracker.
  auto enqueue = [&]() -> int {
    int st2;
    cudaStream_t sdet = track ? s->aux : sm;
    if (track) {
      OFB_CUDA(h, cudaEventRecord(s->ev_up, sm));
      OFB_CUDA(h, cudaStreamWaitEvent(s->aux, s->ev_up, 0));
    }
    if ((st2 = build_pyr(h, s, cur, width, height, lp->win_w, lp->win_h, lp->max_level, false, &S.fp[cur]))) return st2;
    if (track) {
      if ((st2 = lk_launch(h, s, S.fp[prv], S.fp[cur], width, s->corners[prv], bound, s->counters + 4 + prv, lp))) return st2;
      OFB_CUDA(h, cudaMemcpyAsync(h_next, s->lk_out, lk_bytes, cudaMemcpyDeviceToHost, sm));
    }
    if ((st2 = detect_corners(h, s, cur, width, height, gp, nullptr, sdet))) return st2;
    OFB_CUDA(h, cudaMemcpyAsync(hc, s->counters + 4, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, sdet));
    OFB_CUDA(h, cudaMemcpyAsync(h_new, s->corners[cur], (size_t)bound * sizeof(float2), cudaMemcpyDeviceToHost, sdet));
    if ((st2 = build_derivs(h, s, cur, &S.fp[cur]))) return st2;
    if (track) {
      OFB_CUDA(h, cudaEventRecord(s->ev_aux, s->aux));
      OFB_CUDA(h, cudaStreamWaitEvent(sm, s->ev_aux, 0));
    }
    return OFB_OK;
  };
  if (S.graph_out != (const void*)s->lk_out || S.graph_bound != bound) {   // (the result buffers moved or changed layout)
    stream_graphs_drop(s);
    S.graph_out = s->lk_out;
    S.graph_bound = bound;
  }
  if (!track || h->no_graph || h->timing) {
    if ((st = enqueue())) { S.primed = false; return st; }
  } else if (S.graph[cur]) {
    OFB_CUDA(h, cudaGraphLaunch(S.graph[cur], sm));
    h->launches += S.graph_launches[cur];
  } else {
    const uint64_t l0 = h->launches;
    OFB_CUDA(h, cudaStreamBeginCapture(sm, cudaStreamCaptureModeThreadLocal));
    st = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(sm, &graph);
    if (st != OFB_OK || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      S.primed = false;
      return st != OFB_OK ? st : set_error(h, OFB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    }
    S.graph_launches[cur] = h->launches - l0;
    const cudaError_t ie = cudaGraphInstantiate(&S.graph[cur], graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      S.graph[cur] = nullptr;
      S.primed = false;
      return set_error(h, OFB_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
    }
    OFB_CUDA(h, cudaGraphLaunch(S.graph[cur], sm));
  }
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  const int n_cur = (int)std::min<unsigned int>(hc[cur], (unsigned int)bound);
  if (track) {
    const int n_prev = S.n_host[prv];
    if (prev_pts) memcpy(prev_pts, S.host_corners[prv].data(), (size_t)n_prev * sizeof(float2));
    if (next_pts) memcpy(next_pts, h_next, (size_t)n_prev * sizeof(float2));
    if (status) memcpy(status, h_status, (size_t)n_prev);
    if (err) memcpy(err, h_err, (size_t)n_prev * sizeof(float));
    *n_tracked = n_prev;
  }
  S.host_corners[cur].assign(h_new, h_new + 2 * (size_t)n_cur);
  S.n_host[cur] = n_cur;
  if (new_corners) memcpy(new_corners, h_new, (size_t)n_cur * sizeof(float2));
  if (n_new) *n_new = n_cur;
  S.cur = cur;
  S.w = width; S.h = height; S.gp = *gp; S.lp = *lp;
  S.primed = true;
  return OFB_OK;
}

int ofb_lk_stream_reset(ofb_handle* h) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (h->sparse) reinterpret_cast<SparseState*>(h->sparse)->st.primed = false;
  return OFB_OK;
}

}  // extern "C"
