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
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "common.cuh"
#include "fb_device.cuh"

namespace ofb {

constexpr int kMaxLkLevels = 8;
constexpr int kGridSlots = 4;

struct SparseState {
  int cap_w = 0, cap_h = 0;
  uint8_t* img[2] = {nullptr, nullptr};        // staged level-0 images (prev, next), packed pitch = width
  uint8_t* pyr[2] = {nullptr, nullptr};        // levels 1.. of both pyramids, packed one after another
  short2* deriv = nullptr;                     // Scharr (dx,dy) of every level of `prev`
  float* cov = nullptr;                        // 3 planes
  float* eig = nullptr;
  unsigned long long* keys = nullptr;          // 2 x cand_cap (radix sort in/out)
  size_t cand_cap = 0;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  unsigned int* counters = nullptr;            // [0] candidate count, [1] max(eig) bits, [2] corner count
  unsigned int* grid_cnt = nullptr;            // per cell
  ushort2* grid_pts = nullptr;                 // per cell x kGridSlots
  float2* corners = nullptr;                   // accepted corners (device)
  float2* pts_prev = nullptr;                  // LK inputs / outputs (device)
  float2* pts_next = nullptr;
  uint8_t* lk_status = nullptr;
  float* lk_err = nullptr;
  int pts_cap = 0;
  // pinned host staging
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
};

static void sparse_free(SparseState* s) {
  if (!s) return;
  for (int i = 0; i < 2; i++) { cudaFree(s->img[i]); cudaFree(s->pyr[i]); }
  cudaFree(s->deriv); cudaFree(s->cov); cudaFree(s->eig); cudaFree(s->keys); cudaFree(s->cub_tmp);
  cudaFree(s->counters); cudaFree(s->grid_cnt); cudaFree(s->grid_pts); cudaFree(s->corners);
  cudaFree(s->pts_prev); cudaFree(s->pts_next); cudaFree(s->lk_status); cudaFree(s->lk_err);
  if (s->h_stage) cudaFreeHost(s->h_stage);
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
  SP_CUDA(h, cudaMalloc(&s->deriv, 2 * N * sizeof(short2)));
  SP_CUDA(h, cudaMalloc(&s->cov, 3 * N * sizeof(float)));
  SP_CUDA(h, cudaMalloc(&s->eig, N * sizeof(float)));
  s->cand_cap = N / 2 + 1024;
  SP_CUDA(h, cudaMalloc(&s->keys, 2 * s->cand_cap * sizeof(unsigned long long)));
  size_t tmp = 0;
  cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp, s->keys, s->keys + s->cand_cap, (int)s->cand_cap);
  s->cub_tmp_bytes = tmp;
  SP_CUDA(h, cudaMalloc(&s->cub_tmp, tmp));
  SP_CUDA(h, cudaMalloc(&s->counters, 16 * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->grid_cnt, N * sizeof(unsigned int)));
  SP_CUDA(h, cudaMalloc(&s->grid_pts, N * kGridSlots * sizeof(ushort2)));
  // the selection can keep every candidate (minDistance < 1, maxCorners <= 0): corner buffers sized like the candidate list
  SP_CUDA(h, cudaMalloc(&s->corners, s->cand_cap * sizeof(float2)));
  s->h_stage_bytes = std::max<size_t>(2 * N, s->cand_cap * sizeof(float2));
  SP_CUDA(h, cudaHostAlloc(&s->h_stage, s->h_stage_bytes, cudaHostAllocDefault));
  *out = s;
  return OFB_OK;
}

static int sparse_points(ofb_handle* h, SparseState* s, int n) {
  if (n <= s->pts_cap) return OFB_OK;
  cudaFree(s->pts_prev); cudaFree(s->pts_next); cudaFree(s->lk_status); cudaFree(s->lk_err);
  s->pts_prev = s->pts_next = nullptr; s->lk_status = nullptr; s->lk_err = nullptr;
  s->pts_cap = 0;
  const int cap = std::max(n, 4096);
  SP_CUDA(h, cudaMalloc(&s->pts_prev, cap * sizeof(float2)));
  SP_CUDA(h, cudaMalloc(&s->pts_next, cap * sizeof(float2)));
  SP_CUDA(h, cudaMalloc(&s->lk_status, cap));
  SP_CUDA(h, cudaMalloc(&s->lk_err, cap * sizeof(float)));
  s->pts_cap = cap;
  return OFB_OK;
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
// ======================================================================================
// Sobel with the scale folded into the smoothing kernel k = f32([1,2,1]/(4*blockSize*255)):
//   dx = fma(r[y-1] + r[y+1], k0, r[y]*k1),  r = p[x+1] - p[x-1]
//   dy = rw[y+1] - rw[y-1],  rw = fma(k2, p[x+1], fma(k1, p[x], k0*p[x-1]))   (vector body)
//        rw = (p[x-1]*k0 + p[x]*k1) + p[x+1]*k2   for columns past the last full block of 32 (SIMD tail)
__global__ void __launch_bounds__(256) k_sobel_cov(const uint8_t* __restrict__ src, int w, int h, size_t spitch,
                                                   float* __restrict__ cov, float k0, float k1, float k2, int wb) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
  const uint8_t* rows[3] = {src + (size_t)reflect101(y - 1, h) * spitch, src + (size_t)y * spitch,
                            src + (size_t)reflect101(y + 1, h) * spitch};
  float r[3], rw[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const float pm = (float)rows[j][xm], pc = (float)rows[j][x], pp = (float)rows[j][xp];
    r[j] = pp - pm;
    if (x < wb) rw[j] = fmaf(k2, pp, fmaf(k1, pc, __fmul_rn(k0, pm)));
    else rw[j] = __fadd_rn(__fadd_rn(__fmul_rn(pm, k0), __fmul_rn(pc, k1)), __fmul_rn(pp, k2));
  }
  const float dx = fmaf(r[0] + r[2], k0, __fmul_rn(r[1], k1));
  const float dy = rw[2] - rw[0];
  const size_t n = (size_t)w * h, o = (size_t)y * w + x;
  cov[o] = __fmul_rn(dx, dx);
  cov[n + o] = __fmul_rn(dx, dy);
  cov[2 * n + o] = __fmul_rn(dy, dy);
}

// unnormalised blockSize^2 box sum in double (exact for these magnitudes), then
// eig = (a + c) - sqrt((a - c)^2 + b^2), a = Sxx/2, c = Syy/2 — plain mul/add, no FMA.
__global__ void __launch_bounds__(256) k_min_eig(const float* __restrict__ cov, int w, int h, int block_size,
                                                 float* __restrict__ eig, unsigned int* __restrict__ max_bits) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  float e = 0.f;
  if (x < w && y < h) {
    const size_t n = (size_t)w * h;
    const int rb = block_size / 2;
    double sxx = 0, sxy = 0, syy = 0;
    for (int j = -rb; j <= rb; j++) {
      const size_t ro = (size_t)reflect101(y + j, h) * w;
      for (int i = -rb; i <= rb; i++) {
        const size_t o = ro + reflect101(x + i, w);
        sxx += (double)cov[o];
        sxy += (double)cov[n + o];
        syy += (double)cov[2 * n + o];
      }
    }
    const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
    const float d = __fsub_rn(a, c);
    e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
    eig[(size_t)y * w + x] = e;
  }
  // max over the image (minMaxLoc); the map is >= 0 up to rounding, a negative max means "no corners"
  float m = fmaxf(e, 0.f);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));
}

// threshold (THRESH_TOZERO at max*quality, strict >) + 3x3 local maximum, 1-px frame skipped.
// Candidates are packed as (value bits << 32 | linear index): descending sort = value descending,
// ties by DESCENDING address — cv2's greaterThanPtr.
__global__ void __launch_bounds__(256) k_candidates(const float* __restrict__ eig, int w, int h, double quality,
                                                    const unsigned int* __restrict__ max_bits,
                                                    unsigned long long* __restrict__ keys, unsigned int* __restrict__ count,
                                                    unsigned int cap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int y = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (x >= w - 1 || y >= h - 1) return;
  const float thr = (float)((double)__uint_as_float(*max_bits) * quality);
  const float v = eig[(size_t)y * w + x];
  if (!(v > thr)) return;
  float mx = v;
#pragma unroll
  for (int j = -1; j <= 1; j++)
#pragma unroll
    for (int i = -1; i <= 1; i++) mx = fmaxf(mx, eig[(size_t)(y + j) * w + x + i]);
  if (v != mx) return;
  const unsigned int slot = atomicAdd(count, 1u);
  if (slot < cap) keys[slot] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(y * w + x);
}

// Greedy minimum-distance selection in sorted order (lexicographically-first maximal independent set), exactly
// cv2's sequential rule, in rounds of 1024 candidates by one CTA:
//   phase 1 (1024 threads): every candidate of the round is tested against the corners accepted in EARLIER rounds
//            (grid of cell = cvRound(minDistance), +-1 cell); the survivors are compacted in rank order;
//   phase 2 (1024 threads): the survivors decide among themselves by the parallel form of the same rule (see below)
//            and the accepted ones are committed in rank order, up to the corner limit.
constexpr int GS_THREADS = 1024;
constexpr int GS_BUCKETS = 4096;

__device__ __forceinline__ bool gs_far_from_accepted(int x, int y, int cell, int gw, int gh, float md2,
                                                     const unsigned int* grid_cnt, const ushort2* grid_pts) {
  const int xc = x / cell, yc = y / cell;
  for (int yy = max(0, yc - 1); yy <= min(gh - 1, yc + 1); yy++)
    for (int xx = max(0, xc - 1); xx <= min(gw - 1, xc + 1); xx++) {
      const int c = yy * gw + xx;
      const unsigned int cnt = min(grid_cnt[c], (unsigned int)kGridSlots);   // (one CTA: barrier/fence-ordered, L1-coherent)
      for (unsigned int k = 0; k < cnt; k++) {
        const ushort2 p = grid_pts[c * kGridSlots + k];
        const float dx = (float)(x - (int)p.x), dy = (float)(y - (int)p.y);
        if (dx * dx + dy * dy < md2) return false;
      }
    }
  return true;
}

__global__ void __launch_bounds__(GS_THREADS) k_greedy_select(const unsigned long long* __restrict__ keys,
                                                              const unsigned int* __restrict__ count_ptr, unsigned int cap,
                                                              int w, int h, float min_dist, int max_corners,
                                                              unsigned int* grid_cnt, ushort2* grid_pts,
                                                              float2* __restrict__ corners, unsigned int* __restrict__ n_out) {
  __shared__ unsigned int surv[GS_THREADS];          // x | y << 16, rank order
  __shared__ unsigned int warp_cnt[GS_THREADS / 32];
  __shared__ unsigned int s_nsurv, s_accepted;
  __shared__ unsigned int bucket[GS_BUCKETS];        // head of the chain of in-round survivors per hashed grid cell
  __shared__ unsigned short chain[GS_THREADS];
  __shared__ unsigned char state[GS_THREADS];        // 0 undecided, 1 accepted, 2 rejected
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned int total = min(*count_ptr, cap);
  const unsigned int limit = max_corners > 0 ? (unsigned int)max_corners : 0xffffffffu;
  if (min_dist < 1.f) {
    const unsigned int n = min(total, limit);
    for (unsigned int i = tid; i < n; i += GS_THREADS) {
      const unsigned int idx = (unsigned int)(keys[i] & 0xffffffffu);
      corners[i] = make_float2((float)(idx % w), (float)(idx / w));
    }
    if (tid == 0) *n_out = n;
    return;
  }
  const int cell = __float2int_rn(min_dist);
  const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
  const float md2 = min_dist * min_dist;
  if (tid == 0) s_accepted = 0;
  __syncthreads();
  for (unsigned int base = 0; base < total; base += GS_THREADS) {
    if (s_accepted >= limit) break;                  // (uniform: read after the barrier that ends a round)
    // ---- phase 1
    const unsigned int i = base + tid;
    bool alive = i < total;
    int x = 0, y = 0;
    if (alive) {
      const unsigned int idx = (unsigned int)(keys[i] & 0xffffffffu);
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
      // chain inserts in rank order by one warp per 32 survivors is not needed: chains are unordered, the rank test is
      // explicit (j < i).  atomicExch on 32-bit words holding the 16-bit index.
      if ((unsigned int)tid < ns) {
        const unsigned int bk = ((unsigned int)(sy / cell) * 73u + (unsigned int)(sx / cell)) & (GS_BUCKETS - 1);
        chain[tid] = (unsigned short)atomicExch(&bucket[bk], (unsigned int)tid);
      }
      __syncthreads();
      for (;;) {
        bool undecided = false;
        if ((unsigned int)tid < ns && state[tid] == 0) {
          bool blocked = false, rejected = false;
          const int xc = sx / cell, yc = sy / cell;
          for (int yy = max(0, yc - 1); yy <= min(gh - 1, yc + 1) && !rejected; yy++)
            for (int xx = max(0, xc - 1); xx <= min(gw - 1, xc + 1) && !rejected; xx++) {
              unsigned int k = bucket[((unsigned int)yy * 73u + (unsigned int)xx) & (GS_BUCKETS - 1)];
              while (k != 0xffffu) {
                if (k < (unsigned int)tid) {
                  const unsigned int p = surv[k];
                  const float dx = (float)(sx - (int)(p & 0xffffu)), dy = (float)(sy - (int)(p >> 16));
                  if (dx * dx + dy * dy < md2) {
                    const unsigned char sk = *(volatile unsigned char*)&state[k];
                    if (sk == 1) { rejected = true; break; }
                    if (sk == 0) blocked = true;
                  }
                }
                k = chain[k];
              }
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
      const unsigned int pos = s_accepted + before + __popc(bal2 & ((1u << lane) - 1u));
      if (acc && pos < limit) {
        corners[pos] = make_float2((float)sx, (float)sy);
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

__global__ void __launch_bounds__(LK_WARPS * 32) k_lk_track(LkLevels lv, const float2* __restrict__ prev_pts,
                                                            float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                            float* __restrict__ err, int n_points, int ww, int wh,
                                                            int max_count, double eps2, int flags, double min_eig_thr) {
  extern __shared__ short lk_smem[];  // per warp: Iwin[area], dIx[area], dIy[area]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = blockIdx.x * LK_WARPS + warp;
  if (pid >= n_points) return;
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
    float a11 = 0.f, a12 = 0.f, a22 = 0.f;
    __syncwarp();
    for (int e = lane; e < area; e += 32) {
      const int wy = e / ww, wx = e - wy * ww;
      const int x = ipx + wx, y = ipy + wy;
      const int ival = lk_sample_u8(I, pitch, cols, rows, x, y, iw00, iw01, iw10, iw11, in_i);
      const short2 d00 = lk_deriv_at(D, cols, rows, x, y), d01 = lk_deriv_at(D, cols, rows, x + 1, y);
      const short2 d10 = lk_deriv_at(D, cols, rows, x, y + 1), d11 = lk_deriv_at(D, cols, rows, x + 1, y + 1);
      const int ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
      const int iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
      Iwin[e] = (short)ival;
      dIx[e] = (short)ix;
      dIy[e] = (short)iy;
      a11 += (float)(ix * ix);
      a12 += (float)(ix * iy);
      a22 += (float)(iy * iy);
    }
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
      float b1 = 0.f, b2 = 0.f;
      for (int e = lane; e < area; e += 32) {
        const int wy = e / ww, wx = e - wy * ww;
        const int diff = lk_sample_u8(J, pitch, cols, rows, jx + wx, jy + wy, iw00, iw01, iw10, iw11, in_j) - (int)Iwin[e];
        b1 += (float)(diff * (int)dIx[e]);
        b2 += (float)(diff * (int)dIy[e]);
      }
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
      float ev = 0.f;
      for (int e = lane; e < area; e += 32) {
        const int wy = e / ww, wx = e - wy * ww;
        const int diff = lk_sample_u8(J, pitch, cols, rows, jx + wx, jy + wy, iw00, iw01, iw10, iw11, in_j) - (int)Iwin[e];
        ev += fabsf((float)diff);
      }
      er = __fdiv_rn(warp_sum(ev), (float)(32 * ww * wh));
    }
  }
  if (lane == 0) {
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
    uint8_t* stg = reinterpret_cast<uint8_t*>(s->h_stage) + (size_t)which * width * height;
    for (int y = 0; y < height; y++) memcpy(stg + (size_t)y * width, host + (size_t)y * stride, width);
    from = stg;
    from_stride = width;
  }
  SP_CUDA(h, cudaMemcpy2DAsync(s->img[which], width, from, from_stride, width, height, cudaMemcpyHostToDevice, h->stream));
  return OFB_OK;
}

// Builds levels 1.. of pyramid `which` (level 0 = s->img[which]); fills sizes; returns number of levels.
static int build_pyr(ofb_handle* h, SparseState* s, int which, int width, int height, int win_w, int win_h,
                     int max_level, const uint8_t** lv_ptr, int* lv_w, int* lv_h, int* n_levels) {
  lv_ptr[0] = s->img[which];
  lv_w[0] = width;
  lv_h[0] = height;
  int n = 1;
  uint8_t* next = s->pyr[which];
  for (int l = 1; l <= max_level && l < kMaxLkLevels; l++) {
    const int ow = (lv_w[l - 1] + 1) / 2, oh = (lv_h[l - 1] + 1) / 2;
    if (ow <= win_w || oh <= win_h) break;
    dim3 b(32, 8);
    k_pyrdown_u8<<<g2(ow, oh, b), b, 0, h->stream>>>(lv_ptr[l - 1], lv_w[l - 1], lv_h[l - 1], (size_t)lv_w[l - 1], next,
                                                     ow, oh);
    OFB_LAUNCH_CHECK(h);
    lv_ptr[l] = next;
    lv_w[l] = ow;
    lv_h[l] = oh;
    next += (size_t)ow * oh;
    n++;
  }
  *n_levels = n;
  return OFB_OK;
}

static int eigen_map(ofb_handle* h, SparseState* s, int width, int height, int block_size) {
  const double scale = 1.0 / (4.0 * block_size * 255.0);
  const float k0 = (float)(1.0 * scale), k1 = (float)(2.0 * scale), k2 = (float)(1.0 * scale);
  dim3 b(32, 8);
  SP_CUDA(h, cudaMemsetAsync(s->counters, 0, 16 * sizeof(unsigned int), h->stream));
  k_sobel_cov<<<g2(width, height, b), b, 0, h->stream>>>(s->img[0], width, height, (size_t)width, s->cov, k0, k1, k2,
                                                         (width / 32) * 32);
  OFB_LAUNCH_CHECK(h);
  k_min_eig<<<g2(width, height, b), b, 0, h->stream>>>(s->cov, width, height, block_size, s->eig, s->counters + 1);
  OFB_LAUNCH_CHECK(h);
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
  if ((st = upload_image(h, s, 0, image, width, height, stride_bytes))) return st;
  if ((st = eigen_map(h, s, width, height, block_size))) return st;
  OFB_CUDA(h, cudaMemcpyAsync(eig_out, s->eig, (size_t)width * height * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_good_features(ofb_handle* h, const uint8_t* image, int width, int height, size_t stride_bytes,
                      const ofb_gftt_params* p, float* corners_xy, int* n_out) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!image || !p || !corners_xy || !n_out) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (!(p->quality_level > 0) || p->min_distance < 0)
    return set_error(h, OFB_ERR_INVALID_ARG, "qualityLevel must be > 0 and minDistance >= 0");
  if (p->block_size < 1 || p->block_size % 2 == 0 || p->block_size > 31)
    return set_error(h, OFB_ERR_INVALID_ARG, "blockSize must be odd and in [1,31]");
  if (width > 65535 || height > 65535) return set_error(h, OFB_ERR_INVALID_ARG, "image larger than 65535 px");
  if (width < 3 || height < 3) {     // no interior pixel: cv2 returns an empty list
    *n_out = 0;
    return OFB_OK;
  }
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  int st = sparse_get(h, &s);
  if (st) return st;
  if ((st = upload_image(h, s, 0, image, width, height, stride_bytes))) return st;
  if ((st = eigen_map(h, s, width, height, p->block_size))) return st;
  cudaStream_t sm = h->stream;
  dim3 b(32, 8);
  k_candidates<<<g2(width - 2, height - 2, b), b, 0, sm>>>(s->eig, width, height, p->quality_level, s->counters + 1,
                                                          s->keys, s->counters, (unsigned int)s->cand_cap);
  OFB_LAUNCH_CHECK(h);
  // the candidate count is needed on the host to size the sort (a 4-byte D2H)
  unsigned int* hc = reinterpret_cast<unsigned int*>(s->h_stage);
  OFB_CUDA(h, cudaMemcpyAsync(hc, s->counters, sizeof(unsigned int), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  const unsigned int n_cand = std::min<unsigned int>(hc[0], (unsigned int)s->cand_cap);
  const unsigned long long* sorted = s->keys;
  if (n_cand > 1) {
    size_t tmp = s->cub_tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortKeysDescending(s->cub_tmp, tmp, s->keys, s->keys + s->cand_cap, (int)n_cand,
                                                             0, 64, sm);
    if (e != cudaSuccess) return set_error(h, OFB_ERR_CUDA, "radix sort failed: %s", cudaGetErrorString(e));
    h->launches += 4;
    sorted = s->keys + s->cand_cap;
  }
  const int cell = p->min_distance >= 1 ? (int)__builtin_nearbyint(p->min_distance) : 1;
  const size_t cells = (size_t)((width + cell - 1) / cell) * ((height + cell - 1) / cell);
  if (p->min_distance >= 1) OFB_CUDA(h, cudaMemsetAsync(s->grid_cnt, 0, cells * sizeof(unsigned int), sm));
  k_greedy_select<<<1, GS_THREADS, 0, sm>>>(sorted, s->counters, (unsigned int)s->cand_cap, width, height, (float)p->min_distance,
                                    p->max_corners, s->grid_cnt, s->grid_pts, s->corners, s->counters + 2);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemcpyAsync(hc, s->counters + 2, sizeof(unsigned int), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  const unsigned int n = hc[0];
  if (n) {
    float* hp = reinterpret_cast<float*>(s->h_stage);
    OFB_CUDA(h, cudaMemcpyAsync(hp, s->corners, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, sm));
    OFB_CUDA(h, cudaStreamSynchronize(sm));
    memcpy(corners_xy, hp, (size_t)n * sizeof(float2));
  }
  *n_out = (int)n;
  return OFB_OK;
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
  if ((st = upload_image(h, s, 0, image, width, height, stride_bytes))) return st;
  const uint8_t* lp[kMaxLkLevels];
  int lw[kMaxLkLevels], lh[kMaxLkLevels], nl = 0;
  if ((st = build_pyr(h, s, 0, width, height, win_w, win_h, max_level, lp, lw, lh, &nl))) return st;
  short2* d = s->deriv;
  for (int l = 0; l < nl; l++) {
    if (level_out && level_out[l])
      OFB_CUDA(h, cudaMemcpyAsync(level_out[l], lp[l], (size_t)lw[l] * lh[l], cudaMemcpyDeviceToHost, h->stream));
    if (deriv_out && deriv_out[l]) {
      dim3 b(32, 8);
      k_scharr<<<g2(lw[l], lh[l], b), b, 0, h->stream>>>(lp[l], lw[l], lh[l], (size_t)lw[l], d);
      OFB_LAUNCH_CHECK(h);
      OFB_CUDA(h, cudaMemcpyAsync(deriv_out[l], d, (size_t)lw[l] * lh[l] * sizeof(short2), cudaMemcpyDeviceToHost,
                                  h->stream));
      d += (size_t)lw[l] * lh[l];
    }
  }
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_levels_out = nl;
  return OFB_OK;
}

int ofb_pyrlk(ofb_handle* h, const uint8_t* prev, const uint8_t* next, int width, int height, size_t stride_bytes,
              const float* prev_pts, int n_points, float* next_pts, uint8_t* status, float* err,
              const ofb_lk_params* p) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!prev || !next || !p || !status || !next_pts || (!prev_pts && n_points > 0))
    return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (n_points < 0) return set_error(h, OFB_ERR_INVALID_ARG, "negative point count");
  if (p->win_w <= 2 || p->win_h <= 2) return set_error(h, OFB_ERR_INVALID_ARG, "winSize must be > 2x2");
  if (p->max_level < 0) return set_error(h, OFB_ERR_INVALID_ARG, "maxLevel must be >= 0");
  if ((size_t)p->win_w * p->win_h * 3 * sizeof(short) * LK_WARPS > 200 * 1024)
    return set_error(h, OFB_ERR_INVALID_ARG, "winSize too large");
  if (n_points == 0) return OFB_OK;
  OFB_CUDA(h, cudaSetDevice(h->device));
  SparseState* s;
  int st = sparse_get(h, &s);
  if (st) return st;
  if ((st = sparse_points(h, s, n_points))) return st;
  if ((st = upload_image(h, s, 0, prev, width, height, stride_bytes))) return st;
  if ((st = upload_image(h, s, 1, next, width, height, stride_bytes))) return st;
  cudaStream_t sm = h->stream;
  LkLevels lv;
  int nl0 = 0, nl1 = 0;
  if ((st = build_pyr(h, s, 0, width, height, p->win_w, p->win_h, p->max_level, lv.I, lv.w, lv.h, &nl0))) return st;
  int w2[kMaxLkLevels], h2[kMaxLkLevels];
  if ((st = build_pyr(h, s, 1, width, height, p->win_w, p->win_h, p->max_level, lv.J, w2, h2, &nl1))) return st;
  lv.n_levels = nl0;
  lv.pitch0 = (size_t)width;
  short2* d = s->deriv;
  for (int l = 0; l < nl0; l++) {
    dim3 b(32, 8);
    k_scharr<<<g2(lv.w[l], lv.h[l], b), b, 0, sm>>>(lv.I[l], lv.w[l], lv.h[l], (size_t)lv.w[l], d);
    OFB_LAUNCH_CHECK(h);
    lv.D[l] = d;
    d += (size_t)lv.w[l] * lv.h[l];
  }
  OFB_CUDA(h, cudaMemcpyAsync(s->pts_prev, prev_pts, (size_t)n_points * sizeof(float2), cudaMemcpyHostToDevice, sm));
  if (p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)
    OFB_CUDA(h, cudaMemcpyAsync(s->pts_next, next_pts, (size_t)n_points * sizeof(float2), cudaMemcpyHostToDevice, sm));
  const int max_count = std::min(std::max(p->max_count, 0), 100);
  double eps = std::min(std::max(p->epsilon, 0.0), 10.0);
  eps *= eps;
  const size_t smem = (size_t)p->win_w * p->win_h * 3 * sizeof(short) * LK_WARPS;
  if (smem > 48 * 1024)
    OFB_CUDA(h, cudaFuncSetAttribute(k_lk_track, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_lk_track<<<(n_points + LK_WARPS - 1) / LK_WARPS, LK_WARPS * 32, smem, sm>>>(
      lv, s->pts_prev, s->pts_next, s->lk_status, s->lk_err, n_points, p->win_w, p->win_h, max_count, eps, p->flags,
      p->min_eig_threshold);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemcpyAsync(next_pts, s->pts_next, (size_t)n_points * sizeof(float2), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaMemcpyAsync(status, s->lk_status, (size_t)n_points, cudaMemcpyDeviceToHost, sm));
  if (err) OFB_CUDA(h, cudaMemcpyAsync(err, s->lk_err, (size_t)n_points * sizeof(float), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  return OFB_OK;
}

}  // extern "C"
