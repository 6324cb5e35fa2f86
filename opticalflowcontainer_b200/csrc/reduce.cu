// reduce.cu — on-device reduction of the flow field to the node's scalar.
//
// Every reference node collapses the dense field to one number right after the flow call:
//   np.median(flow_np[0])        ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:207
//   np.mean(flow_np[0])          ros2_ws/src/optical_flow/optical_flow/opticalflow_node.py:98
//   np.median(flow_np[0][mask])  ros2_ws/src/liteflownet3/liteflownet3/sub_n_pub_lfn3_node.py:206-210
// Doing it here removes the 8N-byte D2H of the field.  Mean: one pass, double accumulation.
// Median: exact 3-pass radix select (11 + 11 + 10 bits of the order-preserving key) of the two
// middle ranks at once; for an even count the result is the float32 mean of the two, as
// np.median returns for float32 input.
#include <algorithm>

#include "common.cuh"

namespace ofb {

constexpr int SEL_BINS = 2048;
// per-pair selection state in d_sel (uint32 words)
//   [0..2047] histogram A, [2048..4095] histogram B, then:
constexpr int SEL_PREFIX_A = 4096, SEL_PREFIX_B = 4097, SEL_RANK_A = 4098, SEL_RANK_B = 4099, SEL_COUNT = 4100,
              SEL_WORDS = 4104;

__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// pass p (0,1,2): histogram of digit p of the keys whose higher digits equal the selected prefix.
// digits: p0 = bits 31..21 (11), p1 = bits 20..10 (11), p2 = bits 9..0 (10).
__global__ void __launch_bounds__(256) k_sel_hist(const float2* __restrict__ flow, const uint8_t* __restrict__ mask,
                                                  int npix, uint32_t* __restrict__ sel, int pass,
                                                  double* __restrict__ sums) {
  __shared__ uint32_t ha[SEL_BINS], hb[SEL_BINS];
  __shared__ double ssum[8];
  uint32_t* st = sel + (size_t)blockIdx.y * SEL_WORDS;
  const float2* f = flow + (size_t)blockIdx.y * npix;
  for (int i = threadIdx.x; i < SEL_BINS; i += 256) ha[i] = hb[i] = 0;
  __syncthreads();
  const uint32_t pa = st[SEL_PREFIX_A], pb = st[SEL_PREFIX_B];
  const int shift_hi = pass == 1 ? 21 : 10;  // bits above the current digit
  double s = 0.0;
  // A flow field is smooth: the 32 neighbouring pixels of a warp usually fall into the SAME bin of the leading digits
  // (all of them under a uniform pan), and 32 shared-memory atomics on one address serialise.  One vote tells whether
  // the warp agrees; then one lane adds 32.  (Otherwise: one atomic per lane, as before.)
  const int lane = threadIdx.x & 31;
  auto add = [&](uint32_t* hist, bool pred, uint32_t bin) {
    int same;
    __match_all_sync(0xffffffffu, pred ? bin : 0xffffffffu, &same);
    if (same) {
      if (pred && lane == 0) atomicAdd(&hist[bin], 32u);
    } else if (pred) {
      atomicAdd(&hist[bin], 1u);
    }
  };
  // four 32-pixel groups per warp and step, their loads issued together: one load in flight per thread left the
  // kernel at 3 TB/s (a pass is nothing but one read of the field)
  constexpr int UNR = 8;
  const int warp = threadIdx.x >> 5;
  for (int base = (blockIdx.x * 8 + warp) * (32 * UNR); base < npix; base += gridDim.x * 8 * 32 * UNR) {   // (uniform over the warp)
    float u[UNR];
    bool valid[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) {
      const int i = base + j * 32 + lane;
      valid[j] = i < npix && !(mask && !mask[i]);
      u[j] = valid[j] ? f[i].x : 0.f;
    }
#pragma unroll
    for (int j = 0; j < UNR; j++) {
      const uint32_t k = f2key(u[j]);
      if (pass == 0) {
        add(ha, valid[j], k >> 21);
        if (valid[j]) s += (double)u[j];
      } else {
        const uint32_t hi = k >> shift_hi;
        const uint32_t dig = pass == 1 ? (k >> 10) & 0x7ffu : k & 0x3ffu;
        add(ha, valid[j] && hi == pa, dig);
        if (pa != pb) add(hb, valid[j] && hi == pb, dig);     // (same prefix — the usual case: hb is ha, see below)
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SEL_BINS; i += 256) {
    if (ha[i]) atomicAdd(&st[i], ha[i]);
    const uint32_t vb = pa == pb ? ha[i] : hb[i];
    if (pass > 0 && vb) atomicAdd(&st[SEL_BINS + i], vb);
  }
  if (pass == 0) {
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int i = 0; i < 8; i++) t += ssum[i];
      atomicAdd(&sums[blockIdx.y], t);
    }
  }
}

// after pass p: locate the bins holding the two ranks, extend the prefixes, clear the histograms.
__global__ void __launch_bounds__(256) k_sel_scan(uint32_t* __restrict__ sel, int pass) {
  uint32_t* st = sel + (size_t)blockIdx.x * SEL_WORDS;
  __shared__ uint32_t sh[2][SEL_BINS];
  __shared__ uint32_t csum[2][256];
  __shared__ uint32_t res[4];
  for (int i = threadIdx.x; i < SEL_BINS; i += 256) {
    sh[0][i] = st[i];
    sh[1][i] = pass == 0 ? st[i] : st[SEL_BINS + i];
  }
  __syncthreads();
  for (int which = 0; which < 2; which++) {
    uint32_t c = 0;
    for (int i = 0; i < 8; i++) c += sh[which][threadIdx.x * 8 + i];
    csum[which][threadIdx.x] = c;
  }
  __syncthreads();
  // the bin that holds rank r: prefix sums of the 256 chunk counts over the CTA (warp shuffles + the 8 warp totals), the
  // one thread whose chunk [exclusive, inclusive) contains r walks its 8 bins (two threads walking 255 chunks one
  // after the other took ~8 us per launch, three launches per reduction)
  __shared__ uint32_t wtot[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc[2];
  for (int which = 0; which < 2; which++) {
    uint32_t v = csum[which][threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    inc[which] = v;
    if (lane == 31) wtot[which][wid] = v;
  }
  __syncthreads();
  for (int which = 0; which < 2; which++) {
    uint32_t before = 0, cnt = 0;
    for (int k = 0; k < 8; k++) {
      if (k < wid) before += wtot[which][k];
      cnt += wtot[which][k];
    }
    uint32_t rank;
    if (pass == 0) {
      if (which == 0 && threadIdx.x == 0) st[SEL_COUNT] = cnt;
      rank = cnt ? (which == 0 ? (cnt - 1) / 2 : cnt / 2) : 0;
    } else {
      rank = st[which == 0 ? SEL_RANK_A : SEL_RANK_B];
    }
    const uint32_t incl = before + inc[which], excl = incl - csum[which][threadIdx.x];
    // the chunk that contains the rank — or, as the sequential walk did, the last chunk when no count exceeds it
    const bool mine = cnt > rank ? (excl <= rank && rank < incl) : threadIdx.x == 255;
    if (mine) {
      uint32_t acc = excl;
      int b = threadIdx.x * 8;
      for (; b < threadIdx.x * 8 + 7; b++) {
        if (acc + sh[which][b] > rank) break;
        acc += sh[which][b];
      }
      res[which * 2] = (uint32_t)b;
      res[which * 2 + 1] = rank - acc;
    }
  }
  __syncthreads();
  const uint32_t pa = st[SEL_PREFIX_A], pb = st[SEL_PREFIX_B];
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * SEL_BINS; i += 256) st[i] = 0;
  if (threadIdx.x == 0) {
    const int bits = pass == 2 ? 10 : 11;
    st[SEL_PREFIX_A] = pass == 0 ? res[0] : (pa << bits) | res[0];
    st[SEL_PREFIX_B] = pass == 0 ? res[2] : (pb << bits) | res[2];
    st[SEL_RANK_A] = res[1];
    st[SEL_RANK_B] = res[3];
  }
}

__global__ void k_sel_final(const uint32_t* __restrict__ sel, const double* __restrict__ sums, double* mean_out,
                            float* median_out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* st = sel + (size_t)i * SEL_WORDS;
  const uint32_t cnt = st[SEL_COUNT];
  mean_out[i] = cnt ? sums[i] / (double)cnt : __longlong_as_double(0x7ff8000000000000LL);
  const float a = key2f(st[SEL_PREFIX_A]), b = key2f(st[SEL_PREFIX_B]);
  median_out[i] = cnt ? (a + b) * 0.5f : __uint_as_float(0x7fc00000u);
}

static void hand_over(ofb_handle* h, const ofb_handle::PendingStats& ps) {
  const char* slot = h->h_stats + (size_t)ps.slot * h->max_batch * 16;
  const double* hm = reinterpret_cast<const double*>(slot);
  const float* hd = reinterpret_cast<const float*>(slot + (size_t)h->max_batch * 8);
  for (int i = 0; i < ps.n; i++) {
    if (ps.out_mean) ps.out_mean[i] = hm[i];
    if (ps.out_median) ps.out_median[i] = hd[i];
  }
}

int finish_oldest_stats(ofb_handle* h) {
  if (h->pending_stats.empty()) return OFB_OK;
  const ofb_handle::PendingStats ps = h->pending_stats.front();
  OFB_CUDA(h, cudaEventSynchronize(h->stats_ev[ps.slot]));
  hand_over(h, ps);
  h->pending_stats.erase(h->pending_stats.begin());
  return OFB_OK;
}

int finish_pending_stats(ofb_handle* h) {
  if (h->pending_stats.empty()) return OFB_OK;
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (const auto& ps : h->pending_stats) hand_over(h, ps);
  h->pending_stats.clear();
  return OFB_OK;
}

int flow_u_stats(ofb_handle* h, int n, const uint8_t* host_mask, double* out_mean, float* out_median, bool async) {
  if (!h->last_flow || n < 1 || n > h->last_n)
    return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_u_stats: no flow field of %d pair(s) on the device", n);
  OFB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int npix = h->last_w * h->last_h;
  const uint8_t* dmask = nullptr;
  if (host_mask) {
    OFB_CUDA(h, cudaMemcpyAsync(h->d_mask, host_mask, (size_t)npix, cudaMemcpyHostToDevice, st));
    dmask = h->d_mask;
  }
  OFB_CUDA(h, cudaMemsetAsync(h->d_sel, 0, (size_t)n * SEL_WORDS * sizeof(uint32_t), st));
  double* sums = h->d_stats;
  double* mean_d = h->d_stats + h->max_batch;
  float* med_d = reinterpret_cast<float*>(h->d_stats + 2 * (size_t)h->max_batch);
  OFB_CUDA(h, cudaMemsetAsync(sums, 0, sizeof(double) * n, st));
  const int blocks = std::max(1, std::min((npix + 256 * 8 - 1) / (256 * 8), 2 * h->num_sms));
  int s;
  if ((s = timing_begin(h, OFB_STAGE_OTHER))) return s;
  for (int pass = 0; pass < 3; pass++) {
    k_sel_hist<<<dim3(blocks, n), 256, 0, st>>>((const float2*)h->last_flow, dmask, npix, h->d_sel, pass, sums);
    OFB_LAUNCH_CHECK(h);
    k_sel_scan<<<n, 256, 0, st>>>(h->d_sel, pass);
    OFB_LAUNCH_CHECK(h);
  }
  k_sel_final<<<(n + 127) / 128, 128, 0, st>>>(h->d_sel, sums, mean_d, med_d, n);
  OFB_LAUNCH_CHECK(h);
  if ((s = timing_end(h))) return s;
  // results: n doubles + n floats through a pinned slot; the caller's arrays are filled once the stream
  // has got there (right away for the synchronous call, in ofb_wait for the asynchronous one)
  if ((int)h->pending_stats.size() >= kStatSlots && (s = finish_oldest_stats(h))) return s;   // (no stream drain)
  const int slot_i = h->stats_slot;
  h->stats_slot = (h->stats_slot + 1) % kStatSlots;
  char* slot = h->h_stats + (size_t)slot_i * h->max_batch * 16;
  OFB_CUDA(h, cudaMemcpyAsync(slot, mean_d, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  OFB_CUDA(h, cudaMemcpyAsync(slot + (size_t)h->max_batch * 8, med_d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  if (!h->stats_ev[slot_i]) OFB_CUDA(h, cudaEventCreateWithFlags(&h->stats_ev[slot_i], cudaEventDisableTiming));
  OFB_CUDA(h, cudaEventRecord(h->stats_ev[slot_i], st));
  h->pending_stats.push_back({out_mean, out_median, n, slot_i});
  return async ? OFB_OK : finish_pending_stats(h);
}

}  // namespace ofb
