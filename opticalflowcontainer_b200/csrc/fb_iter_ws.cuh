// fb_iter_ws.cuh — warp-specialised fused Farneback iteration kernel (box window).
//
// Same arithmetic and HBM traffic as k_iter_box (fb_iter.cuh): UpdateMatrices + (2m+1)^2 box blur +
// 2x2 solve in one pass, 56 B per pixel-iteration.  The difference is the schedule inside the CTA:
// ablations of k_iter_box showed its three barrier-separated phases (gather -> horizontal sums ->
// vertical sums/solve) each run latency-bound with 16 warps per SM.  Here the phases overlap:
//
//   warps 0-3  PRODUCERS  A1: matrices of a 2-row chunk (4 px per lane, all loads in flight before
//                         first use, next chunk's flow prefetched) -> double-buffered staging rows
//   warps 4-7  CONSUMERS  A2: horizontal window sums of the staged rows -> ring of 2m+1 rows
//                         B : vertical running sums in double + solve + flow store (2 columns/thread)
//
// Producers and consumers hand staging buffers over with named barriers (bar.arrive / bar.sync,
// FULL[2] and EMPTY[2]); consumers order their ring accesses with a 128-thread barrier.  While the
// producers wait on the R1 gather, the consumers of the same CTA (and both roles of the second
// resident CTA) keep the issue slots busy.
#pragma once
#include "fb_device.cuh"

namespace ofb {

constexpr int WS_COLS = 256;     // strip width (matrix columns, 2m of them halo)
constexpr int WS_CH = 2;         // rows per chunk
constexpr int WS_THREADS = 256;  // 4 producer + 4 consumer warps

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

enum { WS_BAR_FULL0 = 1, WS_BAR_FULL1 = 2, WS_BAR_EMPTY0 = 3, WS_BAR_EMPTY1 = 4, WS_BAR_CONS = 5 };

template <int MT>
__global__ void __launch_bounds__(WS_THREADS, 2)
    k_iter_ws(const float4* __restrict__ RA, const float* __restrict__ RB, const float2* __restrict__ flow_in,
              float2* __restrict__ flow_out, int w, int h, int f1_offset, int m_rt, float scale, int seg_rows,
              int strips) {
  constexpr int COLS = WS_COLS;
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int tw = COLS - 2 * m;
  extern __shared__ float smem[];
  float* stage = smem;                               // [2 buffers][WS_CH][5][COLS]
  float* ring = smem + 2 * WS_CH * 5 * COLS;         // [R][5][COLS]

  const int strip = blockIdx.x % strips;
  const int seg = blockIdx.x / strips;
  const int pair = blockIdx.y;
  const int x_base = strip * tw - m;                 // image x of strip column 0
  const int y0 = seg * seg_rows;
  const int y1 = min(y0 + seg_rows, h);              // exclusive
  const int t_first = y0 - m, t_last = y1 - 1 + m;
  const int n_chunks = (t_last - t_first + WS_CH) / WS_CH;

  const size_t n = (size_t)w * h;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (warp < 4) {
    // ------------------------------------------------------------------ PRODUCERS
    const float4* RA0 = RA + (size_t)pair * n;
    const float* RB0 = RB + (size_t)pair * n;
    const float4* RA1 = RA + (size_t)(pair + f1_offset) * n;
    const float* RB1 = RB + (size_t)(pair + f1_offset) * n;
    const float2* fin = flow_in + (size_t)pair * n;
    const int a_row = warp >> 1, a_half = warp & 1;
    float2 fl[4];
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; j++) xs[j] = clampi(x_base + a_half * 128 + lane + 32 * j, 0, w - 1);
    {
      const int y = clampi(t_first + a_row, 0, h - 1);
#pragma unroll
      for (int j = 0; j < 4; j++) fl[j] = __ldg(fin + y * w + xs[j]);
    }
    for (int c = 0; c < n_chunks; c++) {
      const int buf = c & 1;
      const int t = t_first + c * WS_CH + a_row;
      if (c >= 2) named_bar_sync(WS_BAR_EMPTY0 + buf, WS_THREADS);   // consumers released this buffer
      if (t <= t_last) {
        const int y = clampi(t, 0, h - 1);
        float* srow = stage + (buf * WS_CH + a_row) * 5 * COLS + a_half * 128 + lane;
        UmLoads L[4];
#pragma unroll
        for (int j = 0; j < 4; j++) um_issue(L[j], RA0, RB0, RA1, RB1, fl[j], xs[j], y, w, h);
        {  // next chunk's flow (clamped row: always a valid address)
          const int yn = clampi(t + WS_CH, 0, h - 1);
#pragma unroll
          for (int j = 0; j < 4; j++) fl[j] = __ldg(fin + yn * w + xs[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const M5 mm = um_finish(L[j], xs[j], y, w, h);
          srow[0 * COLS + 32 * j] = mm.g11;
          srow[1 * COLS + 32 * j] = mm.g12;
          srow[2 * COLS + 32 * j] = mm.g22;
          srow[3 * COLS + 32 * j] = mm.h1;
          srow[4 * COLS + 32 * j] = mm.h2;
        }
      }
      named_bar_arrive(WS_BAR_FULL0 + buf, WS_THREADS);              // staging rows of chunk c are ready
    }
    return;
  }

  // -------------------------------------------------------------------- CONSUMERS
  float2* fout = flow_out + (size_t)pair * n;
  const int ct = tid - 128;                          // 0..127
  const int q_row = ct >> 6;                         // A2: staged row of this thread's quad
  const int q0 = (ct & 63) * 4;                      // A2: first of its 4 columns
  const int colA = ct, colB = ct + 128;              // B: its two columns
  const bool validA = colA >= m && colA < COLS - m && x_base + colA < w;
  const bool validB = colB >= m && colB < COLS - m && x_base + colB < w;
  double va[5] = {0, 0, 0, 0, 0}, vb[5] = {0, 0, 0, 0, 0};
  float oldA[WS_CH][5], oldB[WS_CH][5];
#pragma unroll
  for (int rr = 0; rr < WS_CH; rr++)
#pragma unroll
    for (int ch = 0; ch < 5; ch++) oldA[rr][ch] = oldB[rr][ch] = 0.f;

  int slot0 = 0;                                     // ring slot of the chunk's first row
  for (int c = 0; c < n_chunks; c++) {
    const int buf = c & 1;
    const int tc = t_first + c * WS_CH;
    int slot[WS_CH];
#pragma unroll
    for (int rr = 0; rr < WS_CH; rr++) {
      const int sl = slot0 + rr;
      slot[rr] = sl >= R ? sl - R : sl;
    }
    named_bar_sync(WS_BAR_FULL0 + buf, WS_THREADS);  // producers finished staging chunk c
    // ---- A2: horizontal window sums of one staged quad-row, kept in registers
    float hq[5][4];
    const bool q_live = tc + q_row <= t_last;
    if (q_live) {
      const float* srow = stage + (buf * WS_CH + q_row) * 5 * COLS;
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        const float* s = srow + ch * COLS;
        float s0, s1, s2, s3;
        if (MT > 0) {
          constexpr int KQ = (MT + 3) / 4;
          float e[(2 * KQ + 1) * 4];                 // e[d + 4*KQ] = staged value at column q0 + d
#pragma unroll
          for (int k = -KQ; k <= KQ; k++) {
            const int cq = min(max(q0 + 4 * k, 0), COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            e[(k + KQ) * 4 + 0] = v.x; e[(k + KQ) * 4 + 1] = v.y; e[(k + KQ) * 4 + 2] = v.z; e[(k + KQ) * 4 + 3] = v.w;
          }
          constexpr int O = 4 * KQ;
          float core = e[O + 3 - MT];                // d in [3-MT, MT] is inside all four windows
#pragma unroll
          for (int d = 4 - MT; d <= MT; d++) core += e[O + d];
          float l = e[O + 2 - MT];
          s2 = core + l;
          l += e[O + 1 - MT];
          s1 = core + l;
          l += e[O - MT];
          s0 = core + l;
          float r = e[O + MT + 1];
          s1 += r;
          r += e[O + MT + 2];
          s2 += r;
          r += e[O + MT + 3];
          s3 = core + r;
        } else {
          s0 = s1 = s2 = s3 = 0.f;
          const int kq = (m + 3) >> 2;
          for (int k = -kq; k <= kq; k++) {
            const int cq = min(max(q0 + 4 * k, 0), COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int d = 4 * k + i;
              if (d >= 0 - m && d <= 0 + m) s0 += e[i];
              if (d >= 1 - m && d <= 1 + m) s1 += e[i];
              if (d >= 2 - m && d <= 2 + m) s2 += e[i];
              if (d >= 3 - m && d <= 3 + m) s3 += e[i];
            }
          }
        }
        hq[ch][0] = s0; hq[ch][1] = s1; hq[ch][2] = s2; hq[ch][3] = s3;
      }
    }
    if (c + 2 < n_chunks) named_bar_arrive(WS_BAR_EMPTY0 + buf, WS_THREADS);  // staging buffer may be refilled
    // every consumer has fetched the rows leaving the window (end of the previous iteration):
    // their ring slots may now be overwritten
    named_bar_sync(WS_BAR_CONS, 128);
    if (q_live) {
      float* rrow = ring + (q_row == 0 ? slot[0] : slot[1]) * 5 * COLS + q0;
#pragma unroll
      for (int ch = 0; ch < 5; ch++)
        *reinterpret_cast<float4*>(rrow + ch * COLS) = make_float4(hq[ch][0], hq[ch][1], hq[ch][2], hq[ch][3]);
    }
    named_bar_sync(WS_BAR_CONS, 128);                // new ring rows visible to all consumers
    // ---- B: vertical running sums (double) + solve, two columns per thread
    const int nrows = min(WS_CH, t_last - tc + 1);
#pragma unroll
    for (int rr = 0; rr < WS_CH; rr++) {
      if (rr < nrows) {
        const float* ra = ring + slot[rr] * 5 * COLS + colA;
        const float* rb = ra + 128;
#pragma unroll
        for (int ch = 0; ch < 5; ch++) {
          va[ch] += (double)ra[ch * COLS] - (double)oldA[rr][ch];
          vb[ch] += (double)rb[ch * COLS] - (double)oldB[rr][ch];
        }
        const int y = tc + rr - m;
        if (y >= y0) {
          if (validA)
            fout[y * w + x_base + colA] = solve2x2((float)va[0] * scale, (float)va[1] * scale, (float)va[2] * scale,
                                                   (float)va[3] * scale, (float)va[4] * scale);
          if (validB)
            fout[y * w + x_base + colB] = solve2x2((float)vb[0] * scale, (float)vb[1] * scale, (float)vb[2] * scale,
                                                   (float)vb[3] * scale, (float)vb[4] * scale);
        }
      }
    }
    // ---- fetch the rows that leave the window in the NEXT chunk (their slots are overwritten there)
    slot0 += WS_CH;
    if (slot0 >= R) slot0 -= R;
    const int n_done = tc + WS_CH - t_first;         // rows in the ring after this chunk
#pragma unroll
    for (int rr = 0; rr < WS_CH; rr++) {
      const bool have_old = n_done + rr >= R;        // row (next tc + rr - R) exists
      const int sl = slot0 + rr;
      const int so = (sl >= R ? sl - R : sl) * 5 * COLS;
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        oldA[rr][ch] = have_old ? ring[so + ch * COLS + colA] : 0.f;
        oldB[rr][ch] = have_old ? ring[so + ch * COLS + colB] : 0.f;
      }
    }
  }
}

}  // namespace ofb
