// fb_iter_z.cuh — high-occupancy warp-specialised fused Farneback iteration kernel (box window).
//
// UpdateMatrices + (2m+1)^2 box blur + 2x2 solve in one pass over the level, 56 B of HBM traffic per
// pixel-iteration, same arithmetic as k_iter_ws2 (fb_iter_ws2.cuh).  ncu and ablations of the
// 128-register kernels (k_iter_box / k_iter_ws / k_iter_ws2: 16 warps per SM) showed every warp
// running at ~0.25 IPC — dependent-issue latency in the consumers, exposed gather latency in the
// producers — and neither trimming instructions (-8 %) nor L2 prefetch moved the time.  The cure is
// thread-level parallelism: this kernel spends registers per THREAD sparingly so that 24-32 warps are
// resident per SM:
//   * producers hold ONE pixel of gather results per lane (30 registers) instead of four;
//   * consumers own ONE column in the vertical phase and write the horizontal sums straight into
//     the ring (no per-thread copy held across a barrier);
//   * COLS producer threads + COLS consumer threads per CTA, <= 64-80 registers each.
// Per chunk of CH rows:  producers  A1: matrices of CH rows -> staging (double-buffered)
//                        consumers  A2: horizontal window sums of the staged rows -> ring of 2m+1 rows
//                                   B : vertical running sums in double + solve + flow store
// hand-over by named barriers FULL[2] / EMPTY[2]; one consumer-only barrier per chunk.
#pragma once
#include "fb_device.cuh"
#include "fb_iter_ws2.cuh"

namespace ofb {

enum { Z_BAR_FULL0 = 1, Z_BAR_FULL1 = 2, Z_BAR_EMPTY0 = 3, Z_BAR_EMPTY1 = 4, Z_BAR_CONS = 5 };

template <int COLS, int CH>
constexpr int iter_z_smem_floats(int m) { return (2 * CH + 2 * m + 1) * 5 * COLS; }

template <int MT, int COLS, int CH, int MINB>
__global__ void __launch_bounds__(2 * COLS, MINB)
    k_iter_z(const float4* __restrict__ RA, const float* __restrict__ RB, const float2* __restrict__ flow_in,
             float2* __restrict__ flow_out, int w, int h, int f1_offset, int m_rt, float reg, int seg_rows,
             int strips) {
  static_assert(CH * (COLS / 4) <= COLS, "A2 needs one consumer thread per (row, quad)");
  constexpr int NT = 2 * COLS;
  constexpr int QUADS = COLS / 4;
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int tw = COLS - 2 * m;
  extern __shared__ float smem[];
  float* stage = smem;                               // [2 buffers][CH][5][COLS]
  float* ring = smem + 2 * CH * 5 * COLS;            // [R][5][COLS]

  const int strip = blockIdx.x % strips;
  const int seg = blockIdx.x / strips;
  const int pair = blockIdx.y;
  const int x_base = strip * tw - m;                 // image x of strip column 0
  const int y0 = seg * seg_rows;
  const int y1 = min(y0 + seg_rows, h);              // exclusive
  const int t_first = y0 - m, t_last = y1 - 1 + m;
  const int n_chunks = (t_last - t_first + CH) / CH;

  const size_t n = (size_t)w * h;
  const int tid = threadIdx.x;

  if (tid < COLS) {
    // ------------------------------------------------------------------ PRODUCERS (one column each)
    const float4* RA0 = RA + (size_t)pair * n;
    const float* RB0 = RB + (size_t)pair * n;
    const float4* RA1 = RA + (size_t)(pair + f1_offset) * n;
    const float* RB1 = RB + (size_t)(pair + f1_offset) * n;
    const float2* fin = flow_in + (size_t)pair * n;
    asm volatile("" : "+l"(RA0), "+l"(RB0), "+l"(RA1), "+l"(RB1), "+l"(fin));   // keep the bases, do not re-derive
    const unsigned uw = (unsigned)w, uh = (unsigned)h;
    const int x = clampi(x_base + tid, 0, w - 1);
    const bool xborder = (unsigned)(x - 5) >= (unsigned)(w - 10);
    float2 fl = __ldg(fin + ((unsigned)clampi(t_first, 0, h - 1) * uw + (unsigned)x));
    for (int c = 0; c < n_chunks; c++) {
      const int buf = c & 1;
      if (c >= 2) named_bar_sync(Z_BAR_EMPTY0 + buf, NT);          // consumers released this buffer
      float* srow = stage + buf * CH * 5 * COLS + tid;
#pragma unroll
      for (int rr = 0; rr < CH; rr++) {
        const int t = t_first + c * CH + rr;
        if (t <= t_last) {
          const int y = clampi(t, 0, h - 1);
          const unsigned yw = (unsigned)y * uw;
          UmLoads2 L;
          um_issue2<false>(L, RA0, RB0, RA1, RB1, fl, x, y, yw, uw, uh);
          fl = __ldg(fin + ((unsigned)clampi(t + 1, 0, h - 1) * uw + (unsigned)x));   // next row's flow
          const M5 mm = um_finish2(L, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h);
          srow[(rr * 5 + 0) * COLS] = mm.g11;
          srow[(rr * 5 + 1) * COLS] = mm.g12;
          srow[(rr * 5 + 2) * COLS] = mm.g22;
          srow[(rr * 5 + 3) * COLS] = mm.h1;
          srow[(rr * 5 + 4) * COLS] = mm.h2;
        }
      }
      named_bar_arrive(Z_BAR_FULL0 + buf, NT);                     // staging rows of chunk c are ready
    }
    return;
  }

  // -------------------------------------------------------------------- CONSUMERS
  float2* fout = flow_out + (size_t)pair * n;
  const int ct = tid - COLS;                         // 0..COLS-1
  const int q_row = ct / QUADS;                      // A2: staged row of this thread's quad
  const int q0 = (ct % QUADS) * 4;                   // A2: first of its 4 columns
  const bool q_thread = q_row < CH;
  const int col = ct;                                // B: its column
  const bool valid = col >= m && col < COLS - m && x_base + col < w;
  double vs[5] = {0, 0, 0, 0, 0};
  float old[CH][5];
#pragma unroll
  for (int rr = 0; rr < CH; rr++)
#pragma unroll
    for (int ch = 0; ch < 5; ch++) old[rr][ch] = 0.f;

  int slot0 = 0;                                     // ring slot of the chunk's first row
  for (int c = 0; c < n_chunks; c++) {
    const int buf = c & 1;
    const int tc = t_first + c * CH;
    // all consumers are past the end of the previous chunk (they have fetched the rows that leave the
    // window), so the ring slots of this chunk's rows may be overwritten once the barrier completes
    named_bar_sync(Z_BAR_FULL0 + buf, NT);           // producers finished staging chunk c
    // ---- A2: horizontal window sums of one staged quad-row -> ring
    if (q_thread && tc + q_row <= t_last) {
      int sl = slot0 + q_row;
      if (sl >= R) sl -= R;
      const float* srow = stage + (buf * CH + q_row) * 5 * COLS;
      float* rrow = ring + sl * 5 * COLS + q0;
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
        *reinterpret_cast<float4*>(rrow + ch * COLS) = make_float4(s0, s1, s2, s3);
      }
    }
    if (c + 2 < n_chunks) named_bar_arrive(Z_BAR_EMPTY0 + buf, NT);   // staging buffer may be refilled
    named_bar_sync(Z_BAR_CONS, COLS);                // new ring rows visible to all consumers
    // ---- B: vertical running sums (double) + solve, one column per thread
    const int nrows = min(CH, t_last - tc + 1);
#pragma unroll
    for (int rr = 0; rr < CH; rr++) {
      if (rr < nrows) {
        int sl = slot0 + rr;
        if (sl >= R) sl -= R;
        const float* ra = ring + sl * 5 * COLS + col;
#pragma unroll
        for (int ch = 0; ch < 5; ch++) vs[ch] += (double)ra[ch * COLS] - (double)old[rr][ch];
        const int y = tc + rr - m;
        if (y >= y0 && valid)
          fout[(unsigned)(y * w + x_base + col)] =
              solve2x2_sums((float)vs[0], (float)vs[1], (float)vs[2], (float)vs[3], (float)vs[4], reg);
      }
    }
    // ---- fetch the rows that leave the window in the NEXT chunk (their slots are overwritten there)
    slot0 += CH;
    if (slot0 >= R) slot0 -= R;
    const int n_done = tc + CH - t_first;            // rows in the ring after this chunk
#pragma unroll
    for (int rr = 0; rr < CH; rr++) {
      const bool have_old = n_done + rr >= R;        // row (next tc + rr - R) exists
      int sl = slot0 + rr;
      if (sl >= R) sl -= R;
#pragma unroll
      for (int ch = 0; ch < 5; ch++) old[rr][ch] = have_old ? ring[(sl * 5 + ch) * COLS + col] : 0.f;
    }
  }
}

}  // namespace ofb
