// fb_iter_v.cuh — fused Farneback iteration kernel (box window), vertical-first, no FP64.
//
// UpdateMatrices + (2m+1)^2 box blur + 2x2 solve in one pass over the level, 56 B of HBM traffic
// per pixel-iteration.  ncu on the earlier kernels (k_iter_ws / ws2 / z) found the common wall: the
// XU pipe at 75 % "realtime" — they kept cv2's vertical running sums in double, and the
// F2F.F64.F32 / F2F.F32.F64 conversions around them (15 per pixel) run at a fraction of the FP32
// rate on B200.  Instruction diets, L2 prefetch and twice the resident warps all left the time
// unchanged.  This kernel removes FP64 altogether:
//
//   * the vertical window sum is done FIRST, on the matrices themselves, by the thread that owns
//     the column, in float, WITHOUT cancellation drift: rows are grouped in blocks of R = 2m+1;
//     P[k] = running prefix sum inside the block, B = sum of the finished block, and the sum of the
//     R rows ending at offset k of the current block is  (B_prev - P_prev[k]) + P_cur[k]
//     (van Herk / Gil-Werman).  P_prev[k] lives in a shared-memory ring that only its owner thread
//     touches (no barrier).  A float running add/subtract sum would drift (measured 0.07 px max
//     EPE on high-contrast frames); this form measured <= 1.4e-2 px max / 3e-5 px mean against cv2
//     on the same frames and 5e-6 px max on textured ones (tools/exp_float_blur.py).
//   * producers (one column per thread) stage the vertically summed rows; consumers only do the
//     horizontal window sums (4 adjacent pixels per thread from float4 reads), the solve and the
//     coalesced flow store.  No consumer-side ring, no consumer-only barrier.
//
// CTA = COLS producer threads + CH*COLS/4 consumer threads; FULL/EMPTY named barriers hand the
// double-buffered staging rows (CH output rows per chunk) over.
#pragma once
#include "fb_device.cuh"
#include "fb_iter_ws2.cuh"

namespace ofb {

template <int COLS, int CH>
constexpr int iter_v_smem_floats(int m) { return (2 * CH + 2 * m + 1) * 5 * COLS; }

// PFD > 0: every producer thread also issues prefetch.global.L2 for what it will load PFD rows later
// (R0, flow, and the new corner row of the R1 gather — the flow field is smooth, so "same
// displacement, PFD rows down" predicts it).  A producer has only one row of loads in flight, so
// without this each row pays a full HBM round trip (~1 us under load); with it the demand loads
// hit in L2.  Prefetches write no register and use no scoreboard.  The R buffers carry kRowPad
// spare rows so the predicted corner row stays inside the allocation (PFD + 1 <= kRowPad).
template <int MT, int COLS, int CH, int MINB, int PFD>
__global__ void __launch_bounds__(COLS + CH * COLS / 4, MINB)
    k_iter_v(const float4* __restrict__ RA, const float* __restrict__ RB, const float2* __restrict__ flow_in,
             float2* __restrict__ flow_out, int w, int h, int f1_offset, int m_rt, float reg, int seg_rows,
             int strips) {
  constexpr int QUADS = COLS / 4;
  constexpr int NCONS = CH * QUADS;
  constexpr int NT = COLS + NCONS;
  enum { BAR_FULL0 = 1, BAR_EMPTY0 = 3 };
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int tw = COLS - 2 * m;
  extern __shared__ float smem[];
  float* stage = smem;                               // [2 buffers][CH][5][COLS]   vertically summed rows
  float* ring = smem + 2 * CH * 5 * COLS;            // [R][5][COLS]               P_prev[k], owner-private columns

  const int strip = blockIdx.x % strips;
  const int seg = blockIdx.x / strips;
  const int pair = blockIdx.y;
  const int x_base = strip * tw - m;                 // image x of strip column 0
  const int y0 = seg * seg_rows;
  const int y1 = min(y0 + seg_rows, h);              // exclusive
  const int t_first = y0 - m;                        // first matrix row the segment needs
  const int n_chunks = (y1 - y0 + CH - 1) / CH;

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
    float* rcol = ring + tid;                        // ring[k][ch][tid]
    float P[5] = {0.f, 0.f, 0.f, 0.f, 0.f};          // prefix sums of the current block
    float Bp[5] = {0.f, 0.f, 0.f, 0.f, 0.f};         // sum of the previous block
    int k = 0;                                       // offset of row t inside its block
    bool have_prev = false;

    // one matrix row: M(t) -> P += M; V = (Bp - P_prev[k]) + P; ring[k] = P; block bookkeeping
    auto row = [&](int t, float (&V)[5]) {
      const int y = clampi(t, 0, h - 1);
      const unsigned yw = (unsigned)y * uw;
#if defined(OFB_DBG) && (OFB_DBG & 1)     // experiment: zero displacement (perfectly regular gathers)
      fl = make_float2(0.f, 0.f);
#endif
#if defined(OFB_DBG) && (OFB_DBG & 2)     // experiment: no loads, no UpdateMatrices
      M5 mm; mm.g11 = 1.f; mm.g12 = 0.f; mm.g22 = 1.f; mm.h1 = 0.5f; mm.h2 = 0.25f;
      float old[5];
#pragma unroll
      for (int ch = 0; ch < 5; ch++) old[ch] = have_prev ? rcol[(k * 5 + ch) * COLS] : 0.f;
#else
      UmLoads2 L;
#if defined(OFB_DBG) && (OFB_DBG & 8)     // experiment: no R1 gather (R0 values stand in for the corners)
      {
        const unsigned o = yw + (unsigned)x;
        L.a0 = __ldg(RA0 + o); L.b0 = __ldg(RB0 + o);
        L.q00 = L.q01 = L.q10 = L.q11 = L.a0; L.s00 = L.s01 = L.s10 = L.s11 = L.b0;
        L.dx = fl.x; L.dy = fl.y; L.fx = 0.25f; L.fy = 0.5f; L.inside = true;
      }
#else
      um_issue2<false>(L, RA0, RB0, RA1, RB1, fl, x, y, yw, uw, uh);
#endif
      if (PFD > 0) {
        static_assert(PFD + 1 <= kRowPad, "prefetch distance exceeds the row padding of the R buffers");
        const unsigned op = (unsigned)clampi(t + PFD, 0, h - 1) * uw + (unsigned)x;
        prefetch_l2(RA0 + op);
        prefetch_l2(RB0 + op);
        prefetch_l2(fin + ((unsigned)clampi(t + PFD + 1, 0, h - 1) * uw + (unsigned)x));
        const unsigned g = L.inside ? (unsigned)__float2int_rd((float)y + L.dy) * uw + (unsigned)__float2int_rd((float)x + L.dx) : 0u;
        prefetch_l2(RA1 + (g + (PFD + 1) * uw));
        prefetch_l2(RB1 + (g + (PFD + 1) * uw));
      }
      fl = __ldg(fin + ((unsigned)clampi(t + 1, 0, h - 1) * uw + (unsigned)x));   // next row's flow
      float old[5];
#pragma unroll
      for (int ch = 0; ch < 5; ch++) old[ch] = have_prev ? rcol[(k * 5 + ch) * COLS] : 0.f;
      const M5 mm = um_finish2(L, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h);
#endif
      P[0] += mm.g11; P[1] += mm.g12; P[2] += mm.g22; P[3] += mm.h1; P[4] += mm.h2;
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        V[ch] = (Bp[ch] - old[ch]) + P[ch];
        rcol[(k * 5 + ch) * COLS] = P[ch];
      }
      if (++k == R) {
        k = 0;
        have_prev = true;
#pragma unroll
        for (int ch = 0; ch < 5; ch++) { Bp[ch] = P[ch]; P[ch] = 0.f; }
      }
    };

    // warm-up: the R-1 rows above the first output row (no hand-over)
    for (int t = t_first; t < t_first + R - 1; t++) {
      float V[5];
      row(t, V);
    }
    for (int c = 0; c < n_chunks; c++) {
      const int buf = c & 1;
      if (c >= 2) named_bar_sync(BAR_EMPTY0 + buf, NT);            // consumers released this buffer
      float* srow = stage + buf * CH * 5 * COLS + tid;
#pragma unroll
      for (int rr = 0; rr < CH; rr++) {
        const int yo = y0 + c * CH + rr;                           // output row; newest matrix row = yo + m
        if (yo < y1) {
          float V[5];
          row(yo + m, V);
#pragma unroll
          for (int ch = 0; ch < 5; ch++) srow[(rr * 5 + ch) * COLS] = V[ch];
        }
      }
      named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
    }
    return;
  }

  // -------------------------------------------------------------------- CONSUMERS (one quad of one row each)
  float2* fout = flow_out + (size_t)pair * n;
  const int ct = tid - COLS;                         // 0..NCONS-1
  const int q_row = ct / QUADS;                      // staged row of this thread's quad
  const int q0 = (ct % QUADS) * 4;                   // first of its 4 strip columns
  const int ox = x_base + q0;                        // image x of that column
  // columns of the quad that are real outputs of this strip
  bool valid[4];
#pragma unroll
  for (int j = 0; j < 4; j++) valid[j] = q0 + j >= m && q0 + j < COLS - m && ox + j < w;
  const bool any_valid = valid[0] || valid[1] || valid[2] || valid[3];
  const bool all_valid = valid[0] && valid[1] && valid[2] && valid[3];

  for (int c = 0; c < n_chunks; c++) {
    const int buf = c & 1;
    const int yo = y0 + c * CH + q_row;
    named_bar_sync(BAR_FULL0 + buf, NT);             // producers finished staging chunk c
#if defined(OFB_DBG) && (OFB_DBG & 4)     // experiment: consumers only hand the buffers back
    if (c == n_chunks - 1 && any_valid) fout[(unsigned)(y0 * w + max(ox, 0))] = make_float2(stage[ct], 0.f);
    if (c + 2 < n_chunks) named_bar_arrive(BAR_EMPTY0 + buf, NT);
    continue;
#endif
    if (yo < y1 && any_valid) {
      const float* srow = stage + (buf * CH + q_row) * 5 * COLS;
      float sum[5][4];
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        const float* s = srow + ch * COLS;
        float s0, s1, s2, s3;
        if (MT > 0) {
          constexpr int KQ = (MT + 3) / 4;
          float e[(2 * KQ + 1) * 4];                 // e[d + 4*KQ] = staged value at column q0 + d
#pragma unroll
          for (int kk = -KQ; kk <= KQ; kk++) {
            const int cq = min(max(q0 + 4 * kk, 0), COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            e[(kk + KQ) * 4 + 0] = v.x; e[(kk + KQ) * 4 + 1] = v.y; e[(kk + KQ) * 4 + 2] = v.z; e[(kk + KQ) * 4 + 3] = v.w;
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
          for (int kk = -kq; kk <= kq; kk++) {
            const int cq = min(max(q0 + 4 * kk, 0), COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int d = 4 * kk + i;
              if (d >= 0 - m && d <= 0 + m) s0 += e[i];
              if (d >= 1 - m && d <= 1 + m) s1 += e[i];
              if (d >= 2 - m && d <= 2 + m) s2 += e[i];
              if (d >= 3 - m && d <= 3 + m) s3 += e[i];
            }
          }
        }
        sum[ch][0] = s0; sum[ch][1] = s1; sum[ch][2] = s2; sum[ch][3] = s3;
      }
      float2 f[4];
#pragma unroll
      for (int j = 0; j < 4; j++) f[j] = solve2x2_sums(sum[0][j], sum[1][j], sum[2][j], sum[3][j], sum[4][j], reg);
      const int oi = yo * w + ox;                    // (oi + j >= 0 for every valid column j)
      float2* o = fout + (unsigned)max(oi, 0);
      if (all_valid && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
        // 16-byte aligned: two 128-bit stores
        *reinterpret_cast<float4*>(o) = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
        *reinterpret_cast<float4*>(o + 2) = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (valid[j]) fout[(unsigned)(oi + j)] = f[j];
      }
    }
    if (c + 2 < n_chunks) named_bar_arrive(BAR_EMPTY0 + buf, NT);   // staging buffer may be refilled
  }
}

}  // namespace ofb
