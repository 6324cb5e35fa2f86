// fb_iter.cuh — fused Farneback iteration kernel (box window):
// FarnebackUpdateMatrices + FarnebackUpdateFlow_Blur in ONE pass over the level.
#pragma once
#include "fb_device.cuh"

namespace ofb {

// HBM traffic per pixel-iteration = R0 (20 B) + R1 gather (20 B) + flow in (8 B) + flow out (8 B).
//
// A CTA (8 warps) owns a strip of COLS = 256 matrix columns (2m of them halo) and a segment of
// `seg_rows` output rows, and marches down it FI_CH = 4 matrix rows at a time:
//   A1  every warp computes M for one half-row (4 px per lane, 32 px apart: coalesced R0/flow loads
//       and L1-friendly gathers; all 40 loads of the 4 px are issued before the first use, and the
//       flow of the next chunk is prefetched) into a shared staging row;
//   A2  horizontal window sums H: each lane owns 4 adjacent columns, reads the 2m+4 staged values
//       it needs as float4s and writes H into a ring of 2m+1 rows in shared memory;
//   B   one thread per column keeps the vertical window sum as a running sum in DOUBLE (add the
//       new H row, subtract the row leaving the window — exactly cv2's vsum scheme, so there is no
//       float cancellation drift), scales, solves the 2x2 system and writes flow.
constexpr int FI_CH = 4;

// COLS = strip width = threads per CTA (128 or 256); 512 threads per SM either way.
template <int MT, int COLS>
__global__ void __launch_bounds__(COLS, 512 / COLS)
    k_iter_box(const float4* __restrict__ RA, const float* __restrict__ RB, const float2* __restrict__ flow_in,
               float2* __restrict__ flow_out, int w, int h, int f1_offset, int m_rt, float scale, int seg_rows,
               int strips) {
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int tw = COLS - 2 * m;
  extern __shared__ float smem[];
  float* stage = smem;                          // [FI_CH][5][COLS]
  float* ring = smem + FI_CH * 5 * COLS;     // [R][5][COLS]

  const int strip = blockIdx.x % strips;
  const int seg = blockIdx.x / strips;
  const int pair = blockIdx.y;
  const int x_base = strip * tw - m;            // image x of strip column 0
  const int y0 = seg * seg_rows;
  const int y1 = min(y0 + seg_rows, h);         // exclusive
  const int t_first = y0 - m, t_last = y1 - 1 + m;

  const size_t n = (size_t)w * h;
  const float4* RA0 = RA + (size_t)pair * n;
  const float* RB0 = RB + (size_t)pair * n;
  const float4* RA1 = RA + (size_t)(pair + f1_offset) * n;
  const float* RB1 = RB + (size_t)(pair + f1_offset) * n;
  const float2* fin = flow_in + (size_t)pair * n;
  float2* fout = flow_out + (size_t)pair * n;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int HALVES = COLS / 128;            // warps per matrix row
  const int a_row = warp / HALVES, a_half = warp % HALVES;
  const int col = tid;                          // phase-B column
  const int out_x = x_base + col;
  const bool col_valid = col >= m && col < COLS - m && out_x < w;

  double vs0 = 0, vs1 = 0, vs2 = 0, vs3 = 0, vs4 = 0;

  // flow of the 4 pixels this lane handles in A1, prefetched one chunk ahead so that the
  // flow -> gather dependency never costs two memory latencies back to back
  float2 fl[4];
  int xs[4];
#pragma unroll
  for (int j = 0; j < 4; j++) xs[j] = clampi(x_base + a_half * 128 + lane + 32 * j, 0, w - 1);
  {
    const int y = clampi(t_first + a_row, 0, h - 1);
#pragma unroll
    for (int j = 0; j < 4; j++) fl[j] = __ldg(fin + y * w + xs[j]);
  }

  int slot0 = 0;  // ring slot of row tc
  for (int tc = t_first; tc <= t_last; tc += FI_CH) {
    // ---------------- A1: matrices of row tc + a_row -> staging
    const int t = tc + a_row;
    if (t <= t_last) {
      const int y = clampi(t, 0, h - 1);
      float* srow = stage + a_row * 5 * COLS + a_half * 128 + lane;
      UmLoads L[4];
#if defined(OFB_DBG) && (OFB_DBG & 2)   // experiment: zero displacement (perfectly coalesced gathers)
#pragma unroll
      for (int j = 0; j < 4; j++) fl[j] = make_float2(0.f, 0.f);
#endif
#pragma unroll
      for (int j = 0; j < 4; j++) um_issue(L[j], RA0, RB0, RA1, RB1, fl[j], xs[j], y, w, h);
      {  // next chunk's flow (clamped row: always a valid address)
        const int yn = clampi(t + FI_CH, 0, h - 1);
#pragma unroll
        for (int j = 0; j < 4; j++) fl[j] = __ldg(fin + yn * w + xs[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
#if defined(OFB_DBG) && (OFB_DBG & 8)   // experiment: no R loads / UpdateMatrices arithmetic
        M5 mm;
        mm.g11 = fl[j].x; mm.g12 = fl[j].y; mm.g22 = 1.f; mm.h1 = fl[j].x; mm.h2 = fl[j].y;
#else
        const M5 mm = um_finish(L[j], xs[j], y, w, h);
#endif
        srow[0 * COLS + 32 * j] = mm.g11;
        srow[1 * COLS + 32 * j] = mm.g12;
        srow[2 * COLS + 32 * j] = mm.g22;
        srow[3 * COLS + 32 * j] = mm.h1;
        srow[4 * COLS + 32 * j] = mm.h2;
      }
    }
    // ring slots of the chunk's rows (slot = (row - t_first) mod R, kept incrementally)
    int slot[FI_CH];
#pragma unroll
    for (int rr = 0; rr < FI_CH; rr++) {
      int sl = slot0 + rr;
      slot[rr] = sl >= R ? sl - R : sl;
    }
    // rows leaving the window: their ring slots are overwritten in A2, so fetch them now
    float old[FI_CH][5];
    const int n_done = tc - t_first;            // rows already in the ring
#pragma unroll
    for (int rr = 0; rr < FI_CH; rr++) {
      const bool have_old = n_done + rr >= R;   // row tc+rr-R exists (its slot is slot[rr]; R > FI_CH)
#pragma unroll
      for (int ch = 0; ch < 5; ch++) old[rr][ch] = have_old ? ring[(slot[rr] * 5 + ch) * COLS + col] : 0.f;
    }
    __syncthreads();
    // ---------------- A2: horizontal window sums of the staged rows -> ring
    if (t <= t_last) {
      const float* srow = stage + a_row * 5 * COLS;
      float* rrow = ring + slot[0] * 5 * COLS;
      if (a_row == 1) rrow = ring + slot[1] * 5 * COLS;
      if (a_row == 2) rrow = ring + slot[2] * 5 * COLS;
      if (a_row == 3) rrow = ring + slot[3] * 5 * COLS;
      const int q0 = a_half * 128 + 4 * lane;   // first of this lane's 4 columns
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        const float* s = srow + ch * COLS;
        float s0, s1, s2, s3;
        if (MT > 0) {
          constexpr int KQ = (MT + 3) / 4;
          float e[(2 * KQ + 1) * 4];            // e[d + 4*KQ] = staged value at column q0 + d
#pragma unroll
          for (int k = -KQ; k <= KQ; k++) {
            const int cq = min(max(q0 + 4 * k, 0), COLS - 4);
            const float4 v = *reinterpret_cast<const float4*>(s + cq);
            e[(k + KQ) * 4 + 0] = v.x; e[(k + KQ) * 4 + 1] = v.y; e[(k + KQ) * 4 + 2] = v.z; e[(k + KQ) * 4 + 3] = v.w;
          }
          constexpr int O = 4 * KQ;
          // d in [3-MT, MT] lies inside all four windows: sum once; then suffix sums on the left
          // (d = 2-MT .. -MT) and prefix sums on the right (d = MT+1 .. MT+3)
          float core = e[O + 3 - MT];
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
          const int kq = (m + 3) >> 2;          // quads to each side
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
#if defined(OFB_DBG) && (OFB_DBG & 1)   // experiment: no horizontal sums (dead-code-eliminates the above)
        {
          const float4 v = *reinterpret_cast<const float4*>(s + q0);
          s0 = v.x; s1 = v.y; s2 = v.z; s3 = v.w;
        }
#endif
        *reinterpret_cast<float4*>(rrow + ch * COLS + q0) = make_float4(s0, s1, s2, s3);
      }
    }
    __syncthreads();
    // ---------------- B: vertical running sums (double) + solve
    const int nrows = min(FI_CH, t_last - tc + 1);
#pragma unroll
    for (int rr = 0; rr < FI_CH; rr++) {
      if (rr < nrows) {
        const float* rrow = ring + slot[rr] * 5 * COLS + col;
#if defined(OFB_DBG) && (OFB_DBG & 4)   // experiment: no double running sums / solve
        if (tc + rr - m >= y0 && col_valid) fout[(tc + rr - m) * w + out_x] = make_float2(rrow[0], rrow[3 * COLS] + old[rr][1]);
        continue;
#endif
        vs0 += (double)rrow[0 * COLS] - (double)old[rr][0];
        vs1 += (double)rrow[1 * COLS] - (double)old[rr][1];
        vs2 += (double)rrow[2 * COLS] - (double)old[rr][2];
        vs3 += (double)rrow[3 * COLS] - (double)old[rr][3];
        vs4 += (double)rrow[4 * COLS] - (double)old[rr][4];
        const int y = tc + rr - m;
        if (y >= y0 && col_valid) {
          fout[y * w + out_x] = solve2x2((float)vs0 * scale, (float)vs1 * scale, (float)vs2 * scale,
                                         (float)vs3 * scale, (float)vs4 * scale);
        }
      }
    }
    slot0 += FI_CH;
    if (slot0 >= R) slot0 -= R;
  }
}

}  // namespace ofb
