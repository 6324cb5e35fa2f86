// fb_iter_ws2.cuh — warp-specialised fused Farneback iteration kernel (box window), second version.
//
// Same arithmetic and HBM traffic as k_iter_box / k_iter_ws: UpdateMatrices + (2m+1)^2 box blur +
// 2x2 solve in one pass over the level, 56 B per pixel-iteration.  ncu showed k_iter_ws bound by
// exposed gather latency in the producer warps (long-scoreboard stalls on the first use of the R1
// corners) and by instruction count.  Against it:
//   * L2 prefetch one chunk ahead: while a producer warp issues the loads of chunk c it also issues
//     prefetch.global.L2 for the lines chunk c+1 will touch for the first time (R0 two rows down, the
//     R1 corner row three rows below this pixel's gather — the flow field is smooth, so the prediction
//     holds), so the demand loads of the next chunk hit in L2 (~300 cycles) instead of HBM.
//     (Register-level software pipelining does not work here: ptxas puts every LDG of the loop on
//     one counting scoreboard, so waiting for the old loads also waits for the ones just issued —
//     measured 1.6x SLOWER.  Prefetches write no register and need no scoreboard.)
//   * 32-bit unsigned element offsets from per-pair base pointers pinned in (uniform) registers — no
//     per-access 64-bit index arithmetic —, floor via F2I + I2FP instead of FRND + F2I, cv2's
//     unsigned inside test, border attenuation only on border pixels, winsize^-2 folded into the
//     regulariser of the solve and an approximate reciprocal (1 ulp) there.
#pragma once
#include "fb_device.cuh"
#include "fb_iter_ws.cuh"

namespace ofb {

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct UmLoads2 {
  float4 a0, q00, q01, q10, q11;
  float b0, s00, s01, s10, s11;
  float fx, fy, dx, dy;
  bool inside;
};

// Starts the loads of one pixel: R0 at (x, y) and the 2x2 neighbourhood of R1 at floor((x,y) + flow).
// yw = y * w.  Outside pixels gather from (0,0) (in bounds, result discarded).
// PF: also prefetch into L2 what the same lane will load one chunk (WS_CH rows) later.  The R buffers
// are allocated with kRowPad spare rows, so the prefetch addresses stay inside the allocation.
template <bool PF>
__device__ __forceinline__ void um_issue2(UmLoads2& L, const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                          const float4* __restrict__ RA1, const float* __restrict__ RB1, float2 fl,
                                          int x, int y, unsigned yw, unsigned w, unsigned h) {
  const unsigned o = yw + (unsigned)x;
  L.a0 = __ldg(RA0 + o);
  L.b0 = __ldg(RB0 + o);
  L.dx = fl.x;
  L.dy = fl.y;
  const float fx = (float)x + fl.x, fy = (float)y + fl.y;
  const int ix = __float2int_rd(fx), iy = __float2int_rd(fy);
  L.fx = fx - (float)ix;
  L.fy = fy - (float)iy;
  // cv2: (unsigned)x1 < (unsigned)(w-1) && (unsigned)y1 < (unsigned)(h-1)  (F2I saturates, so huge
  // flows stay outside)
  L.inside = (unsigned)ix < w - 1u && (unsigned)iy < h - 1u;
  const unsigned g = L.inside ? (unsigned)iy * w + (unsigned)ix : 0u;
  const float4* pa = RA1 + g;
  const float* pb = RB1 + g;
  L.q00 = __ldg(pa);
  L.q01 = __ldg(pa + 1);
  L.q10 = __ldg(pa + w);
  L.q11 = __ldg(pa + w + 1);
  L.s00 = __ldg(pb);
  L.s01 = __ldg(pb + 1);
  L.s10 = __ldg(pb + w);
  L.s11 = __ldg(pb + w + 1);
  if (PF) {
    prefetch_l2(RA0 + (o + WS_CH * w));
    prefetch_l2(RB0 + (o + WS_CH * w));
    prefetch_l2(pa + (WS_CH + 1) * w);
    prefetch_l2(pb + (WS_CH + 1) * w);
  }
}

// border: this pixel lies within 5 px of the level border (attenuation table applies)
__device__ __forceinline__ M5 um_finish2(const UmLoads2& L, bool border, int x, int y, int w, int h) {
  const float fx = L.fx, fy = L.fy;
  const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
  float r2 = a00 * L.q00.x + a01 * L.q01.x + a10 * L.q10.x + a11 * L.q11.x;
  float r3 = a00 * L.q00.y + a01 * L.q01.y + a10 * L.q10.y + a11 * L.q11.y;
  float r4 = a00 * L.q00.z + a01 * L.q01.z + a10 * L.q10.z + a11 * L.q11.z;
  float r5 = a00 * L.q00.w + a01 * L.q01.w + a10 * L.q10.w + a11 * L.q11.w;
  float r6 = a00 * L.s00 + a01 * L.s01 + a10 * L.s10 + a11 * L.s11;
  if (L.inside) {
    r4 = (L.a0.z + r4) * 0.5f;
    r5 = (L.a0.w + r5) * 0.5f;
    r6 = (L.b0 + r6) * 0.25f;
  } else {
    r2 = r3 = 0.f;
    r4 = L.a0.z;
    r5 = L.a0.w;
    r6 = L.b0 * 0.5f;
  }
  r2 = (L.a0.x - r2) * 0.5f;
  r3 = (L.a0.y - r3) * 0.5f;
  r2 += r4 * L.dy + r6 * L.dx;
  r3 += r6 * L.dy + r5 * L.dx;
  if (border) {
    const float s = border_w(x, w) * border_w(y, h);
    r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
  }
  M5 m;
  m.g11 = r4 * r4 + r6 * r6;
  m.g12 = (r4 + r5) * r6;
  m.g22 = r5 * r5 + r6 * r6;
  m.h1 = r4 * r2 + r6 * r3;
  m.h2 = r6 * r2 + r5 * r3;
  return m;
}

// sums are unscaled window sums; reg = 1e-3 / scale^2 (scale = winsize^-2 folded into the regulariser)
__device__ __forceinline__ float2 solve2x2_sums(float g11, float g12, float g22, float h1, float h2, float reg) {
  const float idet = rcp_approx(g11 * g22 - g12 * g12 + reg);
  return make_float2((g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet);
}

constexpr int WS2_PROD_REGS = 152;   // 128 * (152 + 104) = 2^15 = half the register file: 2 CTAs per SM
constexpr int WS2_CONS_REGS = 104;

template <int MT, bool PF>
__global__ void __launch_bounds__(WS_THREADS, 2)
    k_iter_ws2(const float4* __restrict__ RA, const float* __restrict__ RB, const float2* __restrict__ flow_in,
               float2* __restrict__ flow_out, int w, int h, int f1_offset, int m_rt, float reg, int seg_rows,
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
    // registers follow the roles: four pixels of gather results per lane = 120 live registers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS2_PROD_REGS));
    const float4* RA0 = RA + (size_t)pair * n;
    const float* RB0 = RB + (size_t)pair * n;
    const float4* RA1 = RA + (size_t)(pair + f1_offset) * n;
    const float* RB1 = RB + (size_t)(pair + f1_offset) * n;
    const float2* fin = flow_in + (size_t)pair * n;
    // Keep the five per-pair base pointers in registers: without this the compiler re-derives
    // pair * n + offset in 64 bits for every access (~25 instructions per pixel).
    asm volatile("" : "+l"(RA0), "+l"(RB0), "+l"(RA1), "+l"(RB1), "+l"(fin));
    const int a_row = warp >> 1, a_half = warp & 1;
    int soff = a_row * 5 * COLS + a_half * 128 + lane;   // this lane's element offset inside a staging buffer
    asm volatile("" : "+r"(soff));
    const unsigned uw = (unsigned)w, uh = (unsigned)h;
    int xs[4];
    unsigned xborder = 0;                            // bit j: column xs[j] is within 5 px of the border
#pragma unroll
    for (int j = 0; j < 4; j++) {
      xs[j] = clampi(x_base + a_half * 128 + lane + 32 * j, 0, w - 1);
      if ((unsigned)(xs[j] - 5) >= (unsigned)(w - 10)) xborder |= 1u << j;
    }
    asm volatile("" : "+r"(xborder));
    float2 fl[4];
    {
      const unsigned yw = (unsigned)clampi(t_first + a_row, 0, h - 1) * uw;
#pragma unroll
      for (int j = 0; j < 4; j++) fl[j] = __ldg(fin + (yw + (unsigned)xs[j]));
    }
    for (int c = 0; c < n_chunks; c++) {
      const int buf = c & 1;
      const int t = t_first + c * WS_CH + a_row;
      if (c >= 2) named_bar_sync(WS_BAR_EMPTY0 + buf, WS_THREADS);   // consumers released this buffer
      if (t <= t_last) {
        const int y = clampi(t, 0, h - 1);
        const unsigned yw = (unsigned)y * uw;
        const bool yborder = (unsigned)(y - 5) >= (unsigned)(h - 10);
        float* srow = stage + buf * WS_CH * 5 * COLS + soff;
        UmLoads2 L[4];
#if defined(OFB_DBG) && (OFB_DBG & 2)    // experiment: producers do no loads and no arithmetic
#pragma unroll
        for (int j = 0; j < 4; j++) {
          srow[0 * COLS + 32 * j] = 1.f; srow[1 * COLS + 32 * j] = 0.f; srow[2 * COLS + 32 * j] = 1.f;
          srow[3 * COLS + 32 * j] = 0.5f; srow[4 * COLS + 32 * j] = 0.25f;
        }
        named_bar_arrive(WS_BAR_FULL0 + buf, WS_THREADS);
        continue;
#endif
#if defined(OFB_DBG) && (OFB_DBG & 1)    // experiment: zero displacement (perfectly coalesced gathers)
#pragma unroll
        for (int j = 0; j < 4; j++) fl[j] = make_float2(0.f, 0.f);
#endif
#pragma unroll
        for (int j = 0; j < 4; j++) um_issue2<PF>(L[j], RA0, RB0, RA1, RB1, fl[j], xs[j], y, yw, uw, uh);
        {  // next chunk's flow (clamped row: always a valid address)
          const unsigned ywn = (unsigned)clampi(t + WS_CH, 0, h - 1) * uw;
#pragma unroll
          for (int j = 0; j < 4; j++) fl[j] = __ldg(fin + (ywn + (unsigned)xs[j]));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const M5 mm = um_finish2(L[j], yborder || ((xborder >> j) & 1u), xs[j], y, w, h);
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
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS2_CONS_REGS));
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
#if defined(OFB_DBG) && (OFB_DBG & 4)    // experiment: consumers only hand the buffers back
    if (c + 2 < n_chunks) named_bar_arrive(WS_BAR_EMPTY0 + buf, WS_THREADS);
    if (c == n_chunks - 1 && validA) fout[(unsigned)(y0 * w + x_base + colA)] = make_float2(stage[ct], 0.f);
    continue;
#endif
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
          const int o = y * w + x_base;                // (o + col >= 0 for valid columns)
          if (validA)
            fout[(unsigned)(o + colA)] = solve2x2_sums((float)va[0], (float)va[1], (float)va[2], (float)va[3], (float)va[4], reg);
          if (validB)
            fout[(unsigned)(o + colB)] = solve2x2_sums((float)vb[0], (float)vb[1], (float)vb[2], (float)vb[3], (float)vb[4], reg);
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
