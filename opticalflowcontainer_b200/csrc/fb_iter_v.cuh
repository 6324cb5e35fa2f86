// fb_iter_v.cuh — fused Farneback iteration kernel (box window), vertical-first, no FP64.
//
// UpdateMatrices + (2m+1)^2 box blur + 2x2 solve in one pass over the level, 56 B of HBM traffic
// per pixel-iteration.  ncu on the earlier kernels (k_iter_ws and two successors) showed
// the XU pipe at 75 % "realtime": they kept cv2's vertical running sums in double, and the
// F2F.F64.F32 / F2F.F32.F64 conversions around them (15 per pixel) run at a fraction of the FP32 rate
// on B200.  This kernel removes FP64 altogether:
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
// CTA = COLS producer threads + CH*COLS/PXT consumer threads (PXT = 4 or 8 adjacent pixels each); FULL/EMPTY named barriers hand the
// double-buffered staging rows (CH output rows per chunk) over.
#pragma once
#include "fb_device.cuh"
#include "fb_iter_ws.cuh"   // named_bar_sync / named_bar_arrive
#include "fb_um.cuh"

namespace ofb {

// REUSE variant: the kernel is launched with 80 registers/thread (2 CTAs of 384 threads per SM); the consumer
// warpgroup hands 24 of them back and the two producer warpgroups take 12 more each (setmaxnreg), so the
// persistent corner rows do not push the producers' loads behind their first use.
constexpr int REUSE_PROD_REGS = 88, REUSE_CONS_REGS = 56;

template <int COLS, int CH>
constexpr int iter_v_smem_floats(int m) { return (2 * CH + 2 * m + 1) * 5 * COLS; }

// PFD > 0: every producer thread also issues prefetch.global.L2 for what it will load PFD rows later
// (R0, flow, and the new corner row of the R1 gather — the flow field is smooth, so "same
// displacement, PFD rows down" predicts it).  A producer has only one row of loads in flight, so
// without this each row pays a full HBM round trip (~1 us under load); with it the demand loads
// hit in L2.  Prefetches write no register and use no scoreboard (a register-level software
// pipeline does not work: ptxas puts every LDG of the loop on one counting scoreboard).  Measured
// -10 % kernel time; a cp.async.bulk.prefetch.L2 variant (5 requests per strip row) measured +2 %.
// The R buffers carry kRowPad spare rows so the predicted corner row stays inside the allocation.
// RIF = matrix rows a producer thread keeps in flight (1, or CH: the loads of all CH rows of a chunk are
// issued before the first is consumed).  CLOOP = rows a consumer thread handles per chunk (1, or CH).
// TILED (spatially tiled mode, one pair): the CTA grid covers only the level rows [y_begin, y_end) this
// rank owns; R0 / R1 / flow rows owned by other ranks (the 2m halo rows of the blur and whatever the
// displacement reaches) are read from those ranks' buffers through the NVLink peer pointers in `tab` —
// the halo exchange is these loads, issued tile by tile inside the kernel that consumes them.
// REUSE: row-reuse gather (fb_um.cuh): the top corner row of a pixel is taken from the registers of the pixel
// above when the displacement allows it; needs RIF == 1, CH even, untiled.
template <int MT, int COLS, int CH, int MINB, int PFD, int PXT, int RIF = 1, int CLOOP = 1, bool TILED = false,
          bool REUSE = false>
__global__ void __launch_bounds__(COLS + (CH / CLOOP) * COLS / PXT, MINB)
    k_iter_v(const RSet rs, const float2* __restrict__ flow_in,
             float2* __restrict__ flow_out, int w, int h, int m_rt, float reg, int seg_rows,
             int strips, int y_begin, int y_end, PeerTab tab, int my_rank) {
  static_assert(PXT == 4 || PXT == 8, "4 or 8 adjacent pixels per consumer thread");
  constexpr int GROUPS = COLS / PXT;
  static_assert(RIF == 1 || RIF == CH, "rows in flight: 1 or the whole chunk");
  static_assert(CLOOP == 1 || CLOOP == CH, "consumer rows per thread: 1 or the whole chunk");
  constexpr int NCONS = (CH / CLOOP) * GROUPS;
  static_assert(NCONS % 32 == 0, "whole consumer warps");
  static_assert(!REUSE || (!TILED && CH == 2 && (RIF == 1 || RIF == 2)), "row-reuse gather: chunks of two rows, untiled");
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
  const int y0 = y_begin + seg * seg_rows;
  const int y1 = min(y0 + seg_rows, y_end);          // exclusive
  const int t_first = y0 - m;                        // first matrix row the segment needs
  const int n_chunks = (y1 - y0 + CH - 1) / CH;

  const size_t n = (size_t)w * h;
  const int tid = threadIdx.x;

  if (tid < COLS) {
    // ------------------------------------------------------------------ PRODUCERS (one column each)
    if constexpr (REUSE && RIF == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REUSE_PROD_REGS));
    else if constexpr (RIF > 1 && COLS == 256 && PXT == 4 && CLOOP == 1 && MINB == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(96));
    const float4* RA0 = rs.A0 + (size_t)pair * n;
    const float* RB0 = rs.B0 + (size_t)pair * n;
    const float4* RA1 = rs.A1 + (size_t)pair * n;
    const float* RB1 = rs.B1 + (size_t)pair * n;
    const float2* fin = flow_in + (size_t)pair * n;
    asm volatile("" : "+l"(RA0), "+l"(RB0), "+l"(RA1), "+l"(RB1), "+l"(fin));   // keep the bases, do not re-derive
    const unsigned uw = (unsigned)w, uh = (unsigned)h;
    const int x = clampi(x_base + tid, 0, w - 1);
    const bool xborder = (unsigned)(x - 5) >= (unsigned)(w - 10);
    float* rcol = ring + tid;                        // ring[k][ch][tid]
    float P[5] = {0.f, 0.f, 0.f, 0.f, 0.f};          // prefix sums of the current block
    float Bp[5] = {0.f, 0.f, 0.f, 0.f, 0.f};         // sum of the previous block
    int k = 0;                                       // offset of row t inside its block
    bool have_prev = false;

    // start the loads of matrix row t (flow of the row in `f`), prefetch PFD rows ahead
    auto issue = [&](UmLoads2& L, float2 f, int t) {
      const int y = clampi(t, 0, h - 1);
      const unsigned yw = (unsigned)y * uw;
#if defined(OFB_DBG) && (OFB_DBG & 1)     // experiment: zero displacement (perfectly regular gathers)
      f = make_float2(0.f, 0.f);
#endif
#if defined(OFB_DBG) && (OFB_DBG & 8)     // experiment: no R1 gather (R0 values stand in for the corners)
      {
        const unsigned o = yw + (unsigned)x;
        L.a0 = __ldg(RA0 + o); L.b0 = __ldg(RB0 + o);
        L.q00 = L.q01 = L.q10 = L.q11 = L.a0; L.s00 = L.s01 = L.s10 = L.s11 = L.b0;
        L.dx = f.x; L.dy = f.y; L.fx = 0.25f; L.fy = 0.5f; L.inside = true;
      }
#else
      if constexpr (TILED) {
        const int ro = (y >= tab.r_lo && y < tab.r_hi) ? my_rank : tile_owner(y, tab);   // uniform over the CTA
        um_issue2_tiled(L, tab.RA[ro], tab.RB[ro], tab, n, my_rank, f, x, y, yw, uw, uh);
      } else {
        um_issue2(L, RA0, RB0, RA1, RB1, f, x, y, yw, uw, uh);
      }
#endif
      if (PFD > 0) {   // (tiled mode: RA0.. are this rank's own buffers — rows of a neighbour are simply not prefetched usefully)
        static_assert(PFD + 1 <= kRowPad, "prefetch distance exceeds the row padding of the R buffers");
        const unsigned op = (unsigned)clampi(t + PFD, 0, h - 1) * uw + (unsigned)x;
        prefetch_l2(RA0 + op);
        prefetch_l2(RB0 + op);
        prefetch_l2(fin + ((unsigned)clampi(t + PFD + RIF, 0, h - 1) * uw + (unsigned)x));
        const unsigned g = L.inside ? (unsigned)__float2int_rd((float)y + L.dy) * uw + (unsigned)__float2int_rd((float)x + L.dx) : 0u;
        prefetch_l2(RA1 + (g + (PFD + 1) * uw));
        prefetch_l2(RB1 + (g + (PFD + 1) * uw));
      }
    };
    // matrix row t from its loads: M(t) -> P += M; V = (Bp - P_prev[k]) + P; ring[k] = P; block bookkeeping
    auto finish = [&](const UmLoads2& L, int t, float (&V)[5]) {
      const int y = clampi(t, 0, h - 1);
      float old[5];
#pragma unroll
      for (int ch = 0; ch < 5; ch++) old[ch] = have_prev ? rcol[(k * 5 + ch) * COLS] : 0.f;
#if defined(OFB_DBG) && (OFB_DBG & 2)     // experiment: no UpdateMatrices (and, with the loads unused, no loads)
      M5 mm; mm.g11 = 1.f; mm.g12 = 0.f; mm.g22 = 1.f; mm.h1 = 0.5f; mm.h2 = 0.25f;
#else
      const M5 mm = um_finish2(L, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h);
#endif
      P[0] = __fadd_rn(P[0], mm.g11); P[1] = __fadd_rn(P[1], mm.g12); P[2] = __fadd_rn(P[2], mm.g22);
        P[3] = __fadd_rn(P[3], mm.h1); P[4] = __fadd_rn(P[4], mm.h2);
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
    auto flow_at = [&](int t) {
      const int yc = clampi(t, 0, h - 1);
      const float2* f = TILED ? tab.flow[(yc >= tab.f_lo && yc < tab.f_hi) ? my_rank : tile_owner(yc, tab)] : fin;
      if constexpr (RIF > 1) {
        // volatile: keeps the load where it is written (between the gathers and the barrier); ptxas otherwise
        // sinks it to the end of the loop body, right in front of the address arithmetic that needs it
        float2 v;
        asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(f + ((unsigned)yc * uw + (unsigned)x)));
        return v;
      }
      return __ldg(f + ((unsigned)yc * uw + (unsigned)x));
    };

    if constexpr (REUSE && RIF == 2) {
      // Two rows in flight AND row-reuse gather.  Rows A = t, B = t + 1 of a chunk: all loads of both are issued
      // before the first is consumed.  Corner-row register sets: X = top of A, Y = bottom of A (and top of B when B
      // sits exactly one row below A), W = top of B otherwise, Z = bottom of B.  The next chunk's A takes Z as its top
      // row when the displacement allows it (one register copy per chunk), so a smooth field costs two corner-row
      // loads per chunk-row pair... i.e. 8 gather loads per chunk instead of 16.
      UmRow X, Y, Z, W;
      UmPix pa, pb;
      unsigned prev_g = ~0u - uw;
      bool reuse_b = false;
      Z.q0 = Z.q1 = make_float4(0.f, 0.f, 0.f, 0.f); Z.s0 = Z.s1 = 0.f;
      W = Z;
      auto issue2 = [&](float2 fa, float2 fb, int t) {
        const int ya = clampi(t, 0, h - 1), yb = clampi(t + 1, 0, h - 1);
        X = Z;                                                     // bottom row of the previous chunk's row B
        um_pix(pa, RA0, RB0, fa, x, ya, (unsigned)ya * uw, uw, uh);
        if (pa.g != prev_g + uw) um_row_load(X, RA1, RB1, pa.g);
        um_row_load(Y, RA1, RB1, pa.g + uw);
        um_pix(pb, RA0, RB0, fb, x, yb, (unsigned)yb * uw, uw, uh);
        reuse_b = pa.inside && pb.g == pa.g + uw;
        if (!reuse_b) um_row_load(W, RA1, RB1, pb.g);
        um_row_load(Z, RA1, RB1, pb.g + uw);
        prev_g = pb.inside ? pb.g : ~0u - uw;
      };
      auto ring_step = [&](const M5& mm, float (&V)[5]) {
        float old[5];
#pragma unroll
        for (int ch = 0; ch < 5; ch++) old[ch] = have_prev ? rcol[(k * 5 + ch) * COLS] : 0.f;
        P[0] = __fadd_rn(P[0], mm.g11); P[1] = __fadd_rn(P[1], mm.g12); P[2] = __fadd_rn(P[2], mm.g22);
        P[3] = __fadd_rn(P[3], mm.h1); P[4] = __fadd_rn(P[4], mm.h2);
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
      auto finish_a = [&](int t, float (&V)[5]) {
        const int y = clampi(t, 0, h - 1);
        ring_step(um_finish_rows(pa, X, Y, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h), V);
      };
      auto finish_b = [&](int t, float (&V)[5]) {
        const int y = clampi(t, 0, h - 1);
        if (reuse_b) W = Y;                                        // (select per thread: 10 predicated moves)
        ring_step(um_finish_rows(pb, W, Z, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h), V);
      };
      float2 fa = flow_at(t_first), fb = flow_at(t_first + 1);
      // warm-up: the R-1 = 2m rows above the first output row, chunk by chunk
      for (int t = t_first; t < t_first + R - 1; t += 2) {
        float V[5];
        issue2(fa, fb, t);
        fa = flow_at(t + 2); fb = flow_at(t + 3);
        finish_a(t, V);
        finish_b(t + 1, V);
      }
      for (int c = 0; c < n_chunks; c++) {
        const int buf = c & 1;
        const int tc = y0 + c * CH + m;                              // newest matrix row of output row y0 + c*CH
        issue2(fa, fb, tc);
        fa = flow_at(tc + 2); fb = flow_at(tc + 3);
        if (c >= 2) named_bar_sync(BAR_EMPTY0 + buf, NT);            // consumers released this buffer
        float* srow = stage + buf * CH * 5 * COLS + tid;
        float V[5];
        finish_a(tc, V);
#pragma unroll
        for (int ch = 0; ch < 5; ch++) srow[ch * COLS] = V[ch];
        finish_b(tc + 1, V);                                         // (a row past y1 keeps the state consistent; never read)
#pragma unroll
        for (int ch = 0; ch < 5; ch++) srow[(5 + ch) * COLS] = V[ch];
        named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
      }
    } else if constexpr (REUSE) {
      // Row-reuse gather.  Two corner-row register sets alternate as "top" and "bottom" from one row to the
      // next (rows are handled in pairs, so the alternation is in the register names, not in moves).
      UmRow ra, rb;
      UmPix px;
      unsigned prev_g = ~0u - uw;
      auto issue_r = [&](UmRow& top, UmRow& bot, float2 f, int t) {
        const int y = clampi(t, 0, h - 1);
        um_issue_rows(px, top, bot, prev_g, RA0, RB0, RA1, RB1, f, x, y, (unsigned)y * uw, uw, uh);
        if (PFD > 0) {
          static_assert(PFD + 1 <= kRowPad, "prefetch distance exceeds the row padding of the R buffers");
          const unsigned op = (unsigned)clampi(t + PFD, 0, h - 1) * uw + (unsigned)x;
          prefetch_l2(RA0 + op);
          prefetch_l2(RB0 + op);
          prefetch_l2(fin + ((unsigned)clampi(t + PFD + 1, 0, h - 1) * uw + (unsigned)x));
          prefetch_l2(RA1 + (px.g + (PFD + 1) * uw));
          prefetch_l2(RB1 + (px.g + (PFD + 1) * uw));
        }
      };
      auto finish_r = [&](const UmRow& top, const UmRow& bot, int t, float (&V)[5]) {
        const int y = clampi(t, 0, h - 1);
        float old[5];
#pragma unroll
        for (int ch = 0; ch < 5; ch++) old[ch] = have_prev ? rcol[(k * 5 + ch) * COLS] : 0.f;
        const M5 mm = um_finish_rows(px, top, bot, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h);
        P[0] = __fadd_rn(P[0], mm.g11); P[1] = __fadd_rn(P[1], mm.g12); P[2] = __fadd_rn(P[2], mm.g22);
        P[3] = __fadd_rn(P[3], mm.h1); P[4] = __fadd_rn(P[4], mm.h2);
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
      float2 fl = flow_at(t_first);
      // warm-up: the R-1 = 2m rows above the first output row, two at a time
      for (int t = t_first; t < t_first + R - 1; t += 2) {
        float V[5];
        issue_r(ra, rb, fl, t);
        fl = flow_at(t + 1);
        finish_r(ra, rb, t, V);
        issue_r(rb, ra, fl, t + 1);
        fl = flow_at(t + 2);
        finish_r(rb, ra, t + 1, V);
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
            if (rr & 1) issue_r(rb, ra, fl, yo + m); else issue_r(ra, rb, fl, yo + m);
            fl = flow_at(yo + m + 1);
            if (rr & 1) finish_r(rb, ra, yo + m, V); else finish_r(ra, rb, yo + m, V);
#pragma unroll
            for (int ch = 0; ch < 5; ch++) srow[(rr * 5 + ch) * COLS] = V[ch];
          }
        }
        named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
      }
    } else if constexpr (RIF == 1) {
      float2 fl = flow_at(t_first);
      // warm-up: the R-1 rows above the first output row (no hand-over)
      for (int t = t_first; t < t_first + R - 1; t++) {
        UmLoads2 L;
        float V[5];
        issue(L, fl, t);
        fl = flow_at(t + 1);
        finish(L, t, V);
      }
      for (int c = 0; c < n_chunks; c++) {
        const int buf = c & 1;
        if (c >= 2) named_bar_sync(BAR_EMPTY0 + buf, NT);            // consumers released this buffer
        float* srow = stage + buf * CH * 5 * COLS + tid;
#pragma unroll
        for (int rr = 0; rr < CH; rr++) {
          const int yo = y0 + c * CH + rr;                           // output row; newest matrix row = yo + m
          if (yo < y1) {
            UmLoads2 L;
            float V[5];
            issue(L, fl, yo + m);
            fl = flow_at(yo + m + 1);
            finish(L, yo + m, V);
#pragma unroll
            for (int ch = 0; ch < 5; ch++) srow[(rr * 5 + ch) * COLS] = V[ch];
          }
        }
        named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
      }
    } else {
      // CH rows in flight: all loads of a chunk are issued before the first row is consumed
      float2 fl[CH];
#pragma unroll
      for (int rr = 0; rr < CH; rr++) fl[rr] = flow_at(t_first + rr);
      // warm-up rows t_first .. t_first+R-2 in groups of CH (rows past the warm-up are handled by chunk 0,
      // so the warm-up only takes whole groups and chunk 0 starts where it stopped)
      int t = t_first;
      const int t_out = t_first + R - 1;                             // first matrix row that yields an output row
      for (; t + CH <= t_out; t += CH) {
        UmLoads2 L[CH];
#pragma unroll
        for (int rr = 0; rr < CH; rr++) issue(L[rr], fl[rr], t + rr);
#pragma unroll
        for (int rr = 0; rr < CH; rr++) fl[rr] = flow_at(t + CH + rr);
#pragma unroll
        for (int rr = 0; rr < CH; rr++) {
          float V[5];
          finish(L[rr], t + rr, V);
        }
      }
      for (; t < t_out; t++) {                                       // remainder of the warm-up, one row at a time
        UmLoads2 L1;
        float V[5];
        issue(L1, fl[0], t);
#pragma unroll
        for (int rr = 0; rr + 1 < CH; rr++) fl[rr] = fl[rr + 1];
        fl[CH - 1] = flow_at(t + CH);
        finish(L1, t, V);
      }
      for (int c = 0; c < n_chunks; c++) {
        const int buf = c & 1;
        const int tc = y0 + c * CH + m;                              // newest matrix row of output row y0 + c*CH
        UmLoads2 L[CH];
#pragma unroll
        for (int rr = 0; rr < CH; rr++) issue(L[rr], fl[rr], tc + rr);
#pragma unroll
        for (int rr = 0; rr < CH; rr++) fl[rr] = flow_at(tc + CH + rr);
        if (c >= 2) named_bar_sync(BAR_EMPTY0 + buf, NT);            // consumers released this buffer
        float* srow = stage + buf * CH * 5 * COLS + tid;
#pragma unroll
        for (int rr = 0; rr < CH; rr++) {
          float V[5];
          finish(L[rr], tc + rr, V);                                 // (rows past y1 keep the state consistent; never read)
#pragma unroll
          for (int ch = 0; ch < 5; ch++) srow[(rr * 5 + ch) * COLS] = V[ch];
        }
        named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
      }
    }
    return;
  }

  // -------------------------------------------------------------------- CONSUMERS (PXT adjacent pixels of one row each)
  if constexpr (REUSE && RIF == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REUSE_CONS_REGS));
  else if constexpr (RIF > 1 && COLS == 256 && PXT == 4 && CLOOP == 1 && MINB == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(48));
  float2* fout = flow_out + (size_t)pair * n;
  const int ct = tid - COLS;                         // 0..NCONS-1
  const int q_row0 = (ct / GROUPS) * CLOOP;          // first staged row of this thread's pixel group
  const int q0 = (ct % GROUPS) * PXT;                // first of its PXT strip columns
  const int ox = x_base + q0;                        // image x of that column
  // columns of the group that are real outputs of this strip
  unsigned vmask = 0;
#pragma unroll
  for (int j = 0; j < PXT; j++)
    if (q0 + j >= m && q0 + j < COLS - m && ox + j < w) vmask |= 1u << j;
  constexpr unsigned ALL = (1u << PXT) - 1u;

  for (int c = 0; c < n_chunks; c++) {
    const int buf = c & 1;
    named_bar_sync(BAR_FULL0 + buf, NT);             // producers finished staging chunk c
#if defined(OFB_DBG) && (OFB_DBG & 4)     // experiment: consumers only hand the buffers back
    if (c == n_chunks - 1 && vmask) fout[(unsigned)(y0 * w + max(ox, 0))] = make_float2(stage[ct], 0.f);
    if (c + 2 < n_chunks) named_bar_arrive(BAR_EMPTY0 + buf, NT);
    continue;
#endif
#pragma unroll
    for (int cl = 0; cl < CLOOP; cl++) {
      const int q_row = q_row0 + cl;
      const int yo = y0 + c * CH + q_row;
      if (!(yo < y1 && vmask)) continue;
      const float* srow = stage + (buf * CH + q_row) * 5 * COLS;
      float sum[5][PXT];
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        const float* s = srow + ch * COLS;
        if constexpr (MT > 0) {
          // e[i] = staged value at strip column q0 - PAD + i; the window of pixel j is e[PAD+j-MT .. PAD+j+MT]
          constexpr int PAD = (MT + 3) / 4 * 4;
          constexpr int NE = PXT + 2 * PAD;
          float e[NE];
#pragma unroll
          for (int v = 0; v < NE / 4; v++) {
            const int cq = min(max(q0 - PAD + 4 * v, 0), COLS - 4);
            const float4 t4 = *reinterpret_cast<const float4*>(s + cq);
            e[4 * v] = t4.x; e[4 * v + 1] = t4.y; e[4 * v + 2] = t4.z; e[4 * v + 3] = t4.w;
          }
          // columns common to all PXT windows: [PAD + PXT-1 - MT, PAD + MT]
          constexpr int C0 = PAD + PXT - 1 - MT, C1 = PAD + MT;
          static_assert(C0 <= C1, "window narrower than the pixel group");
          float core = e[C0];
#pragma unroll
          for (int i = C0 + 1; i <= C1; i++) core += e[i];
          float l = 0.f;                             // suffix sums on the left of the core
          sum[ch][PXT - 1] = core;
#pragma unroll
          for (int j = PXT - 2; j >= 0; j--) {
            l += e[PAD + j - MT];                    // columns PAD+j-MT .. C0-1 belong to windows <= j
            sum[ch][j] = core + l;
          }
          float r = 0.f;                             // prefix sums on the right of the core
#pragma unroll
          for (int j = 1; j < PXT; j++) {
            r += e[PAD + j + MT];
            sum[ch][j] += r;
          }
        } else {
          const int kq = (m + 3) >> 2;               // quads to each side
#pragma unroll
          for (int j = 0; j < PXT; j++) sum[ch][j] = 0.f;
          for (int v = -kq; v < PXT / 4 + kq; v++) {
            const int cq = min(max(q0 + 4 * v, 0), COLS - 4);
            const float4 t4 = *reinterpret_cast<const float4*>(s + cq);
            const float e[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int d = 4 * v + i;               // column offset from q0
#pragma unroll
              for (int j = 0; j < PXT; j++)
                if (d >= j - m && d <= j + m) sum[ch][j] += e[i];
            }
          }
        }
      }
      float2 f[PXT];
#pragma unroll
      for (int j = 0; j < PXT; j++) f[j] = solve2x2_sums(sum[0][j], sum[1][j], sum[2][j], sum[3][j], sum[4][j], reg);
      const int oi = yo * w + ox;                    // (oi + j >= 0 for every valid column j)
      float2* o = fout + (unsigned)max(oi, 0);
      if (vmask == ALL && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
        // 16-byte aligned: 128-bit stores
#pragma unroll
        for (int j = 0; j < PXT; j += 2)
          *reinterpret_cast<float4*>(o + j) = make_float4(f[j].x, f[j].y, f[j + 1].x, f[j + 1].y);
      } else {
#pragma unroll
        for (int j = 0; j < PXT; j++)
          if ((vmask >> j) & 1u) fout[(unsigned)(oi + j)] = f[j];
      }
    }
    if (c + 2 < n_chunks) named_bar_arrive(BAR_EMPTY0 + buf, NT);   // staging buffer may be refilled
  }
}

}  // namespace ofb
