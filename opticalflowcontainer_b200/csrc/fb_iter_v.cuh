// fb_iter_v.cuh — fused Farneback iteration kernel (box window), vertical-first, no FP64.
//
// UpdateMatrices + (2m+1)^2 box blur + 2x2 solve in one pass over the level, 56 B of HBM traffic
// per pixel-iteration (48 B for the first iteration of a level, whose input flow is the bilinear
// upsample of the coarser level's result computed on the fly — stage a8 fused in).
//
//   * the vertical window sum is done FIRST, on the matrices themselves, by the thread that owns
//     the column, in float, WITHOUT cancellation drift: rows are grouped in blocks of R = 2m+1;
//     P[k] = running prefix sum inside the block, B = sum of the finished block, and the sum of the
//     R rows ending at offset k of the current block is  (B_prev - P_prev[k]) + P_cur[k]
//     (van Herk / Gil-Werman).  A float running add/subtract sum drifts (measured 0.07 px max EPE on
//     high-contrast frames); this form measured <= 1.4e-2 px max / 3e-5 px mean against cv2.
//   * P_prev[k] is private to its column.  TMEM = true keeps it in TENSOR MEMORY: a producer thread is one TMEM lane,
//     ring slot k = 8 columns of that lane (5 used), written with tcgen05.st and read back R rows later with
//     tcgen05.ld.  The ring is a quarter of the kernel's shared-memory wavefronts (read 5 + write 5 of every
//     ~85 per 32 pixels) and the kernel sits on the L1 data pipe (ncu: 77 % of peak), so moving it to the
//     TMEM datapath takes it off the bottleneck and frees 77 KB of shared memory per CTA, spent on a deeper
//     staging pipeline (NBUF buffers).  TMEM = false keeps the ring in shared memory (radii whose ring does not fit
//     the CTA's TMEM columns, 128-column strips, the tiled mode).
//   * producers (one column per thread) stage the vertically summed rows; consumers only do the
//     horizontal window sums (4 adjacent pixels per thread from float4 reads), the solve and the
//     coalesced flow store.  FULL/EMPTY named barriers hand the NBUF staging buffers (CH rows each) over.
//
// Producer load schedules (bit-identical results):
//   REUSE (default): two rows of loads in flight per thread and the row-reuse gather (fb_um.cuh); the kernel is
//     launched at 80 registers/thread and setmaxnreg moves registers from the consumer warpgroup (48) to the
//     two producer warpgroups (96).
//   plain: one row in flight + L2 prefetch PFD rows ahead (tiled mode, run-time radius fallback).
#pragma once
#include "fb_device.cuh"
#include "fb_um.cuh"

// Experiment builds only: bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the R0 / flow rows OFB_EXP_L2PF chunks ahead,
// issued by one thread of the CTA in the two-rows-in-flight schedule (0 = off).
#ifndef OFB_EXP_L2PF
#define OFB_EXP_L2PF 0
#endif

namespace ofb {

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Per-warp staging ring of the regular inputs of the two-rows-in-flight schedule (TMAR): the R0 rows (float4 + float) and
// the input-flow rows of the warp's 32 columns arrive by bulk copies TR_CHUNKS chunks ahead of their use.
#ifndef OFB_EXP_TR_CHUNKS
#define OFB_EXP_TR_CHUNKS 3
#endif
constexpr int TR_CHUNKS = OFB_EXP_TR_CHUNKS;  // chunk slots per warp (2 rows each); shared memory spent here is L1 lost to the gathers
constexpr int TR_W = 36;                     // columns per slot row: the 4-aligned hull of 32 columns
constexpr int TR_ROW_BYTES = TR_W * (16 + 8 + 4);
constexpr int TR_WARP_BYTES = 2 * TR_CHUNKS * TR_ROW_BYTES + 64;     // + the slots' mbarriers

template <int COLS, int CH>
constexpr int iter_v_smem_floats(int m, bool tmem, int nbuf, bool tmar = false) {
  return (nbuf * CH + (tmem ? 0 : 2 * m + 1)) * 5 * COLS + (tmar ? (COLS / 32) * TR_WARP_BYTES / 4 : 0);
}

constexpr int kTmemRingStride = 8;    // TMEM columns per ring slot (5 used; x4 + x1 accesses stay aligned)
constexpr int kTmemWgCols = 128;      // TMEM columns per producer warpgroup

// Coarser level's flow for the fused upsample (prev == nullptr: the launch reads flow_in as it is).
struct UpsSrc {
  const float2* prev;     // [pair][ph][pw]
  const LinTab* tabx;     // cv::resize INTER_LINEAR tables of the level (x: w entries, y: h entries)
  const LinTab* taby;
  int pw, ph;
  float mul;              // 1 / pyr_scale
  int exact2y;            // the level has exactly twice the coarser level's rows: row coordinates without the table
};

// resize(prevFlow, INTER_LINEAR) * mul at one pixel — the arithmetic of k_upsample_flow, shared so the fused and the
// stand-alone upsample produce the same bits.
__device__ __forceinline__ float2 ups_blend(float2 q00, float2 q01, float2 q10, float2 q11, float fx, float fy, float mul) {
  const float ax0 = __fsub_rn(1.f, fx), ay0 = __fsub_rn(1.f, fy);
  const float tx = __fmaf_rn(q01.x, fx, __fmul_rn(q00.x, ax0)), ty = __fmaf_rn(q01.y, fx, __fmul_rn(q00.y, ax0));
  const float bx = __fmaf_rn(q11.x, fx, __fmul_rn(q10.x, ax0)), by = __fmaf_rn(q11.y, fx, __fmul_rn(q10.y, ax0));
  return make_float2(__fmul_rn(__fmaf_rn(bx, fy, __fmul_rn(tx, ay0)), mul), __fmul_rn(__fmaf_rn(by, fy, __fmul_rn(ty, ay0)), mul));
}

__device__ __forceinline__ float2 ld_volatile_f2(const float2* p) {
  float2 v;
  asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// ---- tensor memory as per-thread scratch (tcgen05.alloc / ld / st; one warp allocates for the CTA) ----
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tmem_ld5(uint32_t taddr, float (&v)[5]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v[4]) : "r"(taddr + 4u));
}
// the loaded registers may only be read after this (tied as in/out operands so nothing moves across)
__device__ __forceinline__ void tmem_wait_ld5(float (&v)[5]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4])::"memory");
}
__device__ __forceinline__ void tmem_st5(uint32_t taddr, const float (&v)[5]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + 4u), "f"(v[4]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- bulk asynchronous copies (TMA, cp.async.bulk) completing on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// MT: window radius (0 = run time).  COLS: strip width = producer threads.  CH: rows per chunk.  MINB: CTAs per SM the
// register budget is held to.  PFD: L2 prefetch distance of the plain schedule.  TILED (spatially tiled mode, one pair):
// the CTA grid covers only the level rows [y_begin, y_end) of this rank's band; R1 rows the displacement reaches outside
// the band are read from the owner's buffer through the NVLink peer pointers in `tab` (the flow and R0 are local).
// UPS: first iteration of a level — the input flow is the bilinear upsample of the coarser level's result (UpsSrc),
// computed by the producers: a thread marching down its column keeps the horizontally blended coarse flow of the two
// coarse rows around it and loads a new coarse row only when it crosses one (every second row at pyr_scale 0.5).
// TMAR: the regular inputs of the two-rows-in-flight schedule — the R0 rows and (plain launches) the input-flow rows — are
// staged in shared memory by bulk asynchronous copies (TMA, cp.async.bulk completing on mbarriers), per warp, TR_CHUNKS
// chunks ahead.  They cost no registers while in flight, the flow of the NEXT chunk is at hand early, and the gather of
// chunk v+1 is issued between the two halves of chunk v (its top row as soon as chunk v's top half is consumed), so
// every R1 load has about half an iteration to land instead of being waited for right after its issue (ncu before:
// long-scoreboard the top producer stall, issue 59 %, L1 62 %, DRAM 38 %).  Needs w % 4 == 0 (16-byte copy granules).
template <int MT, int COLS, int CH, int MINB, int PFD, bool TILED, bool REUSE, bool TMEM, int NBUF, bool UPS, bool TMAR = false>
__global__ void __launch_bounds__(COLS + CH * COLS / 4, MINB)
    k_iter_v(const RSet rs, const float2* __restrict__ flow_in, float2* __restrict__ flow_out, int w, int h, int m_rt,
             float reg, int seg_rows, int strips, int y_begin, int y_end, PeerTab tab, int my_rank, UpsSrc ups) {
  constexpr int PXT = 4;                             // adjacent pixels per consumer thread
  constexpr int GROUPS = COLS / PXT;
  constexpr int NCONS = CH * GROUPS;
  static_assert(NCONS % 32 == 0, "whole consumer warps");
  static_assert(!REUSE || (!TILED && CH == 2), "row-reuse schedule: chunks of two rows, untiled");
  static_assert(!TMEM || (COLS == 256 && MT >= 1 && (2 * MT + 1) * kTmemRingStride <= kTmemWgCols),
                "TMEM ring: two producer warpgroups, ring slots within the warpgroup's columns");
  static_assert(NBUF >= 2 && 1 + 2 * NBUF <= 16, "named barriers");
  static_assert(!TMAR || (REUSE && TMEM && MT > 0), "staged inputs: the default schedule only");
  constexpr bool REGMOVE = REUSE && COLS == 256 && MINB == 2;   // setmaxnreg 96 / 48
  constexpr int NT = COLS + NCONS;
  constexpr int BAR_FULL0 = 1, BAR_EMPTY0 = 1 + NBUF;
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int tw = COLS - 2 * m;
  extern __shared__ float smem[];
  float* stage = smem;                               // [NBUF][CH][5][COLS]   vertically summed rows
  float* ring = smem + NBUF * CH * 5 * COLS;         // [R][5][COLS]          P_prev[k] (TMEM = false)
  __shared__ uint32_t tmem_slot;

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

  if constexpr (TMEM) {
    if (tid < 32) tmem_alloc<2 * kTmemWgCols>(&tmem_slot);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (tid < COLS) {
    // ------------------------------------------------------------------ PRODUCERS (one column each)
    if constexpr (REGMOVE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(96));
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
    // TMEM ring: lane = thread within its warpgroup, columns [wg * 128 + k * 8, +5)
    uint32_t tring = 0, tk = 0;
    if constexpr (TMEM) tring = tmem_slot + ((uint32_t)((tid >> 5) & 3) << 21) + (uint32_t)(tid >> 7) * kTmemWgCols;
    float P[5] = {0.f, 0.f, 0.f, 0.f, 0.f};          // prefix sums of the current block
    float Bp[5] = {0.f, 0.f, 0.f, 0.f, 0.f};         // sum of the previous block
    int k = 0;                                       // offset of row t inside its block
    bool have_prev = false;

    // vertical van Herk step: P += M(t); V = (Bp - P_prev[k]) + P; ring[k] = P; block bookkeeping
    auto ring_step = [&](const M5& mm, float (&V)[5]) {
      float old[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if constexpr (TMEM) {
        tmem_wait_st();                              // the slot read below was written R rows ago: long complete
        if (have_prev) tmem_ld5(tring + tk, old);    // (uniform over the CTA: depends on the row count only)
      } else {
#pragma unroll
        for (int ch = 0; ch < 5; ch++) old[ch] = have_prev ? rcol[(k * 5 + ch) * COLS] : 0.f;
      }
      P[0] = __fadd_rn(P[0], mm.g11); P[1] = __fadd_rn(P[1], mm.g12); P[2] = __fadd_rn(P[2], mm.g22);
      P[3] = __fadd_rn(P[3], mm.h1); P[4] = __fadd_rn(P[4], mm.h2);
      if constexpr (TMEM) {
        if (have_prev) tmem_wait_ld5(old);
        tmem_st5(tring + tk, P);
        tk += kTmemRingStride;
      }
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        V[ch] = (Bp[ch] - old[ch]) + P[ch];
        if constexpr (!TMEM) rcol[(k * 5 + ch) * COLS] = P[ch];
      }
      if (++k == R) {
        k = 0;
        tk = 0;
        have_prev = true;
#pragma unroll
        for (int ch = 0; ch < 5; ch++) { Bp[ch] = P[ch]; P[ch] = 0.f; }
      }
    };

    // fused upsample: per-thread column entry of the resize table, cache of the two coarse rows around the current row
    int ux0 = 0, ux1 = 0, u_ca = -1, u_cb = -1;
    float ufx = 0.f, uax0 = 1.f;
    float2 uHa = make_float2(0.f, 0.f), uHb = uHa;
    const float2* uprev = nullptr;
    if constexpr (UPS) {
      const int2 tx = __ldg(reinterpret_cast<const int2*>(ups.tabx + x));   // {i0, f}
      ux0 = tx.x; ux1 = min(tx.x + 1, ups.pw - 1); ufx = __int_as_float(tx.y); uax0 = __fsub_rn(1.f, ufx);
      uprev = ups.prev + (size_t)pair * ups.pw * ups.ph;   // (tiled mode: the rank's own coarse band, one pair)
    }
    // coarse flow of row r blended along x at this column (the horizontal half of ups_blend, same operations)
    auto ups_hrow = [&](int r) -> float2 {
      const float2* p = uprev + (unsigned)r * (unsigned)ups.pw;
      const float2 q0 = __ldg(p + ux0), q1 = __ldg(p + ux1);
      // the next coarse row this column will cross, pulled into L1 now: the blend below waits for q0 / q1 right away
      // (ncu: 20 % of the producers' samples sat on that wait when every new row came from L2)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (r + 1 < ups.ph ? ups.pw : 0) + ux1));
      return make_float2(__fmaf_rn(q1.x, ufx, __fmul_rn(q0.x, uax0)), __fmaf_rn(q1.y, ufx, __fmul_rn(q0.y, uax0)));
    };
    // input flow of matrix row t at this column.  Plain launches: one load (volatile in the two-rows-in-flight schedule:
    // it stays where it is written, between the gathers and the barrier; ptxas otherwise sinks it to the end of the loop
    // body, in front of the address arithmetic).  Rows are visited in increasing order.
    auto flow_at = [&](int t) -> float2 {
      const int yc = clampi(t, 0, h - 1);
      if constexpr (UPS) {
        int r0;
        float fy;
        if (ups.exact2y) {                 // cv::resize coordinates of an exact x2: (y + 0.5) / 2 - 0.5
          r0 = (yc - 1) >> 1;
          fy = (yc & 1) ? 0.25f : 0.75f;
          if (r0 < 0) { r0 = 0; fy = 0.f; }
          if (r0 >= ups.ph - 1) { r0 = ups.ph - 1; fy = 0.f; }
        } else {
          const int2 ty = __ldg(reinterpret_cast<const int2*>(ups.taby + yc));
          r0 = ty.x;
          fy = __int_as_float(ty.y);
        }
        const int r1 = min(r0 + 1, ups.ph - 1);
        // (all of this is uniform over the CTA: it depends on the row only)
        if (r0 != u_ca) {
          if (r0 == u_cb) uHa = uHb; else uHa = ups_hrow(r0);
          u_ca = r0;
        }
        if (r1 != u_cb) {
          if (r1 == u_ca) uHb = uHa; else uHb = ups_hrow(r1);
          u_cb = r1;
        }
        const float ay0 = __fsub_rn(1.f, fy);
        return make_float2(__fmul_rn(__fmaf_rn(uHb.x, fy, __fmul_rn(uHa.x, ay0)), ups.mul),
                           __fmul_rn(__fmaf_rn(uHb.y, fy, __fmul_rn(uHa.y, ay0)), ups.mul));
      } else {
        if constexpr (REUSE) return ld_volatile_f2(fin + ((unsigned)yc * uw + (unsigned)x));
        return __ldg(fin + ((unsigned)yc * uw + (unsigned)x));
      }
    };
    auto stage_row = [&](int buf, int rr) { return stage + ((buf * CH + rr) * 5) * COLS + tid; };

    if constexpr (REUSE && TMAR) {
      // ---- staged inputs + interleaved gathers (see TMAR above).  Virtual chunk v = matrix rows t_first + 2v, + 1:
      // the m warm-up chunks, then the output chunks.
      const int wid = tid >> 5, lane = tid & 31;
      char* wbase = reinterpret_cast<char*>(smem + NBUF * CH * 5 * COLS) + wid * TR_WARP_BYTES;
      float4* s_ra = reinterpret_cast<float4*>(wbase);                                   // [2 * TR_CHUNKS][TR_W]
      float2* s_fl = reinterpret_cast<float2*>(wbase + 2 * TR_CHUNKS * TR_W * 16);
      float* s_rb = reinterpret_cast<float*>(wbase + 2 * TR_CHUNKS * TR_W * 24);
      uint64_t* s_bar = reinterpret_cast<uint64_t*>(wbase + 2 * TR_CHUNKS * TR_ROW_BYTES);
      // the warp's column slice: 4-aligned hull of its (clamped) columns
      const int xmin = clampi(x_base + 32 * wid, 0, w - 1), xmax = clampi(x_base + 32 * wid + 31, 0, w - 1);
      const int xs = xmin & ~3, xn = min((xmax + 4) & ~3, w) - xs;
      const int cx = x - xs;
      const int nv = m + n_chunks;
      auto copy_chunk = [&](int v) {            // lane 0: rows of virtual chunk v -> slot v % TR_CHUNKS
        const int sl = v % TR_CHUNKS;
        uint64_t* bar = s_bar + sl;
        mbar_expect_tx(bar, 2u * (unsigned)xn * (UPS ? 20u : 28u));
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const unsigned o = (unsigned)clampi(t_first + 2 * v + j, 0, h - 1) * uw + (unsigned)xs;
          bulk_g2s(s_ra + (2 * sl + j) * TR_W, RA0 + o, (unsigned)xn * 16u, bar);
          bulk_g2s(s_rb + (2 * sl + j) * TR_W, RB0 + o, (unsigned)xn * 4u, bar);
          if constexpr (!UPS) bulk_g2s(s_fl + (2 * sl + j) * TR_W, fin + o, (unsigned)xn * 8u, bar);
        }
      };
      if (lane == 0) {
        for (int i = 0; i < TR_CHUNKS; i++) mbar_init(s_bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int v = 0; v < min(TR_CHUNKS, nv); v++) copy_chunk(v);
      }
      __syncwarp();
      UmRow X, Y, Z, W;
      struct { float fx, fy; unsigned g; bool inside; } pa, pb;
      float2 fa = make_float2(0.f, 0.f), fb = fa;                    // UPS: the chunk's input flow (computed, not staged)
      bool reuse_b = false;
      Z.q0 = Z.q1 = make_float4(0.f, 0.f, 0.f, 0.f); Z.s0 = Z.s1 = 0.f;
      W = Z; X = Z; Y = Z;
      auto pix = [&](decltype(pa)& P, float2 fl, int y) {            // (um_pix without the R0 loads)
        const float fx = (float)x + fl.x, fy = (float)y + fl.y;
        const int ix = __float2int_rd(fx), iy = __float2int_rd(fy);
        P.fx = fx - (float)ix;
        P.fy = fy - (float)iy;
        P.inside = (unsigned)ix < uw - 1u && (unsigned)iy < uh - 1u;
        P.g = P.inside ? (unsigned)iy * uw + (unsigned)ix : 0u;
      };
      auto wait_chunk = [&](int v) { mbar_wait(s_bar + v % TR_CHUNKS, (unsigned)(v / TR_CHUNKS) & 1u); };
      auto flow_of = [&](int v, int j) -> float2 {
        if constexpr (UPS) return flow_at(t_first + 2 * v + j);
        else return s_fl[(2 * (v % TR_CHUNKS) + j) * TR_W + cx];
      };
      // top half of chunk v: pixel A; top corner row from the previous chunk's bottom row (Z) when it lines up
      auto issue_a = [&](int v, unsigned prev_g) {
        const int ya = clampi(t_first + 2 * v, 0, h - 1);
        fa = flow_of(v, 0);
        pix(pa, fa, ya);
        if (pa.g != prev_g + uw) um_row_load(X, RA1, RB1, pa.g); else X = Z;
        um_row_load(Y, RA1, RB1, pa.g + uw);
      };
      auto issue_b = [&](int v) {
        const int yb = clampi(t_first + 2 * v + 1, 0, h - 1);
        fb = flow_of(v, 1);
        pix(pb, fb, yb);
        reuse_b = pa.inside && pb.g == pa.g + uw;
        if (!reuse_b) um_row_load(W, RA1, RB1, pb.g);
        um_row_load(Z, RA1, RB1, pb.g + uw);
      };
      auto finish_half = [&](int v, int j, const decltype(pa)& P, const UmRow& top, const UmRow& bot, float2 fl, float (&V)[5]) {
        const int y = clampi(t_first + 2 * v + j, 0, h - 1);
        const int si = (2 * (v % TR_CHUNKS) + j) * TR_W + cx;
        const float4 a0 = s_ra[si];
        const float b0 = s_rb[si];
        if constexpr (!UPS) fl = s_fl[si];
        ring_step(um_arith(a0, b0, top.q0, top.q1, bot.q0, bot.q1, top.s0, top.s1, bot.s0, bot.s1, P.fx, P.fy, fl.x, fl.y,
                           P.inside, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h), V);
      };
      wait_chunk(0);
      issue_a(0, ~0u - uw);
      issue_b(0);
      int buf = 0;
      for (int v = 0; v < nv; v++) {
        const int c = v - m;                                         // output chunk (negative: warm-up)
        if (c >= NBUF) named_bar_sync(BAR_EMPTY0 + buf, NT);         // consumers released this buffer
        float V[5];
        finish_half(v, 0, pa, X, Y, fa, V);
        if (c >= 0) {
          float* s0 = stage_row(buf, 0);
#pragma unroll
          for (int ch = 0; ch < 5; ch++) s0[ch * COLS] = V[ch];
        }
        if (reuse_b) W = Y;                                          // B's top row is A's bottom row: keep it, Y is refilled now
        const unsigned g_b = pb.inside ? pb.g : ~0u - uw;
        if (v + 1 < nv) {
          wait_chunk(v + 1);                                         // (requested TR_CHUNKS - 1 chunks ago)
          issue_a(v + 1, g_b);                                       // X <- Z (copy) or load; Y <- load: both free now
        }
        finish_half(v, 1, pb, W, Z, fb, V);                          // (a row past y1 keeps the state consistent; never read)
        if (c >= 0) {
          float* s1 = stage_row(buf, 1);
#pragma unroll
          for (int ch = 0; ch < 5; ch++) s1[ch * COLS] = V[ch];
          named_bar_arrive(BAR_FULL0 + buf, NT);                     // staged rows of chunk c are ready
          if (++buf == NBUF) buf = 0;
        }
        if (v + 1 < nv) issue_b(v + 1);                              // W, Z free now
        // the warp is done with the slot of chunk v: refill it
        __syncwarp();
        if (lane == 0 && v + TR_CHUNKS < nv) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          copy_chunk(v + TR_CHUNKS);
        }
      }
    } else if constexpr (REUSE) {
      // Two rows in flight AND row-reuse gather.  Rows A = t, B = t + 1 of a chunk: all loads of both are issued
      // before the first is consumed.  Corner-row register sets: X = top of A, Y = bottom of A (and top of B when B
      // sits exactly one row below A), W = top of B otherwise, Z = bottom of B.  The next chunk's A takes Z as its top
      // row when the displacement allows it, so a smooth field costs 8 gather loads per chunk instead of 16.
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
      // warm-up: the R-1 = 2m rows above the first output row, chunk by chunk (unrolled for a compile-time radius: the
      // rolled loop measured 3 % slower over a whole 1080p step)
      for (int t = t_first; t < t_first + R - 1; t += 2) {
        float V[5];
        issue2(fa, fb, t);
        fa = flow_at(t + 2); fb = flow_at(t + 3);
        finish_a(t, V);
        finish_b(t + 1, V);
      }
      int buf = 0;
#if OFB_EXP_L2PF > 0
      const int pf_xs = max(x_base & ~3, 0), pf_n = min((x_base + COLS + 3) & ~3, w) - pf_xs;
      const bool pf_on = tid == 0 && (w & 3) == 0 && pf_n > 0;
#endif
      for (int c = 0; c < n_chunks; c++) {
        const int tc = y0 + c * CH + m;                              // newest matrix row of output row y0 + c*CH
#if OFB_EXP_L2PF > 0
        if (pf_on) {
#pragma unroll
          for (int rr = 0; rr < 2; rr++) {
            const unsigned o = (unsigned)clampi(tc + 2 * OFB_EXP_L2PF + rr, 0, h - 1) * uw + (unsigned)pf_xs;
            bulk_prefetch_l2(RA0 + o, (unsigned)pf_n * 16u);
            bulk_prefetch_l2(RB0 + o, (unsigned)pf_n * 4u);
            if constexpr (!UPS) bulk_prefetch_l2(fin + o, (unsigned)pf_n * 8u);
          }
        }
#endif
        issue2(fa, fb, tc);
        fa = flow_at(tc + 2); fb = flow_at(tc + 3);
        if (c >= NBUF) named_bar_sync(BAR_EMPTY0 + buf, NT);         // consumers released this buffer
        float V[5];
        finish_a(tc, V);
        float* s0 = stage_row(buf, 0);
#pragma unroll
        for (int ch = 0; ch < 5; ch++) s0[ch * COLS] = V[ch];
        finish_b(tc + 1, V);                                         // (a row past y1 keeps the state consistent; never read)
        float* s1 = stage_row(buf, 1);
#pragma unroll
        for (int ch = 0; ch < 5; ch++) s1[ch * COLS] = V[ch];
        named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
        if (++buf == NBUF) buf = 0;
      }
    } else {
      // one row in flight; prefetch.global.L2 PFD rows ahead (R0, flow, the predicted corner row of the R1 gather)
      auto issue = [&](UmLoads2& L, float2 f, int t) {
        const int y = clampi(t, 0, h - 1);
        const unsigned yw = (unsigned)y * uw;
        if constexpr (TILED) {
          const int ro = (y >= tab.r_lo && y < tab.r_hi) ? my_rank : tile_owner(y, tab);   // uniform over the CTA
          um_issue2_tiled(L, tab.RA[ro], tab.RB[ro], tab, n, my_rank, f, x, y, yw, uw, uh);
        } else {
          um_issue2(L, RA0, RB0, RA1, RB1, f, x, y, yw, uw, uh);
        }
        if (PFD > 0) {   // (tiled mode: RA0.. are this rank's own buffers — rows of a neighbour are simply not prefetched usefully)
          static_assert(PFD + 1 <= kRowPad, "prefetch distance exceeds the row padding of the R buffers");
          const unsigned op = (unsigned)clampi(t + PFD, 0, h - 1) * uw + (unsigned)x;
          prefetch_l2(RA0 + op);
          prefetch_l2(RB0 + op);
          if constexpr (!UPS) prefetch_l2(fin + ((unsigned)clampi(t + PFD + 1, 0, h - 1) * uw + (unsigned)x));
          const unsigned g = L.inside ? (unsigned)__float2int_rd((float)y + L.dy) * uw + (unsigned)__float2int_rd((float)x + L.dx) : 0u;
          prefetch_l2(RA1 + (g + (PFD + 1) * uw));
          prefetch_l2(RB1 + (g + (PFD + 1) * uw));
        }
      };
      auto finish = [&](const UmLoads2& L, int t, float (&V)[5]) {
        const int y = clampi(t, 0, h - 1);
        ring_step(um_finish2(L, xborder || (unsigned)(y - 5) >= (unsigned)(h - 10), x, y, w, h), V);
      };
      float2 fl = flow_at(t_first);
      // warm-up: the R-1 rows above the first output row (no hand-over)
#pragma unroll 1
      for (int t = t_first; t < t_first + R - 1; t++) {
        UmLoads2 L;
        float V[5];
        issue(L, fl, t);
        fl = flow_at(t + 1);
        finish(L, t, V);
      }
      int buf = 0;
      for (int c = 0; c < n_chunks; c++) {
        if (c >= NBUF) named_bar_sync(BAR_EMPTY0 + buf, NT);         // consumers released this buffer
#pragma unroll
        for (int rr = 0; rr < CH; rr++) {
          const int yo = y0 + c * CH + rr;                           // output row; newest matrix row = yo + m
          if (yo < y1) {
            UmLoads2 L;
            float V[5];
            issue(L, fl, yo + m);
            fl = flow_at(yo + m + 1);
            finish(L, yo + m, V);
            float* srow = stage_row(buf, rr);
#pragma unroll
            for (int ch = 0; ch < 5; ch++) srow[ch * COLS] = V[ch];
          }
        }
        named_bar_arrive(BAR_FULL0 + buf, NT);                       // staged rows of chunk c are ready
        if (++buf == NBUF) buf = 0;
      }
    }
    if constexpr (TMEM) tmem_wait_st();                // drained before the columns are freed
  } else {
    // ------------------------------------------------------------------ CONSUMERS (4 adjacent pixels of one row each)
    if constexpr (REGMOVE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(48));
    float2* fout = flow_out + (size_t)pair * n;
    const int ct = tid - COLS;                         // 0..NCONS-1
    const int q_row = ct / GROUPS;                     // staged row of this thread's pixel group
    const int q0 = (ct % GROUPS) * PXT;                // first of its PXT strip columns
    const int ox = x_base + q0;                        // image x of that column
    // columns of the group that are real outputs of this strip
    unsigned vmask = 0;
#pragma unroll
    for (int j = 0; j < PXT; j++)
      if (q0 + j >= m && q0 + j < COLS - m && ox + j < w) vmask |= 1u << j;
    constexpr unsigned ALL = (1u << PXT) - 1u;

    int buf = 0;
    for (int c = 0; c < n_chunks; c++) {
      named_bar_sync(BAR_FULL0 + buf, NT);             // producers finished staging chunk c
      const int yo = y0 + c * CH + q_row;
      if (yo < y1 && vmask) {
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
      if (c + NBUF < n_chunks) named_bar_arrive(BAR_EMPTY0 + buf, NT);   // staging buffer may be refilled
      if (++buf == NBUF) buf = 0;
    }
  }

  if constexpr (TMEM) {
    // every producer has drained its tensor-memory traffic (wait::st above); the allocating warp frees the columns
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) tmem_dealloc<2 * kTmemWgCols>(tmem_slot);
  }
}

}  // namespace ofb
