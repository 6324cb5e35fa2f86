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
#include <type_traits>
#include "fb_device.cuh"
#include "fb_um.cuh"

// Bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the R0 / flow / R1 rows OFB_EXP_L2PF chunks ahead,
// issued by one thread of the CTA in the two-rows-in-flight schedule (0 = off).  Measured (iteration stage of an 18-pair
// 1080p step): off 1.851 ms, 1 chunk ahead 1.749, 2: 1.767, 3: 1.803, 5: 1.850 (L2 turns over every ~25 us at this
// traffic); one prefetch set per producer WARP instead of per CTA: 2.05 ms (many small bulk requests).
#ifndef OFB_EXP_L2PF
#define OFB_EXP_L2PF 1
#endif

namespace ofb {

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Columns of window halo on each side of a strip.  Radius 7 (winsize 15, the default) and 15 take one column more than the
// window needs: strips then start at a multiple of 8 pixels, so the R0 rows (16 B per pixel) are read in whole 128-byte
// lines, the flow rows in whole sectors, and the consumers' 16-byte flow stores are always aligned (with a halo of 7 every
// strip started at an odd x: 5 lines per LDG.128 instead of 4 and four 8-byte stores per thread instead of two 16-byte ones).
// 1920 / 960 / 480 / 240 / 3840 split into the same number of 240-column strips as of 242-column ones.
__host__ __device__ constexpr int iter_v_halo(int mt, int m) { return (mt == 7 || mt == 15) ? mt + 1 : m; }

// Ring slots a packed TMEM ring holds (5 columns each in a warpgroup's 128); the slots beyond (radii 13..15: 2..6 slots)
// live in shared memory behind the staging buffers.
constexpr int kTmemPackedSlots = 25;
template <int COLS, int CH>
constexpr int iter_v_smem_floats(int m, bool tmem, int nbuf) {
  const int ring_slots = tmem ? (2 * m + 1 > kTmemPackedSlots ? 2 * m + 1 - kTmemPackedSlots : 0) : 2 * m + 1;
  return (nbuf * CH + ring_slots) * 5 * COLS;
}

constexpr int kTmemRingStride = 8;    // TMEM columns per ring slot (5 used; x4 + x1 accesses stay aligned)
constexpr int kTmemWgCols = 128;      // TMEM columns per producer warpgroup
// Rings of more than 16 slots (window radius 8..12) do not fit the warpgroup's 128 columns at 8 columns per slot; they are
// PACKED: four channels of slot k at columns [4k, 4k + 4), the fifth at column 4T + k (T = min(R, 25) slots) — 5 columns per
// slot, every access aligned to its width, two tcgen05.ld / st per row instead of one.  Rings of more than 25 slots
// (radius 13..15) keep their last R - 25 slots in shared memory.
__host__ __device__ constexpr bool tmem_ring_packed(int mt) { return (2 * mt + 1) * kTmemRingStride > kTmemWgCols; }

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
// one ring slot (8 columns, 5 used) in one instruction each way; the three spare columns carry don't-care registers
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld8(float (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4])::"memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[5]) {
  asm volatile("{\n.reg .b32 u;\ntcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, u, u, u};\n}" ::"r"(taddr),
               "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]) : "memory");
}
// packed ring slot: x4 at `ta`, the fifth channel x1 at `tb`
__device__ __forceinline__ void tmem_ld41(uint32_t ta, uint32_t tb, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(ta));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v[4]) : "r"(tb));
}
__device__ __forceinline__ void tmem_st41(uint32_t ta, uint32_t tb, const float (&v)[5]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb), "f"(v[4]) : "memory");
}
__device__ __forceinline__ void lds_f4(uint32_t addr, float& a, float& b, float& c, float& d) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}
__device__ __forceinline__ void stg_f8(void* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e),
               "f"(f), "f"(g), "f"(h) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarriers (shared memory): the FULL / EMPTY hand-over of the staging buffers.  One elected lane per warp arrives
// (after __syncwarp, so the warp's shared-memory accesses are ordered before the release); waiters only wait — unlike
// bar.sync on a named barrier, the producer warps do not wait for EACH OTHER, so they drift apart by up to NBUF chunks
// and hide one another's load latency (ncu on the named-barrier version: 27 % of the producers' samples at the EMPTY
// barrier although the consumers were idle half of the time).
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
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
template <int MT, int COLS, int CH, int MINB, int PFD, bool TILED, bool REUSE, bool TMEM, int NBUF, bool UPS>
__global__ void __launch_bounds__(COLS + CH * COLS / 4, MINB)
    k_iter_v(const RSet rs, const float2* __restrict__ flow_in, float2* __restrict__ flow_out, int w, int h, int m_rt,
             float reg, int seg_rows, int strips, int y_begin, int y_end, PeerTab tab, int my_rank, UpsSrc ups) {
  constexpr int PXT = 4;                             // adjacent pixels per consumer thread
  constexpr int GROUPS = COLS / PXT;
  constexpr int NCONS = CH * GROUPS;
  static_assert(NCONS % 32 == 0, "whole consumer warps");
  static_assert(!REUSE || CH == 2, "row-reuse schedule: chunks of two rows");
  static_assert(!TMEM || (COLS == 256 && MT >= 1), "TMEM ring: two producer warpgroups, radius known at compile time");
  constexpr bool PACKED = TMEM && tmem_ring_packed(MT);
  constexpr int TSLOTS = !PACKED ? 2 * MT + 1 : (2 * MT + 1 < kTmemPackedSlots ? 2 * MT + 1 : kTmemPackedSlots);   // slots in TMEM
  constexpr bool SPILL = PACKED && 2 * MT + 1 > kTmemPackedSlots;      // the slots beyond TSLOTS live in shared memory
  static_assert(!TMEM || TSLOTS * (PACKED ? 5 : kTmemRingStride) <= kTmemWgCols, "ring slots within the warpgroup's columns");
  static_assert(NBUF >= 2 && NBUF <= 8, "staging buffers");
  constexpr bool REGMOVE = REUSE && COLS == 256 && MINB == 2;   // setmaxnreg 96 / 48
  const int m = MT > 0 ? MT : m_rt;
  const int R = 2 * m + 1;
  const int halo = iter_v_halo(MT, m);
  const int tw = COLS - 2 * halo;
  extern __shared__ float smem[];
  float* stage = smem;                               // [NBUF][CH][5][COLS]   vertically summed rows
  float* ring = smem + NBUF * CH * 5 * COLS;         // [R][5][COLS]          P_prev[k] (TMEM = false)
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t handover[2 * NBUF];            // mbarriers: FULL[NBUF] (producer warps arrive), EMPTY[NBUF] (consumer warps)
  const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(handover);
  const uint32_t bar_empty = bar_full + 8u * NBUF;

  const int strip = blockIdx.x % strips;
  const int seg = blockIdx.x / strips;
  const int pair = blockIdx.y;
  const int x_base = strip * tw - halo;              // image x of strip column 0
  const int y0 = y_begin + seg * seg_rows;
  const int y1 = min(y0 + seg_rows, y_end);          // exclusive
  const int t_first = y0 - m;                        // first matrix row the segment needs
  const int n_chunks = (y1 - y0 + CH - 1) / CH;

  const size_t n = (size_t)w * h;
  const int tid = threadIdx.x;

  if (tid == COLS) {                                 // (a consumer thread: the first warp may be busy allocating)
    for (int i = 0; i < NBUF; i++) {
      mbar_init(bar_full + 8u * i, COLS / 32);
      mbar_init(bar_empty + 8u * i, NCONS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (TMEM) {
    if (tid < 32) tmem_alloc<2 * kTmemWgCols>(&tmem_slot);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if constexpr (TMEM) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const bool lane0 = (tid & 31) == 0;

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
    // The ring starts zero-filled, so the first block subtracts zeros: no "have a previous block" flag in the row step.
    if constexpr (TMEM) {
      const float z[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if constexpr (PACKED) {
        for (int i = 0; i < TSLOTS; i++) tmem_st41(tring + 4u * (uint32_t)i, tring + 4u * (uint32_t)TSLOTS + (uint32_t)i, z);
        if constexpr (SPILL)
          for (int i = 0; i < (R - TSLOTS) * 5; i++) rcol[i * COLS] = 0.f;
      } else {
        for (int i = 0; i < R; i++) tmem_st8(tring + (uint32_t)i * kTmemRingStride, z);
      }
    } else {
      for (int i = 0; i < R * 5; i++) rcol[i * COLS] = 0.f;
    }

    // vertical van Herk step: P += M(t); V = (Bp - P_prev[k]) + P; ring[k] = P; block bookkeeping
    auto ring_step = [&](const M5& mm, float (&V)[5]) {
      OFB_DASSERT(k >= 0 && k < R);                                   // ring slot inside the (2m+1)-row ring
      OFB_DASSERT(!TMEM || PACKED || (tk == (uint32_t)k * kTmemRingStride && tk + 5 <= (uint32_t)kTmemWgCols));
      OFB_DASSERT(!PACKED || (tk == 4u * (uint32_t)k && 5 * TSLOTS <= kTmemWgCols));
      float old[8];
      const bool in_tmem = !SPILL || k < TSLOTS;     // (uniform over the warp: k is the row's offset in its block)
      if constexpr (TMEM) {
        if (in_tmem) {
          tmem_wait_st();                            // the slot read below was written R rows ago: long complete
          if constexpr (PACKED) tmem_ld41(tring + tk, tring + 4u * (uint32_t)TSLOTS + (uint32_t)k, old);
          else tmem_ld8(tring + tk, old);
        } else {
#pragma unroll
          for (int ch = 0; ch < 5; ch++) old[ch] = rcol[((k - TSLOTS) * 5 + ch) * COLS];
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < 5; ch++) old[ch] = rcol[(k * 5 + ch) * COLS];
      }
      P[0] = __fadd_rn(P[0], mm.g11); P[1] = __fadd_rn(P[1], mm.g12); P[2] = __fadd_rn(P[2], mm.g22);
      P[3] = __fadd_rn(P[3], mm.h1); P[4] = __fadd_rn(P[4], mm.h2);
      if constexpr (TMEM) {
        if (in_tmem) {
          tmem_wait_ld8(old);
          if constexpr (PACKED) tmem_st41(tring + tk, tring + 4u * (uint32_t)TSLOTS + (uint32_t)k, P);
          else tmem_st8(tring + tk, P);
        }
        tk += PACKED ? 4u : (uint32_t)kTmemRingStride;
      }
#pragma unroll
      for (int ch = 0; ch < 5; ch++) {
        V[ch] = (Bp[ch] - old[ch]) + P[ch];
        if constexpr (!TMEM) rcol[(k * 5 + ch) * COLS] = P[ch];
        if constexpr (SPILL) {
          if (!in_tmem) rcol[((k - TSLOTS) * 5 + ch) * COLS] = P[ch];
        }
      }
      if (++k == R) {
        k = 0;
        tk = 0;
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
    auto stage_row = [&](int buf, int rr) {
      OFB_DASSERT(buf >= 0 && buf < NBUF && rr >= 0 && rr < CH && tid < COLS);   // inside the staging pipeline's buffers
      return stage + ((buf * CH + rr) * 5) * COLS + tid;
    };

    if constexpr (REUSE) {
      // Two rows in flight AND row-reuse gather.  Rows A = t, B = t + 1 of a chunk: all loads of both are issued before
      // the first is consumed.  Corner-row register sets: TA = top of A, Y = bottom of A (and top of B when B sits exactly
      // one row below A; otherwise B's top row is loaded late, see below), BZ = bottom of B.  The next chunk's A takes BZ as its top row when the
      // displacement allows it, so a smooth field costs 8 gather loads per chunk instead of 16.
      //
      // The kernel is ISSUE-bound (ncu: 334 instructions per pixel-iteration, 250 of them in this loop, issue slots 59 %
      // busy, every other unit below that), so the loop is written for instruction count:
      //   * interior chunks (rows 5 .. h-6, no clamps, no border rows) run a body without row clamping, unrolled in
      //     PAIRS with the roles of the two register sets S1 / S2 (top of A <-> bottom of B) and of the flow registers
      //     swapped, so nothing is copied between chunks;
      //   * where every lane of the warp is inside, lines up (B's top row = A's bottom row) and is not a border column
      //     — a warp-uniform vote, the common case — the arithmetic runs without selects and predicated moves;
      //   * the staging address is a 32-bit shared address kept in a register; the ring slot moves in ONE tcgen05.ld /
      //     tcgen05.st (x8) each way; no "first block" flag (zero-filled ring).
      // Bits are those of every other schedule (um_arith is the one arithmetic).
      constexpr unsigned FULLMASK = 0xffffffffu;
      constexpr uint32_t ROWB = 5u * COLS * 4u, BUFB = CH * ROWB;      // bytes of a staged row / staging buffer
      const float xf = (float)x;
      const uint32_t st0 = (uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)tid * 4u;
      UmRow S1, S2, Y;
      S1.q0 = S1.q1 = make_float4(0.f, 0.f, 0.f, 0.f); S1.s0 = S1.s1 = 0.f;
      S2 = S1; Y = S1;
      unsigned prev_g = ~0u - uw;
      float2 rq0 = make_float2(0.f, 0.f), rq1 = rq0;                  // fused upsample: see chunk()
      bool hvalid = false;
      int buf = 0;
      unsigned par = 0;                                               // parity of the staging round (flips when buf wraps)
      uint32_t stb = st0;                                             // staging address of buffer `buf`

#if OFB_EXP_L2PF > 0
      // columns of the strip, 4-pixel aligned hull (bulk prefetches move 16-byte granules; rows of w % 4 == 0 pixels)
      const int pf_xs = max(x_base, 0) & ~3, pf_n = min((x_base + COLS + 3) & ~3, w) - pf_xs;
      const bool pf_on = tid == 0 && (w & 3) == 0 && pf_n >= 4 && h >= 4;
#endif
      // One chunk: matrix rows t, t + 1; output chunk c (negative: warm-up, nothing staged).  FA / FB: input flow of the
      // two rows (loaded by the previous chunk); NA / NB receive the next chunk's.  TA holds the previous chunk's bottom
      // row on entry; BZ receives this chunk's.
      auto chunk = [&](auto interior, const int t, const int c, UmRow& TA, UmRow& BZ, const float2& FA, const float2& FB,
                       float2& NA, float2& NB) {
        constexpr bool INT = decltype(interior)::value;
        const int ya = INT ? t : clampi(t, 0, h - 1), yb = INT ? t + 1 : clampi(t + 1, 0, h - 1);
        const unsigned oa = (unsigned)ya * uw + (unsigned)x, ob = INT ? oa + uw : (unsigned)yb * uw + (unsigned)x;
        const float4 a0a = __ldg(RA0 + oa);
        const float b0a = __ldg(RB0 + oa);
        const float4 a0b = __ldg(RA0 + ob);
        const float b0b = __ldg(RB0 + ob);
        // Input flow of the two rows.  Plain launches: loaded by the previous chunk.  Fused upsample: blended here from the
        // coarse level's rows.  Interior chunks (the level has exactly twice the coarse rows and the chunk starts on an odd
        // row t = 2k + 1, see the loop below) need the horizontally blended coarse rows k and k + 1 only (uHa, uHb): one
        // new coarse row per chunk, whose two loads were issued a whole chunk earlier (rq0, rq1), so nothing here waits for
        // memory (ncu before: long-scoreboard 5.8 warps per issue against 2.4 of the plain launch, 30 % more
        // instructions).  cv::resize rows of an exact x2: y = 2k + 1 blends rows (k, k+1) with weight 0.25 on the second,
        // y = 2k + 2 the same rows with 0.75.  Edge chunks use the general row-by-row form (flow_at).
        float2 fa_, fb_;
        if constexpr (!UPS) {
          fa_ = FA;
          fb_ = FB;
        } else if constexpr (INT) {
          const int r = t >> 1;
          if (!hvalid) {
            uHa = ups_hrow(r); uHb = ups_hrow(r + 1);
            hvalid = true;
          } else {
            uHa = uHb;
            uHb = make_float2(__fmaf_rn(rq1.x, ufx, __fmul_rn(rq0.x, uax0)), __fmaf_rn(rq1.y, ufx, __fmul_rn(rq0.y, uax0)));
          }
          {   // the row the NEXT chunk adds: r + 2 <= ph - 1 for every interior chunk
            const float2* p = uprev + (unsigned)(r + 2) * (unsigned)ups.pw;
            rq0 = __ldg(p + ux0);
            rq1 = __ldg(p + ux1);
          }
          fa_ = make_float2(__fmul_rn(__fmaf_rn(uHb.x, 0.25f, __fmul_rn(uHa.x, 0.75f)), ups.mul),
                            __fmul_rn(__fmaf_rn(uHb.y, 0.25f, __fmul_rn(uHa.y, 0.75f)), ups.mul));
          fb_ = make_float2(__fmul_rn(__fmaf_rn(uHb.x, 0.75f, __fmul_rn(uHa.x, 0.25f)), ups.mul),
                            __fmul_rn(__fmaf_rn(uHb.y, 0.75f, __fmul_rn(uHa.y, 0.25f)), ups.mul));
        } else {
          if (hvalid) { hvalid = false; u_ca = u_cb = -1; }           // (the general form keeps its own row ids)
          fa_ = flow_at(t);
          fb_ = flow_at(t + 1);
        }
        const float2 FA_ = fa_, FB_ = fb_;
        // pixel A
        const float pxa = xf + FA_.x, pya = (float)ya + FA_.y;
        const int ixa = __float2int_rd(pxa), iya = __float2int_rd(pya);
        const float fxa = pxa - (float)ixa, fya = pya - (float)iya;
        const bool ina = (unsigned)ixa < uw - 1u && (unsigned)iya < uh - 1u;
        // (outside pixels gather — and discard — from (0, 0); tiled mode: from the start of their own row, always local)
        const unsigned ga = ina ? (unsigned)iya * uw + (unsigned)ixa : (TILED ? (unsigned)ya * uw : 0u);
        // pixel B
        const float pxb = xf + FB_.x, pyb = (float)yb + FB_.y;
        const int ixb = __float2int_rd(pxb), iyb = __float2int_rd(pyb);
        const float fxb = pxb - (float)ixb, fyb = pyb - (float)iyb;
        const bool inb = (unsigned)ixb < uw - 1u && (unsigned)iyb < uh - 1u;
        const unsigned gb = inb ? (unsigned)iyb * uw + (unsigned)ixb : (TILED ? (unsigned)yb * uw : 0u);
        const bool reuse_b = ina && gb == ga + uw;
        // Tiled mode: corner rows outside the band this rank computed live in their owner's buffer (NVLink peer pointer).
        // A warp whose rows are all local (the common case) takes the plain loads.
        const int rowa = ina ? iya : ya, rowb = inb ? iyb : yb;
        bool all_local = true;
        if constexpr (TILED)
          all_local = __all_sync(FULLMASK, min(rowa, rowb) >= tab.r_lo && max(rowa, rowb) + 1 < tab.r_hi);
        auto ldrow = [&](UmRow& R, unsigned g, int row) {
          if constexpr (TILED) {
            if (!all_local) {
              const int o = (row >= tab.r_lo && row < tab.r_hi) ? my_rank : tile_owner(row, tab);
              um_row_load(R, tab.RA[o] + n, tab.RB[o] + n, g);
              return;
            }
          }
          um_row_load(R, RA1, RB1, g);
        };
        if (ga != prev_g + uw) ldrow(TA, ga, rowa);
        ldrow(Y, ga + uw, rowa + 1);
        ldrow(BZ, gb + uw, rowb + 1);
        prev_g = inb ? gb : ~0u - uw;
        // input flow of the next chunk (volatile: it stays here, in front of the barrier and the arithmetic)
        if constexpr (UPS) {
          // (computed by the next chunk itself)
        } else if constexpr (INT) {
          NA = ld_volatile_f2(fin + (ob + uw));
          NB = ld_volatile_f2(fin + (ob + 2u * uw));
        } else {
          NA = ld_volatile_f2(fin + ((unsigned)clampi(t + 2, 0, h - 1) * uw + (unsigned)x));
          NB = ld_volatile_f2(fin + ((unsigned)clampi(t + 3, 0, h - 1) * uw + (unsigned)x));
        }
#if OFB_EXP_L2PF > 0
        // One thread of the CTA pulls the rows of the chunk OFB_EXP_L2PF chunks ahead into L2 with bulk prefetches: the
        // strip's R0 and flow rows, and the R1 rows where this thread's displacement points (a smooth field moves the
        // whole strip's gather by about the same whole pixels).  With the producers' loads consumed right after their
        // issue (ncu: long scoreboard 6 warps per issue, the top stall), an L2 hit instead of a DRAM access is a third
        // of the wait.
        if constexpr (INT) {
          if (pf_on) {
            const int tp = min(t + 2 * OFB_EXP_L2PF, h - 2);
            const unsigned o0 = (unsigned)tp * uw + (unsigned)pf_xs;
            bulk_prefetch_l2(RA0 + o0, (unsigned)pf_n * 16u);
            bulk_prefetch_l2(RA0 + o0 + uw, (unsigned)pf_n * 16u);
            bulk_prefetch_l2(RB0 + o0, (unsigned)pf_n * 4u);
            bulk_prefetch_l2(RB0 + o0 + uw, (unsigned)pf_n * 4u);
            if constexpr (!UPS) {
              bulk_prefetch_l2(fin + o0, (unsigned)pf_n * 8u);
              bulk_prefetch_l2(fin + o0 + uw, (unsigned)pf_n * 8u);
            }
            // R1: rows tp + (iyb - yb) + 1, + 2 (the two bottom corner rows the chunk will load), columns shifted by ixb - x
            const int ry = min(max(tp + (iyb - yb) + 1, 0), h - 2);
            const int rx = min(max(pf_xs + ((ixb - x) & ~3), 0), w - pf_n);
            const unsigned o1 = (unsigned)ry * uw + (unsigned)rx;
            bulk_prefetch_l2(RA1 + o1, (unsigned)pf_n * 16u);
            bulk_prefetch_l2(RA1 + o1 + uw, (unsigned)pf_n * 16u);
            bulk_prefetch_l2(RB1 + o1, (unsigned)pf_n * 4u);
            bulk_prefetch_l2(RB1 + o1 + uw, (unsigned)pf_n * 4u);
          }
        }
#endif
        // consumers released this buffer (first round: the wait on the preceding phase of a fresh barrier passes)
        if (c >= 0) mbar_wait(bar_empty + 8u * buf, par ^ 1u);
        float V[5];
        // (reuse_b implies that B is inside: its corner offset is A's plus one row, never 0)
        const bool fast = INT && __all_sync(FULLMASK, reuse_b && !xborder);
        if (fast) {
          ring_step(um_arith(a0a, b0a, TA.q0, TA.q1, Y.q0, Y.q1, TA.s0, TA.s1, Y.s0, Y.s1, fxa, fya, FA_.x, FA_.y, true, false,
                             x, ya, w, h), V);
          if (c >= 0) {
#pragma unroll
            for (int ch = 0; ch < 5; ch++) sts_f32(stb + ch * COLS * 4, V[ch]);
          }
          ring_step(um_arith(a0b, b0b, Y.q0, Y.q1, BZ.q0, BZ.q1, Y.s0, Y.s1, BZ.s0, BZ.s1, fxb, fyb, FB_.x, FB_.y, true, false,
                             x, yb, w, h), V);
        } else {
          ring_step(um_arith(a0a, b0a, TA.q0, TA.q1, Y.q0, Y.q1, TA.s0, TA.s1, Y.s0, Y.s1, fxa, fya, FA_.x, FA_.y, ina,
                             xborder || (unsigned)(ya - 5) >= (unsigned)(h - 10), x, ya, w, h), V);
          if (c >= 0) {
#pragma unroll
            for (int ch = 0; ch < 5; ch++) sts_f32(stb + ch * COLS * 4, V[ch]);
          }
          // B's top row where it is not A's bottom row (flow discontinuity, floor crossing, outside pixel): loaded late
          // into A's top-row registers, free now — a wait for memory in the rare case instead of ten registers held
          // through every chunk
          if (!reuse_b) ldrow(TA, gb, rowb);
          UmRow W = TA;
          if (reuse_b) W = Y;                                        // (select per thread: 10 predicated moves)
          ring_step(um_arith(a0b, b0b, W.q0, W.q1, BZ.q0, BZ.q1, W.s0, W.s1, BZ.s0, BZ.s1, fxb, fyb, FB_.x, FB_.y, inb,
                             xborder || (unsigned)(yb - 5) >= (unsigned)(h - 10), x, yb, w, h), V);
        }
        if (c >= 0) {                                                // (a row past y1 keeps the state consistent; never read)
#pragma unroll
          for (int ch = 0; ch < 5; ch++) sts_f32(stb + ROWB + ch * COLS * 4, V[ch]);
          __syncwarp();
          if (lane0) mbar_arrive(bar_full + 8u * buf);               // this warp's columns of chunk c are staged
          stb += BUFB;
          if (++buf == NBUF) { buf = 0; stb = st0; par ^= 1u; }
        }
      };
      const std::true_type kInterior{};
      const std::false_type kEdge{};
      // rows t, t+1 off the border band; t+2, t+3 exist; fused upsample: only the exact x2 row mapping has an interior form
      // (on odd rows: the launcher makes the segments an even number of rows, so t_first = y0 - m is odd for odd m)
      const bool can_int = !UPS || (ups.exact2y != 0 && (t_first & 1) != 0);
      auto interior = [&](int t) { return can_int && t >= 5 && t <= h - 7; };
      float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0, f3 = f0;
      if constexpr (!UPS) {
        f0 = __ldg(fin + ((unsigned)clampi(t_first, 0, h - 1) * uw + (unsigned)x));
        f1 = __ldg(fin + ((unsigned)clampi(t_first + 1, 0, h - 1) * uw + (unsigned)x));
      }
      const int nv = m + n_chunks;                                   // the m warm-up chunks, then the output chunks
      int v = 0;
      while (v < nv) {
        const int t = t_first + 2 * v;
        if (v + 1 < nv && interior(t) && interior(t + 2)) {
          chunk(kInterior, t, v - m, S1, S2, f0, f1, f2, f3);
          chunk(kInterior, t + 2, v + 1 - m, S2, S1, f2, f3, f0, f1);
          v += 2;
        } else {
          chunk(kEdge, t, v - m, S1, S2, f0, f1, f2, f3);
          S1 = S2;
          f0 = f2; f1 = f3;
          v += 1;
        }
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
      unsigned par = 0;
      for (int c = 0; c < n_chunks; c++) {
        mbar_wait(bar_empty + 8u * buf, par ^ 1u);                   // consumers released this buffer
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
        __syncwarp();
        if (lane0) mbar_arrive(bar_full + 8u * buf);                 // this warp's columns of chunk c are staged
        if (++buf == NBUF) { buf = 0; par ^= 1u; }
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
      if (q0 + j >= halo && q0 + j < COLS - halo && ox + j < w) vmask |= 1u << j;
    constexpr unsigned ALL = (1u << PXT) - 1u;

    int buf = 0;
    unsigned par = 0;
    for (int c = 0; c < n_chunks; c++) {
      mbar_wait(bar_full + 8u * buf, par);             // every producer warp has staged chunk c
      const int yo = y0 + c * CH + q_row;
      if (yo < y1 && vmask) {
        OFB_DASSERT(buf >= 0 && buf < NBUF && q_row >= 0 && q_row < CH && q0 >= 0 && q0 + PXT <= COLS);
        OFB_DASSERT(yo >= y_begin && yo < y_end && yo < h);              // an output row of this launch
        const float* srow = stage + (buf * CH + q_row) * 5 * COLS;
        // shared address of this thread's first quad in channel 0 of its staged row (PAD columns left of its pixels).
        // Where the strip halo equals PAD (radius 7 / 15) every thread with a valid output reads inside the row: one
        // base address, immediate offsets.  Otherwise the quads clamp at the row ends (the clamped values are unused).
        constexpr int CPAD = MT > 0 ? (MT + 3) / 4 * 4 : 0;
        constexpr bool NOCLAMP = MT > 0 && iter_v_halo(MT, MT) == CPAD;
        const uint32_t c_row = (uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)((buf * CH + q_row) * 5 * COLS) * 4u;
        const uint32_t c_first = c_row + (uint32_t)(q0 - CPAD) * 4u;
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
              // always a whole 16-byte read: a quad of which only part is used would otherwise be narrowed to LDS.64 +
              // LDS.32, whose 16-byte lane stride is a 2-way / 4-way bank conflict (8 wavefronts instead of 4)
              if constexpr (NOCLAMP) {
                lds_f4(c_first + (uint32_t)(ch * COLS + 4 * v) * 4u, e[4 * v], e[4 * v + 1], e[4 * v + 2], e[4 * v + 3]);
              } else {
                const int cq = min(max(q0 - PAD + 4 * v, 0), COLS - 4);
                lds_f4(c_row + (uint32_t)(ch * COLS + cq) * 4u, e[4 * v], e[4 * v + 1], e[4 * v + 2], e[4 * v + 3]);
              }
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
#pragma unroll
        for (int j = 0; j < PXT; j++) OFB_DASSERT(!((vmask >> j) & 1u) || (ox + j >= 0 && ox + j < w && oi + j >= 0 && oi + j < w * h));
        float2* o = fout + (unsigned)max(oi, 0);
        if (PXT == 4 && vmask == ALL && (reinterpret_cast<uintptr_t>(o) & 31) == 0) {
          // 32-byte aligned (always, where strips start at multiples of 8 pixels and w % 4 == 0): the thread's four flow
          // vectors in ONE 256-bit store — a warp writes 1 KB of contiguous row in a single instruction
          stg_f8(o, f[0].x, f[0].y, f[1].x, f[1].y, f[2].x, f[2].y, f[3].x, f[3].y);
        } else if (vmask == ALL && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
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
      __syncwarp();
      if (lane0) mbar_arrive(bar_empty + 8u * buf);    // staging buffer may be refilled
      if (++buf == NBUF) { buf = 0; par ^= 1u; }
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
