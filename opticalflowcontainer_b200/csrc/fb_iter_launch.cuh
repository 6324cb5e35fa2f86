// fb_iter_launch.cuh — launcher of the fused iteration kernel k_iter_v (grid geometry, shared-memory opt-in), shared by
// farneback.cu, tiled.cuh and the translation units that hold the extra window-size instantiations.
#pragma once
#include <algorithm>
#include <cstring>

#include "fb_iter_v.cuh"

namespace ofb {

// k_iter_v launcher.
template <int MT, int COLS, int CH, int MINB, int PFD, int PXT, int RIF = 1, int CLOOP = 1, bool TILED = false,
          bool REUSE = false>
static cudaError_t launch_iter_v(ofb_handle* h, const float2* fin, float2* fout, int w, int hh, int n_pairs,
                                 const RSet& rs, int m, float reg, cudaStream_t st, int y_begin = 0, int y_end = -1,
                                 const PeerTab* tab = nullptr, int my_rank = 0) {
  if (y_end < 0) y_end = hh;
  if (y_end <= y_begin) return cudaSuccess;
  PeerTab t;
  if (tab) t = *tab; else memset(&t, 0, sizeof(t));
  const int smem = iter_v_smem_floats<COLS, CH>(m) * (int)sizeof(float);
  // largest dynamic smem configured for this instantiation, per device (function attributes are per device)
  static int configured[64] = {0};
  const int dev = h->device & 63;
  if (smem > configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_iter_v<MT, COLS, CH, MINB, PFD, PXT, RIF, CLOOP, TILED, REUSE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured[dev] = smem;
  }
  const int tw = COLS - 2 * m;
  const int strips = (w + tw - 1) / tw;
  const int slots = MINB * h->num_sms * h->iter_waves;
  const int per = strips * n_pairs;
  const int rows = y_end - y_begin;
  int segs = per >= slots ? 1 : slots / per;
  int seg_rows = std::max(16, (rows + segs - 1) / segs);
  segs = (rows + seg_rows - 1) / seg_rows;
  dim3 g(strips * segs, n_pairs);
  k_iter_v<MT, COLS, CH, MINB, PFD, PXT, RIF, CLOOP, TILED, REUSE><<<g, COLS + (CH / CLOOP) * COLS / PXT, smem, st>>>(
      rs, fin, fout, w, hh, m, reg, seg_rows, strips, y_begin, y_end, t, my_rank);
  return cudaGetLastError();
}


// A setmaxnreg schedule (MINB == 2 instantiations) is only safe if the kernel really launches with 80 registers per
// thread: 2 x 128 producers x 96 + 128 consumers x 48 = 384 x 80.
template <typename K>
static bool iter_regs_ok(K kernel) {
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, kernel) == cudaSuccess && a.numRegs >= 80;
}

}  // namespace ofb
