// fb_iter_launch.cuh — launcher of the fused iteration kernel k_iter_v (grid geometry, shared-memory opt-in), shared by
// farneback.cu, tiled.cuh and the translation units that hold the extra window-size instantiations.
#pragma once
#include <algorithm>
#include <cstring>

#include "fb_iter_v.cuh"

// Fewest rows of a whole-frame row segment (experiment builds: -DOFB_EXP_MIN_SEG_ROWS=n; measured 8 / 12 / 16).
#ifndef OFB_EXP_MIN_SEG_ROWS
#define OFB_EXP_MIN_SEG_ROWS 8
#endif

namespace ofb {

// k_iter_v launcher.  ups != nullptr: the launch is the first iteration of a level and upsamples its input flow from
// the coarser level on the fly (fin is not read).
template <int MT, int COLS, int CH, int MINB, int PFD, bool TILED, bool REUSE, bool TMEM, int NBUF, bool UPS = false>
static cudaError_t launch_iter_v(ofb_handle* h, const float2* fin, float2* fout, int w, int hh, int n_pairs,
                                 const RSet& rs, int m, float reg, cudaStream_t st, const UpsSrc* ups = nullptr,
                                 int y_begin = 0, int y_end = -1, const PeerTab* tab = nullptr, int my_rank = 0) {
  if (y_end < 0) y_end = hh;
  if (y_end <= y_begin) return cudaSuccess;
  PeerTab t;
  if (tab) t = *tab; else memset(&t, 0, sizeof(t));
  UpsSrc u;
  if (ups) u = *ups; else memset(&u, 0, sizeof(u));
  if (UPS != (ups != nullptr)) return cudaErrorInvalidValue;   // the instantiation and the launch must agree
  auto kern = k_iter_v<MT, COLS, CH, MINB, PFD, TILED, REUSE, TMEM, NBUF, UPS>;
  const int smem = iter_v_smem_floats<COLS, CH>(m, TMEM, NBUF) * (int)sizeof(float);
  // largest dynamic smem configured for this instantiation, per device (function attributes are per device)
  static int configured[64] = {0};
  const int dev = h->device & 63;
  if (smem > configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (REUSE && COLS == 256 && MINB == 2) {
      // The setmaxnreg schedule moves registers between the warpgroups of a CTA: 2 x 128 producers x 96 + 128 consumers
      // x 48 = 384 x 80.  It only works if the kernel really launches with 80 registers per thread (a smaller pool
      // would leave the producers waiting for registers for ever): checked once per instantiation.
      cudaFuncAttributes a;
      e = cudaFuncGetAttributes(&a, kern);
      if (e != cudaSuccess) return e;
      if (a.numRegs < 80) return cudaErrorLaunchOutOfResources;
    }
    configured[dev] = smem;
  }
  const int tw = COLS - 2 * iter_v_halo(MT, m);
  const int strips = (w + tw - 1) / tw;
  const int slots = MINB * h->num_sms;
  const int per = strips * n_pairs;
  const int rows = y_end - y_begin;
  // Row segments: a CTA marches its segment row by row (plus 2m warm-up rows), the CTAs run in waves of `slots`, so the
  // launch takes about  waves(segs) * (seg_rows + 2m)  row steps.  The number of segments minimises that (fewest
  // segments on a tie): whole frames at the benchmark batch keep one full wave (1080p x 18 pairs: 2 segments, as before),
  // while small frames in large batches — VGA x 72: 216 strip columns, 0.73 of a wave — are cut so that the waves are
  // full (4 segments: 2.9 waves of 134 rows instead of one of 494).
  // Tiled mode (one pair, narrow bands): at least 6 rows per segment; whole frames: at least 8 (a single VGA pair:
  // 0.211 -> 0.180 ms against a minimum of 16; nothing changes at the benchmark batch).  The vertical block sums restart
  // with a segment, so their float rounding — the last bits of the flow — follows the segmentation, and with it the
  // batch size: calls over different batch sizes agree to ~1e-6 px, not bit for bit.
  const int min_rows = REUSE && TILED ? 6 : OFB_EXP_MIN_SEG_ROWS;
  int segs = 1, seg_rows = rows + (rows & 1);
  {
    long best = -1;
    const int max_segs = std::min(64, std::max(1, rows / min_rows));
    for (int sgm = 1; sgm <= max_segs; sgm++) {
      int sr = std::max(min_rows, (rows + sgm - 1) / sgm);
      sr += sr & 1;       // even: the chunks of the fused upsample then start on odd rows (fb_iter_v.cuh)
      const int ns = (rows + sr - 1) / sr;
      const long waves = ((long)per * ns + slots - 1) / slots;
      const long cost = waves * (sr + 2 * m);
      if (best < 0 || cost < best) { best = cost; segs = ns; seg_rows = sr; }
    }
  }
  dim3 g(strips * segs, n_pairs);
  kern<<<g, COLS + CH * COLS / 4, smem, st>>>(rs, fin, fout, w, hh, m, reg, seg_rows, strips, y_begin, y_end, t, my_rank, u);
  return cudaGetLastError();
}

}  // namespace ofb
