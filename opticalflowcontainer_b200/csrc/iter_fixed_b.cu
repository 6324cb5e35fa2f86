// iter_fixed_b.cu — k_iter_v with the window radius as a template argument for the window sizes 19..31 (see
// fb_iter_launch.cuh, farneback.cu): the default geometry (256-column strips, 2 CTAs/SM) with the ring packed into
// tensor memory; radii 13..15 keep the ring slots beyond the 25 that fit there in shared memory.
#include "common.cuh"
#include "fb_iter_launch.cuh"

namespace ofb {

cudaError_t launch_iter_fixed_b(ofb_handle* h, int m, const float2* fin, float2* fout, int w, int hh, int n_pairs,
                                const RSet& rs, float reg, cudaStream_t st, bool* served) {
  *served = true;
  switch (m) {
    case 9: return launch_iter_v<9, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 10: return launch_iter_v<10, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 11: return launch_iter_v<11, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 12: return launch_iter_v<12, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 13: return launch_iter_v<13, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 14: return launch_iter_v<14, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 15: return launch_iter_v<15, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    default: break;
  }
  *served = false;
  return cudaSuccess;
}

}  // namespace ofb
