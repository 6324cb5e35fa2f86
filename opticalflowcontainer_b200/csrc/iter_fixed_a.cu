// iter_fixed_a.cu — k_iter_v with the window radius as a template argument (256-column strips, 2 CTAs/SM) for the
// window sizes 5..17 other than the default 15 (see fb_iter_launch.cuh, farneback.cu).  Radii up to 7 keep the
// vertical ring in tensor memory like the default kernel; radius 8 (17 ring slots) keeps it in shared memory.
#include "common.cuh"
#include "fb_iter_launch.cuh"

namespace ofb {

cudaError_t launch_iter_fixed_a(ofb_handle* h, int m, const float2* fin, float2* fout, int w, int hh, int n_pairs,
                                const RSet& rs, float reg, cudaStream_t st, bool* served) {
  *served = true;
  switch (m) {
    case 2: return launch_iter_v<2, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 3: return launch_iter_v<3, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 4: return launch_iter_v<4, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 5: return launch_iter_v<5, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 6: return launch_iter_v<6, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    case 8: return launch_iter_v<8, 256, 2, 2, 0, false, true, false, 2>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    default: break;
  }
  *served = false;
  return cudaSuccess;
}

}  // namespace ofb
