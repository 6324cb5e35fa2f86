// iter_fixed_a.cu — k_iter_v with the window radius as a template argument (256-column strips, 2 CTAs/SM) for the
// window sizes 5..17 other than the default 15 (see fb_iter_launch.cuh, farneback.cu).  The vertical ring lives in tensor
// memory like the default kernel's; radius 8 (17 ring slots) packs it (fb_iter_v.cuh, tmem_ring_packed).
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
    case 8: return launch_iter_v<8, 256, 2, 2, 0, false, true, true, 4>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st, nullptr);
    default: break;
  }
  *served = false;
  return cudaSuccess;
}

}  // namespace ofb
