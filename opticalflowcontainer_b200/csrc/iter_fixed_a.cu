// iter_fixed_a.cu — k_iter_v with the window radius as a template argument, default schedule (two rows of loads in flight, row-reuse
// gather), for window sizes other than the default 15 (see fb_iter_launch.cuh, farneback.cu).
#include "common.cuh"
#include "fb_iter_launch.cuh"

namespace ofb {

cudaError_t launch_iter_fixed_a(ofb_handle* h, int m, const float2* fin, float2* fout, int w, int hh, int n_pairs, const RSet& rs,
                      float reg, cudaStream_t st, bool* served) {
  *served = true;
  switch (m) {
    case 2: {
      static const bool ok = iter_regs_ok(k_iter_v<2, 256, 2, 2, 0, 4, 2, 1, false, true>);
      if (!ok) break;
      return launch_iter_v<2, 256, 2, 2, 0, 4, 2, 1, false, true>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st);
    }
    case 3: {
      static const bool ok = iter_regs_ok(k_iter_v<3, 256, 2, 2, 0, 4, 2, 1, false, true>);
      if (!ok) break;
      return launch_iter_v<3, 256, 2, 2, 0, 4, 2, 1, false, true>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st);
    }
    case 4: {
      static const bool ok = iter_regs_ok(k_iter_v<4, 256, 2, 2, 0, 4, 2, 1, false, true>);
      if (!ok) break;
      return launch_iter_v<4, 256, 2, 2, 0, 4, 2, 1, false, true>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st);
    }
    case 5: {
      static const bool ok = iter_regs_ok(k_iter_v<5, 256, 2, 2, 0, 4, 2, 1, false, true>);
      if (!ok) break;
      return launch_iter_v<5, 256, 2, 2, 0, 4, 2, 1, false, true>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st);
    }
    case 6: {
      static const bool ok = iter_regs_ok(k_iter_v<6, 256, 2, 2, 0, 4, 2, 1, false, true>);
      if (!ok) break;
      return launch_iter_v<6, 256, 2, 2, 0, 4, 2, 1, false, true>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st);
    }
    case 8: {
      static const bool ok = iter_regs_ok(k_iter_v<8, 256, 2, 2, 0, 4, 2, 1, false, true>);
      if (!ok) break;
      return launch_iter_v<8, 256, 2, 2, 0, 4, 2, 1, false, true>(h, fin, fout, w, hh, n_pairs, rs, m, reg, st);
    }
    default: break;
  }
  *served = false;
  return cudaSuccess;
}

}  // namespace ofb
