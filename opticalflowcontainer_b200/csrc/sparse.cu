// sparse.cu — sparse path (Shi-Tomasi + pyramidal LK).  Placeholder until the kernels land.
#include "common.cuh"

namespace ofb {
void sparse_destroy(ofb_handle*) {}
}  // namespace ofb

using namespace ofb;
extern "C" {
int ofb_good_features(ofb_handle* h, const uint8_t*, int, int, size_t, const ofb_gftt_params*, float*, int*) {
  return set_error(h, OFB_ERR_INVALID_ARG, "ofb_good_features: not implemented yet");
}
int ofb_corner_min_eigenval(ofb_handle* h, const uint8_t*, int, int, size_t, int, float*) {
  return set_error(h, OFB_ERR_INVALID_ARG, "ofb_corner_min_eigenval: not implemented yet");
}
int ofb_lk_pyramid(ofb_handle* h, const uint8_t*, int, int, size_t, int, int, int, uint8_t* const*, int16_t* const*, int*) {
  return set_error(h, OFB_ERR_INVALID_ARG, "ofb_lk_pyramid: not implemented yet");
}
int ofb_pyrlk(ofb_handle* h, const uint8_t*, const uint8_t*, int, int, size_t, const float*, int, float*, uint8_t*, float*,
              const ofb_lk_params*) {
  return set_error(h, OFB_ERR_INVALID_ARG, "ofb_pyrlk: not implemented yet");
}
}
