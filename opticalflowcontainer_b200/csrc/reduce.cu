// reduce.cu — on-device reduction of the flow field to the node's scalar (mean / median of u).
#include "common.cuh"

namespace ofb {
int flow_u_stats(ofb_handle* h, int n, const uint8_t* host_mask, double* out_mean, float* out_median) {
  (void)n; (void)host_mask; (void)out_mean; (void)out_median;
  return set_error(h, OFB_ERR_INVALID_ARG, "ofb_flow_u_stats: not implemented yet");
}
}  // namespace ofb
