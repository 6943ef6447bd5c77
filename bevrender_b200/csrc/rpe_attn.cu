// K3 placeholder while the fused kernel is being written (replaced in the next commit).
#include "bevr_common.cuh"
extern "C" int bevr_rpe_attn_fwd(const float*, const float*, const float*, const float*, const float*, float*, float*,
                                 int, int, int, int, int, int, int, int, int, float, void*) {
  bevr::set_error("rpe_attn_fwd: not built into this library yet");
  return -2;
}
extern "C" int bevr_rpe_attn_bwd(const float*, const float*, const float*, const float*, const float*, const float*,
                                 const float*, const float*, float*, float*, float*, float*, float*, float*, int, int,
                                 int, int, int, int, int, int, int, float, void*) {
  bevr::set_error("rpe_attn_bwd: not built into this library yet");
  return -2;
}
