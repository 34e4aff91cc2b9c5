#pragma once
#include <vector>

#include "common.cuh"

namespace irb {

struct DncnnLayer { int cin, cout, k, kp; long long w, b; int p_w, p_b, p_bn; bool tc; bool tma; };   // tma: tma_conv3.cu (weights in its layout)
struct DncnnPlan {
  IrDncnnCfg cfg;
  bool half = false;
  std::vector<DncnnLayer> layers;
  long long bn_scale = 0, bn_shift = 0;   // scratch inside the packed buffer used while folding BatchNorm
  int n_params = 0;
  long long packed_floats = 0;
};

int build_dncnn_plan(DncnnPlan& pl, const IrDncnnCfg& cfg, int engine);
long long dncnn_param_numel(const DncnnPlan& pl, int index);
int dncnn_pack(const DncnnPlan& pl, const float* const* params, float* packed, cudaStream_t s);
size_t dncnn_workspace_bytes(const DncnnPlan& pl, int B, int H, int W);
int dncnn_forward(const DncnnPlan& pl, const float* packed, const float* x, float* y, int B, int H, int W,
                  void* workspace, size_t workspace_bytes, cudaStream_t s);

}  // namespace irb
