// Whole GDFN (ffn_fused.cu): project_in + depthwise 3x3 + GELU gate + project_out + residual in one tcgen05 kernel.
#pragma once
#include "common.cuh"

namespace irb {

struct FfnFusedArgs {
  const void* xn;          // [B*H*W][C] fp16: LayerNorm(x) (norm2, restormer.py:148), the project_in operand
  float* x;                // [B*H*W][C] fp32 residual stream, updated in place (x += ffn(xn))
  const void* w_in;        // project_in, fp16 SWIZZLE_128B operand image (PackMat fmt 4): [Kpad/64][2*hp][128 B]
  const void* w_out;       // project_out, fmt 4: [hp/64][C][128 B]
  const float* dw_chunked; // depthwise taps [hp/64][2][9][64] (launch_pack_dw_chunked, kc = 64)
  int B, H, W, C, hp;
  // optional: also write LayerNorm(x_new) -- the NEXT block's norm1 (restormer.py:147) -- as an fp16 tensor; the epilogue then
  // loads the residual instead of reducing into it
  void* xn_next = nullptr;          // [B*H*W][C] fp16
  const float* ln_w_next = nullptr; // [C]
  const float* ln_b_next = nullptr; // [C] (WithBias) or nullptr
  int ln_mode_next = 0;             // LN_BIASFREE / LN_WITHBIAS
};

bool ffn_fused_supported(int C, int hp);
int  launch_ffn_fused(const FfnFusedArgs& a, cudaStream_t s);

}  // namespace irb
