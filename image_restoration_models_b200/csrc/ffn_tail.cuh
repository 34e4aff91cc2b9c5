// Fused GDFN tail (ffn_tail.cu): depthwise 3x3 + GELU gate + project_out + residual in one tcgen05 kernel.
#pragma once
#include "common.cuh"

namespace irb {

struct FfnTailArgs {
  const void* hidden;      // [B*H*W][2*hp] project_in output (fp32, or fp16 when half)
  int half;
  float* x;                // [B*H*W][C] residual stream, updated in place
  const void* w_out;       // project_out in the SWIZZLE_128B operand image (PackMat fmt 3 / 4), K = hp
  const float* dw_chunked; // depthwise taps [hp/kc][2][9][kc] (launch_pack_dw_chunked), kc = ffn_tail_kc(half)
  const float* bias;       // project_out bias or nullptr
  int B, H, W, C, hp;
};

bool ffn_tail_supported(int C, int hp, bool half);
int  ffn_tail_kc(bool half);
int  launch_ffn_tail(const FfnTailArgs& a, cudaStream_t s);
// dst[((chunk*nsets + set)*9 + tap)*kc + c] = src[(set*h + chunk*kc + c)*9 + tap]   (zero for padded channels >= h;
// hp = h padded to a multiple of kc; nsets = 2 for the GDFN halves, 1 for the qkv depthwise conv)
int  launch_pack_dw_chunked(const float* src, float* dst, int h, int hp, int kc, int nsets, cudaStream_t s);

}  // namespace irb
