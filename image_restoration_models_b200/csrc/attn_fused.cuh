// Whole MDTA front (attn_fused.cu) behind norm1: qkv 1x1 + depthwise 3x3 + q.k^T Gram partials + squared norms + v store.
#pragma once
#include "common.cuh"

namespace irb {

struct AttnFusedArgs {
  const void* xn;          // [B*H*W][C] fp16: norm1(x)
  const void* w_qkv;       // qkv 1x1, fp16 SWIZZLE_128B operand image (PackMat fmt 4): [Kpad/64][attn_fused_wrows(C)][128 B]
  const float* dw_chunked; // taps [ceil(3C/32)][9][32] (launch_pack_dw_chunked with one set)
  void* v;                 // [B*H*W][C] fp16: depthwise-convolved v
  float* s_part;           // [B][heads][parts][ch][ch]
  float* n_part;           // [B][heads][parts][2][ch]
  int parts;               // must equal attn_fused_parts(B, H, W)
  int B, H, W, C, heads;
};

bool attn_fused_supported(int C, int heads);
int  attn_fused_wrows(int C);                 // rows of the packed W_qkv image (3C rounded up to whole 32-channel units)
int  attn_fused_parts(int B, int H, int W);
int  launch_attn_fused(const AttnFusedArgs& a, cudaStream_t s);

}  // namespace irb
