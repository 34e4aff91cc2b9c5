// Fused MDTA front (attn_front.cu): depthwise 3x3 over qkv + q.k^T Gram partials + squared norms + v store.
#pragma once
#include "common.cuh"

namespace irb {

struct AttnFrontArgs {
  const void* qkv;         // [B*H*W][3C] qkv 1x1 output (fp32, or fp16 when half)
  int half;
  int v_half;              // fp32 qkv, but v stored as fp16 (the attention-output contraction then runs on fp16 operands)
  void* v;                 // [B*H*W][C] depthwise-convolved v (fp16 when half || v_half, else fp32 rounded to tf32)
  const float* dw_chunked; // taps [ceil(3C/32)][9][32] (launch_pack_dw_chunked with one set)
  float* s_part;           // [B][heads][parts][ch][ch]
  float* n_part;           // [B][heads][parts][2][ch]
  int parts;               // must equal attn_front_parts(B, H, W, C, heads)
  int B, H, W, C, heads;
};

bool attn_front_supported(int C, int heads, bool half);
int  attn_front_parts(int B, int H, int W, int C, int heads);
int  launch_attn_front(const AttnFrontArgs& a, cudaStream_t s);

}  // namespace irb
