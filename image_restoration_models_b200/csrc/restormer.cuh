// Plan / workspace / launch-sequence declarations shared by restormer.cu, dncnn.cu and api.cu.
#pragma once
#include <vector>

#include "common.cuh"

namespace irb {

// One weight-packing step: parameter `param` (state_dict order) -> packed[dst ...].
struct PackOp {
  enum Kind { VEC = 0, MAT1 = 1, MAT3 = 2, DW = 3, DWC = 4 };   // DWC: depthwise taps chunked for ffn_tail.cu (k_dst = chunk width)
  int kind;
  int param;
  long long dst;       // float offset into the packed buffer
  int a, b, c;         // src_half, dst_half, n_halves (split-pad row/channel mapping)
  int k_src, k_dst;    // MAT1/MAT3: logical / padded reduction length
  int cin;             // MAT3
  int fmt;             // MAT1: 0 row-major fp32, 1 tcgen05 operand layout (PackMat::fmt)
};

struct BlockPlan {     // one TransformerBlock (restormer.py:137-150); offsets in floats, -1 = absent
  int C, heads, h, hp;
  long long ln1_w, ln1_b, temp, qkv_w, qkv_b, qkvdw_w, qkvdw_b, proj_w, proj_b;
  long long ln2_w, ln2_b, pin_w, pin_b, ffdw_w, ffdw_b, pout_w, pout_b;
  bool tc_qkv, tc_attn, tc_pin, tc_pout;   // which 1x1 contractions run on the tcgen05 kernel
  bool tma_qkv, tma_attn, tma_pin, tma_pout;   // ... and of those, which on the TMA-fed kernel (tma_gemm.cu; weights in its layout)
  bool fuse_front;                         // qkv dwconv + Gram + norms + v store in one kernel (attn_front.cu)
  bool fuse_attn;                          // norm1 + qkv 1x1 + all of the above in one kernel (attn_fused.cu): qkv never exists in HBM
  bool fuse_tail;                          // dwconv + gate + project_out + residual in one kernel (ffn_tail.cu)
  bool fuse_ffn;                           // the whole GDFN in one kernel behind a standalone LayerNorm (ffn_fused.cu)
  bool v_half;                             // fp32 mode: v and the folded attention matrix are fp16 operands (same mantissa as tf32)
  bool k4_xn;                              // the attention-output contraction also emits norm2(x) as the fused GDFN's fp16 operand
  int kp_attn;                             // K pitch of the folded attention matrix W_eff (padded for the TMA kernel)
  bool ref_kernels;                        // ENGINE_SIMT: reference CUDA-core kernels everywhere
  bool half;                               // 16-bit plan: fp16 intermediates (qkv, v, hidden, gated) and fp16 operands
  bool wide16 = false;                     // ... chosen for this block by IR_MODE_FP32 (C > 128), not by the mode
};

struct ConvPlan { int cout, cin, k, kp; long long w, b; bool tc; int cout_p; bool tma = false; };   // cout_p: rows in the packed weight (>= cout, zero rows)

struct RestormerPlan {
  IrRestormerCfg cfg;
  bool half = false;
  std::vector<PackOp> ops;
  int n_params = 0;
  long long packed_floats = 0;
  ConvPlan patch_embed, down[3], up[3], reduce[3], skip, output;
  std::vector<BlockPlan> enc[4], dec[3], refine;
};

struct BlockScratchNeed { long long qkv = 0, hidden = 0, gated = 0, s_part = 0, n_part = 0, w_eff = 0, xhat = 0, xhat2 = 0, v16 = 0; };   // floats (v16: fp16 elements)   // v16: fp16 elements of v where the fused front writes it
struct BlockScratch { void *qkv, *qkv_dw, *hidden, *gated; float *s_part, *n_part; void *w_eff, *xhat, *xhat2; };   // xhat: norm2 / wide-level LayerNorm output, xhat2: fp16 norm1 output of the fused MDTA front
struct RestormerWs { float* e[4]; float* e1_in; float* d[3]; float* up_tmp; BlockScratch bs; };

int  block_param_count(int bias, int ln_bias);
// engine: 0 = tcgen05 contractions with tf32 operands and fp32 intermediates; 1 = CUDA-core fp32 contractions and
// reference kernels (on-device second oracle); 2 = tcgen05 contractions with fp16 operands and fp16 intermediates
// (the residual stream, LayerNorm statistics, Gram accumulation, softmax and GELU stay fp32)
// 3 = as 0 but without any fp16 tensor (v, norm2 output and the GDFN hidden tensor stay fp32 / tf32): IR_MODE_FP32_STRICT
enum Engine { ENGINE_TC = 0, ENGINE_SIMT = 1, ENGINE_TC_HALF = 2, ENGINE_TC_STRICT = 3 };
int  build_block_plan(BlockPlan& bp, std::vector<PackOp>& ops, long long& packed_floats, int C, int heads, double ffn,
                      int bias, int ln_bias, int engine);
int  build_restormer_plan(RestormerPlan& pl, const IrRestormerCfg& cfg, int engine);
bool tc_gemm_supported(int K, int N, bool half);
bool tc_conv3_supported(int cin, int cout, bool half);
// 3x3 convolution (stride 1, zero pad 1) on the tcgen05 kernel: fp32 channels-last in, fp32 out (o_mode rows / shuffle scatter)
int  run_conv3_tc(const float* in, int ld_in, int cin, const float* w_packed, const float* bias, int cout, int cout_valid,
                  int B, int H, int W, float* out, int ld_out, int o_mode, int relu, bool half, cudaStream_t s);
long long pack_op_src_numel(const PackOp& op);
int  run_pack_ops(const std::vector<PackOp>& ops, const float* const* params, float* packed, cudaStream_t s);

size_t block_workspace_bytes(const BlockPlan& bp, int B, int H, int W);
int  engine_of_mode(int mode);
size_t restormer_workspace_bytes(const RestormerPlan& pl, int B, int H, int W);
// next: the block that consumes x_out (same C, same extent) or nullptr; when both blocks are fused, this block's GDFN also
// writes the next block's norm1 output (bs.xhat2) and *xn1_ready is set for the next call.  xn1_ready (in): bs.xhat2
// already holds norm1(x_in).
int  run_block(const BlockPlan& bp, const float* packed, const float* x_in, float* x_out, int B, int H, int W,
               const BlockScratch& bs, int ln_with_bias, cudaStream_t s, const BlockPlan* next = nullptr,
               bool* xn1_ready = nullptr);
// true when `bp`'s GDFN can emit norm1 of `next` (both on the fused kernels, same width)
bool block_chains_norm1(const BlockPlan& bp, const BlockPlan& next);
int  block_forward(const BlockPlan& bp, const float* packed, float* x, int B, int H, int W, void* workspace,
                   size_t workspace_bytes, int ln_with_bias, cudaStream_t s);
int  restormer_launch_count(const RestormerPlan& pl);
int  restormer_forward(const RestormerPlan& pl, const float* packed, const float* x, float* y, int B, int H, int W,
                       void* workspace, size_t workspace_bytes, cudaStream_t s);

}  // namespace irb
