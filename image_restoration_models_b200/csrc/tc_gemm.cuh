// Parameter block of the tcgen05 contraction (tc_gemm.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace irb {

struct TcGemmParams {
  // A operand: rows = pixels, source 1 channels [0,k1) then optional source 2 channels [k1, k1+k2) (channel concat)
  // (element type: fp32, or fp16 when a_half)
  const void* a1; int lda1; int k1;
  const void* a2; int lda2; int k2;
  int B, HW;
  // weights packed [K/epc][N][epc] (epc = 4 tf32-rounded fp32, or 8 fp16 when op_half; pack.cu, fold kernel);
  // w_bstride != 0 (in elements): one matrix per image
  const void* w; long long w_bstride;
  int N, K;
  const float* bias;
  int ln_mode; const float* ln_w; const float* ln_b;     // LayerNorm prologue (K <= 128, single source)
  const float* r; int ldr;                               // residual added in the epilogue (may alias y)
  void* y; int ldy;                                      // fp32, or fp16 when y_half
  int tag;
  int a_pad;                                             // 0/1: extra row per K-chunk slab of A in smem
  int a_half, op_half, y_half;                           // element types (0 = fp32 / tf32 operand, 1 = fp16)
  // 3x3 convolution as implicit GEMM: a_mode 1 gathers k = tap*cin + c from the pixel shifted by the tap
  // (zero padding 1); k1 = cin, K = 9*cin, H x W = spatial extent of one image (HW = H*W)
  int a_mode, H, W;
  int o_mode;                                            // OMode: plain rows, or PixelUnshuffle / PixelShuffle scatter
  int relu;
  int n_valid;                                           // scatter epilogue: output channels >= n_valid are padding (0 = N)
  float acc_sign;                                        // y = r + acc_sign * act(acc + bias)   (0 is treated as +1)
  // optional second output of the TMA-fed kernel: xn = LayerNorm(y) as fp16 [rows][N] (the next kernel's tensor-core
  // operand), computed in the epilogue while the row is in registers (N <= 96, one N-chunk, fp32 y)
  void* xn; int ldxn; int xn_ln_mode; const float* xn_w; const float* xn_b;
  int w_stream;                                          // filled by configure: weights streamed per K-chunk
  // filled by tc_gemm_configure / launch_gemm_tc
  // NC columns per CTA = nsub sub-chunks of NS <= 256 columns (one tcgen05.mma each); nacc accumulator buffers
  int NC, NS, nsub, sub_stride, nacc, KC, lpp, upl, unr, stages, acc_stride, tmem_cols, tiles_per_img, ntiles;
};

size_t tc_gemm_configure(TcGemmParams& p);
int launch_gemm_tc(TcGemmParams p, cudaStream_t s);

// TMA-fed second-generation kernel (tma_gemm.cu).  Weights are packed [K/kc][N][128 B, 16-byte chunks XOR-swizzled by
// n & 7] (kc = 32 tf32 / 64 fp16 elements; PackMat fmt 3 / 4, K zero-padded to tma_gemm_kpad).  launch_gemm_tma returns
// IR_UNSUPPORTED_SHAPE when the problem has to go to launch_gemm_tc instead.
constexpr int IR_UNSUPPORTED_SHAPE = 1000;
bool tma_gemm_shape_supported(int K, int N, bool op_half, bool ln, bool has_r, bool y_half);
int tma_gemm_kpad(int K, bool op_half);
int launch_gemm_tma(const TcGemmParams& t, cudaStream_t s);
// can launch_gemm_tma also emit xn = LayerNorm(y) for this residual contraction (K = N = C)?
bool tma_gemm_xn_supported(int C, bool op_half);

// TMA-fed 3x3 implicit GEMM (tma_conv3.cu): PixelUnshuffle / PixelShuffle scatter, or plain rows with bias + ReLU
// (DnCNN body); fp32 channels-last in and out; weights packed with PackMat kind 2, fmt 3 / 4, K pitch 9 * tma_conv3_kpt.
bool tma_conv3_supported(int cin, int cout_p, bool half);
int  tma_conv3_kpt(int cin, bool half);
int  launch_conv3_tma(const float* in, int ld_in, int cin, const void* w_packed, const float* bias, int relu, int cout_p,
                      int cout_valid, int B, int H, int W, float* out, int ld_out, int o_mode, bool half, cudaStream_t s);

// Row-strip variant for Cin <= 64 (tma_conv3_row.cu): every image row is loaded once and the nine taps are row-shifted
// descriptors into the strips.  Same weight layout as launch_conv3_tma; plain rows (bias + ReLU) or PixelUnshuffle.
bool conv3_row_supported(int cin, int cout_p, bool half);
int  launch_conv3_row(const float* in, int ld_in, int cin, const void* w_packed, const float* bias, int relu, int cout_p,
                      int cout_valid, int B, int H, int W, float* out, int ld_out, int o_mode, bool half, cudaStream_t s);

}  // namespace irb
