/*
 * irb200_testing.h -- test hooks of libirb200.so: NOT part of the product ABI (include/irb200.h).
 *
 * Per-kernel entry points for unit tests and ncu (one contraction kernel in isolation with reference-layout weights), and
 * the hardware probe behind the row-strip 3x3 convolution.  They are built into the library when it is compiled with
 * -DIRB200_TESTING (the in-tree build does, the tests call them through ctypes); a deployment build without the flag
 * exports only irb200.h.
 */
#ifndef IRB200_TESTING_H_
#define IRB200_TESTING_H_

#include "irb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One 1x1 convolution y[pix, n] = sum_k LN?(a)[pix, k] * w[n, k] (+bias) (+r) on channels-last rows, with the
 * reference's row-major weight [N][K] (K = k1 + k2, second source = channel concat).  engine 0 = tcgen05
 * kernel, 1 = CUDA-core fp32 kernel.  ln_mode: 0 none, 1 BiasFree, 2 WithBias.  scratch >= (N*K + B*HW*K)*4 bytes. */
int    ir_test_conv1x1(int engine, const void* a1, int lda1, int k1, const void* a2, int lda2, int k2,
                       const float* w_rowmajor, const float* bias, int ln_mode, const float* ln_w, const float* ln_b,
                       const float* r, int ldr, void* y, int ldy, int B, int HW, int N, int a_pad,
                       int a_half, int op_half, int y_half,   /* element types: a1/a2, tensor-core operands, y */
                       void* scratch, size_t scratch_bytes, void* stream);

/* One 3x3 convolution (stride 1, zero padding 1, PyTorch [cout][cin][3][3] weight) on channels-last fp32 rows.
 * o_mode: 0 plain rows y[pix*ldy + n]; 1 PixelUnshuffle(2) folded into the store; 2 PixelShuffle(2) folded into the
 * store (restormer.py:176,186).  engine 0 = tcgen05 implicit GEMM, 1 = CUDA-core fp32.  scratch >= cout*9*cin*4 B. */
int    ir_test_conv3x3(int engine, const float* x_nhwc, int ldx, int cin, const float* w_oihw, const float* bias,
                       int cout, int B, int H, int W, float* y, int ldy, int o_mode, int relu, int op_half,
                       void* scratch, size_t scratch_bytes, void* stream);

/* Hardware probe (bring-up): D[128][32] = A[shift : shift+128][32] . W[32][32]^T with the A operand descriptor's start
 * address shifted by `shift` rows inside one TMA-written SWIZZLE_128B box and `base_off` in its base-offset field. */
int    ir_probe_shifted_descriptor(const float* a /* [160][32] */, const float* w /* [32][32] */, float* d /* [128][32] */,
                                   int shift, int base_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IRB200_TESTING_H_ */
