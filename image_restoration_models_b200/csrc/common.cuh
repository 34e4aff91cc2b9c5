// Shared declarations for the sm_100a Restormer / DnCNN forward library.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

#include "../../include/irb200.h"

namespace irb {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, int status; SURVEY.md §8b "Errors")
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define IRB_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return ::irb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define IRB_LAUNCH_CHECK() IRB_CUDA(cudaGetLastError())

#define IRB_REQUIRE(cond, msg)                                                \
  do {                                                                        \
    if (!(cond)) { ::irb::set_error(std::string("invalid argument: ") + (msg)); return IR_ERR_INVALID; } \
  } while (0)

#define IRB_TRY(expr)                                                         \
  do { int _s = (expr); if (_s != IR_OK) return _s; } while (0)

__host__ __device__ static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// fp32 pair -> fp16 pair, round to nearest, SATURATING: +-65504 instead of inf when an activation leaves fp16's range
// (one F2FP.SATFINITE instruction, the cost of the plain conversion).  Every activation conversion in the library goes
// through this; see IR_MODE_FP32_STRICT / the Python range guard for models that need more than fp16's 5 exponent bits.
__device__ __forceinline__ __half2 f2h2_sat(float lo, float hi) {
  uint32_t r;
#ifdef IRB_BF16_BUILD
  asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#else
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#endif
  return *reinterpret_cast<__half2*>(&r);
}
#endif

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (opt-in, see pdl_enabled).  Every kernel of the forward can be launched with the
// programmatic-stream-serialisation attribute (launch_pdl) and follows ONE rule: its set-up (barrier initialisation, tensor-memory allocation, tensor-map
// prefetch: nothing that reads or writes global memory another kernel of the forward touches) runs first, then
// pdl_wait() -- which returns once the preceding grid has completed and flushed -- and only then the first global access.
// pdl_trigger() sits next to it: the next grid's blocks may then take an SM the moment this grid's block leaves it and run
// their own set-up under this grid's tail.  Both instructions are no-ops in a launch without the attribute.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

// OFF by default: measured on B200 (profiles/r02_pdl_ab.json) it shortens back-to-back plain launches of batch-1 forwards
// (256x256 tile 5.31 -> 4.82 ms) but not the CUDA-graph replay the latency path uses (4.73 ms either way), and at batch 8 the
// step is 0.1-0.8 % SLOWER (the early-resident blocks of the next grid cost more than the set-up they hide).  IRB_PDL=1 turns it on.
static inline bool pdl_enabled() {
  static const bool on = getenv("IRB_PDL") != nullptr;
  return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

// Opt a kernel into `bytes` of dynamic shared memory once per DEVICE (the attribute is per device, and one process may
// drive several); `done` is the caller's per-kernel table.
struct SmemOptIn { bool done[64] = {}; };
template <typename K>
static inline int opt_in_smem(K kernel, SmemOptIn& st, int bytes = 227 * 1024) {
  int dev = 0;
  IRB_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && st.done[dev]) return IR_OK;
  IRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev >= 0 && dev < 64) st.done[dev] = true;
  return IR_OK;
}
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// optional per-launch device timing (api.cu); tags name the kernel family a launch belongs to
// ---------------------------------------------------------------------------------------------
enum KernelTag {
  TAG_OTHER = 0, TAG_LN_QKV, TAG_DW_QKV, TAG_GRAM, TAG_FOLD, TAG_ATTN_OUT, TAG_LN_PIN, TAG_DW_GATE, TAG_FFN_OUT,
  TAG_CONV3, TAG_REDUCE, TAG_COPY, TAG_LAYERNORM, TAG_FFN_TAIL, TAG_ATTN_FRONT, TAG_FFN_FUSED, TAG_ATTN_FUSED, TAG_COUNT
};
struct ProfScope {
  ProfScope(int tag, double bytes, double flops, cudaStream_t s);
  ~ProfScope();
  int idx; cudaStream_t stream;
};

// ---------------------------------------------------------------------------------------------
// generic contraction (1x1 conv, 3x3 conv as implicit GEMM) parameter block
// rows = pixels of image b (row = b*H*W + y*W + x), columns = output channels
// ---------------------------------------------------------------------------------------------
enum AMode { A_PLAIN = 0, A_IM2COL_NHWC = 1, A_IM2COL_NCHW = 2 };
enum OMode { O_NHWC = 0, O_UNSHUFFLE = 1, O_SHUFFLE = 2, O_NCHW = 3 };
enum LnMode { LN_NONE = 0, LN_BIASFREE = 1, LN_WITHBIAS = 2 };

struct GemmParams {
  // A operand: source 1 (+ optional concat source 2, plain mode only)
  const float* a1; int lda1; int k1;      // plain: channels of source 1; im2col: Cin
  const float* a2; int lda2; int k2;
  int a_mode;
  int B, H, W;                            // spatial extent of the rows (conv input == conv output extent)
  // weights, row-major [N][Kp]; w_bstride != 0 selects per-image weights
  const float* w; long long w_bstride;
  int N, K, Kp;
  const float* bias;                      // [N] or nullptr
  // LayerNorm prologue over the k1 channels of source 1
  int ln_mode; const float* ln_w; const float* ln_b;
  // epilogue: y = (r ? r : 0) + acc_sign * act(acc + bias)
  int relu;
  const float* r; int ldr; float acc_sign;
  float* y; int ldy; int o_mode;
  int tag;                                // KernelTag for the profiler
};

struct DwParams {
  const float* in; int ldi;
  float* out; int ldo;
  const float* w;                         // [9][Cw] tap-major
  const float* bias;                      // [Cw] or nullptr
  int Cw;                                 // channel pitch of w / bias
  int B, H, W;
  int C;                                  // channels produced
  int gate;                               // 1: out[c] = gelu(dw(in[c])) * dw(in[c + gate_off])
  int gate_off;
  int tag;
  int in_half, out_half;                  // element types of in / out (0 = fp32, 1 = fp16); in/out are then __half*
  int round_tf32;                         // fp32 output rounded to tf32 (it is read by a tensor-core kernel without a pass
                                          // through registers, and the hardware would otherwise truncate)
};

struct GramParams {
  const float* qkv; int ld;               // depthwise-convolved qkv [B*HW, 3C]
  int B, HW, C, heads, nparts;
  float* s_part;                          // [B][heads][nparts][ch][ch]
  float* n_part;                          // [B][heads][nparts][2][ch]  (sum q^2, sum k^2)
  int in_half;                            // qkv element type: 0 fp32, 1 fp16 (qkv is then a __half*)
};

struct FoldParams {
  const float* s_part; const float* n_part;
  int B, C, heads, nparts;
  const float* temperature;               // [heads]
  const float* w_proj;                    // [C][C] row-major (project_out)
  float* w_eff; long long w_eff_bstride;  // [B][C][C] row-major (fmt 0), [B][C/4][C][4] tf32 (fmt 1), [B][C/8][C][8] fp16 (fmt 2),
                                          // SWIZZLE_128B images (fmt 3 tf32 / 4 fp16, K padded: see PackMat::fmt);
                                          // stride in elements of the output type
  int fmt;
};

// launchers (simt_kernels.cu)
int launch_gemm_simt(const GemmParams& p, cudaStream_t s);
int launch_dwconv(const DwParams& p, cudaStream_t s);       // dwconv.cu: rolling-window kernel
int launch_dwconv_ref(const DwParams& p, cudaStream_t s);   // simt_kernels.cu: one thread per output vector (reference)
int launch_gram(const GramParams& p, cudaStream_t s);         // gram.cu: mma.sync tensor-core kernel
int launch_gram_ref(const GramParams& p, cudaStream_t s);     // simt_kernels.cu: CUDA-core reference
int launch_fold(const FoldParams& p, cudaStream_t s);
// 3x3 conv with 1..4 output channels, channels-last fp32 in, NCHW out: y = r + sign * (conv + bias)
int launch_conv3x3_small(const float* in, int ld, int cin, const float* w, int kp, const float* bias, int cout, int B,
                         int H, int W, const float* r, float sign, float* y, cudaStream_t s);
// first 3x3 conv of a network: NCHW image with 1 / 3 / 6 channels -> channels-last rows (+ bias, ReLU); w row-major [cout][kp], k = tap*cin + c
bool conv3x3_first_supported(int cin, int cout);
int launch_conv3x3_first(const float* x_nchw, int cin, const float* w, int kp, const float* bias, int relu, int cout, int B,
                         int H, int W, float* y, int ldy, cudaStream_t s);
// standalone channel LayerNorm (levels whose C does not fit the contraction's register-resident prologue)
int launch_layernorm(const float* x, int ldx, void* y, int ldy, int y_half, long long rows, int C, int ln_mode,
                     const float* w, const float* b, cudaStream_t s);
int launch_copy_channels(const float* src, int lds, float* dst, int ldd, long long rows, int C, cudaStream_t s);
int launch_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t s);
int launch_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t s);

// device-side tiling / blending of the inference harness (tiling.cu)
int launch_tile_gather(const void* img, int dtype, float divisor, int H, int W, int C, const int* tile_xy, int T, int th,
                       int tw, int TH, int TW, const double* noise, float* out, cudaStream_t s);
int launch_tile_blend(const float* pred, const int* tile_xy, int T, int th, int tw, int TH, int TW, const float* window,
                      int win_ld, int H, int W, int C, void* out, int dtype, float scale, float lo, float hi,
                      cudaStream_t s);

// PSNR / SSIM of an HWC image pair (metrics.cu): out = {psnr, ssim, mse} as float64
size_t image_metrics_workspace_bytes(int H, int W, int C);
int launch_image_metrics(const void* pred, const void* target, int dtype, int H, int W, int C, double data_range,
                         double* out, void* ws, size_t ws_bytes, cudaStream_t s);

// weight packing (pack.cu)
struct PackMat {
  const float* src; float* dst;
  int kind;            // 0: src[n][k] (1x1 / linear), 1: src[n][cin][3][3] -> k = tap*cin + c,
                       // 2: src[n][cin][3][3] -> k = tap*(k_dst/9) + c (taps padded to whole operand boxes)
  int cin;             // kind 1 only
  int n_src_half, n_dst_half, n_halves;   // row split-pad mapping (GDFN hidden padding)
  int k_src, k_dst;                        // logical / padded reduction length (kind 0: zero-pad; kind 1: k_src = 9*cin)
  const float* row_scale;                  // optional per-source-row scale (BatchNorm folding)
  int fmt;                                 // 0: dst[n][k] fp32;  1: dst[k/4][n][k%4] rounded to tf32;  2: dst[k/8][n][k%8] fp16
                                           // (1, 2: tcgen05 no-swizzle operand layouts); 3 / 4: tf32 / fp16 in the
                                           // SWIZZLE_128B image of tma_gemm.cu (k_dst a multiple of 32 / 64)
};
int launch_pack_mat(const PackMat& p, cudaStream_t s);
// dst[t][map(c)] = src[c][t]  (depthwise 3x3 [C][1][3][3] -> [9][Cdst])
int launch_pack_dw(const float* src, float* dst, int c_src_half, int c_dst_half, int n_halves, cudaStream_t s);
// dst[map(i)] = src[i]*scale[i] + shift[i] (scale/shift optional)
int launch_pack_vec(const float* src, float* dst, int src_half, int dst_half, int n_halves,
                    const float* scale, const float* shift, cudaStream_t s);
// BatchNorm eval folding: scale = g/sqrt(var+eps), shift = b - mean*scale
int launch_bn_fold(const float* g, const float* b, const float* mean, const float* var, float eps,
                   float* scale, float* shift, int n, cudaStream_t s);

}  // namespace irb
