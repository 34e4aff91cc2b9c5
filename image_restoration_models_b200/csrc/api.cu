// extern "C" boundary (include/irb200.h).  Plain pointers and sizes only; no torch types.
#include "dncnn.cuh"
#include "restormer.cuh"

namespace irb {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error: %s (%s) at %s:%d [%s]", cudaGetErrorString(e), cudaGetErrorName(e), file,
           line, what);
  g_err = buf;
  if (e == cudaErrorMemoryAllocation) { g_err = std::string("CUDA out of memory: ") + buf; return IR_ERR_OOM; }
  return IR_ERR_CUDA;
}

static int check_mode(int mode) {
  IRB_REQUIRE(mode == IR_MODE_FP32, "mode: only IR_MODE_FP32 is available in this build");
  return IR_OK;
}

}  // namespace irb

using namespace irb;

extern "C" {
#pragma GCC visibility push(default)

int ir_abi_version(void) { return IRB200_ABI_VERSION; }
const char* ir_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------ Restormer
int ir_restormer_param_count(const IrRestormerCfg* cfg) {
  if (!cfg) { set_error("invalid argument: null cfg"); return -1; }
  RestormerPlan pl;
  if (build_restormer_plan(pl, *cfg) != IR_OK) return -1;
  return pl.n_params;
}

long long ir_restormer_param_numel(const IrRestormerCfg* cfg, int index) {
  if (!cfg) { set_error("invalid argument: null cfg"); return -1; }
  RestormerPlan pl;
  if (build_restormer_plan(pl, *cfg) != IR_OK) return -1;
  for (const PackOp& op : pl.ops)
    if (op.param == index) return pack_op_src_numel(op);
  set_error("invalid argument: parameter index out of range");
  return -1;
}

size_t ir_restormer_packed_bytes(const IrRestormerCfg* cfg, int mode) {
  if (!cfg || check_mode(mode) != IR_OK) return 0;
  RestormerPlan pl;
  if (build_restormer_plan(pl, *cfg) != IR_OK) return 0;
  return (size_t)pl.packed_floats * sizeof(float);
}

int ir_restormer_pack_weights(const IrRestormerCfg* cfg, const float* const* h_params, int n_params, void* packed,
                              size_t packed_bytes, int mode, void* stream) {
  IRB_REQUIRE(cfg && h_params && packed, "pack: null argument");
  IRB_TRY(check_mode(mode));
  RestormerPlan pl;
  IRB_TRY(build_restormer_plan(pl, *cfg));
  IRB_REQUIRE(n_params == pl.n_params, "pack: parameter count does not match the configuration's state_dict");
  if (packed_bytes < (size_t)pl.packed_floats * sizeof(float)) { set_error("packed buffer too small"); return IR_ERR_WORKSPACE; }
  IRB_CUDA(cudaMemsetAsync(packed, 0, (size_t)pl.packed_floats * sizeof(float), (cudaStream_t)stream));
  return run_pack_ops(pl.ops, h_params, (float*)packed, (cudaStream_t)stream);
}

size_t ir_restormer_workspace_bytes(const IrRestormerCfg* cfg, int B, int H, int W, int mode) {
  if (!cfg || check_mode(mode) != IR_OK || B <= 0 || H <= 0 || W <= 0) return 0;
  RestormerPlan pl;
  if (build_restormer_plan(pl, *cfg) != IR_OK) return 0;
  return restormer_workspace_bytes(pl, B, H, W);
}

int ir_restormer_forward(const IrRestormerCfg* cfg, const void* packed, const float* x, float* y, int B, int H, int W,
                         void* workspace, size_t workspace_bytes, int mode, void* stream) {
  IRB_REQUIRE(cfg && packed && x && y && workspace, "forward: null argument");
  IRB_TRY(check_mode(mode));
  RestormerPlan pl;
  IRB_TRY(build_restormer_plan(pl, *cfg));
  return restormer_forward(pl, (const float*)packed, x, y, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ir_restormer_launch_count(const IrRestormerCfg* cfg) {
  if (!cfg) return -1;
  RestormerPlan pl;
  if (build_restormer_plan(pl, *cfg) != IR_OK) return -1;
  return restormer_launch_count(pl);
}

// ------------------------------------------------------------------------------------------ DnCNN
int ir_dncnn_param_count(const IrDncnnCfg* cfg) {
  if (!cfg) { set_error("invalid argument: null cfg"); return -1; }
  DncnnPlan pl;
  if (build_dncnn_plan(pl, *cfg) != IR_OK) return -1;
  return pl.n_params;
}

long long ir_dncnn_param_numel(const IrDncnnCfg* cfg, int index) {
  if (!cfg) { set_error("invalid argument: null cfg"); return -1; }
  DncnnPlan pl;
  if (build_dncnn_plan(pl, *cfg) != IR_OK) return -1;
  return dncnn_param_numel(pl, index);
}

size_t ir_dncnn_packed_bytes(const IrDncnnCfg* cfg, int mode) {
  if (!cfg || check_mode(mode) != IR_OK) return 0;
  DncnnPlan pl;
  if (build_dncnn_plan(pl, *cfg) != IR_OK) return 0;
  return (size_t)pl.packed_floats * sizeof(float);
}

int ir_dncnn_pack_weights(const IrDncnnCfg* cfg, const float* const* h_params, int n_params, void* packed,
                          size_t packed_bytes, int mode, void* stream) {
  IRB_REQUIRE(cfg && h_params && packed, "pack: null argument");
  IRB_TRY(check_mode(mode));
  DncnnPlan pl;
  IRB_TRY(build_dncnn_plan(pl, *cfg));
  IRB_REQUIRE(n_params == pl.n_params, "pack: parameter count does not match the configuration's state_dict");
  if (packed_bytes < (size_t)pl.packed_floats * sizeof(float)) { set_error("packed buffer too small"); return IR_ERR_WORKSPACE; }
  IRB_CUDA(cudaMemsetAsync(packed, 0, (size_t)pl.packed_floats * sizeof(float), (cudaStream_t)stream));
  return dncnn_pack(pl, h_params, (float*)packed, (cudaStream_t)stream);
}

size_t ir_dncnn_workspace_bytes(const IrDncnnCfg* cfg, int B, int H, int W, int mode) {
  if (!cfg || check_mode(mode) != IR_OK || B <= 0 || H <= 0 || W <= 0) return 0;
  DncnnPlan pl;
  if (build_dncnn_plan(pl, *cfg) != IR_OK) return 0;
  return dncnn_workspace_bytes(pl, B, H, W);
}

int ir_dncnn_forward(const IrDncnnCfg* cfg, const void* packed, const float* x, float* y, int B, int H, int W,
                     void* workspace, size_t workspace_bytes, int mode, void* stream) {
  IRB_REQUIRE(cfg && packed && x && y && workspace, "forward: null argument");
  IRB_TRY(check_mode(mode));
  DncnnPlan pl;
  IRB_TRY(build_dncnn_plan(pl, *cfg));
  return dncnn_forward(pl, (const float*)packed, x, y, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ir_dncnn_launch_count(const IrDncnnCfg* cfg) { return cfg ? cfg->nb : -1; }

// ------------------------------------------------------------------------------------------ single block
size_t ir_block_workspace_bytes(int C, int heads, float ffn, int B, int H, int W, int mode) {
  if (check_mode(mode) != IR_OK || B <= 0 || H <= 0 || W <= 0) return 0;
  BlockPlan bp; std::vector<PackOp> ops; long long pf;
  if (build_block_plan(bp, ops, pf, C, heads, ffn, 0, 0) != IR_OK) return 0;
  return block_workspace_bytes(bp, B, H, W);
}

size_t ir_block_packed_bytes(int C, int heads, float ffn, int bias, int ln_with_bias, int mode) {
  if (check_mode(mode) != IR_OK) return 0;
  BlockPlan bp; std::vector<PackOp> ops; long long pf;
  if (build_block_plan(bp, ops, pf, C, heads, ffn, bias, ln_with_bias) != IR_OK) return 0;
  return (size_t)pf * sizeof(float);
}

int ir_block_pack_weights(int C, int heads, float ffn, int bias, int ln_with_bias, const float* const* h_params,
                          int n_params, void* packed, size_t packed_bytes, int mode, void* stream) {
  IRB_REQUIRE(h_params && packed, "pack: null argument");
  IRB_TRY(check_mode(mode));
  BlockPlan bp; std::vector<PackOp> ops; long long pf;
  IRB_TRY(build_block_plan(bp, ops, pf, C, heads, ffn, bias, ln_with_bias));
  IRB_REQUIRE(n_params == block_param_count(bias, ln_with_bias), "pack: wrong parameter count for a TransformerBlock");
  if (packed_bytes < (size_t)pf * sizeof(float)) { set_error("packed buffer too small"); return IR_ERR_WORKSPACE; }
  IRB_CUDA(cudaMemsetAsync(packed, 0, (size_t)pf * sizeof(float), (cudaStream_t)stream));
  return run_pack_ops(ops, h_params, (float*)packed, (cudaStream_t)stream);
}

int ir_block_forward(int C, int heads, float ffn, int bias, int ln_with_bias, const void* packed, float* x_nhwc, int B,
                     int H, int W, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  IRB_REQUIRE(packed && x_nhwc && workspace, "forward: null argument");
  IRB_TRY(check_mode(mode));
  BlockPlan bp; std::vector<PackOp> ops; long long pf;
  IRB_TRY(build_block_plan(bp, ops, pf, C, heads, ffn, bias, ln_with_bias));
  return block_forward(bp, (const float*)packed, x_nhwc, B, H, W, workspace, workspace_bytes, ln_with_bias,
                       (cudaStream_t)stream);
}

int ir_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  IRB_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "layout: bad argument");
  return launch_nchw_to_nhwc(src, dst, B, C, H, W, (cudaStream_t)stream);
}
int ir_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  IRB_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "layout: bad argument");
  return launch_nhwc_to_nchw(src, dst, B, C, H, W, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
