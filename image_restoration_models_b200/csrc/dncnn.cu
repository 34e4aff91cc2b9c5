// Host-side plan + launch sequence of the DnCNN forward.
// Restates /root/reference/src/dncnn/models/network_dncnn.py:63-71: nb 3x3 convs (bias) with ReLU between them
// (eval-mode BatchNorm, basicblock.py:69, folded into the preceding conv at pack time) and `x - model(x)`.
#include "dncnn.cuh"
#include "restormer.cuh"
#include "tc_gemm.cuh"

namespace irb {

static inline int round_up4(int v) { return (v + 3) / 4 * 4; }

int build_dncnn_plan(DncnnPlan& pl, const IrDncnnCfg& c, int engine) {
  IRB_REQUIRE(c.in_nc > 0 && c.out_nc > 0 && c.nc > 0 && c.nc % 4 == 0, "dncnn: nc must be a positive multiple of 4");
  IRB_REQUIRE(c.nb >= 2, "dncnn: nb must be >= 2");
  IRB_REQUIRE(c.in_nc == c.out_nc, "dncnn: in_nc must equal out_nc (x - model(x), network_dncnn.py:71)");
  pl.cfg = c;
  pl.half = engine == ENGINE_TC_HALF;
  pl.layers.clear();
  long long off = 0;
  auto alloc = [&](long long n) { const long long o = off; off += (n + 63) / 64 * 64; return o; };
  int pidx = 0;
  for (int l = 0; l < c.nb; ++l) {
    DncnnLayer L{};
    L.cin = l == 0 ? c.in_nc : c.nc;
    L.cout = l == c.nb - 1 ? c.out_nc : c.nc;
    L.k = 9 * L.cin; L.kp = round_up4(L.k);
    // the nc -> nc layers run on the TMA-fed implicit GEMM when their shape allows (taps padded to whole operand boxes)
    L.tma = engine != ENGINE_SIMT && l > 0 && l < c.nb - 1 && L.cout % 16 == 0 && tma_conv3_supported(L.cin, L.cout, pl.half);
    if (L.tma) L.kp = 9 * tma_conv3_kpt(L.cin, pl.half);
    L.w = alloc((long long)L.cout * L.kp);
    L.b = alloc(L.cout);
    L.p_w = pidx++; L.p_b = pidx++;
    L.p_bn = -1;
    // the nc -> nc layers run as implicit GEMM on the tcgen05 kernel; head (cin = in_nc) and tail (cout = out_nc)
    // have 1..3 channels on one side and stay on the CUDA-core kernel
    L.tc = L.tma || (engine != ENGINE_SIMT && l > 0 && l < c.nb - 1 && tc_conv3_supported(L.cin, L.cout, pl.half));
    if (c.has_bn && l > 0 && l < c.nb - 1) { L.p_bn = pidx; pidx += 5; }  // weight, bias, mean, var, num_batches_tracked
    pl.layers.push_back(L);
  }
  pl.bn_scale = alloc(c.nc);
  pl.bn_shift = alloc(c.nc);
  pl.n_params = pidx;
  pl.packed_floats = off;
  return IR_OK;
}

long long dncnn_param_numel(const DncnnPlan& pl, int index) {
  for (const auto& L : pl.layers) {
    if (index == L.p_w) return (long long)L.cout * L.cin * 9;
    if (index == L.p_b) return L.cout;
    if (L.p_bn >= 0 && index >= L.p_bn && index < L.p_bn + 4) return L.cout;
    if (L.p_bn >= 0 && index == L.p_bn + 4) return 1;
  }
  return -1;
}

int dncnn_pack(const DncnnPlan& pl, const float* const* params, float* packed, cudaStream_t s) {
  for (const auto& L : pl.layers) {
    const float* scale = nullptr; const float* shift = nullptr;
    if (L.p_bn >= 0) {
      // BatchNorm2d(momentum=0.9, eps=1e-4) in eval mode: y = (x-mean)/sqrt(var+eps)*g + b  (basicblock.py:69)
      IRB_TRY(launch_bn_fold(params[L.p_bn], params[L.p_bn + 1], params[L.p_bn + 2], params[L.p_bn + 3], 1e-4f,
                             packed + pl.bn_scale, packed + pl.bn_shift, L.cout, s));
      scale = packed + pl.bn_scale; shift = packed + pl.bn_shift;
    }
    PackMat pm{params[L.p_w], packed + L.w, L.tma ? 2 : 1, L.cin, L.cout, L.cout, 1, L.k, L.kp, scale,
               !L.tc ? 0 : L.tma ? (pl.half ? 4 : 3) : pl.half ? 2 : 1};
    IRB_TRY(launch_pack_mat(pm, s));
    IRB_TRY(launch_pack_vec(params[L.p_b], packed + L.b, L.cout, L.cout, 1, scale, shift, s));
  }
  return IR_OK;
}

size_t dncnn_workspace_bytes(const DncnnPlan& pl, int B, int H, int W) {
  return 2 * align_up((size_t)B * H * W * pl.cfg.nc * sizeof(float), 256);
}

int dncnn_forward(const DncnnPlan& pl, const float* packed, const float* x, float* y, int B, int H, int W,
                  void* workspace, size_t workspace_bytes, cudaStream_t s) {
  IRB_REQUIRE(B > 0 && H > 0 && W > 0, "dncnn: empty input");
  if (dncnn_workspace_bytes(pl, B, H, W) > workspace_bytes) { set_error("workspace too small"); return IR_ERR_WORKSPACE; }
  const size_t half = align_up((size_t)B * H * W * pl.cfg.nc * sizeof(float), 256);
  float* buf[2] = {(float*)workspace, (float*)((char*)workspace + half)};
  const int nb = (int)pl.layers.size();
  for (int l = 0; l < nb; ++l) {
    const DncnnLayer& L = pl.layers[l];
    if (L.tma && W >= 96 && conv3_row_supported(L.cin, L.cout, pl.half)) {
      IRB_TRY(launch_conv3_row(buf[(l + 1) & 1], pl.cfg.nc, L.cin, packed + L.w, packed + L.b, 1, L.cout, L.cout, B, H, W,
                               buf[l & 1], pl.cfg.nc, O_NHWC, pl.half, s));
      continue;
    }
    if (L.tma) {
      IRB_TRY(launch_conv3_tma(buf[(l + 1) & 1], pl.cfg.nc, L.cin, packed + L.w, packed + L.b, 1, L.cout, L.cout, B, H, W,
                               buf[l & 1], pl.cfg.nc, O_NHWC, pl.half, s));
      continue;
    }
    if (L.tc) {
      IRB_TRY(run_conv3_tc(buf[(l + 1) & 1], pl.cfg.nc, L.cin, packed + L.w, packed + L.b, L.cout, L.cout, B, H, W, buf[l & 1],
                           pl.cfg.nc, O_NHWC, 1, pl.half, s));
      continue;
    }
    if (l == nb - 1 && l > 0 && L.cout <= 4 && L.cin % 4 == 0 && pl.cfg.nc == L.cin) {
      // tail conv: y = x - conv(n)  (network_dncnn.py:70-71)
      IRB_TRY(launch_conv3x3_small(buf[(l + 1) & 1], pl.cfg.nc, L.cin, packed + L.w, L.kp, packed + L.b, L.cout, B, H, W,
                                   x, -1.f, y, s));
      continue;
    }
    if (l == 0 && l < nb - 1 && conv3x3_first_supported(L.cin, L.cout)) {
      // head conv + bias + ReLU straight from the NCHW image (network_dncnn.py:63)
      IRB_TRY(launch_conv3x3_first(x, L.cin, packed + L.w, L.kp, packed + L.b, 1, L.cout, B, H, W, buf[l & 1], pl.cfg.nc, s));
      continue;
    }
    GemmParams g{};
    g.B = B; g.H = H; g.W = W;
    g.k1 = L.cin;
    if (l == 0) { g.a1 = x; g.lda1 = 0; g.a_mode = A_IM2COL_NCHW; }
    else { g.a1 = buf[(l + 1) & 1]; g.lda1 = pl.cfg.nc; g.a_mode = A_IM2COL_NHWC; }
    g.w = packed + L.w; g.N = L.cout; g.K = L.k; g.Kp = L.kp; g.bias = packed + L.b;
    g.ln_mode = LN_NONE;
    if (l == nb - 1) {            // tail conv, no activation; y = x - n  (network_dncnn.py:70-71)
      g.relu = 0; g.acc_sign = -1.f; g.r = x; g.y = y; g.o_mode = O_NCHW;
    } else {
      g.relu = 1; g.acc_sign = 1.f; g.r = nullptr; g.y = buf[l & 1]; g.ldy = pl.cfg.nc; g.o_mode = O_NHWC;
    }
    g.tag = TAG_CONV3;
    IRB_TRY(launch_gemm_simt(g, s));
  }
  return IR_OK;
}

}  // namespace irb
