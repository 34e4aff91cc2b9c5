// Host-side plan + launch sequence of the Restormer forward.
// Restates the control flow of /root/reference/src/restormer/restormer.py:245-284 (Restormer.forward)
// and :146-150 (TransformerBlock.forward) as a fixed sequence of kernel launches on one stream.
// All activations are channels-last fp32 ([pixels][C]); the NCHW<->channels-last conversion is folded
// into the first (patch_embed) and last (output) 3x3 convs.
#include "restormer.cuh"
#include "tc_gemm.cuh"
#include "ffn_tail.cuh"
#include "attn_front.cuh"
#include "ffn_fused.cuh"
#include "attn_fused.cuh"

#include <algorithm>
#include <cstdlib>

namespace irb {

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// -----------------------------------------------------------------------------------------------
// plan
// -----------------------------------------------------------------------------------------------
bool tc_gemm_supported(int K, int N, bool half) {
  TcGemmParams t{};
  t.K = K; t.N = N; t.k1 = K; t.k2 = 0; t.ln_mode = LN_NONE; t.a_pad = 1;
  t.a_half = half; t.op_half = half; t.y_half = 0;
  return tc_gemm_configure(t) != 0;
}

bool tc_conv3_supported(int cin, int cout, bool half) {
  TcGemmParams t{};
  t.K = 9 * cin; t.N = cout; t.k1 = cin; t.k2 = 0; t.ln_mode = LN_NONE; t.a_pad = 1; t.a_mode = 1;
  t.a_half = 0; t.op_half = half; t.y_half = 0;
  return tc_gemm_configure(t) != 0;
}

int run_conv3_tc(const float* in, int ld_in, int cin, const float* w_packed, const float* bias, int cout, int cout_valid,
                 int B, int H, int W, float* out, int ld_out, int o_mode, int relu, bool half, cudaStream_t s) {
  TcGemmParams t{};
  t.a1 = in; t.lda1 = ld_in; t.k1 = cin; t.k2 = 0; t.a_mode = 1; t.H = H; t.W = W;
  t.B = B; t.HW = H * W;
  t.w = w_packed; t.N = cout; t.K = 9 * cin; t.bias = bias; t.ln_mode = LN_NONE;
  t.y = out; t.ldy = ld_out; t.o_mode = o_mode; t.relu = relu; t.acc_sign = 1.f; t.n_valid = cout_valid;
  t.tag = TAG_CONV3; t.a_pad = 1; t.a_half = 0; t.op_half = half; t.y_half = 0;
  return launch_gemm_tc(t, s);
}

int engine_of_mode(int mode) {
  return mode == IR_MODE_FP32_SIMT ? ENGINE_SIMT : mode == IR_MODE_HALF ? ENGINE_TC_HALF
       : mode == IR_MODE_FP32_STRICT ? ENGINE_TC_STRICT : ENGINE_TC;
}

struct Builder {
  std::vector<PackOp>& ops;
  long long off = 0;       // floats
  int pidx = 0;            // running state_dict index
  int engine = ENGINE_TC;
  Builder(std::vector<PackOp>& o, int e) : ops(o), engine(e) {}
  bool half() const { return engine == ENGINE_TC_HALF; }
  bool tc(int K, int N) const { return engine != ENGINE_SIMT && tc_gemm_supported(K, N, half()); }
  bool tma(int K, int N, bool ln, bool has_r, bool y_half) const {
    return engine != ENGINE_SIMT && tma_gemm_shape_supported(K, N, half(), ln, has_r, y_half);
  }
  int fmt(bool tc_layer, bool tma_layer = false) const { return !tc_layer ? 0 : tma_layer ? (half() ? 4 : 3) : half() ? 2 : 1; }
  long long alloc(long long n) { const long long o = off; off += (n + 63) / 64 * 64; return o; }
};

static void plan_block_engine(Builder& bl, BlockPlan& bp, int C, int heads, double ffn, int bias, int ln_bias);

// IR_MODE_FP32 at the two low-resolution levels (C > 128: 1/16 and 1/64 of the pixels, 13 % of the bytes, but 28 % of the
// step as fp32 tensors): the block runs the 16-bit plan -- qkv, v, the GDFN hidden / gated tensors and the LayerNorm outputs
// are fp16 (tf32's mantissa), the contractions take fp16 operands; the residual stream, statistics, softmax, GELU and every
// accumulator stay fp32 as in every mode.  The high-resolution levels already keep those tensors on chip.  The pack-time
// range guard covers these tensors (restormer.py _block_fp16_bound); IR_MODE_FP32_STRICT keeps fp32 everywhere.
static void plan_block(Builder& bl, BlockPlan& bp, int C, int heads, double ffn, int bias, int ln_bias) {
  static const bool no_wide16 = getenv("IRB_NO_WIDE16") != nullptr;           // A/B switch for benchmarks
  const bool wide16 = !no_wide16 && bl.engine == ENGINE_TC && C > 128 && !bias;
  if (wide16) bl.engine = ENGINE_TC_HALF;
  plan_block_engine(bl, bp, C, heads, ffn, bias, ln_bias);
  if (wide16) bl.engine = ENGINE_TC;
  bp.wide16 = wide16;
}

static void plan_block_engine(Builder& bl, BlockPlan& bp, int C, int heads, double ffn, int bias, int ln_bias) {
  bp.C = C; bp.heads = heads;
  bp.h = (int)(C * ffn);           // int(dim*ffn_expansion_factor) in double precision, like Python (restormer.py:80)
  // python: int(48*2.66)=127, int(96*2.66)=255, int(192*2.66)=510, int(384*2.66)=1021
  bp.hp = round_up(bp.h, 16);
  auto vec = [&](long long& dst, int n) {
    dst = bl.alloc(n);
    bl.ops.push_back(PackOp{PackOp::VEC, bl.pidx++, dst, n, n, 1, 0, 0, 0, 0});
  };
  auto vec_split = [&](long long& dst, int src_half, int dst_half, int halves) {
    dst = bl.alloc((long long)dst_half * halves);
    bl.ops.push_back(PackOp{PackOp::VEC, bl.pidx++, dst, src_half, dst_half, halves, 0, 0, 0, 0});
  };
  // f16_image: the fp16 SWIZZLE_128B operand image (fmt 4) whatever the mode (the fused GDFN kernel's operands)
  auto mat = [&](long long& dst, int n_src_half, int n_dst_half, int halves, int k_src, int k_dst, bool tc, bool tma = false,
                 bool f16_image = false) {
    if (tma || f16_image) k_dst = tma_gemm_kpad(k_dst, bl.half() || f16_image);
    dst = bl.alloc((long long)n_dst_half * halves * k_dst);
    bl.ops.push_back(PackOp{PackOp::MAT1, bl.pidx++, dst, n_src_half, n_dst_half, halves, k_src, k_dst, 0,
                            f16_image ? 4 : bl.fmt(tc, tma)});
  };
  auto dw = [&](long long& dst, int src_half, int dst_half, int halves) {
    dst = bl.alloc(9LL * dst_half * halves);
    bl.ops.push_back(PackOp{PackOp::DW, bl.pidx++, dst, src_half, dst_half, halves, 0, 0, 0, 0});
  };
  bp.ln1_b = bp.qkv_b = bp.qkvdw_b = bp.proj_b = bp.ln2_b = bp.pin_b = bp.ffdw_b = bp.pout_b = -1;
  bp.ref_kernels = bl.engine == ENGINE_SIMT;
  bp.half = bl.half();
  bp.tc_qkv = bl.tc(C, 3 * C);
  bp.tc_attn = bl.tc(C, C);
  bp.tc_pin = bl.tc(C, 2 * bp.hp);
  bp.tc_pout = bl.tc(bp.hp, C);
  // the TMA-fed kernel takes the layers whose A tile is shared by at most two N-chunks (the high-resolution levels)
  // (LayerNorm is fused into the contraction up to C = 128; wider levels normalise into a scratch tensor first)
  const bool ln_fused = C <= 128;
  bp.tma_qkv = bp.tc_qkv && bl.tma(C, 3 * C, ln_fused, false, bl.half());
  bp.tma_attn = bp.tc_attn && bl.tma(C, C, false, true, false);
  bp.tma_pin = bp.tc_pin && bl.tma(C, 2 * bp.hp, ln_fused, false, bl.half());
  bp.tma_pout = bp.tc_pout && bl.tma(bp.hp, C, false, true, false);
  // MDTA front / GDFN tail in one kernel each (no biases: every shipped configuration has bias=False)
  bp.fuse_front = bl.engine != ENGINE_SIMT && !bias && bp.tc_attn && attn_front_supported(C, heads, bl.half());
  // fp32 mode: v is only ever a tensor-core operand (rounded to a 10-bit mantissa either way), so the fused front stores it
  // as fp16 and the attention-output contraction takes fp16 operands (v and the per-image folded matrix)
  static const bool no_vhalf = getenv("IRB_NO_V_HALF") != nullptr;         // A/B switch for benchmarks
  bp.v_half = !no_vhalf && bl.engine == ENGINE_TC && bp.fuse_front && tma_gemm_shape_supported(C, C, true, false, true, false);
  if (bp.v_half) bp.tma_attn = true;
  bp.kp_attn = bp.tma_attn ? tma_gemm_kpad(C, bl.half() || bp.v_half) : C;
  bp.fuse_tail = bl.engine != ENGINE_SIMT && !bias && ffn_tail_supported(C, bp.hp, bl.half());
  // the whole GDFN in one kernel (ffn_fused.cu).  Its operands are fp16 images in BOTH tensor-core modes: the hidden
  // tensor lives only in shared memory, where fp32 storage does not fit; fp16 carries the same 10-bit mantissa as the
  // tf32 operands it replaces, and accumulation stays fp32.
  static const bool no_fused = getenv("IRB_NO_FFN_FUSED") != nullptr;      // A/B switch for benchmarks
  bp.fuse_ffn = !no_fused && bl.engine != ENGINE_SIMT && bl.engine != ENGINE_TC_STRICT && !bias && ffn_fused_supported(C, bp.hp);
  if (bp.fuse_ffn) bp.fuse_tail = false;
  static const bool no_k4xn = getenv("IRB_NO_K4_XN") != nullptr;           // A/B switch for benchmarks
  bp.k4_xn = !no_k4xn && bp.fuse_ffn && bp.tma_attn && !bias && tma_gemm_xn_supported(C, bl.half() || bp.v_half);
  // the whole MDTA front in one kernel (attn_fused.cu): its v output and the attention-output contraction's operands are
  // fp16 (as with v_half), its qkv weights the fp16 operand image padded to whole 32-channel units
  static const bool no_attn_fused = getenv("IRB_NO_ATTN_FUSED") != nullptr;   // A/B switch for benchmarks
  bp.fuse_attn = !no_attn_fused && bp.fuse_front && !bias && attn_fused_supported(C, heads) &&
                 (bl.engine == ENGINE_TC ? bp.v_half : bl.half()) && bp.tma_attn;
  vec(bp.ln1_w, C);
  if (ln_bias) vec(bp.ln1_b, C);
  vec(bp.temp, heads);
  if (bp.fuse_attn) {
    const int np = attn_fused_wrows(C);
    mat(bp.qkv_w, 3 * C, np, 1, C, C, true, false, true);
  } else {
    mat(bp.qkv_w, 3 * C, 3 * C, 1, C, C, bp.tc_qkv, bp.tma_qkv);
  }
  if (bias) vec(bp.qkv_b, 3 * C);
  if (bp.fuse_front) {
    const int cpad = round_up(3 * C, 32);
    bp.qkvdw_w = bl.alloc(9LL * cpad);
    bl.ops.push_back(PackOp{PackOp::DWC, bl.pidx++, bp.qkvdw_w, 3 * C, cpad, 1, 0, 32, 0, 0});
  } else {
    dw(bp.qkvdw_w, 3 * C, 3 * C, 1);
  }
  if (bias) vec(bp.qkvdw_b, 3 * C);
  mat(bp.proj_w, C, C, 1, C, C, false);   // consumed by the softmax/fold kernel, never a GEMM operand
  if (bias) vec(bp.proj_b, C);
  vec(bp.ln2_w, C);
  if (ln_bias) vec(bp.ln2_b, C);
  mat(bp.pin_w, bp.h, bp.hp, 2, C, C, bp.tc_pin, bp.tma_pin, bp.fuse_ffn);
  if (bias) vec_split(bp.pin_b, bp.h, bp.hp, 2);
  if (bp.fuse_tail || bp.fuse_ffn) {
    bp.ffdw_w = bl.alloc(18LL * bp.hp);
    bl.ops.push_back(PackOp{PackOp::DWC, bl.pidx++, bp.ffdw_w, bp.h, bp.hp, 2, 0, bp.fuse_ffn ? 64 : ffn_tail_kc(bl.half()), 0, 0});
  } else {
    dw(bp.ffdw_w, bp.h, bp.hp, 2);
  }
  if (bias) vec_split(bp.ffdw_b, bp.h, bp.hp, 2);
  mat(bp.pout_w, C, C, 1, bp.h, bp.hp, bp.tc_pout || bp.fuse_tail, bp.tma_pout || bp.fuse_tail, bp.fuse_ffn);
  if (bias) vec(bp.pout_b, C);
}

int block_param_count(int bias, int ln_bias) { return 9 + (ln_bias ? 2 : 0) + (bias ? 6 : 0); }

int build_block_plan(BlockPlan& bp, std::vector<PackOp>& ops, long long& packed_floats, int C, int heads, double ffn,
                     int bias, int ln_bias, int engine) {
  IRB_REQUIRE(C > 0 && heads > 0 && C % heads == 0, "block: C must be divisible by heads");
  IRB_REQUIRE((C / heads) % 16 == 0 && C / heads <= 128, "block: head dim must be a multiple of 16 and <= 128");
  IRB_REQUIRE(ffn > 0.0, "block: ffn_expansion_factor must be positive");
  Builder bl(ops, engine);
  plan_block(bl, bp, C, heads, ffn, bias, ln_bias);
  packed_floats = bl.off;
  IRB_REQUIRE(!bp.half || bp.wide16 || (bp.tc_qkv && bp.tc_attn && bp.tc_pin && bp.tc_pout),
              "half mode: channel widths must be multiples of 16 (tensor-core operand granularity)");
  return IR_OK;
}

static void plan_conv3(Builder& bl, ConvPlan& cp, int cout, int cin, int bias, bool scatter = false) {
  cp.cout = cout; cp.cin = cin; cp.k = 9 * cin; cp.kp = round_up(9 * cin, 4);
  // the tensor-core kernel wants N % 16 == 0: pad with zero rows (only used with the shuffle-scatter epilogue,
  // which skips the padding channels), e.g. down1_2: 48 -> 24 channels
  cp.cout_p = scatter ? round_up(cout, 16) : cout;
  cp.tc = bl.engine != ENGINE_SIMT && bias == 0 && tc_conv3_supported(cin, cp.cout_p, bl.half());
  // the shuffle-scatter convolutions run on the TMA-fed implicit GEMM (taps padded to whole operand boxes)
  cp.tma = cp.tc && scatter && tma_conv3_supported(cin, cp.cout_p, bl.half());
  if (cp.tma) cp.kp = 9 * tma_conv3_kpt(cin, bl.half());
  if (!cp.tc) cp.cout_p = cout;
  cp.w = bl.alloc((long long)cp.cout_p * cp.kp);
  bl.ops.push_back(PackOp{PackOp::MAT3, bl.pidx++, cp.w, cout, cp.cout_p, 1, 9 * cin, cp.kp, cin, bl.fmt(cp.tc, cp.tma)});
  cp.b = -1;
  if (bias) {
    cp.b = bl.alloc(cout);
    bl.ops.push_back(PackOp{PackOp::VEC, bl.pidx++, cp.b, cout, cout, 1, 0, 0, 0, 0});
  }
}
static void plan_conv1(Builder& bl, ConvPlan& cp, int cout, int cin, int bias) {
  cp.cout = cout; cp.cin = cin; cp.k = cin; cp.kp = cin;
  cp.w = bl.alloc((long long)cout * cin);
  cp.tc = bl.tc(cin, cout);
  bl.ops.push_back(PackOp{PackOp::MAT1, bl.pidx++, cp.w, cout, cout, 1, cin, cin, 0, bl.fmt(cp.tc)});
  cp.b = -1;
  if (bias) {
    cp.b = bl.alloc(cout);
    bl.ops.push_back(PackOp{PackOp::VEC, bl.pidx++, cp.b, cout, cout, 1, 0, 0, 0, 0});
  }
}

int build_restormer_plan(RestormerPlan& pl, const IrRestormerCfg& c, int engine) {
  IRB_REQUIRE(c.inp_channels > 0 && c.out_channels > 0, "restormer: channel counts must be positive");
  IRB_REQUIRE(c.dim > 0 && c.dim % 8 == 0, "restormer: dim must be a multiple of 8");
  for (int i = 0; i < 4; ++i) {
    IRB_REQUIRE(c.num_blocks[i] >= 0 && c.heads[i] > 0, "restormer: bad num_blocks / heads");
    const int C = c.dim << i;
    IRB_REQUIRE(C % c.heads[i] == 0 && (C / c.heads[i]) % 16 == 0 && C / c.heads[i] <= 128,
                "restormer: head dim (C/heads) must be a multiple of 16 and <= 128");
  }
  IRB_REQUIRE((2 * c.dim) % c.heads[0] == 0 && (2 * c.dim / c.heads[0]) % 16 == 0 && 2 * c.dim / c.heads[0] <= 128,
              "restormer: level-1 decoder head dim must be a multiple of 16 and <= 128");
  IRB_REQUIRE(c.num_refinement_blocks >= 0 && c.ffn_expansion_factor > 0.0, "restormer: bad refinement / ffn factor");
  IRB_REQUIRE(c.dual_pixel_task || c.inp_channels == c.out_channels,
              "restormer: inp_channels must equal out_channels unless dual_pixel_task (residual add, restormer.py:281)");
  pl.cfg = c;
  pl.half = engine == ENGINE_TC_HALF;
  pl.ops.clear();
  Builder bl(pl.ops, engine);
  const int d = c.dim, bias = c.bias, lnb = c.layernorm_with_bias;
  auto stage = [&](std::vector<BlockPlan>& v, int C, int heads, int n) {
    v.resize(n);
    for (int i = 0; i < n; ++i) plan_block(bl, v[i], C, heads, c.ffn_expansion_factor, bias, lnb);
  };
  plan_conv3(bl, pl.patch_embed, d, c.inp_channels, 0);
  stage(pl.enc[0], d, c.heads[0], c.num_blocks[0]);
  plan_conv3(bl, pl.down[0], d / 2, d, 0, true);
  stage(pl.enc[1], 2 * d, c.heads[1], c.num_blocks[1]);
  plan_conv3(bl, pl.down[1], d, 2 * d, 0, true);
  stage(pl.enc[2], 4 * d, c.heads[2], c.num_blocks[2]);
  plan_conv3(bl, pl.down[2], 2 * d, 4 * d, 0, true);
  stage(pl.enc[3], 8 * d, c.heads[3], c.num_blocks[3]);
  plan_conv3(bl, pl.up[2], 16 * d, 8 * d, 0, true);
  plan_conv1(bl, pl.reduce[2], 4 * d, 8 * d, bias);
  stage(pl.dec[2], 4 * d, c.heads[2], c.num_blocks[2]);
  plan_conv3(bl, pl.up[1], 8 * d, 4 * d, 0, true);
  plan_conv1(bl, pl.reduce[1], 2 * d, 4 * d, bias);
  stage(pl.dec[1], 2 * d, c.heads[1], c.num_blocks[1]);
  plan_conv3(bl, pl.up[0], 4 * d, 2 * d, 0, true);
  stage(pl.dec[0], 2 * d, c.heads[0], c.num_blocks[0]);
  stage(pl.refine, 2 * d, c.heads[0], c.num_refinement_blocks);
  if (c.dual_pixel_task) plan_conv1(bl, pl.skip, 2 * d, d, bias);
  plan_conv3(bl, pl.output, c.out_channels, 2 * d, bias);
  pl.n_params = bl.pidx;
  pl.packed_floats = bl.off;
  if (engine == ENGINE_TC_HALF) {
    auto ok = [](const std::vector<BlockPlan>& v) {
      for (const auto& bp : v) if (!(bp.tc_qkv && bp.tc_attn && bp.tc_pin && bp.tc_pout)) return false;
      return true;
    };
    bool all = ok(pl.refine);
    for (int l = 0; l < 4; ++l) all = all && ok(pl.enc[l]);
    for (int l = 0; l < 3; ++l) all = all && ok(pl.dec[l]);
    IRB_REQUIRE(all, "half mode: channel widths must be multiples of 16 (tensor-core operand granularity)");
  }
  return IR_OK;
}

long long pack_op_src_numel(const PackOp& op) {
  switch (op.kind) {
    case PackOp::VEC: return (long long)op.a * op.c;
    case PackOp::MAT1: return (long long)op.a * op.c * op.k_src;
    case PackOp::MAT3: return (long long)op.a * op.c * op.k_src;   // a = source rows (cout), b = padded rows
    case PackOp::DW: return (long long)op.a * op.c * 9;
    case PackOp::DWC: return (long long)op.a * op.c * 9;
    default: return op.a;  // DnCNN-specific kinds carry their numel in a
  }
}

int run_pack_ops(const std::vector<PackOp>& ops, const float* const* params, float* packed, cudaStream_t s) {
  for (const PackOp& op : ops) {
    const float* src = params[op.param];
    IRB_REQUIRE(src != nullptr, "pack: null parameter pointer");
    float* dst = packed + op.dst;
    switch (op.kind) {
      case PackOp::VEC:
        IRB_TRY(launch_pack_vec(src, dst, op.a, op.b, op.c, nullptr, nullptr, s));
        break;
      case PackOp::MAT1: {
        PackMat pm{src, dst, 0, 0, op.a, op.b, op.c, op.k_src, op.k_dst, nullptr, op.fmt};
        IRB_TRY(launch_pack_mat(pm, s));
        break;
      }
      case PackOp::MAT3: {
        PackMat pm{src, dst, op.fmt >= 3 ? 2 : 1, op.cin, op.a, op.b, op.c, op.k_src, op.k_dst, nullptr, op.fmt};
        IRB_TRY(launch_pack_mat(pm, s));
        break;
      }
      case PackOp::DW:
        IRB_TRY(launch_pack_dw(src, dst, op.a, op.b, op.c, s));
        break;
      case PackOp::DWC:
        IRB_TRY(launch_pack_dw_chunked(src, dst, op.a, op.b, op.k_dst, op.c, s));
        break;
      default:
        IRB_REQUIRE(false, "pack: unknown op");
    }
  }
  return IR_OK;
}

// -----------------------------------------------------------------------------------------------
// workspace
// -----------------------------------------------------------------------------------------------
static int gram_parts(int B, int heads, int HW) {
  int want = cdiv(2 * 148, B * heads);
  int cap = std::max(1, HW / 64);
  return std::max(1, std::min(want, cap));
}

void block_scratch_need(BlockScratchNeed& n, const BlockPlan& bp, int B, int H, int W) {
  const long long P = (long long)B * H * W;
  const int ch = bp.C / bp.heads;
  // intermediates are fp32 or 16-bit PER BLOCK (the wide levels of the fp32 mode run the 16-bit plan): sizes in floats
  const long long es = bp.half ? 2 : 4;
  auto fl = [&](long long elems) { return (elems * es + 3) / 4; };
  if (bp.fuse_attn) n.v16 = std::max(n.v16, P * bp.C);      // qkv stays on chip; only v (fp16) is written
  else n.qkv = std::max(n.qkv, fl(P * 3 * bp.C));
  if (!bp.fuse_ffn) n.hidden = std::max(n.hidden, fl(P * 2 * bp.hp));        // fused GDFN: the hidden tensor stays on chip
  if (!bp.fuse_ffn && !bp.fuse_tail) n.gated = std::max(n.gated, fl(P * bp.hp));
  const int parts = bp.fuse_attn ? attn_fused_parts(B, H, W)
                  : bp.fuse_front ? attn_front_parts(B, H, W, bp.C, bp.heads) : gram_parts(B, bp.heads, H * W);
  n.s_part = std::max(n.s_part, (long long)B * bp.heads * parts * ch * ch);
  n.n_part = std::max(n.n_part, (long long)B * bp.heads * parts * 2 * ch);
  n.w_eff = std::max(n.w_eff, fl((long long)B * bp.C * bp.kp_attn));
  if (bp.C > 128 || bp.fuse_ffn) n.xhat = std::max(n.xhat, fl(P * bp.C));
  if (bp.fuse_attn) n.xhat2 = std::max(n.xhat2, (P * bp.C + 1) / 2);      // fp16 norm1 output, counted in floats
}

struct Carver {
  char* base; size_t off = 0; size_t cap;
  Carver(void* b, size_t c) : base((char*)b), cap(c) {}
  float* take(long long nfloats) {
    const size_t bytes = align_up((size_t)nfloats * sizeof(float), 256);
    float* p = base ? (float*)(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

void carve_block_scratch(Carver& cv, BlockScratch& bs, const BlockScratchNeed& n) {
  // sizes are in floats already (block_scratch_need)
  bs.qkv = cv.take(n.qkv);
  bs.qkv_dw = cv.take(std::max(n.qkv, (n.v16 + 1) / 2));
  bs.hidden = cv.take(n.hidden);
  bs.gated = cv.take(n.gated);
  bs.s_part = cv.take(n.s_part);
  bs.n_part = cv.take(n.n_part);
  bs.w_eff = cv.take(n.w_eff);
  bs.xhat = cv.take(n.xhat);
  bs.xhat2 = cv.take(n.xhat2);
}

size_t block_workspace_bytes(const BlockPlan& bp, int B, int H, int W) {
  BlockScratchNeed n;
  block_scratch_need(n, bp, B, H, W);
  Carver cv(nullptr, 0);
  BlockScratch bs;
  carve_block_scratch(cv, bs, n);
  return cv.off;
}

static void restormer_carve(const RestormerPlan& pl, int B, int H, int W, Carver& cv, RestormerWs& ws) {
  const int d = pl.cfg.dim;
  BlockScratchNeed need;
  for (int l = 0; l < 4; ++l) {
    const int h = H >> l, w = W >> l;
    for (const auto& bp : pl.enc[l]) block_scratch_need(need, bp, B, h, w);
    if (l < 3) for (const auto& bp : pl.dec[l]) block_scratch_need(need, bp, B, h, w);
  }
  for (const auto& bp : pl.refine) block_scratch_need(need, bp, B, H, W);
  const long long P0 = (long long)B * H * W;
  for (int l = 0; l < 4; ++l) ws.e[l] = cv.take((P0 >> (2 * l)) * (d << l));
  ws.e1_in = pl.cfg.dual_pixel_task ? cv.take(P0 * d) : nullptr;
  ws.d[2] = cv.take((P0 >> 4) * (4 * d));
  ws.d[1] = cv.take((P0 >> 2) * (2 * d));
  ws.d[0] = cv.take(P0 * (2 * d));
  ws.up_tmp = cv.take(std::max((P0 >> 4) * (4 * d), (P0 >> 2) * (2 * d)));
  carve_block_scratch(cv, ws.bs, need);
}

size_t restormer_workspace_bytes(const RestormerPlan& pl, int B, int H, int W) {
  Carver cv(nullptr, 0);
  RestormerWs ws;
  restormer_carve(pl, B, H, W, cv, ws);
  return cv.off;
}

// -----------------------------------------------------------------------------------------------
// 1x1 contraction dispatch: tcgen05 kernel when the layer was packed for it, CUDA-core kernel otherwise
// -----------------------------------------------------------------------------------------------
// Element types in half mode: a_half / y_half say whether the A source / the output are fp16 buffers (the
// GemmParams pointers are then reinterpreted); weights were packed as fp16 operands at plan time.
struct XnOut { void* xn = nullptr; int ld = 0; int ln_mode = 0; const float* w = nullptr; const float* b = nullptr; };

static int run_1x1(const GemmParams& g, bool tc, bool half, bool a_half, bool y_half, void* xhat, cudaStream_t s,
                   bool tma = false, const XnOut* xo = nullptr) {
  if (!tc) return launch_gemm_simt(g, s);
  TcGemmParams t{};
  t.a1 = g.a1; t.lda1 = g.lda1; t.k1 = g.k1; t.a2 = g.a2; t.lda2 = g.lda2; t.k2 = g.k2;
  t.B = g.B; t.HW = g.H * g.W;
  t.w = g.w; t.w_bstride = g.w_bstride; t.N = g.N; t.K = g.K; t.bias = g.bias;
  t.ln_mode = g.ln_mode; t.ln_w = g.ln_w; t.ln_b = g.ln_b;
  t.r = g.r; t.ldr = g.ldr; t.y = g.y; t.ldy = g.ldy; t.tag = g.tag; t.a_pad = 1;
  t.a_half = a_half; t.op_half = half; t.y_half = y_half;
  if (tma) {
    // the layer's weights were packed for the TMA-fed kernel at plan time; the plan only selects shapes it supports
    if (g.ln_mode != LN_NONE && g.K > 128) {
      // fp32 operands reach the kind::tf32 MMA straight from the TMA box: round them here (the tensor core truncates)
      IRB_TRY(launch_layernorm(g.a1, g.lda1, xhat, g.K, half ? 1 : 2, (long long)g.B * g.H * g.W, g.K, g.ln_mode,
                               g.ln_w, g.ln_b, s));
      t.a1 = xhat; t.lda1 = g.K; t.ln_mode = LN_NONE; t.a_half = half;
    }
    if (xo && xo->xn) { t.xn = xo->xn; t.ldxn = xo->ld; t.xn_ln_mode = xo->ln_mode; t.xn_w = xo->w; t.xn_b = xo->b; }
    const int st = launch_gemm_tma(t, s);
    if (st == IR_UNSUPPORTED_SHAPE) { set_error("internal: layer planned for the TMA kernel is not launchable (alignment)"); return IR_ERR_INVALID; }
    return st;
  }
  if (g.ln_mode != LN_NONE) {
    TcGemmParams probe = t;
    if (tc_gemm_configure(probe) == 0) {
      // wide levels: the row does not fit the register-resident prologue -> normalise once into scratch
      IRB_TRY(launch_layernorm(g.a1, g.lda1, xhat, g.K, half ? 1 : 0, (long long)g.B * g.H * g.W, g.K, g.ln_mode,
                               g.ln_w, g.ln_b, s));
      t.a1 = xhat; t.lda1 = g.K; t.ln_mode = LN_NONE; t.a_half = half;
    }
  }
  return launch_gemm_tc(t, s);
}

// -----------------------------------------------------------------------------------------------
// one TransformerBlock: x_out = block(x_in)   (x_in may equal x_out)
// -----------------------------------------------------------------------------------------------
bool block_chains_norm1(const BlockPlan& bp, const BlockPlan& next) {
  // OFF by default.  Measured (profiles/r02_norm1_chain_ab.json): the LayerNorm passes it removes cost 2.3 ms per step, the
  // heavier epilogue -- one warp per lane quarter holds the pixel's whole row, on the critical path of the lock-step
  // depthwise phases -- costs the GDFN 3.6 ms.  IRB_NORM1_CHAIN=1 turns it on (parity-tested either way).
  static const bool on = getenv("IRB_NORM1_CHAIN") != nullptr;
  return on && bp.fuse_ffn && next.fuse_attn && bp.C == next.C;
}

int run_block(const BlockPlan& bp, const float* packed, const float* x_in, float* x_out, int B, int H, int W,
              const BlockScratch& bs, int ln_with_bias, cudaStream_t s, const BlockPlan* next, bool* xn1_ready) {
  const int C = bp.C, hp = bp.hp;
  const int ln = ln_with_bias ? LN_WITHBIAS : LN_BIASFREE;
  const bool hf = bp.half;                 // fp16 intermediates: qkv, qkv_dw (v), hidden, gated, W_eff
  const size_t es = hf ? 2 : 4;
  auto P = [&](long long off) -> const float* { return off >= 0 ? packed + off : nullptr; };

  const bool xn1_in = xn1_ready && *xn1_ready;
  if (xn1_ready) *xn1_ready = false;        // consumed below; set again if this block's GDFN emits the next norm1
  GemmParams g{};
  if (!bp.fuse_attn) {
  // (1) norm1 + qkv 1x1   (restormer.py:147 norm1, :114 qkv)
  g.a1 = x_in; g.lda1 = C; g.k1 = C; g.a2 = nullptr; g.lda2 = 0; g.k2 = 0; g.a_mode = A_PLAIN;
  g.B = B; g.H = H; g.W = W;
  g.w = P(bp.qkv_w); g.w_bstride = 0; g.N = 3 * C; g.K = C; g.Kp = C; g.bias = P(bp.qkv_b);
  g.ln_mode = ln; g.ln_w = P(bp.ln1_w); g.ln_b = P(bp.ln1_b);
  g.relu = 0; g.r = nullptr; g.ldr = 0; g.acc_sign = 1.f;
  g.y = (float*)bs.qkv; g.ldy = 3 * C; g.o_mode = O_NHWC; g.tag = TAG_LN_QKV;
  IRB_TRY(run_1x1(g, bp.tc_qkv, hf, false, hf, bs.xhat, s, bp.tma_qkv));
  }

  DwParams dwp{};
  GramParams gp{};
  const void* v_ptr = (const char*)bs.qkv_dw + (size_t)2 * C * es;
  int v_ld = 3 * C;
  if (bp.fuse_attn) {
    // (1+2+3) norm1 + qkv 1x1 + depthwise 3x3 + q.k^T Gram partials + squared norms in one kernel; only v (fp16) is
    // written (:147, :114-115, :121-124)
    AttnFusedArgs fa{};
    fa.w_qkv = P(bp.qkv_w); fa.dw_chunked = P(bp.qkvdw_w); fa.v = bs.qkv_dw;
    fa.s_part = bs.s_part; fa.n_part = bs.n_part; fa.parts = attn_fused_parts(B, H, W);
    fa.B = B; fa.H = H; fa.W = W; fa.C = C; fa.heads = bp.heads;
    // norm1 as an fp16 tensor (bs.xhat2): written by the previous block's GDFN epilogue, else by a LayerNorm pass
    if (!xn1_in)
      IRB_TRY(launch_layernorm(x_in, C, bs.xhat2, C, 1, (long long)B * H * W, C, ln, P(bp.ln1_w), P(bp.ln1_b), s));
    fa.xn = bs.xhat2;
    IRB_TRY(launch_attn_fused(fa, s));
    gp.nparts = fa.parts;
    v_ptr = bs.qkv_dw; v_ld = C;
  } else if (bp.fuse_front) {
    // (2+3) depthwise 3x3 + q.k^T Gram partials + squared norms; only v is written (:114-115, :121-124)
    AttnFrontArgs fa{};
    fa.qkv = bs.qkv; fa.half = hf; fa.v_half = bp.v_half; fa.v = bs.qkv_dw; fa.dw_chunked = P(bp.qkvdw_w);
    fa.s_part = bs.s_part; fa.n_part = bs.n_part; fa.parts = attn_front_parts(B, H, W, C, bp.heads);
    fa.B = B; fa.H = H; fa.W = W; fa.C = C; fa.heads = bp.heads;
    IRB_TRY(launch_attn_front(fa, s));
    gp.nparts = fa.parts;
    v_ptr = bs.qkv_dw; v_ld = C;
  } else {
  // (2) depthwise 3x3 over the 3C channels (restormer.py:114 qkv_dwconv)
  dwp.in = (const float*)bs.qkv; dwp.ldi = 3 * C; dwp.out = (float*)bs.qkv_dw; dwp.ldo = 3 * C;
  dwp.in_half = hf; dwp.out_half = hf;
  dwp.w = P(bp.qkvdw_w); dwp.bias = P(bp.qkvdw_b); dwp.Cw = 3 * C;
  dwp.B = B; dwp.H = H; dwp.W = W; dwp.C = 3 * C; dwp.gate = 0; dwp.gate_off = 0; dwp.tag = TAG_DW_QKV;
  dwp.round_tf32 = bp.tma_attn && !hf;     // v is read by the tensor core straight from the TMA box
  IRB_TRY(bp.ref_kernels ? launch_dwconv_ref(dwp, s) : launch_dwconv(dwp, s));

  // (3) Gram q.k^T and row norms per (image, head) (restormer.py:121-124), split over pixel slices
  gp.qkv = (const float*)bs.qkv_dw; gp.ld = 3 * C; gp.B = B; gp.HW = H * W; gp.C = C; gp.heads = bp.heads;
  gp.in_half = hf;
  gp.nparts = gram_parts(B, bp.heads, H * W);
  gp.s_part = bs.s_part; gp.n_part = bs.n_part;
  IRB_TRY(bp.ref_kernels ? launch_gram_ref(gp, s) : launch_gram(gp, s));
  }

  // (4) normalise, temperature, softmax; fold project_out into a per-image C x C matrix (:124-131)
  FoldParams fp{};
  fp.s_part = bs.s_part; fp.n_part = bs.n_part; fp.B = B; fp.C = C; fp.heads = bp.heads; fp.nparts = gp.nparts;
  fp.temperature = P(bp.temp); fp.w_proj = P(bp.proj_w); fp.w_eff = (float*)bs.w_eff; fp.w_eff_bstride = (long long)C * bp.kp_attn;
  const bool ah = hf || bp.v_half;          // operand type of the attention-output contraction
  fp.fmt = !bp.tc_attn ? 0 : bp.tma_attn ? (ah ? 4 : 3) : hf ? 2 : 1;
  IRB_TRY(launch_fold(fp, s));

  // (5) x_out = x_in + W_eff[b] . v  (+ project_out bias)   (:127-131, :147)
  g = GemmParams{};
  g.a1 = (const float*)v_ptr; g.lda1 = v_ld; g.k1 = C; g.a_mode = A_PLAIN;
  g.B = B; g.H = H; g.W = W;
  g.w = (const float*)bs.w_eff; g.w_bstride = (long long)C * bp.kp_attn; g.N = C; g.K = C; g.Kp = C; g.bias = P(bp.proj_b);
  g.ln_mode = LN_NONE; g.acc_sign = 1.f;
  g.r = x_in; g.ldr = C; g.y = x_out; g.ldy = C; g.o_mode = O_NHWC; g.tag = TAG_ATTN_OUT;
  XnOut xo;
  if (bp.k4_xn) { xo.xn = bs.xhat; xo.ld = C; xo.ln_mode = ln; xo.w = P(bp.ln2_w); xo.b = P(bp.ln2_b); }
  IRB_TRY(run_1x1(g, bp.tc_attn, ah, ah, false, bs.xhat, s, bp.tma_attn, &xo));

  if (bp.fuse_ffn) {
    // (6-8) norm2 into an fp16 operand tensor, then project_in + depthwise 3x3 + gate + project_out + residual
    // (:148, :89-92) in one kernel: the hidden tensor never exists in HBM
    if (!bp.k4_xn)
      IRB_TRY(launch_layernorm(x_out, C, bs.xhat, C, 1, (long long)B * H * W, C, ln, P(bp.ln2_w), P(bp.ln2_b), s));
    FfnFusedArgs fa{};
    fa.xn = bs.xhat; fa.x = x_out; fa.w_in = P(bp.pin_w); fa.w_out = P(bp.pout_w); fa.dw_chunked = P(bp.ffdw_w);
    fa.B = B; fa.H = H; fa.W = W; fa.C = C; fa.hp = hp;
    if (next && xn1_ready && block_chains_norm1(bp, *next)) {
      // the next block's norm1 from this kernel's epilogue (the attention front of THIS block has long read bs.xhat2)
      fa.xn_next = bs.xhat2; fa.ln_w_next = P(next->ln1_w); fa.ln_b_next = P(next->ln1_b); fa.ln_mode_next = ln;
      *xn1_ready = true;
    }
    return launch_ffn_fused(fa, s);
  }

  // (6) norm2 + project_in 1x1 (:148, :89)
  g = GemmParams{};
  g.a1 = x_out; g.lda1 = C; g.k1 = C; g.a_mode = A_PLAIN;
  g.B = B; g.H = H; g.W = W;
  g.w = P(bp.pin_w); g.N = 2 * hp; g.K = C; g.Kp = C; g.bias = P(bp.pin_b);
  g.ln_mode = ln; g.ln_w = P(bp.ln2_w); g.ln_b = P(bp.ln2_b); g.acc_sign = 1.f;
  g.y = (float*)bs.hidden; g.ldy = 2 * hp; g.o_mode = O_NHWC; g.tag = TAG_LN_PIN;
  IRB_TRY(run_1x1(g, bp.tc_pin, hf, false, hf, bs.xhat, s, bp.tma_pin));

  if (bp.fuse_tail) {
    // (7+8) depthwise 3x3 + gelu(x1)*x2 + project_out + residual (:90-92, :148) in one kernel
    FfnTailArgs fa{};
    fa.hidden = bs.hidden; fa.half = hf; fa.x = x_out; fa.w_out = P(bp.pout_w); fa.dw_chunked = P(bp.ffdw_w);
    fa.bias = nullptr; fa.B = B; fa.H = H; fa.W = W; fa.C = C; fa.hp = hp;
    return launch_ffn_tail(fa, s);
  }
  // (7) depthwise 3x3 + gelu(x1)*x2 (:90-91)
  dwp = DwParams{};
  dwp.in = (const float*)bs.hidden; dwp.ldi = 2 * hp; dwp.out = (float*)bs.gated; dwp.ldo = hp;
  dwp.in_half = hf; dwp.out_half = hf;
  dwp.w = P(bp.ffdw_w); dwp.bias = P(bp.ffdw_b); dwp.Cw = 2 * hp;
  dwp.B = B; dwp.H = H; dwp.W = W; dwp.C = hp; dwp.gate = 1; dwp.gate_off = hp; dwp.tag = TAG_DW_GATE;
  dwp.round_tf32 = bp.tma_pout && !hf;
  IRB_TRY(bp.ref_kernels ? launch_dwconv_ref(dwp, s) : launch_dwconv(dwp, s));

  // (8) x_out += project_out . gated (:92, :148)
  g = GemmParams{};
  g.a1 = (const float*)bs.gated; g.lda1 = hp; g.k1 = hp; g.a_mode = A_PLAIN;
  g.B = B; g.H = H; g.W = W;
  g.w = P(bp.pout_w); g.N = C; g.K = hp; g.Kp = hp; g.bias = P(bp.pout_b);
  g.ln_mode = LN_NONE; g.acc_sign = 1.f;
  g.r = x_out; g.ldr = C; g.y = x_out; g.ldy = C; g.o_mode = O_NHWC; g.tag = TAG_FFN_OUT;
  IRB_TRY(run_1x1(g, bp.tc_pout, hf, hf, false, bs.xhat, s, bp.tma_pout));
  return IR_OK;
}

int block_forward(const BlockPlan& bp, const float* packed, float* x, int B, int H, int W, void* workspace,
                  size_t workspace_bytes, int ln_with_bias, cudaStream_t s) {
  IRB_REQUIRE(B > 0 && H > 0 && W > 0, "block: empty input");
  BlockScratchNeed n;
  block_scratch_need(n, bp, B, H, W);
  Carver cv(workspace, workspace_bytes);
  BlockScratch bs;
  carve_block_scratch(cv, bs, n);
  if (cv.off > workspace_bytes) { set_error("workspace too small"); return IR_ERR_WORKSPACE; }
  return run_block(bp, packed, x, x, B, H, W, bs, ln_with_bias, s);
}

// after: the first block of the stage that follows on the same tensor (refinement behind decoder level 1) or nullptr;
// xn1_ready: carried from / to that neighbouring stage
static int run_stage(const std::vector<BlockPlan>& blocks, const float* packed, const float* x_in, float* x_out, int B,
                     int H, int W, int C, const BlockScratch& bs, int lnb, cudaStream_t s, const BlockPlan* after = nullptr,
                     bool* xn1_ready = nullptr) {
  bool local = false;
  bool* ready = xn1_ready ? xn1_ready : &local;
  if (blocks.empty()) {   // nn.Sequential() of zero blocks is the identity
    if (x_in != x_out) IRB_TRY(launch_copy_channels(x_in, C, x_out, C, (long long)B * H * W, C, s));
    *ready = false;
    return IR_OK;
  }
  const float* cur = x_in;
  for (size_t i = 0; i < blocks.size(); ++i) {
    const BlockPlan* next = i + 1 < blocks.size() ? &blocks[i + 1] : after;
    IRB_TRY(run_block(blocks[i], packed, cur, x_out, B, H, W, bs, lnb, s, next, ready));
    cur = x_out;
  }
  return IR_OK;
}

static int conv3(const ConvPlan& cp, const float* packed, const float* in, int ld_in, int a_mode, int B, int H, int W,
                 float* out, int ld_out, int o_mode, const float* r, bool half, cudaStream_t s) {
  // wide rows of a narrow input (down1_2): every row crosses L2 once instead of nine times
  if (cp.tma && a_mode == A_IM2COL_NHWC && o_mode == O_UNSHUFFLE && r == nullptr && W >= 96 &&
      conv3_row_supported(cp.cin, cp.cout_p, half))
    return launch_conv3_row(in, ld_in, cp.cin, packed + cp.w, nullptr, 0, cp.cout_p, cp.cout, B, H, W, out, ld_out, o_mode, half, s);
  if (cp.tma && a_mode == A_IM2COL_NHWC && (o_mode == O_UNSHUFFLE || o_mode == O_SHUFFLE) && r == nullptr)
    return launch_conv3_tma(in, ld_in, cp.cin, packed + cp.w, nullptr, 0, cp.cout_p, cp.cout, B, H, W, out, ld_out, o_mode, half, s);
  if (cp.tc && a_mode == A_IM2COL_NHWC && o_mode != O_NCHW && r == nullptr)
    return run_conv3_tc(in, ld_in, cp.cin, packed + cp.w, cp.b >= 0 ? packed + cp.b : nullptr, cp.cout_p, cp.cout, B, H, W,
                        out, ld_out, o_mode, 0, half, s);
  if (!cp.tc && a_mode == A_IM2COL_NCHW && o_mode == O_NHWC && r == nullptr && conv3x3_first_supported(cp.cin, cp.cout))
    return launch_conv3x3_first(in, cp.cin, packed + cp.w, cp.kp, cp.b >= 0 ? packed + cp.b : nullptr, 0, cp.cout, B, H, W, out,
                                ld_out, s);
  if (!cp.tc && a_mode == A_IM2COL_NHWC && o_mode == O_NCHW && cp.cout <= 4 && cp.cin % 4 == 0 &&
      (size_t)cp.cout * 9 * cp.cin * sizeof(float) <= 48 * 1024)
    return launch_conv3x3_small(in, ld_in, cp.cin, packed + cp.w, cp.kp, cp.b >= 0 ? packed + cp.b : nullptr, cp.cout, B, H,
                                W, r, 1.f, out, s);
  GemmParams g{};
  g.a1 = in; g.lda1 = ld_in; g.k1 = cp.cin; g.a_mode = a_mode;
  g.B = B; g.H = H; g.W = W;
  g.w = packed + cp.w; g.N = cp.cout; g.K = cp.k; g.Kp = cp.kp; g.bias = cp.b >= 0 ? packed + cp.b : nullptr;
  g.ln_mode = LN_NONE; g.acc_sign = 1.f; g.r = r; g.ldr = 0;
  g.y = out; g.ldy = ld_out; g.o_mode = o_mode; g.tag = TAG_CONV3;
  return launch_gemm_simt(g, s);
}

int restormer_launch_count(const RestormerPlan& pl) {
  int n = 0;
  // chain: blocks that follow each other on one tensor; `prev` = the block before the first one of `v` (or nullptr)
  auto blocks = [&](const std::vector<BlockPlan>& v, const BlockPlan* prev) {
    for (const auto& bp : v) {
      // (the fused MDTA front replaces LN + qkv and the front kernel by a LayerNorm pass and itself; that pass disappears
      // when the previous block's GDFN emits norm1 from its epilogue)
      n += 8 - (bp.fuse_tail || bp.fuse_ffn ? 1 : 0) - (bp.fuse_front ? 1 : 0) - (bp.k4_xn ? 1 : 0);
      if (prev && bp.fuse_attn && block_chains_norm1(*prev, bp)) n -= 1;
      // standalone LayerNorm where the contraction cannot take it as a prologue (the wide levels)
      auto ln_standalone = [&](bool tc, bool tma, int N) {
        if (!tc) return false;
        if (tma) return bp.C > 128;
        TcGemmParams t{};
        t.K = bp.C; t.N = N; t.k1 = bp.C; t.k2 = 0; t.ln_mode = LN_BIASFREE; t.a_pad = 1;
        t.a_half = 0; t.op_half = bp.half; t.y_half = bp.half;
        return tc_gemm_configure(t) == 0;
      };
      n += (!bp.fuse_attn && ln_standalone(bp.tc_qkv, bp.tma_qkv, 3 * bp.C) ? 1 : 0) + (ln_standalone(bp.tc_pin, bp.tma_pin, 2 * bp.hp) ? 1 : 0);
      prev = &bp;
    }
  };
  for (int l = 0; l < 4; ++l) blocks(pl.enc[l], nullptr);
  for (int l = 2; l >= 1; --l) blocks(pl.dec[l], nullptr);
  blocks(pl.dec[0], nullptr);
  blocks(pl.refine, pl.dec[0].empty() ? nullptr : &pl.dec[0].back());
  // patch_embed, 3 down, 3 up, 2 reduce, 1 concat copy, output (+ skip_conv)
  return n + 11 + (pl.cfg.dual_pixel_task ? 1 : 0);
}

int restormer_forward(const RestormerPlan& pl, const float* packed, const float* x, float* y, int B, int H, int W,
                      void* workspace, size_t workspace_bytes, cudaStream_t s) {
  IRB_REQUIRE(B > 0 && H > 0 && W > 0, "restormer: empty input");
  IRB_REQUIRE(H % 8 == 0 && W % 8 == 0, "restormer: H and W must be multiples of 8 (three PixelUnshuffle(2) stages)");
  Carver cv(workspace, workspace_bytes);
  RestormerWs ws;
  restormer_carve(pl, B, H, W, cv, ws);
  if (cv.off > workspace_bytes) { set_error("workspace too small"); return IR_ERR_WORKSPACE; }
  const IrRestormerCfg& c = pl.cfg;
  const int d = c.dim, lnb = c.layernorm_with_bias;
  const bool dual = c.dual_pixel_task != 0;

  // patch_embed (:247): NCHW image -> channels-last [P, dim]
  float* e1_in = dual ? ws.e1_in : ws.e[0];
  IRB_TRY(conv3(pl.patch_embed, packed, x, 0, A_IM2COL_NCHW, B, H, W, e1_in, d, O_NHWC, nullptr, pl.half, s));
  IRB_TRY(run_stage(pl.enc[0], packed, e1_in, ws.e[0], B, H, W, d, ws.bs, lnb, s));                    // :248
  // encoder levels 2..4 (:250-257): 3x3 conv C->C/2 with PixelUnshuffle folded into the store
  for (int l = 1; l < 4; ++l) {
    const int hi = H >> (l - 1), wi = W >> (l - 1);
    IRB_TRY(conv3(pl.down[l - 1], packed, ws.e[l - 1], d << (l - 1), A_IM2COL_NHWC, B, hi, wi, ws.e[l], d << l,
                  O_UNSHUFFLE, nullptr, pl.half, s));
    IRB_TRY(run_stage(pl.enc[l], packed, ws.e[l], ws.e[l], B, hi / 2, wi / 2, d << l, ws.bs, lnb, s));
  }
  // decoder levels 3, 2 (:259-267): 3x3 conv C->2C with PixelShuffle folded into the store, then the
  // concat [upsampled, skip] is a two-source K loop of reduce_chan
  const float* below = ws.e[3];
  for (int l = 2; l >= 1; --l) {
    const int hi = H >> (l + 1), wi = W >> (l + 1);    // extent of the level below
    const int C = d << l;
    IRB_TRY(conv3(pl.up[l], packed, below, 2 * C, A_IM2COL_NHWC, B, hi, wi, ws.up_tmp, C, O_SHUFFLE, nullptr, pl.half, s));
    GemmParams g{};
    g.a1 = ws.up_tmp; g.lda1 = C; g.k1 = C; g.a2 = ws.e[l]; g.lda2 = C; g.k2 = C; g.a_mode = A_PLAIN;
    g.B = B; g.H = hi * 2; g.W = wi * 2;
    g.w = packed + pl.reduce[l].w; g.N = C; g.K = 2 * C; g.Kp = 2 * C;
    g.bias = pl.reduce[l].b >= 0 ? packed + pl.reduce[l].b : nullptr;
    g.ln_mode = LN_NONE; g.acc_sign = 1.f; g.y = ws.d[l]; g.ldy = C; g.o_mode = O_NHWC; g.tag = TAG_REDUCE;
    IRB_TRY(run_1x1(g, pl.reduce[l].tc, pl.half, false, false, nullptr, s));
    IRB_TRY(run_stage(pl.dec[l], packed, ws.d[l], ws.d[l], B, hi * 2, wi * 2, C, ws.bs, lnb, s));
    below = ws.d[l];
  }
  // level 1 (:269-273): up2_1 writes channels [0,dim) of the 2*dim-wide stream, the skip fills [dim,2dim)
  IRB_TRY(conv3(pl.up[0], packed, below, 2 * d, A_IM2COL_NHWC, B, H / 2, W / 2, ws.d[0], 2 * d, O_SHUFFLE, nullptr, pl.half, s));
  IRB_TRY(launch_copy_channels(ws.e[0], d, ws.d[0] + d, 2 * d, (long long)B * H * W, d, s));
  bool xn1 = false;     // refinement continues on decoder level 1's tensor: its first norm1 comes from the last decoder GDFN
  IRB_TRY(run_stage(pl.dec[0], packed, ws.d[0], ws.d[0], B, H, W, 2 * d, ws.bs, lnb, s,
                    pl.refine.empty() ? nullptr : &pl.refine[0], &xn1));
  IRB_TRY(run_stage(pl.refine, packed, ws.d[0], ws.d[0], B, H, W, 2 * d, ws.bs, lnb, s, nullptr, &xn1));
  if (dual) {
    // out += skip_conv(patch_embed output) (:276-277), then output conv without image residual (:278)
    GemmParams g{};
    g.a1 = e1_in; g.lda1 = d; g.k1 = d; g.a_mode = A_PLAIN; g.B = B; g.H = H; g.W = W;
    g.w = packed + pl.skip.w; g.N = 2 * d; g.K = d; g.Kp = d; g.bias = pl.skip.b >= 0 ? packed + pl.skip.b : nullptr;
    g.ln_mode = LN_NONE; g.acc_sign = 1.f; g.r = ws.d[0]; g.ldr = 2 * d; g.y = ws.d[0]; g.ldy = 2 * d; g.o_mode = O_NHWC; g.tag = TAG_REDUCE;
    IRB_TRY(run_1x1(g, pl.skip.tc, pl.half, false, false, nullptr, s));
    IRB_TRY(conv3(pl.output, packed, ws.d[0], 2 * d, A_IM2COL_NHWC, B, H, W, y, 0, O_NCHW, nullptr, pl.half, s));
  } else {
    IRB_TRY(conv3(pl.output, packed, ws.d[0], 2 * d, A_IM2COL_NHWC, B, H, W, y, 0, O_NCHW, x, pl.half, s));   // :281
  }
  return IR_OK;
}

}  // namespace irb
