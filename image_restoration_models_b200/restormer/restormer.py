"""Drop-in ``Restormer`` whose forward runs in hand-written sm_100a CUDA (libirb200.so).

Mirrors the contract of the reference class (leducthanhig/image-restoration-models,
src/restormer/restormer.py:193-284): identical constructor kwargs, identical ``state_dict`` keys and
shapes, ``forward(x)`` taking and returning contiguous fp32 NCHW.  The sub-modules
only hold parameters; the arithmetic is one C-ABI call, ``ir_restormer_forward``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _native
from .._params import AffineParams, ConvParams, Holder, ordered_tensors

_MODES = {"fp32": _native.MODE_FP32, "half": _native.MODE_HALF, "fp32_simt": _native.MODE_FP32_SIMT,
          "fp32_strict": _native.MODE_FP32_STRICT, "bf16": _native.MODE_BF16}

# Range guard of the fast fp32 mode.  IR_MODE_FP32 keeps the tensors that are only ever tensor-core operands (norm1 / norm2
# output, v, the fused kernels' on-chip operands) as fp16, and runs the two low-resolution levels (C > 128) on the 16-bit
# plan (qkv, hidden, gated as fp16): tf32's mantissa but 5 exponent bits.  The device
# conversions saturate (+-65504, never inf).  At pack time `fp16_range_bound` estimates how large those tensors get from
# the weights alone; a model whose estimate comes near fp16's range runs in IR_MODE_FP32_STRICT (tf32 operands and
# fp32 tensors everywhere) instead.  The estimate is statistical, not worst-case (an L1 worst case is ~1000x above what a
# network produces and would always trip): LayerNorm outputs are taken as zero-mean with standard deviation m * |w_k|,
# m = 1 for WithBias and sqrt(1 + r^2) for BiasFree (which divides x, not x - mean, by the std; r = |mean| / std), a
# contraction output n then has std sqrt(sum_k W[n,k]^2 var_k) (+ the bias path), the depthwise conv scales it by at
# most sum |taps|, and every tensor is bounded at RANGE_SIGMAS standard deviations.
FP16_MAX = 65504.0
FP16_GUARD_FRACTION = 0.25          # switch to the strict mode when the estimate exceeds a quarter of fp16's range
BIASFREE_MEAN_OVER_STD = 4.0
RANGE_SIGMAS = 6.0


def _block_fp16_bound(blk, with_bias: bool, fused_gdfn: bool, wide: bool = False) -> torch.Tensor:
    """RANGE_SIGMAS-sigma estimate of max(|xn2|, |hidden|, |gated|, |v|) of one TransformerBlock (restormer.py:88-93,
    :111-116); xn2 / hidden / gated count where the block keeps them as fp16: under the fused GDFN (C <= 128) and at the
    wide levels (C > 128), which IR_MODE_FP32 runs on the 16-bit plan -- there qkv before and after its depthwise conv
    and norm1's output count too.  0-d device tensor."""
    m = 1.0 if with_bias else (1.0 + BIASFREE_MEAN_OVER_STD ** 2) ** 0.5

    def ln_stats(body):
        w = body.weight.detach().double()
        bias = getattr(body, "bias", None)
        return (m * w) ** 2, (bias.detach().abs().double() if bias is not None else torch.zeros_like(w))

    def conv_out(weight, var, mean_abs):            # std and |mean| bound per output channel of a 1x1 conv
        w = weight.detach().double().flatten(1)
        return torch.sqrt((w * w) @ var), w.abs() @ mean_abs

    k = RANGE_SIGMAS
    var1, mu1 = ln_stats(blk.norm1.body)
    c = var1.numel()
    sq, mq = conv_out(blk.attn.qkv.weight, var1, mu1)
    dq = blk.attn.qkv_dwconv.weight.detach().abs().double().flatten(1).sum(1)
    out = [((k * sq + mq) * dq)[2 * c:].max()]                                   # v
    if wide:
        out += [(k * torch.sqrt(var1) + mu1).max(), (k * sq + mq).max(), ((k * sq + mq) * dq).max()]   # xn1, qkv, dw(qkv)
    if fused_gdfn or wide:
        var2, mu2 = ln_stats(blk.norm2.body)
        sh, mh = conv_out(blk.ffn.project_in.weight, var2, mu2)
        di = blk.ffn.dwconv.weight.detach().abs().double().flatten(1).sum(1)
        hid = k * sh + mh
        dwh = hid * di
        h = dwh.numel() // 2
        out += [(k * torch.sqrt(var2) + mu2).max(), hid.max(), (dwh[:h] * dwh[h:]).max()]   # xn2, hidden, |gelu(a) b| <= |a||b|
    return torch.stack(out).max()


def _transformer_block(dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type):
    """Parameter tree of one TransformerBlock (reference :137-144, Attention :99-107, FeedForward :76-86)."""
    hidden = int(dim * ffn_expansion_factor)
    with_bias = LayerNorm_type != "BiasFree"
    blk = Holder()
    blk.norm1 = Holder()
    blk.norm1.body = AffineParams(dim, with_bias)
    blk.attn = Holder()
    blk.attn.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
    blk.attn.qkv = ConvParams(dim, dim * 3, 1, bias)
    blk.attn.qkv_dwconv = ConvParams(dim * 3, dim * 3, 3, bias, groups=dim * 3)
    blk.attn.project_out = ConvParams(dim, dim, 1, bias)
    blk.norm2 = Holder()
    blk.norm2.body = AffineParams(dim, with_bias)
    blk.ffn = Holder()
    blk.ffn.project_in = ConvParams(dim, hidden * 2, 1, bias)
    blk.ffn.dwconv = ConvParams(hidden * 2, hidden * 2, 3, bias, groups=hidden * 2)
    blk.ffn.project_out = ConvParams(hidden, dim, 1, bias)
    return blk


def _resample(c_in, c_out):
    m = Holder()
    m.body = nn.ModuleList([ConvParams(c_in, c_out, 3, False)])   # key: <name>.body.0.weight
    return m


class Restormer(nn.Module):
    def __init__(self,
                 inp_channels=3,
                 out_channels=3,
                 dim=48,
                 num_blocks=[4, 6, 6, 8],
                 num_refinement_blocks=4,
                 heads=[1, 2, 4, 8],
                 ffn_expansion_factor=2.66,
                 bias=False,
                 LayerNorm_type='WithBias',
                 dual_pixel_task=False,
                 ):
        super().__init__()
        num_blocks, heads = list(num_blocks), list(heads)
        if len(num_blocks) != 4 or len(heads) != 4:
            raise ValueError("num_blocks and heads must have four entries")

        def stage(c, h, n):
            return nn.ModuleList([_transformer_block(c, h, ffn_expansion_factor, bias, LayerNorm_type)
                                  for _ in range(n)])

        # registration order == the reference's (it fixes state_dict order, which the C side relies on)
        self.patch_embed = Holder()
        self.patch_embed.proj = ConvParams(inp_channels, dim, 3, False)
        self.encoder_level1 = stage(dim, heads[0], num_blocks[0])
        self.down1_2 = _resample(dim, dim // 2)
        self.encoder_level2 = stage(dim * 2, heads[1], num_blocks[1])
        self.down2_3 = _resample(dim * 2, dim)
        self.encoder_level3 = stage(dim * 4, heads[2], num_blocks[2])
        self.down3_4 = _resample(dim * 4, dim * 2)
        self.latent = stage(dim * 8, heads[3], num_blocks[3])
        self.up4_3 = _resample(dim * 8, dim * 16)
        self.reduce_chan_level3 = ConvParams(dim * 8, dim * 4, 1, bias)
        self.decoder_level3 = stage(dim * 4, heads[2], num_blocks[2])
        self.up3_2 = _resample(dim * 4, dim * 8)
        self.reduce_chan_level2 = ConvParams(dim * 4, dim * 2, 1, bias)
        self.decoder_level2 = stage(dim * 2, heads[1], num_blocks[1])
        self.up2_1 = _resample(dim * 2, dim * 4)
        self.decoder_level1 = stage(dim * 2, heads[0], num_blocks[0])
        self.refinement = stage(dim * 2, heads[0], num_refinement_blocks)
        self.dual_pixel_task = dual_pixel_task
        if self.dual_pixel_task:
            self.skip_conv = ConvParams(dim, dim * 2, 1, bias)
        self.output = ConvParams(dim * 2, out_channels, 3, bias)

        self.inp_channels, self.out_channels = int(inp_channels), int(out_channels)
        self._cfg = _native.IrRestormerCfg(
            int(inp_channels), int(out_channels), int(dim), (C.c_int32 * 4)(*num_blocks), int(num_refinement_blocks),
            (C.c_int32 * 4)(*heads), float(ffn_expansion_factor), int(bool(bias)),
            int(LayerNorm_type != "BiasFree"), int(bool(dual_pixel_task)))
        self._with_bias_ln = LayerNorm_type != "BiasFree"
        self._conv_bias = bool(bias)
        self._mode = "fp32"
        self._range_guard = True     # fp32 mode: fall back to fp32_strict when fp16_range_bound() nears fp16's range
        self._graphs = "auto"        # CUDA-graph replay of the launch sequence: "auto" (small inputs), True, False
        self._native_mode = None     # the IrMode the packed weights were built for (fp32 may resolve to fp32_strict)
        self._packed = None          # (device, mode, tensor)
        self._workspace = None       # (key, tensor)

    # ------------------------------------------------------------------ packed-weight lifetime
    def set_mode(self, mode: str):
        """'fp32': tf32 tensor-core operands, fp32 intermediates.  'half': fp16 tensor-core operands and fp16
        intermediates (same 10-bit mantissa as tf32; residual stream, statistics and accumulation stay fp32).
        'fp32_simt': every contraction on CUDA cores in exact fp32 (the on-device reference used by the tests).
        'bf16': the half mode with bfloat16 instead of float16 (IR_MODE_BF16; 8 mantissa bits: outside the 1e-3 parity
        bar, reported separately)."""
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {sorted(_MODES)}")
        self._mode = mode
        self._packed = None
        return self

    GRAPH_AUTO_MAX_PIXELS = 1 << 20

    def set_cuda_graphs(self, enabled="auto"):
        """Replay the forward's ~280 launches from a cached CUDA graph (ir_restormer_forward_graph).  "auto" (default): for
        inputs of at most GRAPH_AUTO_MAX_PIXELS pixels per call -- the launch-bound regime of the reference harness, which
        feeds one 256x256 / 512x512 tile at a time (src/utils.py:403-419).  Results are bit-identical either way."""
        if enabled not in ("auto", True, False):
            raise ValueError("enabled must be 'auto', True or False")
        self._graphs = enabled
        self._workspace = None
        return self

    def _use_graph(self, B, H, W):
        return self._graphs is True or (self._graphs == "auto" and B * H * W <= self.GRAPH_AUTO_MAX_PIXELS)

    def set_range_guard(self, enabled: bool):
        """fp32 mode only: enable / disable the pack-time fp16 range guard (on by default)."""
        self._range_guard = bool(enabled)
        self._packed = None
        return self

    def _blocks(self):
        for stage in (self.encoder_level1, self.encoder_level2, self.encoder_level3, self.latent, self.decoder_level3,
                      self.decoder_level2, self.decoder_level1, self.refinement):
            yield from stage

    def fp16_range_bound(self) -> float:
        """Largest magnitude any fp16-held tensor of IR_MODE_FP32 can reach, bounded from the weights alone."""
        if self._conv_bias:
            return 0.0            # conv biases disable every fused kernel: no fp16 tensor exists in fp32 mode
        def fused(blk):           # mirrors ffn_fused_supported (csrc/ffn_fused.cu): C <= 128, hidden padded to 64s
            c = blk.norm2.body.weight.numel()
            hp = -(-(blk.ffn.project_out.weight.shape[1]) // 16) * 16
            return c % 16 == 0 and c <= 128 and hp % 64 == 0 and hp >= 128
        wide = lambda blk: blk.norm2.body.weight.numel() > 128      # mirrors plan_block (csrc/restormer.cu): 16-bit plan
        bounds = [_block_fp16_bound(b, self._with_bias_ln, fused(b), wide(b)) for b in self._blocks()]
        return float(torch.stack(bounds).max()) if bounds else 0.0

    def resolved_mode(self) -> str:
        """The mode the next forward runs in: 'fp32' resolves to 'fp32_strict' when the range guard trips."""
        if self._mode == "fp32" and self._range_guard and self.fp16_range_bound() > FP16_GUARD_FRACTION * FP16_MAX:
            return "fp32_strict"
        return self._mode

    def invalidate_packed(self):
        """Call after modifying parameters in place."""
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._workspace = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _pack(self, device):
        lib = _native.lib()
        mode = _MODES[self.resolved_mode()]
        tensors = ordered_tensors(self)
        n = lib.ir_restormer_param_count(C.byref(self._cfg))
        if n < 0:
            raise ValueError(_native.last_error())
        if n != len(tensors):
            raise RuntimeError(f"internal: {len(tensors)} tensors vs {n} expected by the native plan")
        keep = []
        for i, t in enumerate(tensors):
            _native.require_cuda(t, "Restormer parameter")
            if t.device != device:
                raise RuntimeError("parameters and input are on different devices")
            want = lib.ir_restormer_param_numel(C.byref(self._cfg), i)
            if want != t.numel():
                raise RuntimeError(f"internal: parameter {i} has {t.numel()} elements, native plan expects {want}")
            keep.append(t.detach().contiguous())
        nbytes = lib.ir_restormer_packed_bytes(C.byref(self._cfg), mode)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        _native.check(lib.ir_restormer_pack_weights(C.byref(self._cfg), _native.ptr_array(keep), n, packed.data_ptr(),
                                                    nbytes, mode, stream))
        self._packed = (device, self._mode, packed)
        self._native_mode = mode
        return packed

    def workspace_bytes(self, B, H, W):
        return _native.lib().ir_restormer_workspace_bytes(C.byref(self._cfg), B, H, W, _MODES[self.resolved_mode()])

    def launches_per_forward(self):
        return _native.lib().ir_restormer_launch_count(C.byref(self._cfg))

    # ------------------------------------------------------------------ forward
    def forward(self, inp_img):
        _native.require_cuda(inp_img, "Restormer.forward(inp_img)")
        if inp_img.dim() != 4 or inp_img.shape[1] != self.inp_channels:
            raise ValueError(f"expected [B, {self.inp_channels}, H, W], got {tuple(inp_img.shape)}")
        B, _, H, W = inp_img.shape
        if H % 8 or W % 8:
            raise ValueError("H and W must be multiples of 8 (the caller pads, src/utils.py:174-181)")
        x = inp_img.detach().contiguous()
        dev = x.device
        with torch.cuda.device(dev):
            lib = _native.lib()
            pk = self._packed
            packed = pk[2] if pk is not None and pk[0] == dev and pk[1] == self._mode else self._pack(dev)
            mode = self._native_mode
            stream = torch.cuda.current_stream(dev).cuda_stream
            graph = self._use_graph(B, H, W)
            key = (dev, mode, B, H, W, stream, graph)   # scratch is per stream: forwards on two streams must not share it
            if self._workspace is None or self._workspace[0] != key:
                self._workspace = None      # release before allocating the new one
                size_fn = lib.ir_restormer_graph_workspace_bytes if graph else lib.ir_restormer_workspace_bytes
                nbytes = size_fn(C.byref(self._cfg), B, H, W, mode)
                self._workspace = (key, torch.empty(nbytes, dtype=torch.uint8, device=dev))
            ws = self._workspace[1]
            y = torch.empty((B, self.out_channels, H, W), dtype=torch.float32, device=dev)
            fwd = lib.ir_restormer_forward_graph if graph else lib.ir_restormer_forward
            _native.check(fwd(C.byref(self._cfg), packed.data_ptr(), x.data_ptr(), y.data_ptr(), B, H, W, ws.data_ptr(),
                              ws.numel(), mode, stream))
        return y
