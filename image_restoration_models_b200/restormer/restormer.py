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

_MODES = {"fp32": _native.MODE_FP32, "half": _native.MODE_HALF, "fp32_simt": _native.MODE_FP32_SIMT}


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
        self._mode = "fp32"
        self._packed = None          # (device, mode, tensor)
        self._workspace = None       # (key, tensor)

    # ------------------------------------------------------------------ packed-weight lifetime
    def set_mode(self, mode: str):
        """'fp32': tf32 tensor-core operands, fp32 intermediates.  'half': fp16 tensor-core operands and fp16
        intermediates (same 10-bit mantissa as tf32; residual stream, statistics and accumulation stay fp32).
        'fp32_simt': every contraction on CUDA cores in exact fp32 (the on-device reference used by the tests)."""
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {sorted(_MODES)}")
        self._mode = mode
        self._packed = None
        return self

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
        mode = _MODES[self._mode]
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
        return packed

    def workspace_bytes(self, B, H, W):
        return _native.lib().ir_restormer_workspace_bytes(C.byref(self._cfg), B, H, W, _MODES[self._mode])

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
            mode = _MODES[self._mode]
            pk = self._packed
            packed = pk[2] if pk is not None and pk[0] == dev and pk[1] == self._mode else self._pack(dev)
            key = (dev, self._mode, B, H, W)
            if self._workspace is None or self._workspace[0] != key:
                self._workspace = None      # release before allocating the new one
                nbytes = lib.ir_restormer_workspace_bytes(C.byref(self._cfg), B, H, W, mode)
                self._workspace = (key, torch.empty(nbytes, dtype=torch.uint8, device=dev))
            ws = self._workspace[1]
            y = torch.empty((B, self.out_channels, H, W), dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.ir_restormer_forward(C.byref(self._cfg), packed.data_ptr(), x.data_ptr(), y.data_ptr(),
                                                   B, H, W, ws.data_ptr(), ws.numel(), mode, stream))
        return y
