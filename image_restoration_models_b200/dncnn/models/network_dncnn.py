"""Drop-in ``DnCNN`` whose forward runs in hand-written sm_100a CUDA (libirb200.so).

Mirrors the reference class (src/dncnn/models/network_dncnn.py:40-71 with the layer factory of
src/dncnn/models/basicblock.py:15-35,61-98): constructor ``DnCNN(in_nc, out_nc, nc, nb, act_mode)``,
``state_dict`` keys ``model.<i>.{weight,bias,...}`` with the flattened-Sequential indices, and
``forward(x) = x - model(x)`` on fp32 NCHW.  Only the mode letters the harness uses are supported:
'R' (what src/dncnn/__init__.py:8 builds) and 'BR' (the class default; eval-mode BatchNorm is folded
into the preceding conv when the weights are packed).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ... import _native
from ..._params import ConvParams, Holder, ordered_tensors

_MODES = {"fp32": _native.MODE_FP32, "half": _native.MODE_HALF, "fp32_simt": _native.MODE_FP32_SIMT,
          "bf16": _native.MODE_BF16}


class _BatchNormParams(nn.Module):
    """State of nn.BatchNorm2d(nc, momentum=0.9, eps=1e-4, affine=True) (basicblock.py:69)."""

    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class DnCNN(nn.Module):
    def __init__(self, in_nc=1, out_nc=1, nc=64, nb=17, act_mode='BR'):
        super().__init__()
        if act_mode not in ("R", "BR"):
            raise NotImplementedError("this implementation supports act_mode 'R' and 'BR' (conv [+BN] + ReLU)")
        has_bn = "B" in act_mode
        # flattened Sequential indices: conv, [BN], ReLU per layer (B.sequential, basicblock.py:15-35)
        layers = []
        for layer in range(nb):
            c_in = in_nc if layer == 0 else nc
            c_out = out_nc if layer == nb - 1 else nc
            layers.append(ConvParams(c_in, c_out, 3, True))
            if layer == nb - 1:
                break
            if has_bn and layer > 0:
                layers.append(_BatchNormParams(nc))
            layers.append(Holder())          # ReLU slot: keeps the reference's numbering, holds nothing
        self.model = nn.ModuleList(layers)
        self.in_nc, self.out_nc = int(in_nc), int(out_nc)
        self._cfg = _native.IrDncnnCfg(int(in_nc), int(out_nc), int(nc), int(nb), int(has_bn))
        self._mode = "fp32"
        self._graphs = "auto"        # CUDA-graph replay of the 17-20 launches: "auto" (inputs up to 1 Mpix), True, False
        self._packed = None
        self._workspace = None

    def set_cuda_graphs(self, enabled="auto"):
        """Replay the forward from a cached CUDA graph (ir_dncnn_forward_graph); see Restormer.set_cuda_graphs."""
        if enabled not in ("auto", True, False):
            raise ValueError("enabled must be 'auto', True or False")
        self._graphs = enabled
        self._workspace = None
        return self

    def set_mode(self, mode: str):
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {sorted(_MODES)}")
        self._mode = mode
        self._packed = None
        return self

    def invalidate_packed(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._workspace = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def train(self, mode: bool = True):
        if mode and self._cfg.has_bn:
            raise RuntimeError("inference-only implementation: BatchNorm runs on its running statistics (eval mode)")
        return super().train(mode)

    def launches_per_forward(self):
        return _native.lib().ir_dncnn_launch_count(C.byref(self._cfg))

    def _pack(self, device):
        lib = _native.lib()
        mode = _MODES[self._mode]
        tensors = ordered_tensors(self)
        n = lib.ir_dncnn_param_count(C.byref(self._cfg))
        if n < 0:
            raise ValueError(_native.last_error())
        if n != len(tensors):
            raise RuntimeError(f"internal: {len(tensors)} tensors vs {n} expected by the native plan")
        keep = []
        for i, t in enumerate(tensors):
            if t.dtype == torch.long:          # num_batches_tracked: not read by the native side
                keep.append(None)
                continue
            _native.require_cuda(t, "DnCNN parameter")
            if t.device != device:
                raise RuntimeError("parameters and input are on different devices")
            want = lib.ir_dncnn_param_numel(C.byref(self._cfg), i)
            if want != t.numel():
                raise RuntimeError(f"internal: parameter {i} has {t.numel()} elements, native plan expects {want}")
            keep.append(t.detach().contiguous())
        nbytes = lib.ir_dncnn_packed_bytes(C.byref(self._cfg), mode)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        _native.check(lib.ir_dncnn_pack_weights(C.byref(self._cfg), _native.ptr_array(keep), n, packed.data_ptr(),
                                                nbytes, mode, stream))
        self._packed = (device, self._mode, packed)
        return packed

    def forward(self, x):
        _native.require_cuda(x, "DnCNN.forward(x)")
        if x.dim() != 4 or x.shape[1] != self.in_nc:
            raise ValueError(f"expected [B, {self.in_nc}, H, W], got {tuple(x.shape)}")
        if self.training and self._cfg.has_bn:
            raise RuntimeError("inference-only implementation: call .eval() first")
        B, _, H, W = x.shape
        xin = x.detach().contiguous()
        dev = xin.device
        with torch.cuda.device(dev):
            lib = _native.lib()
            mode = _MODES[self._mode]
            pk = self._packed
            packed = pk[2] if pk is not None and pk[0] == dev and pk[1] == self._mode else self._pack(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            graph = self._graphs is True or (self._graphs == "auto" and B * H * W <= (1 << 20))
            key = (dev, self._mode, B, H, W, stream, graph)   # scratch is per stream
            if self._workspace is None or self._workspace[0] != key:
                self._workspace = None
                size_fn = lib.ir_dncnn_graph_workspace_bytes if graph else lib.ir_dncnn_workspace_bytes
                nbytes = size_fn(C.byref(self._cfg), B, H, W, mode)
                self._workspace = (key, torch.empty(nbytes, dtype=torch.uint8, device=dev))
            ws = self._workspace[1]
            y = torch.empty((B, self.out_nc, H, W), dtype=torch.float32, device=dev)
            fwd = lib.ir_dncnn_forward_graph if graph else lib.ir_dncnn_forward
            _native.check(fwd(C.byref(self._cfg), packed.data_ptr(), xin.data_ptr(), y.data_ptr(), B, H, W, ws.data_ptr(),
                              ws.numel(), mode, stream))
        return y
