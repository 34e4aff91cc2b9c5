"""Shared case list + runner for the tcgen05 1x1-contraction unit tests (GPU)."""
import numpy as np
import torch

from image_restoration_models_b200 import _native

# (k1, k2, N, B, HW, ln_mode, bias, residual, a_pad)
CASES = [
    # the shapes of the Restormer block at dim=48
    (48, 0, 144, 1, 256, 1, False, False, 1),     # K1 enc1: LN(BiasFree)+qkv, one 16-column tail
    (48, 0, 144, 1, 256, 1, False, False, 0),     # same, unpadded A slabs
    (48, 0, 144, 2, 200, 2, True, False, 1),      # WithBias LN, conv bias, ragged tiles (200 = 128 + 72), batch 2
    (48, 0, 256, 1, 384, 1, False, False, 1),     # K5 enc1: LN + project_in (2*hp = 256)
    (128, 0, 48, 1, 300, 0, False, True, 1),      # K6 enc1: project_out K=hp=128 + residual in place
    (48, 0, 48, 3, 130, 0, False, True, 1),       # K4 enc1 shape (shared weights here)
    (96, 0, 288, 1, 512, 1, False, False, 1),     # K1 level 2 / dec1: N split in chunks
    (96, 0, 288, 1, 512, 2, False, False, 0),
    (96, 0, 512, 2, 256, 1, False, False, 1),     # K5 level 2
    (256, 0, 96, 1, 256, 0, True, True, 1),       # K6 level 2: K chunked (4 x 64)
    (96, 0, 96, 1, 128, 0, False, True, 1),       # K4 level 2
    (192, 0, 576, 1, 256, 2, False, False, 1),    # K1 level 3: standalone LN + chunked K
    (192, 0, 1024, 1, 128, 1, False, False, 1),   # K5 level 3
    (512, 0, 192, 1, 200, 0, False, True, 1),     # K6 level 3
    (384, 0, 1152, 2, 64, 1, False, False, 1),    # K1 latent
    (384, 0, 2048, 1, 64, 2, True, False, 1),     # K5 latent
    (1024, 0, 384, 1, 96, 0, False, True, 1),     # K6 latent
    (192, 192, 192, 1, 256, 0, False, False, 1),  # reduce_chan_level3: concat of two sources
    (96, 96, 96, 2, 144, 0, True, False, 1),      # reduce_chan_level2 (chunk straddles the two sources)
    (48, 0, 96, 1, 64, 0, False, True, 1),        # skip_conv
    (32, 0, 96, 1, 100, 1, False, False, 1),      # dim=32 family
    (64, 0, 128, 1, 1, 2, False, False, 1),       # single pixel (latent of an 8x8 image)
    (88, 0, 32, 1, 77, 0, False, False, 1),       # odd multiple-of-8 K
]


# fp16-operand variants: (k1, k2, N, B, HW, ln_mode, bias, residual, a_pad, a_half, y_half)
HALF_CASES = [
    (48, 0, 144, 1, 256, 1, False, False, 1, 0, 1),     # K1: fp32 x -> LN -> fp16 operands -> fp16 qkv
    (96, 0, 288, 2, 200, 2, True, False, 1, 0, 1),      # K1 level 2: two 144-column sub-chunks
    (96, 0, 512, 1, 300, 1, False, False, 1, 0, 1),     # K5 level 2
    (192, 0, 576, 1, 128, 2, False, False, 1, 0, 1),    # K1 level 3: LN fused up to K = 256 in fp16 mode
    (384, 0, 1152, 1, 64, 1, False, False, 1, 0, 1),    # K1 latent: standalone LN (fp16 xhat) + chunked K
    (48, 0, 48, 2, 130, 0, False, True, 1, 1, 0),       # K4: fp16 v -> fp32 residual stream
    (128, 0, 48, 1, 256, 0, False, True, 1, 1, 0),      # K6 level 1: fp16 gated hidden
    (256, 0, 96, 1, 384, 0, True, True, 1, 1, 0),       # K6 level 2
    (1024, 0, 384, 1, 96, 0, False, True, 1, 1, 0),     # K6 latent: 4 K-chunks of 256
    (192, 192, 192, 1, 256, 0, False, False, 1, 0, 0),  # reduce_chan: fp32 sources, fp16 operands, fp32 out
    (48, 0, 96, 1, 64, 0, False, True, 1, 0, 0),        # skip_conv
]


# cases for the TMA-fed kernel (engine 3): same tuple layouts as CASES / HALF_CASES; leading dimensions keep rows
# 16-byte aligned for fp16 outputs as well (a TMA requirement, always true inside the network)
TMA_CASES = [
    (48, 0, 144, 1, 256, 1, False, False, 1),              # K1 enc1: LN(BiasFree) in place, tail group of 16 columns
    (48, 0, 144, 2, 200, 2, True, False, 1),               # WithBias LN + conv bias, ragged tiles, batch 2
    (48, 0, 256, 1, 384, 1, False, False, 1),              # K5 enc1
    (128, 0, 48, 1, 300, 0, False, True, 1),               # K6 enc1: residual ring, in place
    (48, 0, 48, 3, 130, 0, False, True, 1),                # K4 enc1
    (96, 0, 288, 1, 512, 1, False, False, 1),              # K1 level 2: two N-chunks (160 + 128)
    (96, 0, 512, 2, 256, 2, False, False, 1),              # K5 level 2: two N-chunks of 256
    (256, 0, 96, 1, 256, 0, True, True, 1),                # K6 level 2: 8 K boxes, bias + residual
    (96, 0, 96, 1, 128, 0, False, True, 1),                # K4 level 2
    (192, 0, 192, 2, 200, 0, False, True, 1),              # K4 level 3
    (88, 0, 32, 1, 77, 0, False, False, 1),                # K not a multiple of the box width
    (96, 0, 288, 4, 4096, 1, False, False, 1),             # many tiles per CTA: ring wrap-around, both accumulators
    (256, 0, 96, 3, 5000, 0, False, True, 1),              # residual ring wrap-around
    (48, 0, 144, 1, 256, 1, False, False, 1, 0, 1),        # fp16: LN -> fp16 operand ring -> fp16 out
    (96, 0, 288, 2, 200, 2, True, False, 1, 0, 1),
    (96, 0, 512, 1, 300, 1, False, False, 1, 0, 1),
    (48, 0, 48, 2, 130, 0, False, True, 1, 1, 0),          # fp16 v -> fp32 residual stream
    (128, 0, 48, 1, 256, 0, False, True, 1, 1, 0),
    (256, 0, 96, 1, 384, 0, True, True, 1, 1, 0),
    (512, 0, 192, 1, 200, 0, False, True, 1, 1, 0),        # K6 level 3 in fp16: two N-chunks
    (96, 0, 512, 3, 4096, 1, False, False, 1, 0, 1),       # many tiles, fp16
    (192, 0, 576, 1, 256, 2, False, False, 1),             # K1 level 3: standalone LN (tf32-rounded store) + TMA contraction
    (384, 0, 1152, 1, 128, 1, False, False, 1),            # K1 latent, same path
]
# the two wide-LayerNorm cases above also run with a tight bound: unrounded fp32 operands would be TRUNCATED by the
# kind::tf32 MMA (twice the error of round-to-nearest, and a bias towards zero)
TMA_WIDE_LN_CASES = [len(TMA_CASES) - 2, len(TMA_CASES) - 1]


def run_case(case, engine, seed=0):
    """Returns (y, y_ref64) as numpy arrays; y from the C-ABI test entry on the given engine (0 tc, 1 simt)."""
    a_half = y_half = op_half = 0
    if len(case) == 11:
        a_half, y_half = case[9], case[10]
        op_half = 1
        case = case[:9]
    k1, k2, N, B, HW, ln_mode, bias, resid, a_pad = case
    g = torch.Generator().manual_seed(1000 + seed)
    K = k1 + k2
    rows = B * HW
    lda1, lda2, ldy = k1 + 8, (k2 + 4 if k2 else 0), N + (16 if engine == 3 else 12)   # non-trivial leading dimensions
    a1 = torch.randn(rows, lda1, generator=g) * 1.5 + 0.3
    a2 = torch.randn(rows, lda2, generator=g) if k2 else None
    if a_half:      # the source tensor itself is fp16: the reference sees the same rounded values
        a1 = a1.half().float()
        a2 = a2.half().float() if k2 else None
    w = (torch.rand(N, K, generator=g) * 2 - 1) / np.sqrt(K)
    bvec = (torch.rand(N, generator=g) - 0.5) if bias else None
    lw = torch.rand(k1, generator=g) + 0.5
    lb = (torch.rand(k1, generator=g) - 0.5) * 0.2
    r = torch.randn(rows, ldy, generator=g) if resid else None

    # fp64 reference
    x = a1[:, :k1].double()
    if ln_mode:
        var = x.var(dim=1, keepdim=True, unbiased=False)
        if ln_mode == 1:
            x = x / torch.sqrt(var + 1e-5) * lw.double()
        else:
            x = (x - x.mean(dim=1, keepdim=True)) / torch.sqrt(var + 1e-5) * lw.double() + lb.double()
    if k2:
        x = torch.cat([x, a2[:, :k2].double()], 1)
    y_ref = x @ w.double().t()
    if bias:
        y_ref = y_ref + bvec.double()
    if resid:
        y_ref = y_ref + r[:, :N].double()

    dev = "cuda"
    lib = _native.lib()
    d = lambda t: None if t is None else t.to(dev).contiguous()
    a1d, a2d, wd, bd, lwd, lbd = d(a1), d(a2), d(w), d(bvec), d(lw), d(lb)
    if a_half:
        a1d = a1d.half()
        a2d = a2d.half() if k2 else None
    y = d(r) if resid else torch.full((rows, ldy), float("nan"), device=dev)
    if y_half:
        y = y.half()
    scratch = torch.empty((N * (K + 64) + rows * K) * 4 + 1024, dtype=torch.uint8, device=dev)
    P = lambda t: 0 if t is None else t.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    st = lib.ir_test_conv1x1(engine, P(a1d), lda1, k1, P(a2d), lda2, k2, P(wd), P(bd), ln_mode, P(lwd), P(lbd),
                             P(y) if resid else 0, ldy, P(y), ldy, B, HW, N, a_pad, a_half, op_half, y_half,
                             P(scratch), scratch.numel(), stream)
    _native.check(st)
    torch.cuda.synchronize()
    return y[:, :N].float().cpu().numpy(), y_ref.numpy()


def tolerance(case, y_ref):
    """tf32 / fp16 operands (10-bit mantissa, RN) with fp32 accumulation: relative 2^-11 per operand
    (+ 2^-11 output rounding when y is fp16)."""
    return (5e-3 if len(case) == 11 else 4e-3) * float(np.abs(y_ref).max()) + 1e-5


# 3x3 convolution cases: (cin, cout, B, H, W, o_mode, bias, relu, op_half)
CONV3_CASES = [
    (64, 64, 1, 24, 40, 0, True, True, 0),       # DnCNN body layer: bias + ReLU, plain rows
    (64, 64, 2, 19, 31, 0, True, True, 0),       # odd extent, batch 2
    (48, 32, 1, 16, 24, 1, False, False, 0),     # Downsample level 1 shape family (48 -> 24 is below the N%16 rule; 32 here)
    (96, 48, 1, 16, 16, 1, False, False, 0),     # down2_3: 96 -> 48 + PixelUnshuffle
    (192, 96, 2, 8, 8, 1, False, False, 0),      # down3_4
    (384, 768, 1, 8, 8, 2, False, False, 0),     # up4_3: 384 -> 768 + PixelShuffle, streamed weights, 3 N-chunks
    (192, 384, 1, 8, 16, 2, False, False, 0),    # up3_2
    (96, 192, 1, 16, 16, 2, False, False, 0),    # up2_1
    (96, 192, 1, 16, 16, 2, False, False, 1),    # fp16 operands
    (64, 64, 1, 32, 32, 0, True, True, 1),
]


# shuffle-scatter cases for the TMA-fed implicit GEMM (engine 3): (cin, cout, B, H, W, o_mode, bias, relu, op_half)
TMA_CONV3_CASES = [
    (48, 32, 1, 16, 24, 1, False, False, 0),     # down1_2 family: 1.5 boxes per tap, ragged patch grid
    (96, 48, 1, 16, 16, 1, False, False, 0),     # down2_3
    (192, 96, 2, 8, 8, 1, False, False, 0),      # down3_4, batch 2, one patch per image
    (384, 768, 1, 8, 8, 2, False, False, 0),     # up4_3: three N-chunks of 256
    (192, 384, 1, 8, 16, 2, False, False, 0),    # up3_2: two N-chunks
    (96, 192, 1, 16, 16, 2, False, False, 0),    # up2_1
    (96, 192, 2, 40, 72, 2, False, False, 0),    # several patches per CTA, partial patches on both edges
    (48, 32, 1, 64, 96, 1, False, False, 1),     # fp16 operands
    (96, 192, 1, 24, 40, 2, False, False, 1),
    (384, 768, 1, 8, 8, 2, False, False, 1),
    (64, 64, 2, 19, 31, 0, True, True, 0),       # DnCNN body layer on the TMA kernel: plain rows, bias + ReLU, odd extent
    (64, 64, 1, 32, 32, 0, True, True, 1),
    (64, 48, 1, 16, 24, 0, True, False, 0),      # 1.5 output groups, no activation
    # tf32 patch mode (8 x 14 output tiles cut out of one (8+2) x (14+2) patch load): widths around the multiples of 14
    (96, 48, 1, 18, 30, 1, False, False, 0),     # three tile columns, the last two pixels wide; ragged bottom row of tiles
    (192, 384, 1, 9, 29, 2, False, False, 0),    # odd extent, PixelShuffle, two N-chunks
    (96, 192, 2, 17, 57, 2, False, False, 0),    # batch 2, five tile columns (the last one pixel wide)
    (64, 64, 1, 8, 14, 0, True, True, 0),        # exactly one tile, plain rows
]


# row-strip kernel (engine 4): Cin <= 64, plain rows or PixelUnshuffle
ROW_CONV3_CASES = [
    (64, 64, 1, 24, 200, 0, True, True, 0),      # DnCNN body: two strips, the second ragged; one row segment
    (64, 64, 2, 40, 130, 0, True, True, 0),      # batch 2, strip of 2 pixels
    (48, 32, 1, 32, 256, 1, False, False, 0),    # down1_2 family: 1.5 boxes per row, PixelUnshuffle
    (64, 48, 1, 9, 96, 0, True, False, 0),       # odd height, 1.5 output groups
    (64, 64, 1, 300, 128, 0, True, True, 0),     # many row segments, ring wrap-around
    (64, 64, 1, 33, 160, 0, True, True, 1),      # fp16 operand rows
    (48, 32, 1, 16, 128, 1, False, False, 1),
]


def run_conv3_case(case, engine, seed=0):
    import torch.nn.functional as F
    cin, cout, B, H, W, o_mode, bias, relu, op_half = case
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(B, cin, H, W, generator=g)
    w = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) / np.sqrt(9 * cin)
    bvec = (torch.rand(cout, generator=g) - 0.5) if bias else None
    ref = F.conv2d(x.double(), w.double(), bvec.double() if bias else None, padding=1)
    if relu:
        ref = F.relu(ref)
    if o_mode == 1:
        ref = F.pixel_unshuffle(ref, 2)
    elif o_mode == 2:
        ref = F.pixel_shuffle(ref, 2)
    ref = ref.permute(0, 2, 3, 1).contiguous()          # channels-last rows
    Co = ref.shape[-1]
    dev = "cuda"
    lib = _native.lib()
    xl = x.permute(0, 2, 3, 1).contiguous().to(dev)
    wd = w.to(dev).contiguous()
    bd = bvec.to(dev) if bias else None
    y = torch.full(tuple(ref.shape), float("nan"), device=dev)
    scratch = torch.empty(cout * 9 * (cin + 64) * 4 + 1024, dtype=torch.uint8, device=dev)
    P = lambda t: 0 if t is None else t.data_ptr()
    st = lib.ir_test_conv3x3(engine, P(xl), cin, cin, P(wd), P(bd), cout, B, H, W, P(y), Co, o_mode, int(relu), op_half,
                             P(scratch), scratch.numel(), torch.cuda.current_stream().cuda_stream)
    _native.check(st)
    torch.cuda.synchronize()
    return y.cpu().numpy().reshape(-1, Co), ref.numpy().reshape(-1, Co)
