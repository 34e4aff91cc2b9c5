"""CPU restatement of the reference Restormer forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Pure functions over a ``state_dict``; the arithmetic primitives are the same
third-party ATen ops the reference calls (conv2d, gelu, softmax, matmul), which
are not part of /root/reference.  Works in fp32 or fp64 (pass a .double()
state dict and input for a noise-free reference).

Every function cites the reference lines it restates
(paths relative to /root/reference/src/restormer/restormer.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _conv(sd, name, x, padding=0, groups=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), padding=padding, groups=groups)


def layer_norm(sd, prefix, x):
    """LayerNorm.forward :68-70 with BiasFree (:37-39) or WithBias (:54-57) body.

    Normalises over the channel axis per pixel (to_3d :19-20).  BiasFree does NOT
    subtract the mean in the numerator; variance is the population variance, eps inside the sqrt.
    """
    w = sd[prefix + ".body.weight"].view(1, -1, 1, 1)
    b = sd.get(prefix + ".body.bias")
    var = x.var(dim=1, keepdim=True, unbiased=False)
    if b is None:
        return x / torch.sqrt(var + 1e-5) * w
    mu = x.mean(dim=1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * w + b.view(1, -1, 1, 1)


def attention(sd, prefix, x):
    """Attention.forward (MDTA) :111-132."""
    b, c, h, w = x.shape
    temperature = sd[prefix + ".temperature"]
    heads = temperature.shape[0]
    qkv = _conv(sd, prefix + ".qkv_dwconv", _conv(sd, prefix + ".qkv", x), padding=1, groups=3 * c)  # :114
    q, k, v = qkv.chunk(3, dim=1)                                                                  # :115
    q = q.reshape(b, heads, c // heads, h * w)                                                     # :117-119
    k = k.reshape(b, heads, c // heads, h * w)
    v = v.reshape(b, heads, c // heads, h * w)
    q = F.normalize(q, dim=-1)                                                                     # :121-122
    k = F.normalize(k, dim=-1)
    attn = (q @ k.transpose(-2, -1)) * temperature                                                 # :124
    attn = attn.softmax(dim=-1)                                                                    # :125
    out = (attn @ v).reshape(b, c, h, w)                                                           # :127-129
    return _conv(sd, prefix + ".project_out", out)                                                 # :131


def feed_forward(sd, prefix, x):
    """FeedForward.forward (GDFN) :88-93."""
    y = _conv(sd, prefix + ".project_in", x)
    y = _conv(sd, prefix + ".dwconv", y, padding=1, groups=y.shape[1])
    x1, x2 = y.chunk(2, dim=1)
    return _conv(sd, prefix + ".project_out", F.gelu(x1) * x2)


def transformer_block(sd, prefix, x):
    """TransformerBlock.forward :146-150."""
    x = x + attention(sd, prefix + ".attn", layer_norm(sd, prefix + ".norm1", x))
    x = x + feed_forward(sd, prefix + ".ffn", layer_norm(sd, prefix + ".norm2", x))
    return x


def _stage(sd, name, x):
    i = 0
    while f"{name}.{i}.norm1.body.weight" in sd:
        x = transformer_block(sd, f"{name}.{i}", x)
        i += 1
    return x


def downsample(sd, name, x):
    """Downsample.forward :178-179 (3x3 conv C->C/2, PixelUnshuffle(2))."""
    return F.pixel_unshuffle(_conv(sd, name + ".body.0", x, padding=1), 2)


def upsample(sd, name, x):
    """Upsample.forward :188-189 (3x3 conv C->2C, PixelShuffle(2))."""
    return F.pixel_shuffle(_conv(sd, name + ".body.0", x, padding=1), 2)


def restormer_forward(sd, inp_img, taps=None):
    """Restormer.forward :245-284.  ``taps`` (optional dict) receives named intermediates."""
    def tap(k, v):
        if taps is not None:
            taps[k] = v
        return v

    dual = "skip_conv.weight" in sd
    e1_in = tap("patch_embed", _conv(sd, "patch_embed.proj", inp_img, padding=1))     # :247
    e1 = tap("encoder_level1", _stage(sd, "encoder_level1", e1_in))                   # :248
    e2 = tap("encoder_level2", _stage(sd, "encoder_level2", downsample(sd, "down1_2", e1)))   # :250-251
    e3 = tap("encoder_level3", _stage(sd, "encoder_level3", downsample(sd, "down2_3", e2)))   # :253-254
    lat = tap("latent", _stage(sd, "latent", downsample(sd, "down3_4", e3)))          # :256-257
    d3 = torch.cat([upsample(sd, "up4_3", lat), e3], 1)                               # :259-260
    d3 = tap("decoder_level3", _stage(sd, "decoder_level3", _conv(sd, "reduce_chan_level3", d3)))  # :261-262
    d2 = torch.cat([upsample(sd, "up3_2", d3), e2], 1)                                # :264-265
    d2 = tap("decoder_level2", _stage(sd, "decoder_level2", _conv(sd, "reduce_chan_level2", d2)))  # :266-267
    d1 = torch.cat([upsample(sd, "up2_1", d2), e1], 1)                                # :269-270
    d1 = tap("decoder_level1", _stage(sd, "decoder_level1", d1))                      # :271
    d1 = tap("refinement", _stage(sd, "refinement", d1))                              # :273
    if dual:                                                                          # :276-278
        d1 = d1 + _conv(sd, "skip_conv", e1_in)
        return _conv(sd, "output", d1, padding=1)
    return _conv(sd, "output", d1, padding=1) + inp_img                               # :281
