"""Unified operator namespace over the CPU oracle (C + numpy) -- TEST INFRASTRUCTURE ONLY.

Mirrors the names of `lele::kernels::*` / `lele::features::*` so the same KAT replay and
parity tests can drive either this oracle or the CUDA product (`lele_b200.kernels`).
"""
from __future__ import annotations

import numpy as np

from . import binding as B
from . import np_ops as N

# f32 numeric kernels (C)
matmul = B.matmul
matmul_fused_add = B.matmul_fused_add
gemm = B.gemm
softmax = lambda x, axis=-1: B.softmax(x)
layer_norm = B.layer_norm
batch_norm = B.batch_norm
rms_norm = B.rms_norm
dynamic_quantize_linear = B.dynamic_quantize_linear
mat_mul_integer = B.mat_mul_integer
fused_quantized_linear = B.fused_quantized_linear
conv1d = B.conv1d
conv2d = B.conv2d
conv_transpose = B.conv_transpose
max_pool2d = B.max_pool2d
lstm = B.lstm
gru = B.gru
stft = B.stft
hann_window = B.hann_window
rfft = B.rfft
mel_filterbank = B.mel_filterbank
frontend = B.frontend
lfr = B.lfr
cmvn = B.cmvn
hz_to_mel = lambda hz: float(B.lib().lo_hz_to_mel_htk(hz))
relu = lambda x: B.unary("relu", x)
sigmoid = lambda x: B.unary("sigmoid", x)
tanh = lambda x: B.unary("tanh", x)
silu = lambda x: B.unary("silu", x)
erf = lambda x: B.unary("erf", x)
gelu = lambda x: B.unary("gelu", x)
exp = lambda x: B.unary("exp", x)
softplus = lambda x: B.unary("softplus", x)
# indexing / element-wise (numpy)
concat = N.concat
slice = N.slice_
pad = N.pad
gather = N.gather
transpose = N.transpose
split = N.split
where = N.where_op
expand = N.expand
tile = N.tile
reshape = N.reshape
flatten = N.flatten
squeeze = N.squeeze
unsqueeze = N.unsqueeze
topk = N.topk
gather_elements = N.gather_elements
resize_nearest = N.resize_nearest
add, sub, mul, div = N.add, N.sub, N.mul, N.div
maximum, neg, sqrt, reciprocal, clip, mod_f32, prelu = N.maximum, N.neg, N.sqrt, N.reciprocal, N.clip, N.mod_f32, N.prelu
pow, log, sin, cos, equal, less, not_ = N.pow, N.log, N.sin, N.cos, N.equal, N.less, N.not_  # noqa: A001
reduce = N.reduce
SenseVoiceRef = B.SenseVoiceRef


def conv_integer(x, w, x_zp, w_zp, dilations, group, pads, strides):
    """conv_integer (conv2d.rs:2216 -> conv2d_with_zero_points :1507-2000): an f32 convolution of (x - x_zp) with (w - w_zp); the
    im2col pads with RAW zeros, so a padded position contributes (0 - x_zp) (conv2d.rs:2025): pad first, shift second, convolve
    without padding."""
    p = list(pads) if len(pads) >= 4 else (list(pads) * 2 if len(pads) == 2 else [0, 0, 0, 0])
    x = np.asarray(x, np.float32)
    if any(p):
        x = N.pad(x, [0, 0, p[0], p[1], 0, 0, p[2], p[3]], 0.0, "constant")
    if x_zp != 0.0:
        x = N.sub(x, np.array([x_zp], np.float32))
    return B.conv2d(x, np.asarray(w, np.float32) - np.float32(w_zp), None, dilations, group, [0, 0, 0, 0], strides, 0)
