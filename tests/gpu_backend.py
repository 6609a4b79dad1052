"""Adapter exposing lele_b200 (the CUDA product, through the C ABI) under the same operator
names as oracle.reference_api, so the KAT replay and parity tests drive both identically."""
import numpy as np

from lele_b200 import features as F
from lele_b200 import kernels as K

matmul = K.matmul
matmul_fused_add = K.matmul_fused_add
gemm = K.gemm
layer_norm = K.layer_norm
softmax = K.softmax
batch_norm = K.batch_norm
rms_norm = K.rms_norm
dynamic_quantize_linear = K.dynamic_quantize_linear
mat_mul_integer = K.mat_mul_integer


def fused_quantized_linear(x, w_u8, w_scale, w_zp, bias, relu=False):
    return K.fused_quantized_linear(x, w_u8, w_scale, np.array([w_zp], np.float32), bias, relu)


def conv1d(x, w, bias=None, dilations=(1,), group=1, pads=(0, 0), strides=(1,), relu=False):
    return K.conv1d_fused(x, w, bias, dilations, group, pads, strides, relu)


def conv2d(x, w, bias=None, dilations=(1, 1), group=1, pads=(0, 0, 0, 0), strides=(1, 1), act=0):
    return K.conv2d(x, w, bias, dilations, group, pads, strides, act)


def conv_transpose(x, w, bias=None, dilations=(1, 1), pads=(0, 0, 0, 0), strides=(1, 1)):
    return K.conv_transpose(x, w, bias, dilations, pads, strides)


def max_pool2d(x, kernel, pads=(0, 0, 0, 0), strides=(1, 1), dilations=(1, 1), ceil_mode=False):
    return K.max_pool2d(x, kernel, pads, strides, dilations, ceil_mode)


def lstm(x, w, r, bias=None, h0=None, c0=None):
    return K.lstm(x, w, r, bias, None, h0, c0)


def gru(x, w, r, bias=None, h0=None):
    return K.gru(x, w, r, bias, h0)


def stft(sig, n_fft, hop, win, window=None, power=False):
    return K.stft(np.asarray(sig, np.float32).reshape(-1), n_fft, hop, win, window, power)


hann_window = F.hann_window
mel_filterbank = F.mel_filterbank
hz_to_mel = F.hz_to_mel_htk


def rfft(x):
    re, im = F.RealFft(len(x)).process(x)
    return re[0], im[0]


def frontend(pcm, want_mel=False):
    return F.SenseVoiceFrontend().compute(pcm, want_mel)


def lfr(x, m=7, n=6):
    return F.Lfr(m, n).compute(x)


def cmvn(x, eps=1e-5):
    return F.Cmvn(eps).compute(x)


relu, sigmoid, tanh, silu, erf, gelu, exp, softplus = K.relu, K.sigmoid, K.tanh_kernel, K.silu, K.erf, K.gelu, K.exp, K.softplus
concat, pad, gather, transpose, split, expand, tile, reshape = K.concat, K.pad, K.gather, K.transpose, K.split, K.expand, K.tile, K.reshape
slice = K.slice
where = K.where_op
topk = lambda x, k: K.topk(x, k)
gather_elements = K.gather_elements
add, sub, mul, div = K.add, K.sub, K.mul, K.div
maximum, neg, sqrt, reciprocal, clip, mod_f32, prelu = K.max, K.neg, K.sqrt, K.reciprocal, K.clip, K.mod_f32, K.prelu
pow, log, sin, cos, equal, less, not_, flatten = K.pow, K.log, K.sin, K.cos, K.equal, K.less, K.not_, K.flatten  # noqa: A001


def resize_nearest(x, scales=None, sizes=None, mode="asymmetric"):
    return K.resize_nearest(x, scales, sizes, mode)


def reduce(x, axes, keepdims, kind):
    return {"sum": K.reduce_sum, "mean": K.reduce_mean, "max": K.reduce_max, "l2": K.reduce_l2}[kind](x, axes, keepdims)
