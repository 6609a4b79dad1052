"""Writes tests/golden/reference_kats.json.

The reference (miuda-ai/lele) is Rust and cannot be compiled or imported in this
container, so its golden vectors cannot be *generated* here.  What it does ship is a set
of known-answer tests with literal inputs and expected outputs.  This script is the
transcription of those literals (each entry cites the reference file:line it was copied
from, relative to /root/reference) into one JSON fixture that both the oracle tests
(-m "not gpu") and the CUDA parity tests (-m gpu) replay.

Run:  python tests/golden/make_reference_kats.py
"""
import json
import os

K = []


def kat(op, cite, tol, inputs, expect, **attrs):
    K.append(dict(op=op, cite=cite, tol=tol, inputs=inputs, expect=expect, attrs=attrs))


# ---- tests/verify_operators.rs ----
kat("matmul", "tests/verify_operators.rs:6-31", 1e-5,
    dict(a=[[1, 2, 3], [4, 5, 6]], b=[[7, 8], [9, 10], [11, 12]]), [[58, 64], [139, 154]])
kat("layer_norm", "tests/verify_operators.rs:34-53", 1e-4,
    dict(x=[[1.0, 2.0, 3.0]], gamma=[1, 1, 1], beta=[0, 0, 0]), [[-1.2247356, 0.0, 1.2247356]], eps=1e-5)
kat("softmax", "tests/verify_operators.rs:56-70", 1e-5,
    dict(x=[[1.0, 2.0, 3.0]]), [[0.09003057, 0.24472847, 0.66524096]])
kat("mat_mul_integer", "tests/verify_operators.rs:73-107", 1e-5,
    dict(a=[[10.0, 20.0]], b=[[1, 2], [3, 4]]), [[35.0, 200.0]], scale=[0.5, 2.0])
# ---- tests/kernel_accuracy.rs ----
kat("softmax", "tests/kernel_accuracy.rs:27-49", 1e-6,
    dict(x=[[1, 2, 3, 4], [5, 6, 7, 8]]),
    [[0.0320586, 0.08714432, 0.23688284, 0.6439143], [0.0320586, 0.08714432, 0.23688284, 0.6439143]])
kat("mat_mul_integer", "tests/kernel_accuracy.rs:52-97", 1e-5,
    dict(a=[[10, 20, 30], [40, 50, 60]], b=[[1, 2], [3, 4], [5, 6]]), [[130, 175], [310, 445]],
    a_zp=5.0, b_zp=1.0)
kat("matmul", "tests/kernel_accuracy.rs:135-151", 1e-5,
    dict(a=[[1, 2, 3], [4, 5, 6]], b=[[1, 2], [3, 4], [5, 6]]), [[22, 28], [49, 64]])
kat("mat_mul_integer", "tests/kernel_accuracy.rs:250-272", 1e-5,
    dict(a=[[1, 2, 3], [4, 5, 6]], b=[[7, 8], [9, 10], [11, 12]]), [[58, 64], [139, 154]],
    a_zp=0.0, b_zp=0.0)
kat("concat", "tests/kernel_accuracy.rs:153-169", 0,
    dict(xs=[[[1, 2], [3, 4]], [[5, 6]]]), [[1, 2], [3, 4], [5, 6]], axis=0)
kat("where", "tests/kernel_accuracy.rs:171-189", 0,
    dict(cond=[[1, 0], [0, 1]], x=[[1, 2], [3, 4]], y=[[5, 6], [7, 8]]), [[1, 6], [7, 4]])
kat("expand", "tests/kernel_accuracy.rs:191-205", 0,
    dict(x=[[1], [2], [3]]), [[1, 1, 1, 1], [2, 2, 2, 2], [3, 3, 3, 3]], shape=[3, 4])
kat("split", "tests/kernel_accuracy.rs:274-288", 0,
    dict(x=[[1, 2, 3, 4, 5, 6]]), [[[1, 2]], [[3, 4]], [[5, 6]]], axis=1, splits=[2, 2, 2])
kat("transpose", "tests/kernel_accuracy.rs:290-306", 0,
    dict(x=[[1, 2, 3], [4, 5, 6]]), [[1, 4], [2, 5], [3, 6]], perm=[1, 0])
kat("add", "tests/kernel_accuracy.rs:308-323", 1e-6,
    dict(a=[[1, 2], [3, 4]], b=[[5, 6], [7, 8]]), [[6, 8], [10, 12]])
kat("mul", "tests/kernel_accuracy.rs:325-340", 1e-6,
    dict(a=[[1, 2], [3, 4]], b=[[2, 3], [4, 5]]), [[2, 6], [12, 20]])
kat("relu", "tests/kernel_accuracy.rs:342-354", 1e-6,
    dict(x=[[-2, -1, 0], [1, 2, 3]]), [[0, 0, 0], [1, 2, 3]])
kat("gather", "tests/kernel_accuracy.rs:356-374", 0,
    dict(x=[[1, 2, 3], [4, 5, 6], [7, 8, 9]], idx=[0, 2]), [[1, 2, 3], [7, 8, 9]], axis=0)
# ---- tests/regression_kernels.rs ----
kat("gemm", "tests/regression_kernels.rs:898-905", 1e-5,
    dict(a=[[1, 2], [3, 4]], b=[[1, 3], [2, 4]]), [[7, 10], [15, 22]], trans_b=True)
kat("gemm", "tests/regression_kernels.rs:908-916", 1e-5,
    dict(a=[[1, 2]], b=[[1, 2], [3, 4]], c=[0.5, -0.5]), [[7.5, 9.5]])
kat("matmul_fused_add", "tests/regression_kernels.rs:919-926", 1e-4,
    dict(a=[[1, 2, 3], [4, 5, 6]], b=[[1, 2], [3, 4], [5, 6]], bias=[0.1, 0.2]), [[22.1, 28.2], [49.1, 64.2]])
# ---- tests/verify_features.rs + src/kernels/fft.rs tests ----
kat("hann_window", "tests/verify_features.rs:6-24", 1e-6, dict(n=4), [0.0, 0.75, 0.75, 0.0])
kat("rfft", "tests/verify_features.rs:27-38", 1e-6, dict(x=[1, 0, 0, 0]), dict(re=[1, 1, 1], im=[0, 0, 0]))
kat("rfft", "tests/verify_features.rs:41-52", 1e-6, dict(x=[1, 1, 1, 1]), dict(re=[4, 0, 0], im=[0, 0, 0]))
kat("rfft", "src/kernels/fft.rs:273-298", 1e-3, dict(x=[1, 2, 3, 4, 5, 6, 7, 8]),
    dict(re=[36.0, -4.0, -4.0, -4.0, -4.0], im=[0.0, 9.656854, 4.0, 1.656854, 0.0]))
kat("hz_to_mel", "tests/verify_features.rs:55-63", 1e-3, dict(hz=700.0), 781.1728)
# ---- src/kernels/conv2d.rs in-module tests ----
kat("conv_transpose", "src/kernels/conv2d.rs:3476-3495", 1e-6,
    dict(x=[[[[1.0]]]], w=[[[[2.0]]]]), [[[[2.0]]]], strides=[1, 1], pads=[0, 0, 0, 0])
kat("max_pool2d", "src/kernels/conv2d.rs:3630-3652", 1e-6,
    dict(x=[[[[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]]]]), [[[[5, 7], [13, 15]]]],
    kernel=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0])
kat("max_pool2d", "src/kernels/conv2d.rs:3655-3677", 1e-6,
    dict(x=[[[[0, 1, 2], [3, 4, 5], [6, 7, 8]]]]), [[[[4, 5], [7, 8]]]],
    kernel=[2, 2], strides=[1, 1], pads=[0, 0, 0, 0])

# ---- tests/regression_kernels.rs: element-wise, reductions, shape (literal inputs; where the reference computes its expectation with
# Rust's f32 libm call, the same function is evaluated here in float64 and rounded to f32 -- the tests' tolerances are 1e-4..1e-6) ----
import math

import numpy as np


def _f32(vals):
    return [float(np.float32(v)) for v in vals]


def kat2(op, cite, tol, inputs, expect=None, checks=None, **attrs):
    K.append(dict(op=op, cite=cite, tol=tol, inputs=inputs, expect=expect, attrs=attrs, checks=checks or {}))


kat2("sub", "tests/regression_kernels.rs:744-750", 1e-6, dict(a=[5, 3, 1, -2], b=[1, 2, 3, 4]), [4, 1, -2, -6])
kat2("div", "tests/regression_kernels.rs:753-759", 1e-6, dict(a=[10, 6, 0, -8], b=[2, 3, 5, 4]), [5, 2, 0, -2])
kat2("clip", "tests/regression_kernels.rs:762-769", 1e-6, dict(x=[-5, -1, 0, 0.5, 3, 10]), [-1, -1, 0, 0.5, 3, 3], lo=-1.0, hi=3.0)
_x = [i * 0.25 for i in range(21)]
kat2("sqrt", "tests/regression_kernels.rs:772-778", 1e-5, dict(x=_x), _f32(math.sqrt(v) for v in _x))
_x = [i * 0.5 for i in range(1, 21)]
kat2("log", "tests/regression_kernels.rs:781-787", 1e-5, dict(x=_x), _f32(math.log(v) for v in _x))
_x = [i * 0.5 for i in range(-10, 11)]
kat2("exp", "tests/regression_kernels.rs:790-796", 1e-4, dict(x=_x), _f32(math.exp(v) for v in _x))
_x = [i * 0.5 for i in range(11)]
kat2("pow", "tests/regression_kernels.rs:799-806", 1e-4, dict(a=_x, b=[2.0] * 11), _f32(v ** 2.0 for v in _x))
kat2("reduce", "tests/regression_kernels.rs:809-814", 0, dict(x=[[1, 2, 3], [4, 5, 6]]), [6, 15], axes=[1], keepdims=False, kind="sum")
kat2("reduce", "tests/regression_kernels.rs:817-823", 0, dict(x=[[1, 2, 3], [4, 5, 6]]), [[6], [15]], axes=[1], keepdims=True, kind="sum")
kat2("reduce", "tests/regression_kernels.rs:826-831", 1e-6, dict(x=[[2, 4, 6], [8, 10, 12]]), [4, 10], axes=[1], keepdims=False, kind="mean")
kat2("reduce", "tests/regression_kernels.rs:834-839", 1e-6, dict(x=[[3, 1, 4], [1, 5, 9]]), [4, 9], axes=[1], keepdims=False, kind="max")
kat2("reduce", "tests/regression_kernels.rs:842-847", 1e-4, dict(x=[3, 4]), 5.0, axes=[0], keepdims=False, kind="l2")
_x = [float(i) for i in range(-5, 6)]
kat2("tanh", "tests/regression_kernels.rs:850-856", 1e-5, dict(x=_x), _f32(math.tanh(v) for v in _x))
kat2("neg", "tests/regression_kernels.rs:859-864", 1e-6, dict(x=[1, -2, 0, 3.5]), [-1, 2, 0, -3.5])
_x = [i * 0.5 for i in range(-10, 11)]
kat2("sigmoid", "tests/regression_kernels.rs:867-873", 1e-5, dict(x=_x), _f32(1.0 / (1.0 + math.exp(-v)) for v in _x))
_x = [i * 0.5 for i in range(-5, 6)]
kat2("gelu", "tests/regression_kernels.rs:876-882", 1e-5, dict(x=_x), _f32(v * 0.5 * (1.0 + math.erf(v / math.sqrt(2.0))) for v in _x))
_x = [1, 2, 4, -2, 0.5]
kat2("reciprocal", "tests/regression_kernels.rs:885-891", 1e-5, dict(x=_x), _f32(1.0 / v for v in _x))
kat2("reshape", "tests/regression_kernels.rs:934-938", 0, dict(x=[[1, 2, 3], [4, 5, 6]]), [[1, 2], [3, 4], [5, 6]], shape=[-1, 2])
kat2("reshape", "tests/regression_kernels.rs:941-945", 0, dict(x=np.arange(1, 13).reshape(2, 2, 3).tolist()),
     np.arange(1, 13).reshape(2, 3, 2).tolist(), shape=[2, -1, 2])
kat2("transpose", "tests/regression_kernels.rs:948-954", 0, dict(x=[[[1, 2, 3], [4, 5, 6]]]), [[[1, 4], [2, 5], [3, 6]]], perm=[0, 2, 1])
kat2("transpose", "tests/regression_kernels.rs:957-963", 0, dict(x=np.arange(24).reshape(1, 2, 3, 4).tolist()), None,
     dict(shape=[1, 4, 2, 3]), perm=[0, 3, 1, 2])
kat2("add", "tests/regression_kernels.rs:966-970", 1e-6, dict(a=[1, 2, 3], b=[10]), [11, 12, 13])


def _ref_maxpool(x, kh, kw, sh, sw, pt, pl, pb, pr):
    """tests/regression_kernels.rs:258-289: naive loops, padding positions never win (-inf)."""
    x = np.asarray(x, np.float32); n, c, ih, iw = x.shape
    oh = (ih + pt + pb - kh) // sh + 1; ow = (iw + pl + pr - kw) // sw + 1
    xp = np.full((n, c, ih + pt + pb, iw + pl + pr), -np.inf, np.float32); xp[:, :, pt:pt + ih, pl:pl + iw] = x
    out = np.empty((n, c, oh, ow), np.float32)
    for i in range(oh):
        for j in range(ow):
            out[:, :, i, j] = xp[:, :, i * sh:i * sh + kh, j * sw:j * sw + kw].max(axis=(2, 3))
    return out


def _pool_kat(cite, shape, gen, kernel, strides, pads):
    n = int(np.prod(shape))
    x = np.array([gen(i) for i in range(n)], np.float32).reshape(shape)
    kat2("max_pool2d", cite, 1e-6, dict(x=x.tolist()), _ref_maxpool(x, kernel[0], kernel[1], strides[0], strides[1], *pads).tolist(),
         kernel=kernel, strides=strides, pads=pads)


_pool_kat("tests/regression_kernels.rs:292-302", (1, 3, 8, 8), lambda i: np.float32(i) * np.float32(0.1), [2, 2], [2, 2], [0, 0, 0, 0])
_pool_kat("tests/regression_kernels.rs:305-315", (1, 4, 32, 10), lambda i: np.float32((i * 7 + 3) % 100) * np.float32(0.1), [8, 1], [8, 1], [0, 0, 0, 0])
_pool_kat("tests/regression_kernels.rs:318-328", (1, 2, 6, 6), lambda i: np.float32((i * 13 + 7) % 50) * np.float32(0.2), [3, 3], [1, 1], [1, 1, 1, 1])
_pool_kat("tests/regression_kernels.rs:331-341", (1, 2, 32, 300), lambda i: np.float32((i * 7 + 3) % 97) * np.float32(0.05), [2, 2], [2, 2], [0, 0, 0, 0])
kat2("max_pool2d", "tests/regression_kernels.rs:344-358", 1e-6,
     dict(x=[[[[-10, -5, -3, -1], [-8, -2, -6, -4], [-7, -9, -11, -12], [-13, -14, -15, -16]]]]), [[[[-2, -1], [-7, -11]]]],
     kernel=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0])
# pad: the reflect test asserts the shape, a positive sum and the untouched centre; the constant tests assert every element
kat2("pad", "tests/regression_kernels.rs:365-384", 0, dict(x=[[[[1, 2, 3], [4, 5, 6]]]]), None,
     dict(shape=[1, 1, 4, 5], at={"6": 1, "7": 2, "8": 3, "11": 4, "12": 5, "13": 6}, sum_positive=True), pads=[0, 0, 1, 1, 0, 0, 1, 1], value=0.0, mode="reflect")
kat2("pad", "tests/regression_kernels.rs:387-400", 1e-6, dict(x=[[[[1, 2, 3, 4, 5]]]]), [[[[0, 0, 1, 2, 3, 4, 5, 0, 0]]]],
     pads=[0, 0, 0, 2, 0, 0, 0, 2], value=0.0, mode="constant")
kat2("pad", "tests/regression_kernels.rs:403-420", 1e-6, dict(x=[[[[1, 2], [3, 4]]]]),
     [[[[99, 99, 99, 99], [99, 1, 2, 99], [99, 3, 4, 99], [99, 99, 99, 99]]]], pads=[0, 0, 1, 1, 0, 0, 1, 1], value=99.0, mode="constant")
# ---- in-module tests of src/kernels/{math,shape,conv1d,manipulation,gemm,conv2d}.rs ----
kat2("expand", "src/kernels/math.rs:2446-2454", 0, dict(x=[[[1, 2, 3]]]), [[[1, 2, 3], [1, 2, 3]]], shape=[1, 2, 3])
kat2("add", "src/kernels/math.rs:2456-2465", 0, dict(a=[[10, 20, 30]], b=[[1], [2]]), [[11, 21, 31], [12, 22, 32]])
kat2("mod_f32", "src/kernels/math.rs:2477-2483", 0, dict(a=[9, 13, 7, 0, 25], b=[5, 5, 3, 2, 10]), [4, 3, 1, 0, 5])
kat2("mod_f32", "src/kernels/math.rs:2486-2493", 0, dict(a=[3, 9, 13, 18, 23], b=5.0), [3, 4, 3, 3, 3])
kat2("mod_f32", "src/kernels/math.rs:2496-2503", 0, dict(a=[[3, 9], [13, 18]], b=5.0), [[3, 4], [3, 3]])
kat2("mod_f32", "src/kernels/math.rs:2506-2512", 0, dict(a=[5, 10], b=0.0), [0, 0])
kat2("reshape", "src/kernels/shape.rs:194-203", 0, dict(x=[[1, 2], [3, 4]]), [1, 2, 3, 4], shape=[4])
kat2("reshape", "src/kernels/shape.rs:194-203", 0, dict(x=[[1, 2], [3, 4]]), [[1, 2, 3, 4]], shape=[1, -1])
kat2("flatten", "src/kernels/shape.rs:205-212", 0, dict(x=np.ones((2, 3, 4)).tolist()), None, dict(shape=[2, 12]), axis=1)
kat2("flatten", "src/kernels/shape.rs:205-212", 0, dict(x=np.ones((2, 3, 4)).tolist()), None, dict(shape=[6, 4]), axis=2)
kat2("conv1d", "src/kernels/conv1d.rs:1622-1631", 0, dict(x=np.ones((1, 2, 3)).tolist(), w=np.ones((2, 1, 1)).tolist()), np.ones((1, 2, 3)).tolist(),
     dilations=[1], group=2, pads=[0, 0], strides=[1])
kat2("conv1d", "src/kernels/conv1d.rs:1633-1643", 0, dict(x=[[[1, 2, 3]]], w=[[[1, 1]]]), [[[3, 5]]], dilations=[1], group=1, pads=[0, 0], strides=[1])
kat2("conv1d", "src/kernels/conv1d.rs:1645-1674", 0, dict(x=[[list(range(10))]], w=[[[1, 1, 1]]]), None,
     dict(shape=[1, 1, 10], at={"0": 1, "1": 3, "5": 15, "9": 17}), dilations=[1], group=1, pads=[1, 1], strides=[1])
kat2("concat", "src/kernels/manipulation.rs:1389-1399", 0, dict(xs=[[[1, 2], [3, 4]], [[5], [6]]]), [[1, 2, 5], [3, 4, 6]], axis=1)
kat2("matmul", "src/kernels/gemm.rs:788-800", 0.01, dict(a=[[1, 2, 3], [4, 5, 6]], b=[[1, 2], [3, 4], [5, 6]]), [[22, 28], [49, 64]])
kat2("matmul_fused_add", "src/kernels/gemm.rs:803-817", 0.01, dict(a=[[1, 2, 3], [4, 5, 6]], b=[[1, 2], [3, 4], [5, 6]], bias=[100, 200]),
     [[122, 228], [149, 264]])
kat2("conv_transpose", "src/kernels/conv2d.rs:3391-3419", 0, dict(x_zeros=[1, 64, 80, 80], w_ones=[64, 64, 2, 2], bias_zeros=64), None,
     dict(shape=[1, 64, 160, 160]), strides=[2, 2], pads=[0, 0, 0, 0])
kat2("conv_transpose", "src/kernels/conv2d.rs:3422-3446", 0, dict(x_zeros=[1, 32, 40, 40], w_ones=[32, 32, 3, 3], bias_zeros=32), None,
     dict(shape=[1, 32, 79, 79]), strides=[2, 2], pads=[1, 1, 1, 1])
kat2("conv_transpose", "src/kernels/conv2d.rs:3449-3473", 0, dict(x_zeros=[1, 16, 10, 10], w_ones=[16, 16, 3, 3], bias_zeros=16), None,
     dict(shape=[1, 16, 12, 12]), strides=[1, 1], pads=[0, 0, 0, 0])
kat2("resize_nearest", "src/kernels/conv2d.rs:3500-3517", 1e-6, dict(x=[[[[1, 2], [3, 4]]]]), [[[[1, 2], [3, 4]]]], scales=[1.0, 1.0, 1.0, 1.0], mode="asymmetric")
kat2("resize_nearest", "src/kernels/conv2d.rs:3520-3548", 1e-6, dict(x=[[[[1, 2], [3, 4]]]]),
     [[[[1, 1, 2, 2], [1, 1, 2, 2], [3, 3, 4, 4], [3, 3, 4, 4]]]], scales=[1.0, 1.0, 2.0, 2.0], mode="asymmetric")
kat2("resize_nearest", "src/kernels/conv2d.rs:3551-3560", 0, dict(x=[[[[1, 2], [3, 4]]]]), None, dict(shape=[1, 1, 3, 3]), sizes=[1, 1, 3, 3], mode="asymmetric")
kat2("resize_nearest", "src/kernels/conv2d.rs:3562-3577", 0, dict(x=np.arange(8).reshape(1, 2, 2, 2).tolist()), None, dict(shape=[1, 2, 4, 4]),
     scales=[1.0, 1.0, 2.0, 2.0], mode="asymmetric")
kat2("resize_nearest", "src/kernels/conv2d.rs:3579-3598", 0, dict(x=[[[[1, 2], [3, 4]]]]), None, dict(shape=[1, 1, 4, 4], range=[1.0, 4.0]),
     scales=[1.0, 1.0, 2.0, 2.0], mode="half_pixel")
kat2("resize_nearest", "src/kernels/conv2d.rs:3600-3618", 1e-6, dict(x=[[[[42.0]]]]), np.full((1, 1, 100, 100), 42.0).tolist(), sizes=[1, 1, 100, 100], mode="asymmetric")
kat2("max_pool2d", "src/kernels/conv2d.rs:3680-3703", 1e-6, dict(x=[[[[1, 2], [3, 4]]]]), None,
     dict(shape=[1, 1, 3, 3], at={"0": 1, "4": 4, "8": 4}), kernel=[2, 2], strides=[1, 1], pads=[1, 1, 1, 1])
kat2("max_pool2d", "src/kernels/conv2d.rs:3705-3722", 0, dict(x=np.arange(32).reshape(1, 2, 4, 4).tolist()), None, dict(shape=[1, 2, 2, 2]),
     kernel=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0])
kat2("max_pool2d", "src/kernels/conv2d.rs:3724-3749", 1e-6, dict(x_ones=[1, 32, 80, 80]), None, dict(shape=[1, 32, 40, 40], all=1.0),
     kernel=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0])

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(K, f, indent=1)
    print(f"wrote {len(K)} KATs to {out}")
