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

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(K, f, indent=1)
    print(f"wrote {len(K)} KATs to {out}")
