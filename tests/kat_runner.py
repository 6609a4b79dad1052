"""Replays the reference's own known-answer tests (tests/golden/reference_kats.json)
against a backend namespace (oracle.reference_api or lele_b200.kernels)."""
import json
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")
KATS = json.load(open(_PATH))


def kat_id(k):
    return f"{k['op']}@{k['cite']}"


def _check(got, k):
    """`checks` = what the reference test asserts when it does not compare every element: shape, single elements (flat index),
    a value range, a constant, a positive sum."""
    c = k.get("checks") or {}
    got = np.asarray(got, dtype=np.float32)
    if "shape" in c:
        assert list(got.shape) == c["shape"], (got.shape, c["shape"])
    flat = got.reshape(-1)
    for idx, v in (c.get("at") or {}).items():
        assert abs(float(flat[int(idx)]) - v) <= k["tol"], (idx, flat[int(idx)], v)
    if "range" in c:
        assert flat.min() >= c["range"][0] and flat.max() <= c["range"][1]
    if "all" in c:
        np.testing.assert_allclose(flat, c["all"], atol=k["tol"], rtol=0)
    if c.get("sum_positive"):
        assert float(flat.sum()) > 0


def run_kat(be, k):
    op, i, a = k["op"], k["inputs"], k["attrs"]
    f = lambda x: np.asarray(x, dtype=np.float32)
    if "x_zeros" in i: i = dict(i, x=np.zeros(i["x_zeros"], np.float32))
    if "x_ones" in i: i = dict(i, x=np.ones(i["x_ones"], np.float32))
    if "w_ones" in i: i = dict(i, w=np.ones(i["w_ones"], np.float32))
    if op == "matmul": got = be.matmul(f(i["a"]), f(i["b"]))
    elif op == "layer_norm": got = be.layer_norm(f(i["x"]), f(i["gamma"]), f(i["beta"]), -1, a["eps"])
    elif op == "softmax": got = be.softmax(f(i["x"]), -1)
    elif op == "mat_mul_integer":
        got = be.mat_mul_integer(f(i["a"]), f(i["b"]), a.get("a_zp", 0.0), a.get("b_zp", 0.0),
                                 None if "scale" not in a else f(a["scale"]), None, False)
    elif op == "concat": got = be.concat([f(x) for x in i["xs"]], a["axis"])
    elif op == "where": got = be.where(f(i["cond"]), f(i["x"]), f(i["y"]))
    elif op == "expand": got = be.expand(f(i["x"]), a["shape"])
    elif op == "split":
        got = be.split(f(i["x"]), a["axis"], a["splits"])
        for g, e in zip(got, k["expect"]):
            np.testing.assert_array_equal(np.asarray(g), f(e))
        return
    elif op == "transpose": got = be.transpose(f(i["x"]), a["perm"])
    elif op == "add": got = be.add(f(i["a"]), f(i["b"]))
    elif op == "mul": got = be.mul(f(i["a"]), f(i["b"]))
    elif op == "relu": got = be.relu(f(i["x"]))
    elif op == "gather": got = be.gather(f(i["x"]), f(i["idx"]), a["axis"])
    elif op == "gemm":
        got = be.gemm(f(i["a"]), f(i["b"]), None if "c" not in i else f(i["c"]), 1.0, 1.0,
                      a.get("trans_a", False), a.get("trans_b", False))
    elif op == "matmul_fused_add": got = be.matmul_fused_add(f(i["a"]), f(i["b"]), f(i["bias"]))
    elif op == "hann_window": got = be.hann_window(i["n"])
    elif op == "rfft":
        re, im = be.rfft(f(i["x"]))
        np.testing.assert_allclose(re, f(k["expect"]["re"]), atol=k["tol"], rtol=0)
        np.testing.assert_allclose(im, f(k["expect"]["im"]), atol=k["tol"], rtol=0)
        return
    elif op == "hz_to_mel": got = be.hz_to_mel(i["hz"])
    elif op == "conv_transpose":
        bias = np.zeros(i["bias_zeros"], np.float32) if "bias_zeros" in i else None
        got = be.conv_transpose(f(i["x"]), f(i["w"]), bias, (1, 1), a["pads"], a["strides"])
    elif op == "max_pool2d": got = be.max_pool2d(f(i["x"]), a["kernel"], a["pads"], a["strides"], (1, 1), False)
    elif op in ("sub", "div", "pow", "mod_f32"): got = getattr(be, op)(f(i["a"]), f(i["b"]))
    elif op in ("sqrt", "log", "exp", "tanh", "neg", "sigmoid", "gelu", "reciprocal"): got = getattr(be, op)(f(i["x"]))
    elif op == "clip": got = be.clip(f(i["x"]), a["lo"], a["hi"])
    elif op == "reduce": got = be.reduce(f(i["x"]), a["axes"], a["keepdims"], a["kind"])
    elif op == "reshape": got = be.reshape(f(i["x"]), a["shape"])
    elif op == "flatten": got = be.flatten(f(i["x"]), a["axis"])
    elif op == "pad": got = be.pad(f(i["x"]), a["pads"], a["value"], a["mode"])
    elif op == "conv1d": got = be.conv1d(f(i["x"]), f(i["w"]), None, a["dilations"], a["group"], a["pads"], a["strides"], False)
    elif op == "resize_nearest": got = be.resize_nearest(f(i["x"]), a.get("scales"), a.get("sizes"), a["mode"])
    else: raise KeyError(op)
    _check(got, k)
    if k["expect"] is None:
        return
    exp = f(k["expect"])
    got = np.asarray(got, dtype=np.float32)
    assert got.shape == exp.shape, (got.shape, exp.shape)
    if k["tol"] == 0:
        np.testing.assert_array_equal(got, exp)
    else:
        np.testing.assert_allclose(got, exp, atol=k["tol"], rtol=0)
