"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for
the hot path, plus independent cross-checks (numpy / torch CPU)."""
import numpy as np
import pytest

from oracle import reference_api as R
from tests.kat_runner import KATS, kat_id, run_kat


@pytest.mark.parametrize("k", KATS, ids=kat_id)
def test_reference_kat(k):
    run_kat(R, k)


def test_layer_norm_mean_std():  # tests/kernel_accuracy.rs:100-132
    out = R.layer_norm(np.array([[1, 2, 3], [4, 5, 6]], np.float32), np.ones(3), np.zeros(3), -1, 1e-5)
    row = out[0]
    assert abs(row.mean()) < 1e-5
    var = 2.0 / 3.0
    assert abs(row.std() - np.sqrt(var / (var + 1e-5))) < 1e-5


def test_dql_dequant_error():  # tests/kernel_accuracy.rs:205-247
    x = np.arange(1, 9, dtype=np.float32).reshape(2, 4)
    q, s, z = R.dynamic_quantize_linear(x)
    assert np.abs((q - z) * s - x).max() < s + 0.1
    assert z == 0.0 and abs(s - 8.0 / 255.0) < 1e-7


def test_silu_erf_vs_libm():  # tests/kernel_accuracy.rs:377-408 (17/19 elems hit the SIMD tail)
    x = np.arange(-8, 9, dtype=np.float32) * 0.5
    np.testing.assert_allclose(R.silu(x), x / (1 + np.exp(-x.astype(np.float64))), atol=1e-5)
    import math
    x = np.arange(-9, 10, dtype=np.float32) * 0.5
    np.testing.assert_allclose(R.erf(x), [math.erf(v) for v in x], atol=2e-6)


def _lcg_inputs(n, state):  # src/kernels/fft.rs:302-307
    out = []
    for _ in range(n):
        state = (state * 1103515245 + 12345) & 0xFFFFFFFF
        out.append(np.float32(state) / np.float32(0xFFFFFFFF) * 2.0 - 1.0)
    return np.array(out, np.float32), state


def test_rfft_vs_numpy_lcg():  # src/kernels/fft.rs:302-351 sizes and seed
    st = 12345
    for log_n in range(3, 11):
        x, st = _lcg_inputs(1 << log_n, st)
        re, im = R.rfft(x)
        ref = np.fft.rfft(x.astype(np.float64))
        np.testing.assert_allclose(re, ref.real, atol=1e-4)
        np.testing.assert_allclose(im, ref.imag, atol=1e-4)


def test_fft_parseval_linearity():  # tests/regression_kernels.rs:524-596
    rng = np.random.default_rng(0)
    x = rng.standard_normal(256).astype(np.float32)
    y = rng.standard_normal(256).astype(np.float32)
    rx, ix = R.rfft(x); ry, iy = R.rfft(y); rs, is_ = R.rfft(x + y)
    np.testing.assert_allclose(rs, rx + ry, atol=1e-3)
    np.testing.assert_allclose(is_, ix + iy, atol=1e-3)
    full = np.concatenate([rx**2 + ix**2, (rx**2 + ix**2)[1:-1]])
    assert abs(full.sum() / 256 - (x.astype(np.float64) ** 2).sum()) < 1e-4 * (x**2).sum() + 1e-2


def test_mel_filterbank_shape_and_cmvn():  # verify_features.rs:66-80, features/cmvn.rs:98-110
    w = R.mel_filterbank(16000.0, 512, 10, 0.0)
    assert w.shape == (10, 257) and w.sum() > 0
    out = R.cmvn(np.array([[1, 10], [2, 20], [3, 30]], np.float32))
    assert abs(out[:, 0].mean()) < 1e-5 and out[0, 0] < 0 and abs(out[1, 0]) < 1e-5 and out[2, 0] > 0


def test_stft_reference_properties():  # tests/regression_kernels.rs:426-517
    p = R.stft(np.ones(512, np.float32), 256, 64, 256, power=True)
    assert (p[:, 0] > 1000).all() and (p[:, 2:] < 1.0).all()
    sig = np.sin(np.arange(800, dtype=np.float32) * np.float32(0.01))
    c = R.stft(sig, 256, 128, 256); p = R.stft(sig, 256, 128, 256, power=True)
    np.testing.assert_allclose(c[..., 0] ** 2 + c[..., 1] ** 2, p, atol=1e-3)
    t = np.arange(1600, dtype=np.float32)
    s = np.sin(np.float32(2 * np.pi) * 1000.0 * t / 16000.0).astype(np.float32)
    p = R.stft(s, 256, 160, 256, power=True)
    mid = p[p.shape[0] // 2]
    other = np.concatenate([mid[:5], mid[-4:]])
    assert mid[16] > 5 * other.max()
    # shape rules of math.rs:2313-2316, :2362-2367, :2381-2384: empty signal -> empty tensor; shorter than a window -> one frame;
    # rank >= 2 input -> leading batch dim
    assert R.stft(np.zeros(0, np.float32), 256, 64, 256).shape == (0, 0, 129, 2)
    assert R.stft(np.zeros(0, np.float32), 256, 64, 256, power=True).shape == (0, 0, 129)
    assert R.stft(sig[:100], 256, 64, 256).shape == (1, 129, 2)
    assert R.stft(sig[None, :], 256, 128, 256).shape == (1, 5, 129, 2) and R.stft(sig[None, :], 256, 128, 256, power=True).shape == (1, 5, 129)
    np.testing.assert_array_equal(R.stft(sig[None, :], 256, 128, 256)[0], c)


def _ref_gru(x, w, r, bw, br, hs):  # tests/regression_kernels.rs:602-633 (f64 restatement)
    h = np.zeros(hs); ys = []
    sig = lambda v: 1 / (1 + np.exp(-v))
    for xt in x:
        wc = w @ xt; rc = r @ h
        z = sig(wc[:hs] + rc[:hs] + bw[:hs] + br[:hs])
        rg = sig(wc[hs:2 * hs] + rc[hs:2 * hs] + bw[hs:2 * hs] + br[hs:2 * hs])
        hg = np.tanh(wc[2 * hs:] + bw[2 * hs:] + rg * (rc[2 * hs:] + br[2 * hs:]))
        h = (1 - z) * hg + z * h
        ys.append(h.copy())
    return np.array(ys), h


GRU_CASES = [  # tests/regression_kernels.rs:636-737
    dict(sl=1, is_=4, hs=8, x=[0.1, 0.2, -0.1, 0.3], w=lambda i: i * 0.01 - 0.1, r=lambda i: i * 0.02 - 0.2, b=lambda i: i * 0.005 - 0.05, tol=1e-4),
    dict(sl=5, is_=3, hs=6, x=lambda i: ((i * 7 + 3) % 20) * 0.1 - 0.5, w=lambda i: i * 0.03 - 0.2, r=lambda i: i * 0.01 - 0.1, b=lambda i: 0.1 + 0 * i, tol=1e-3),
    dict(sl=3, is_=4, hs=8, x=lambda i: i * 0.15 - 0.3, w=lambda i: i * 0.01, r=lambda i: i * 0.02 - 0.1, b=lambda i: 0.05 + 0 * i, tol=1e-3),
    dict(sl=2, is_=3, hs=4, x=[0.5, -0.3, 0.1, -0.2, 0.4, 0.6], w=lambda i: i * 0.05, r=lambda i: i * 0.03, b=None, tol=1e-4),
]


def gru_case(c):
    sl, is_, hs = c["sl"], c["is_"], c["hs"]
    gen = lambda f, n: np.array(f, np.float32) if isinstance(f, list) else f(np.arange(n, dtype=np.float32)).astype(np.float32)
    x = gen(c["x"], sl * is_).reshape(sl, 1, is_)
    w = gen(c["w"], 3 * hs * is_).reshape(1, 3 * hs, is_)
    r = gen(c["r"], 3 * hs * hs).reshape(1, 3 * hs, hs)
    b = None if c["b"] is None else gen(c["b"], 6 * hs).reshape(1, 6 * hs)
    return x, w, r, b


@pytest.mark.parametrize("c", GRU_CASES)
def test_gru_vs_ref_gru_step(c):
    x, w, r, b = gru_case(c)
    hs = c["hs"]
    y, h = R.gru(x, w, r, b)
    bz = np.zeros(6 * hs) if b is None else b.reshape(-1).astype(np.float64)
    yr, hr = _ref_gru(x[:, 0].astype(np.float64), w[0].astype(np.float64), r[0].astype(np.float64), bz[:3 * hs], bz[3 * hs:], hs)
    np.testing.assert_allclose(y.reshape(-1, hs), yr, atol=c["tol"])
    np.testing.assert_allclose(h.reshape(-1), hr, atol=c["tol"])


def test_lstm_shape_finite():  # tests/regression_kernels.rs:972-997
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 1, 8)).astype(np.float32)
    w = (0.1 * rng.standard_normal((1, 64, 8))).astype(np.float32)
    r = (0.1 * rng.standard_normal((1, 64, 16))).astype(np.float32)
    y, h, c = R.lstm(x, w, r, None)
    assert y.shape == (4, 1, 1, 16) and np.isfinite(y).all() and np.isfinite(c).all()


def test_conv_vs_torch():  # tests/regression_kernels.rs:76-252 shapes; torch CPU as independent check
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    rng = np.random.default_rng(2)
    for (ic, oc, k, s, p, g) in [(3, 8, 3, 1, 1, 1), (4, 8, 3, 2, 1, 1), (4, 4, 3, 1, 1, 4), (8, 16, 1, 1, 0, 1)]:
        x = rng.standard_normal((2, ic, 9, 11)).astype(np.float32)
        w = rng.standard_normal((oc, ic // g, k, k)).astype(np.float32)
        b = rng.standard_normal(oc).astype(np.float32)
        for act in (0, 1, 2):
            got = R.conv2d(x, w, b, (1, 1), g, (p, p, p, p), (s, s), act)
            ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride=s, padding=p, groups=g)
            ref = ref if act == 0 else (F.relu(ref) if act == 1 else F.silu(ref))
            np.testing.assert_allclose(got, ref.numpy(), atol=1e-3)
    x = rng.standard_normal((1, 6, 5, 7)).astype(np.float32)
    w = rng.standard_normal((6, 4, 3, 3)).astype(np.float32)
    b = rng.standard_normal(4).astype(np.float32)
    got = R.conv_transpose(x, w, b, (1, 1), (1, 1, 1, 1), (2, 2))
    ref = F.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride=2, padding=1)
    np.testing.assert_allclose(got, ref.numpy(), atol=1e-3)
    assert R.conv_transpose(np.zeros((1, 8, 10, 10), np.float32), np.ones((8, 8, 2, 2), np.float32), None, (1, 1), (0, 0, 0, 0), (2, 2)).shape == (1, 8, 20, 20)  # conv2d.rs:3391
    x = rng.standard_normal((2, 6, 40)).astype(np.float32)
    w = rng.standard_normal((6, 1, 11)).astype(np.float32)
    got = R.conv1d(x, w, None, (1,), 6, (5, 5), (1,))
    ref = F.conv1d(torch.from_numpy(x), torch.from_numpy(w), padding=5, groups=6)
    np.testing.assert_allclose(got, ref.numpy(), atol=1e-4)


def test_int8_linear_matches_unfused_and_wasm_bench_shapes():
    """fused_quantized_linear == DQL -> MatMulInteger -> scale -> bias (patterns.rs:122-190),
    on the shapes src/bin/wasm_bench.rs:293-317 benchmarks (tol k*0.01 there; exact here)."""
    rng = np.random.default_rng(3)
    for (m, k, n) in [(93, 512, 1536), (93, 512, 512), (17, 560, 512), (5, 2048, 512), (3, 20, 9)]:
        x = ((np.arange(m * k, dtype=np.float32) * 0.01) % 2.0 - 1.0).reshape(1, m, k)  # wasm_bench.rs:765
        w = rng.integers(0, 256, (k, n), dtype=np.uint8)
        ws = rng.uniform(0.002, 0.006, n).astype(np.float32)
        b = rng.standard_normal(n).astype(np.float32)
        fused = R.fused_quantized_linear(x, w, ws, 128, b, True)
        q, s, z = R.dynamic_quantize_linear(x)
        unf = R.mat_mul_integer(q, w.astype(np.float32), z, 128.0, (s * ws).astype(np.float32), b, True)
        if k % 8 == 0:
            np.testing.assert_array_equal(fused, unf)
        else:  # row tail takes the scalar rounding in the fused kernel (avx/quantization.rs:208)
            np.testing.assert_allclose(fused, unf, atol=float(s * ws.max() * 255 * 2))
        # integer core exactness vs int64 numpy
        acc = (q[0].astype(np.int64) - int(z)) @ (w.astype(np.int64) - 128)
        ref = np.maximum(acc.astype(np.float32) * (s * ws) + b, 0)
        np.testing.assert_array_equal(unf[0], ref.astype(np.float32))


def test_indexing_ops_numpy_semantics():
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    np.testing.assert_array_equal(R.slice(x, [1], [2**63 - 1], [2], []), x[:, :, 1:])
    np.testing.assert_array_equal(R.slice(x, [-1], [-(2**63)], [1], [-1]), x[:, ::-1])
    np.testing.assert_array_equal(R.slice(x, [0], [1], [0], []), x[:1])
    np.testing.assert_array_equal(R.pad(x[0], [1, 2], 0.0, "constant").shape, (3, 7))  # short pads -> trailing dims
    np.testing.assert_array_equal(R.pad(x[0], [0, 1, 0, 1], mode="reflect"), np.pad(x[0], ((0, 0), (1, 1)), mode="reflect"))
    v, i = R.topk(np.array([[1, 3, 3, 2]], np.float32), 2)
    np.testing.assert_array_equal(i, [[1, 2]])  # stable: lower index first on ties (conv2d.rs:1385)
    np.testing.assert_array_equal(R.gather(x, np.array([-1], np.float32), 1), x[:, [2]])
    r = R.resize_nearest(np.array([[[[1, 2], [3, 4]]]], np.float32), scales=[1, 1, 2, 2])
    np.testing.assert_array_equal(r[0, 0], [[1, 1, 2, 2], [1, 1, 2, 2], [3, 3, 4, 4], [3, 3, 4, 4]])  # conv2d.rs:3520


def test_frontend_shapes_and_edges():
    """SenseVoiceFrontend::compute (pipeline.rs:67): zh.wav-sized and 16 s clips."""
    from lele_b200.sensevoice_weights import synth_pcm
    assert R.frontend(np.zeros(399, np.float32)).shape == (0, 560)       # shorter than a frame -> empty
    mel, out = R.frontend(synth_pcm(0, 89472), want_mel=True)            # fixtures/zh.wav length
    assert mel.shape == (557, 80) and out.shape == (93, 560)
    np.testing.assert_array_equal(out[0, :80], mel[0]); np.testing.assert_array_equal(out[0, 240:320], mel[0])
    np.testing.assert_array_equal(out[1, :80], mel[3]); np.testing.assert_array_equal(out[92, 480:], mel[555])
    assert np.isfinite(out).all() and out.min() >= np.log(np.float32(1e-5)) - 1e-6
    # independent f64 check of one frame
    pcm = synth_pcm(1, 4000).astype(np.float64) * 32768.0
    fr = pcm[160:560].copy(); fr -= fr.mean(); fr[1:] -= 0.97 * fr[:-1]
    fr *= 0.5 * (1 - np.cos(2 * np.pi * np.arange(400) / 399))
    pw = np.abs(np.fft.rfft(fr, 512)) ** 2
    wts = R.mel_filterbank(16000.0, 512, 80, 20.0).astype(np.float64)
    ref = np.log(np.maximum(wts @ pw, 1e-5))
    mel = R.frontend(synth_pcm(1, 4000), want_mel=True)[0]
    np.testing.assert_allclose(mel[1], ref, rtol=2e-4, atol=2e-4)


def test_reduce_axes_edge_semantics():  # math.rs:1611-1650 (same prologue in reduce_mean / reduce_max / reduce_l2)
    """Axes are resolved, sorted and de-duplicated; an empty list reduces NOTHING (the mask stays all-false): sum and mean return
    0 + x, max returns x, l2 returns |x| -- unlike ONNX, where empty axes mean "all"."""
    x = np.array([[1.0, -2.0], [3.0, -0.0]], np.float32)
    for kind, want in (("sum", x + 0.0), ("mean", x + 0.0), ("max", x), ("l2", np.abs(x))):
        for keep in (True, False):
            got = R.reduce(x, [], keep, kind)
            assert got.shape == (2, 2)
            np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(R.reduce(x, [1, 1, -1], False, "sum"), [-1.0, 3.0])
    np.testing.assert_array_equal(R.reduce(x, [1, 0], False, "max"), 3.0)


def _ref_lstm(x, w, r, b, h0, c0):
    """ONNX LSTM, forward, no peepholes, gate order i, o, f, c (rnn.rs:67-230) in float64."""
    x, w, r = x.astype(np.float64), w[0].astype(np.float64), r[0].astype(np.float64)
    hs = r.shape[1]
    wb, rb = (b[0, :4 * hs].astype(np.float64), b[0, 4 * hs:].astype(np.float64)) if b is not None else (np.zeros(4 * hs), np.zeros(4 * hs))
    h = np.zeros(hs) if h0 is None else h0.reshape(-1).astype(np.float64)
    c = np.zeros(hs) if c0 is None else c0.reshape(-1).astype(np.float64)
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    ys = []
    for t in range(x.shape[0]):
        g = w @ x[t, 0] + r @ h + wb + rb
        i, o, f, cc = sig(g[:hs]), sig(g[hs:2 * hs]), sig(g[2 * hs:3 * hs]), np.tanh(g[3 * hs:])
        c = f * c + i * cc
        h = o * np.tanh(c)
        ys.append(h.copy())
    return np.array(ys), h, c


@pytest.mark.parametrize("hs,isz,seq,with_state", [(4, 3, 1, False), (4, 3, 6, True), (20, 7, 9, True), (128, 64, 5, False)])
def test_lstm_vs_float64_restatement(hs, isz, seq, with_state):
    """An independent check of the LSTM semantics (gate order, bias split Wb | Rb, state hand-over): the oracle against the ONNX
    formulas in float64, 1e-5 absolute (the first case is the reference's own test input, tests/regression_kernels.rs:977-984)."""
    if (hs, isz, seq) == (4, 3, 1):
        x = np.array([0.1, -0.2, 0.3], np.float32).reshape(1, 1, 3)
        w = (np.arange(4 * hs * isz) * 0.01).astype(np.float32).reshape(1, 4 * hs, isz)
        r = (np.arange(4 * hs * hs) * 0.02 - 0.1).astype(np.float32).reshape(1, 4 * hs, hs)
        b = (np.arange(8 * hs) * 0.005).astype(np.float32).reshape(1, 8 * hs)
    else:
        rng = np.random.default_rng(hs + seq)
        x = rng.standard_normal((seq, 1, isz)).astype(np.float32)
        w = (rng.standard_normal((1, 4 * hs, isz)) / np.sqrt(isz)).astype(np.float32); r = (rng.standard_normal((1, 4 * hs, hs)) / np.sqrt(hs)).astype(np.float32)
        b = (0.1 * rng.standard_normal((1, 8 * hs))).astype(np.float32)
    h0 = c0 = None
    if with_state:
        rng = np.random.default_rng(99)
        h0 = rng.standard_normal((1, 1, hs)).astype(np.float32); c0 = rng.standard_normal((1, 1, hs)).astype(np.float32)
    y, h, c = R.lstm(x, w, r, b, h0, c0)
    yr, hr, cr = _ref_lstm(x, w, r, b, h0, c0)
    assert y.shape == (seq, 1, 1, hs) and h.shape == (1, 1, hs) and c.shape == (1, 1, hs)
    np.testing.assert_allclose(y.reshape(seq, hs), yr, atol=1e-5, rtol=0)
    np.testing.assert_allclose(h.reshape(hs), hr, atol=1e-5, rtol=0); np.testing.assert_allclose(c.reshape(hs), cr, atol=2e-5, rtol=0)
    if b is not None:
        y2, _, _ = R.lstm(x, w, r, None, h0, c0)
        np.testing.assert_allclose(y2.reshape(seq, hs), _ref_lstm(x, w, r, None, h0, c0)[0], atol=1e-5, rtol=0)


def test_stft_vs_numpy_rfft():  # math.rs:2304-2370: default window = periodic Hann over win_length, zero beyond win / signal end
    rng = np.random.default_rng(6)
    sig = rng.standard_normal(700).astype(np.float32)
    n_fft, hop, win = 128, 48, 96
    got = R.stft(sig, n_fft, hop, win)
    frames = (700 - win) // hop + 1
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(win) / win))
    want = np.zeros((frames, n_fft // 2 + 1, 2))
    for f in range(frames):
        fr = np.zeros(n_fft); seg = sig[f * hop:f * hop + win].astype(np.float64); fr[:seg.size] = seg * hann[:seg.size]
        sp = np.fft.rfft(fr); want[f, :, 0], want[f, :, 1] = sp.real, sp.imag
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=2e-4)
    w = rng.standard_normal(win).astype(np.float32)                         # explicit window
    got = R.stft(sig, n_fft, hop, win, w, power=True)
    for f in (0, frames - 1):
        fr = np.zeros(n_fft); fr[:win] = sig[f * hop:f * hop + win].astype(np.float64) * w
        np.testing.assert_allclose(got[f], np.abs(np.fft.rfft(fr)) ** 2, rtol=1e-3, atol=1e-3)


def test_norms_vs_float64_formulas():  # norm.rs: batch_norm (per channel, inference form), rms_norm, softmax
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 3, 5, 4)).astype(np.float32)
    sc, bi, mu = (rng.standard_normal(3).astype(np.float32) for _ in range(3)); var = rng.uniform(0.5, 2.0, 3).astype(np.float32)
    want = (x.astype(np.float64) - mu[None, :, None, None]) / np.sqrt(var[None, :, None, None].astype(np.float64) + 1e-5) * sc[None, :, None, None] + bi[None, :, None, None]
    np.testing.assert_allclose(R.batch_norm(x, sc, bi, mu, var, 1e-5), want, atol=1e-5)
    y = rng.standard_normal((7, 33)).astype(np.float32); g = rng.standard_normal(33).astype(np.float32)
    want = y.astype(np.float64) / np.sqrt((y.astype(np.float64) ** 2).mean(-1, keepdims=True) + 1e-6) * g
    np.testing.assert_allclose(R.rms_norm(y, g, 1e-6), want, atol=1e-5)
    e = np.exp(y.astype(np.float64) - y.max(-1, keepdims=True))
    np.testing.assert_allclose(R.softmax(y, -1), e / e.sum(-1, keepdims=True), atol=1e-6)


def test_frontend_vs_float64_restatement():
    """The whole front-end of src/features/pipeline.rs:67-196 written independently in float64 numpy (x32768, per-frame mean removal,
    in-frame pre-emphasis 0.97 with the first sample untouched, symmetric Hann, zero-pad 400 -> 512, |rfft|^2, HTK triangular mel
    bank from 20 Hz to Nyquist, ln(max(., 1e-5)), LFR m=7 n=6 with edge clamping) against the oracle: log-mel within 5e-4
    absolute (f32 FFT + f32 sums of ~1e9-sized powers), LFR exact on the oracle's own mel."""
    from lele_b200.sensevoice_weights import synth_batch
    pcm = synth_batch(3, 1, 16000 + 123)[0]
    mel, out = R.frontend(pcm, want_mel=True)
    x = pcm.astype(np.float64) * 32768.0
    n_frames = (x.size - 400) // 160 + 1
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(400) / 399.0))
    hz2mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    mel2hz = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    pts = mel2hz(hz2mel(20.0) + np.arange(82) * (hz2mel(8000.0) - hz2mel(20.0)) / 81.0)
    freqs = np.arange(257) * 16000.0 / 512.0
    bank = np.zeros((80, 257))
    for i in range(80):
        lo, ce, hi = pts[i], pts[i + 1], pts[i + 2]
        up = (freqs > lo) & (freqs < ce); dn = (freqs >= ce) & (freqs < hi)
        bank[i, up] = (freqs[up] - lo) / (ce - lo); bank[i, dn] = (hi - freqs[dn]) / (hi - ce)
    want = np.zeros((n_frames, 80))
    for f in range(n_frames):
        fr = x[f * 160:f * 160 + 400].copy()
        fr -= fr.mean()
        fr[1:] -= 0.97 * fr[:-1]
        buf = np.zeros(512); buf[:400] = fr * hann
        want[f] = np.log(np.maximum(bank @ (np.abs(np.fft.rfft(buf)) ** 2), 1e-5))
    assert mel.shape == want.shape == (n_frames, 80)
    np.testing.assert_allclose(mel, want, atol=5e-4)          # measured 1e-4 (f32 FFT and mel sums of ~1e9-sized powers)
    t_lfr = -(-n_frames // 6)
    idx = np.clip(np.arange(t_lfr)[:, None] * 6 + np.arange(7)[None, :] - 3, 0, n_frames - 1)
    np.testing.assert_array_equal(out, mel[idx].reshape(t_lfr, 560))


def test_network_oracle_equals_composition_of_pinned_operators():
    """oracle/sensevoice_ref.c (the whole-network CPU restatement the GPU runner is compared with) against the same network composed
    in Python from the operator-level oracle functions, which are the ones pinned to the reference's KATs: prompt embedding +
    sqrt(d) scaling + positions; per layer LayerNorm, fused int8 QKV, FSMN memory (depthwise conv k=11 on V, + V), 4-head softmax
    attention, int8 out-projection + FSMN (+ input) residual, LayerNorm, int8 FFN (ReLU) + residual; after_norm at the stage
    boundary; tp_norm + int8 CTC head; greedy arg-max with the last-max tie rule.  Synthetic weights (no model file exists)."""
    from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, synth_batch
    cfg = SenseVoiceConfig(n_layers=3, vocab=1000, n_stage1=2, max_t=128)
    blob = build_blob(cfg, seed=7)
    hdr = blob[:256].view(np.int32); nt = int(hdr[12])
    table = blob[256:256 + 16 * nt].view(np.uint64).reshape(-1, 2)

    def T_(i, dt=np.float32):
        o, n = int(table[i, 0]), int(table[i, 1])
        return blob[o:o + n].view(dt)

    L = lambda l, w, dt=np.float32: T_(10 + l * 21 + w, dt)
    qlin = lambda x, w, k, n, sc, zp, bi, relu=False: R.fused_quantized_linear(x[None], w.reshape(k, n), sc, int(zp[0]), bi, relu)[0]

    def network(feats, lang, textnorm, n_layers):
        T = feats.shape[0] + 4
        embed = T_(0).reshape(16, 560); pos = T_(1).reshape(-1, 560)
        x = np.concatenate([embed[[lang, 1, 2, textnorm]], feats], 0) * np.float32(np.sqrt(np.float32(512))) + pos[:T]
        for l in range(n_layers):
            cur = x.shape[1]
            h = R.layer_norm(x, L(l, 0), L(l, 1), -1, 1e-5)
            qkv = qlin(h, L(l, 2, np.uint8), cur, 1536, L(l, 3), L(l, 5, np.uint8), L(l, 4))
            q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
            fs = R.conv1d(v.T[None], L(l, 6).reshape(512, 1, 11), None, (1,), 512, (5, 5), (1,))[0].T + v
            qh = (q.reshape(T, 4, 128).transpose(1, 0, 2) * np.float32(1.0 / np.sqrt(np.float32(128)))).astype(np.float32)
            kh = k.reshape(T, 4, 128).transpose(1, 2, 0); vh = v.reshape(T, 4, 128).transpose(1, 0, 2)
            o = R.matmul(R.softmax(R.matmul(qh, kh)), vh).transpose(1, 0, 2).reshape(T, 512)
            att = qlin(o, L(l, 7, np.uint8), 512, 512, L(l, 8), L(l, 10, np.uint8), L(l, 9)) + fs
            x = att + x if cur == 512 else att                          # the 560-wide first layer has no input residual
            h2 = R.layer_norm(x, L(l, 11), L(l, 12), -1, 1e-5)
            f1 = qlin(h2, L(l, 13, np.uint8), 512, 2048, L(l, 14), L(l, 16, np.uint8), L(l, 15), True)
            x = x + qlin(f1, L(l, 17, np.uint8), 2048, 512, L(l, 18), L(l, 20, np.uint8), L(l, 19))
            if l == cfg.n_stage1 - 1:
                x = R.layer_norm(x, T_(2), T_(3), -1, 1e-5)
        if n_layers < cfg.n_layers:
            return x
        h = R.layer_norm(x, T_(4), T_(5), -1, 1e-5)
        return qlin(h, T_(6, np.uint8), 512, cfg.vocab, T_(7), T_(9, np.uint8), T_(8))

    rng = np.random.default_rng(0)
    ref = R.SenseVoiceRef(blob)
    for scale, lang, tn in ((1.0, 3, 0), (2.5, 0, 15)):
        feats = (rng.standard_normal((40, 560)) * scale).astype(np.float32)
        for n_layers in (1, 2, 3):
            got, want = ref.forward(feats, lang, tn, n_layers=n_layers), network(feats, lang, tn, n_layers)
            assert got.shape == want.shape == (44, 1000 if n_layers == 3 else 512)
            err = float(np.abs(got - want).max() / np.abs(want).max())
            assert err < 2e-4 * n_layers, (n_layers, err)   # same operators in the same order; int8 re-quantisation may move a code by one step
    # raw audio -> ids: front-end + CMVN + network + arg-max keeping the LAST maximum (tokenizer.rs:55-59)
    pcm = synth_batch(1, 1, 16000)[0]
    ids, logits = ref.pcm_to_ids(pcm, 3, 0, want_logits=True)
    lfr = R.frontend(pcm)
    want_logits = network(R.cmvn(lfr), 3, 0, 3)
    assert logits.shape == want_logits.shape
    assert float(np.abs(logits - want_logits).max() / np.abs(want_logits).max()) < 1e-3
    last_max = logits.shape[1] - 1 - np.argmax(logits[:, ::-1], axis=1)
    np.testing.assert_array_equal(ids, last_max)


def test_network_sensitivity_floor_of_the_oracle_itself():
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    """The end-to-end bar of tests/test_gpu_sensevoice.py::test_benched_configuration_vs_oracle, measured where it can be measured
    without a GPU: the 70-layer seed-1234 network run twice ON THE ORACLE, the second time on features multiplied by
    (1 + 1e-7 N(0,1)) (about one ulp).  281 dynamic quantisers turn that into flipped u8 codes and the two runs decorrelate to a
    saturation level set by the quantisation step, not by the size of the perturbation: ~3 % of the mean |logit|, ~93 % shared
    greedy ids.  Any implementation that differs from the reference by one rounding anywhere sits on this floor."""
    import json
    from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, synth_batch
    blob = build_blob(SenseVoiceConfig(), seed=1234)
    ref = R.SenseVoiceRef(blob)
    feats = R.cmvn(R.frontend(synth_batch(0, 1, 256000)[0]))
    rng = np.random.default_rng(0)
    last = lambda l: l.shape[-1] - 1 - np.argmax(l[..., ::-1], -1)
    base10, base70 = ref.forward(feats, 3, 0, n_layers=10), ref.forward(feats, 3, 0)
    report = []
    for eps in (1e-7, 1e-5):
        twin = (feats * (1 + eps * rng.standard_normal(feats.shape))).astype(np.float32)
        t10, t70 = ref.forward(twin, 3, 0, n_layers=10), ref.forward(twin, 3, 0)
        report.append({"relative_input_noise": eps, "rel_mean_layer10": float(np.abs(t10 - base10).mean() / np.abs(base10).mean()),
                       "rel_mean_logits": float(np.abs(t70 - base70).mean() / np.abs(base70).mean()), "logits_mae": float(np.abs(t70 - base70).mean()),
                       "ids_agreement": float((last(t70) == last(base70)).mean())})
    with open(os.path.join(ROOT, "profiles", "r02_oracle_self_sensitivity.json"), "w") as fh:
        json.dump({"what": "CPU oracle vs itself on ~1-ulp / 1e-5 perturbed features, clip 0, 70 layers, seed 1234", "rows": report}, fh, indent=1)
    for r in report:
        assert 0.005 < r["rel_mean_logits"] < 0.06 and 0.8 < r["ids_agreement"] < 0.995 and r["logits_mae"] < 0.06, r
    assert report[1]["rel_mean_logits"] < 1.5 * report[0]["rel_mean_logits"]      # saturation: 100x the noise, the same floor
