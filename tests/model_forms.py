"""Generated-style model.rs snippets shared by the CPU and GPU replay tests: statement forms of src/compiler/ops/{nn,math,tensor}.rs
beyond the Yolo fixture (layer_norm, gemm, conv1d, math, reductions, pad, expand, squeeze, where_op, LSTM / GRU tuples)."""
import numpy as np

from oracle import reference_api as R  # noqa: F401  (tests only)

H, I = 8, 6

MATH_TEXT = """
pub struct T2Workspace { pub buf_0: Vec<f32>, }
pub struct T2<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T2Workspace, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {
        let a = lele::kernels::layer_norm(&x, &self.weight_f32(0, 32, &[8]), &self.weight_f32(32, 32, &[8]), -1, 0.00001, &mut ws.buf_0);
        let b = lele::kernels::gemm(&a, &self.weight_f32(64, 192, &[6, 8]), Some(&self.weight_f32(256, 24, &[6])), 1.0, 1.0, false, true, &mut ws.buf_0);
        let c = lele::kernels::tanh_kernel(&b, &mut ws.buf_0);
        let d = lele::kernels::max(&c, &self.weight_f32(280, 4, &[1]), &mut ws.buf_0);
        let e = lele::kernels::reduce_mean(&d, &[1], true, &mut ws.buf_0);
        let f = lele::kernels::exp(&e, &mut ws.buf_0);
        let g = lele::kernels::expand(&f, &[5, 6], &mut ws.buf_0);
        let h = lele::kernels::where_op(&d, &g, &b, &mut ws.buf_0);
        let p = lele::kernels::pad(&h, &[0, 1, 0, 2], 0.5, "constant", &mut ws.buf_0);
        let q = lele::kernels::unsqueeze(&p, &[0]);
        let r = lele::kernels::conv1d(&q, &self.weight_f32(284, 60, &[3, 5, 1]), None, &[1], 1, &[0, 0], &[1], &mut ws.buf_0);
        let s = lele::kernels::squeeze(&r, &[0]);
        (s.to_owned(), e.to_owned())
    }
"""


def math_forms(m):
    prog = m.parse_model_rs(MATH_TEXT)
    blob = m.synth_blob(prog, 3, {280: [0.0]})
    x = np.random.default_rng(1).standard_normal((5, 8)).astype(np.float32)
    return prog, blob, x


def math_forms_direct(m, blob, x):
    W = lambda off, ln, shp: m.weight_view(blob, "weight_f32", off, ln, shp)
    a = R.layer_norm(x, W(0, 32, [8]), W(32, 32, [8]), -1, 1e-5)
    b = R.gemm(a, W(64, 192, [6, 8]), W(256, 24, [6]).reshape(-1), 1.0, 1.0, False, True)
    d = R.maximum(R.tanh(b), W(280, 4, [1]))
    e = R.reduce(d, [1], True, "mean")
    h = R.where(d, R.expand(R.exp(e), [5, 6]), b)
    p = R.pad(h, [0, 1, 0, 2], 0.5, "constant")
    r = R.conv1d(p[None], W(284, 60, [3, 5, 1]), None, (1,), 1, (0, 0), (1,), False)
    return r[0], e


def recurrent_text():
    return f"""
pub struct T3Workspace {{ pub buf_0: Vec<f32>, }}
pub struct T3<'a> {{ data: &'a [u8] }}
    fn run_chunk_0<'w>(&self, ws: &'w mut T3Workspace, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>, TensorView<'static, f32>) {{
        let mut buf_y_h = Vec::<f32>::new();
        let mut buf_y_c = Vec::<f32>::new();
        let (y, yh, yc) = lele::kernels::lstm(&x, &self.weight_f32(0, {4*H*I*4}, &[1, {4*H}, {I}]), &self.weight_f32({4*H*I*4}, {4*H*H*4}, &[1, {4*H}, {H}]), Some(&self.weight_f32({4*H*(I+H)*4}, {8*H*4}, &[1, {8*H}])), None, None, None, &mut ws.buf_0, &mut buf_y_h, &mut buf_y_c);
        let y2 = lele::kernels::reshape(&y, &[0, 1, {H}]);
        let mut buf_g_h = Vec::<f32>::new();
        let mut buf_g = Vec::<f32>::new();
        let (g, _) = lele::kernels::gru(&y2, &self.weight_f32(2000, {3*H*H*4}, &[1, {3*H}, {H}]), &self.weight_f32(3000, {3*H*H*4}, &[1, {3*H}, {H}]), None, Some(&yh), false, &mut buf_g, &mut buf_g_h);
        let (n, _, _) = (lele::kernels::layer_norm(&g, &self.weight_f32(4000, {H*4}, &[{H}]), &self.weight_f32(4100, {H*4}, &[{H}]), -1, 0.00001, &mut ws.buf_0), lele::tensor::TensorView::empty(), lele::tensor::TensorView::empty());
        (n.to_owned(), yh.to_owned(), yc.to_owned())
    }}
"""


def recurrent_forms(m):
    prog = m.parse_model_rs(recurrent_text())
    blob = m.synth_blob(prog, 5)
    x = np.random.default_rng(2).standard_normal((7, 1, I)).astype(np.float32)
    return prog, blob, x


def recurrent_forms_direct(m, blob, x):
    W = lambda off, ln, shp: m.weight_view(blob, "weight_f32", off, ln, shp)
    y, yh, yc = R.lstm(x, W(0, 4*H*I*4, [1, 4*H, I]), W(4*H*I*4, 4*H*H*4, [1, 4*H, H]), W(4*H*(I+H)*4, 8*H*4, [1, 8*H]))
    g, _ = R.gru(y.reshape(7, 1, H), W(2000, 3*H*H*4, [1, 3*H, H]), W(3000, 3*H*H*4, [1, 3*H, H]), None, yh)
    n = R.layer_norm(g, W(4000, H*4, [H]), W(4100, H*4, [H]), -1, 1e-5)
    return n, yh, yc


QUANT_TEXT = """
pub struct T4Workspace { pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, }
pub struct T4<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T4Workspace, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {
        let a = self.layer_norm(&x, self.weight_f32(0, 64, &[16]), self.weight_f32(64, 64, &[16]), self.weight_f32(128, 4, &[]), self.weight_f32(132, 4, &[]), &mut ws.buf_0);
        #[cfg(target_arch = "aarch64")]
        let b = self.linear_quantized_relu_arm(&a, 200, 384, 16, 24, self.weight_f32(600, 4, &[]), self.weight_u8(604, 1, &[]), self.weight_f32(608, 96, &[24]), &mut ws.buf_1);
        #[cfg(not(target_arch = "aarch64"))]
        let b = self.linear_quantized_relu(&a, self.weight_u8(200, 384, &[16, 24]), self.weight_f32(600, 4, &[]), self.weight_u8(604, 1, &[]), self.weight_f32(608, 96, &[24]), &mut ws.buf_1);
        let c = self.linear_quantized(&b, self.weight_u8(704, 384, &[24, 16]), self.weight_f32(1088, 64, &[16]), self.weight_u8(1152, 1, &[]), self.weight_f32(1156, 64, &[16]), &mut ws.buf_0);
        let mut buf_q = Vec::<f32>::new();
        let mut buf_qs = Vec::<f32>::new();
        let mut buf_qz = Vec::<f32>::new();
        let (q_ref, qs_ref, qz_ref) = lele::kernels::dynamic_quantize_linear(&c, &mut buf_q, &mut buf_qs, &mut buf_qz);
        let q = q_ref.to_owned();
        let qs = qs_ref.to_owned();
        let qz = qz_ref.to_owned();
        #[cfg(target_arch = "aarch64")]
        let d = self.mat_mul_integer_arm(&q, 1220, 384, 16, 24, Some(&qz), Some(&self.weight_u8(1604, 1, &[])), &mut ws.buf_1);
        #[cfg(not(target_arch = "aarch64"))]
        let d = lele::kernels::mat_mul_integer(&q, &self.weight_u8(1220, 384, &[16, 24]), Some(&qz), Some(&self.weight_u8(1604, 1, &[])), &mut ws.buf_1);
        let e = lele::kernels::mul(&d, &qs, &mut ws.buf_0);
        let f = lele::kernels::clip(&e, Some(&self.weight_f32(1608, 4, &[])), None, &mut ws.buf_1);
        let g = self.linear(&f, &self.weight_f32(1612, 768, &[24, 8]), &self.weight_f32(2380, 32, &[8]), &mut ws.buf_0);
        (g.to_owned(), d.to_owned())
    }
"""


def quant_forms(m):
    """The int8 statement pairs of patterns.rs:383-430 / ops/math.rs:43-95 (aarch64 arm skipped, portable arm taken) and the helper
    methods of snippets/default_methods.rs."""
    prog = m.parse_model_rs(QUANT_TEXT)
    rng = np.random.default_rng(11)
    consts = {128: [1e-5], 132: [2.0], 200: rng.integers(0, 256, 384), 600: [0.02], 604: [128], 704: rng.integers(0, 256, 384),
              1088: 0.01 + 0.02 * rng.random(16), 1152: [121], 1220: rng.integers(0, 256, 384), 1604: [130], 1608: [-1.5]}
    blob = m.synth_blob(prog, 9, consts)
    x = rng.standard_normal((2, 5, 16)).astype(np.float32)
    return prog, blob, x


def quant_forms_direct(m, blob, x):
    W = lambda kind, off, ln, shp: m.weight_view(blob, "weight_" + kind, off, ln, shp)
    a = R.layer_norm(x, W("f32", 0, 64, [16]), W("f32", 64, 64, [16]), -1, 1e-5)
    b = R.fused_quantized_linear(a, W("u8", 200, 384, [16, 24]), W("f32", 600, 4, []), 128, W("f32", 608, 96, [24]), True)
    c = R.fused_quantized_linear(b, W("u8", 704, 384, [24, 16]), W("f32", 1088, 64, [16]), 121, W("f32", 1156, 64, [16]), False)
    q, qs, qz = R.dynamic_quantize_linear(c)
    d = R.mat_mul_integer(q, W("u8", 1220, 384, [16, 24]), float(qz), 130.0)
    f = R.clip(R.mul(d, qs), -1.5, float("inf"))
    g = R.matmul_fused_add(f, W("f32", 1612, 768, [24, 8]), W("f32", 2380, 32, [8]))
    return g, d


SHAPE_TEXT = """
pub struct T5Workspace { pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, }
pub struct T5<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T5Workspace, x: TensorView<'w, f32>, len: TensorView<'w, i64>) -> (TensorView<'static, f32>, TensorView<'static, i64>, TensorView<'static, i64>) {
        let s = lele::kernels::shape(&x);
        let b = lele::kernels::gather(&s, &self.weight_i64(0, 8, &[]), 0, &mut ws.buf_0);
        let t = lele::kernels::gather(&s, &self.weight_i64(8, 8, &[]), 0, &mut ws.buf_0);
        let bu = lele::kernels::unsqueeze(&b, &[0]);
        let tu = lele::kernels::unsqueeze(&t, &[0]);
        let c = lele::kernels::concat(&[&bu, &tu, &self.weight_i64(16, 16, &[2])], 0, &mut ws.buf_0);
        let temp_i64_1 = lele::kernels::to_i64_vec(&c);
        let r = lele::kernels::reshape(&x, &temp_i64_1);
        let p = lele::kernels::transpose(&r, &[0, 2, 1, 3], &mut ws.buf_1);
        let mut temp_cast_buf_tf = Vec::<f32>::new();
        let tf = lele::kernels::utils::cast_to_f32(&t, &mut temp_cast_buf_tf);
        let sq = lele::kernels::sqrt(&tf, &mut ws.buf_0);
        let y = lele::kernels::div(&p, &sq, &mut ws.buf_0);
        let mut buf_rg = Vec::<i64>::new();
        let rg = lele::kernels::range_i64(&self.weight_i64(32, 8, &[]), &t, &self.weight_i64(40, 8, &[]), &mut buf_rg);
        let ru = lele::kernels::unsqueeze(&rg, &[0]);
        let lu = lele::kernels::unsqueeze(&len, &[1]);
        let mk = lele::kernels::less_i64(&ru, &lu, &mut ws.buf_1);
        let mut temp_cast_buf_mf = Vec::<f32>::new();
        let mf = lele::kernels::utils::cast_to_f32(&mk, &mut temp_cast_buf_mf);
        let m4 = lele::kernels::unsqueeze(&mf, &[1, 3]);
        let z = lele::kernels::mul(&y, &m4, &mut ws.buf_1);
        let h = lele::kernels::mul(&t, &self.weight_i64(48, 8, &[]), &mut ws.buf_0);
        let zs = lele::kernels::slice(&z, &[1], &t.data[..], &[3], &[1], &mut ws.buf_0);
        let k = lele::kernels::constant_of_shape(&c, 0.0, &mut ws.buf_1);
        let sz = lele::kernels::size(&k);
        let e = lele::kernels::reduce_sum(&zs, &lele::kernels::to_i64_vec(&self.weight_i64(56, 8, &[1])), false, &mut ws.buf_1);
        (e.to_owned(), h.to_owned(), sz.to_owned())
    }
"""


def shape_forms(m):
    """Dynamic-shape plumbing of an exported transformer block: Shape -> Gather -> Unsqueeze -> Concat -> Reshape on i64 host tensors,
    a length mask from Range / Less, Cast back to f32 for the device multiply (ops/tensor.rs:10-70, :191-291; ops/math.rs:360-405)."""
    prog = m.parse_model_rs(SHAPE_TEXT)
    blob = m.synth_blob(prog, 4, {0: [0], 8: [1], 16: [4, 4], 32: [0], 40: [1], 48: [2], 56: [-1]})
    x = np.random.default_rng(5).standard_normal((2, 5, 16)).astype(np.float32)
    return prog, blob, [x, np.array([3, 5], np.int64)]


def shape_forms_direct(x, lens):
    bsz, t, _ = x.shape
    y = R.div(R.transpose(x.reshape(bsz, t, 4, 4), [0, 2, 1, 3]), R.sqrt(np.float32(t)))
    mask = (np.arange(t)[None, :] < lens[:, None]).astype(np.float32)[:, None, :, None]
    z = R.mul(y, mask)
    return R.reduce(z[..., 1:], [-1], False, "sum"), np.array(2 * t, np.int64), np.array(bsz * t * 16, np.int64)


CONST_TEXT = """
pub struct T6Workspace { pub buf_0: Vec<f32>, }
pub struct T6<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T6Workspace, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {
        let win = self.weight_f32(0, 256, &[64]);
        let two = lele::tensor::TensorView::from_owned(vec![2.0], vec![1]);
        let big = lele::tensor::TensorView::empty(); // Large
        let sp = lele::kernels::stft(&x, 64, 16, 64, Some(&win), &mut ws.buf_0);
        let p = lele::kernels::pow(&sp, &two, &mut ws.buf_0);
        let e = lele::kernels::reduce_sum(&p, &[-1], false, &mut ws.buf_0);
        let l = lele::kernels::log(&e, &mut ws.buf_0);
        let s = lele::kernels::sin(&l, &mut ws.buf_0);
        let c = lele::kernels::cos(&l, &mut ws.buf_0);
        let lt = lele::kernels::less(&s, &c, &mut ws.buf_0);
        let nt = lele::kernels::not(&lt, &mut ws.buf_0);
        let eq = lele::kernels::equal(&nt, &lt, &mut ws.buf_0);
        let y = lele::kernels::where_op(&lt, &s, &c, &mut ws.buf_0);
        (y.to_owned(), eq.to_owned())
    }
"""


def const_forms(m):
    """Constant / Identity statements (stored tensor, inline literal, empty), STFT with a stored window, and the libm-backed math."""
    prog = m.parse_model_rs(CONST_TEXT)
    blob = m.synth_blob(prog, 6, {0: np.hanning(64)})
    x = np.random.default_rng(8).standard_normal(400).astype(np.float32)
    return prog, blob, x


def const_forms_direct(m, blob, x):
    win = m.weight_view(blob, "weight_f32", 0, 256, [64])
    l = R.log(R.reduce(R.pow(R.stft(x, 64, 16, 64, win), np.float32(2.0)), [-1], False, "sum"))
    s, c = R.sin(l), R.cos(l)
    lt = R.less(s, c)
    return R.where(lt, s, c), R.equal(R.not_(lt), lt)


VH = 16   # hidden size of the synthetic streaming model (Silero: 128)


def vad_text(chunk=512, hidden=None):
    """A recurrent streaming model with Silero's calling convention (examples/silero/src/main.rs:121): (input [1, chunk], state [2,1,H],
    sr [1] i64) -> (probability [1,1], new state [2,1,H]).  STFT power frames -> Conv1d+ReLU -> LSTM over the frames with the carried
    (h, c) -> Gemm + Sigmoid on the last hidden state."""
    H, nfr = (hidden or VH), 33
    o_cw = 0; o_cb = o_cw + 8 * nfr * 4; o_w = o_cb + 8 * 4; o_r = o_w + 4 * H * 8 * 4; o_b = o_r + 4 * H * H * 4; o_g = o_b + 8 * H * 4; o_gb = o_g + H * 4
    return f"""
pub struct SynthVadWorkspace {{ pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, }}
pub struct SynthVad<'a> {{ data: &'a [u8] }}
    fn run_chunk_0<'w>(&self, ws: &'w mut SynthVadWorkspace, input: TensorView<'w, f32>, sr: TensorView<'w, i64>, state: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {{
        let k = lele::tensor::TensorView::from_owned(vec![0.000030517578], vec![1]);
        let sc = lele::kernels::mul(&input, &k, &mut ws.buf_0);
        let sp = lele::kernels::stft(&sc, 64, 64, 64, None, &mut ws.buf_1);
        let pw = lele::kernels::mul(&sp, &sp, &mut ws.buf_0);
        let en = lele::kernels::reduce_sum(&pw, &[-1], false, &mut ws.buf_1);
        let et = lele::kernels::transpose(&en, &[0, 2, 1], &mut ws.buf_0);
        let cv = self.conv1d_relu(et, self.weight_f32({o_cw}, {8 * nfr * 4}, &[8, {nfr}, 1]), Some(&self.weight_f32({o_cb}, 32, &[8])), 1, 1, 1, 0, &mut ws.buf_1);
        let ct = lele::kernels::transpose(&cv, &[2, 0, 1], &mut ws.buf_0);
        let h0 = lele::kernels::slice(&state, &[0], &[1], &[0], &[1], &mut ws.buf_1);
        let c0 = lele::kernels::slice(&state, &[1], &[2], &[0], &[1], &mut ws.buf_1);
        let mut buf_y_h = Vec::<f32>::new();
        let mut buf_y_c = Vec::<f32>::new();
        let (y, yh, yc) = lele::kernels::lstm(&ct, &self.weight_f32({o_w}, {4 * H * 8 * 4}, &[1, {4 * H}, 8]), &self.weight_f32({o_r}, {4 * H * H * 4}, &[1, {4 * H}, {H}]), Some(&self.weight_f32({o_b}, {8 * H * 4}, &[1, {8 * H}])), None, Some(&h0), Some(&c0), &mut ws.buf_0, &mut buf_y_h, &mut buf_y_c);
        let stateN = lele::kernels::concat(&[&yh, &yc], 0, &mut ws.buf_1);
        let hf = lele::kernels::reshape(&yh, &[1, {H}]);
        let lg = lele::kernels::gemm(&hf, &self.weight_f32({o_g}, {H * 4}, &[1, {H}]), Some(&self.weight_f32({o_gb}, 4, &[1])), 1.0, 1.0, false, true, &mut ws.buf_0);
        let output = lele::kernels::sigmoid(&lg, &mut ws.buf_1);
        (output.to_owned(), stateN.to_owned())
    }}

    pub fn forward_with_workspace<'w>(&self, ws: &'w mut SynthVadWorkspace, input: TensorView<'w>, state: TensorView<'w>, sr: TensorView<'w, i64>) -> (TensorView<'w>, TensorView<'w>) {{
        let (output, stateN) = self.run_chunk_0(ws, input, sr, state);
        (output, stateN)
    }}
}}
"""


def vad_model(m, hidden=None):
    prog = m.parse_model_rs(vad_text(hidden=hidden))
    if hidden is None:
        return prog, m.synth_blob(prog, 21)
    # Silero-sized stand-in driven by real audio for 175 chunks x 8 frames: LSTM weights scaled by fan-in and a negative forget
    # bias keep the carried cell state bounded (the reference's SIMD tanh is (1 - e^-2x) / (1 + e^-2x) with e^-2x overflowing to
    # inf below x = -44: a cell state that drifts there is NaN upstream as well, avx/math.rs:81-97)
    H, nfr = hidden, 33
    o_cw = 0; o_cb = o_cw + 8 * nfr * 4; o_w = o_cb + 8 * 4; o_r = o_w + 4 * H * 8 * 4; o_b = o_r + 4 * H * H * 4
    rng = np.random.default_rng(22)
    bias = 0.1 * rng.standard_normal(8 * H); bias[2 * H:3 * H] -= 1.0          # gate order i, o, f, c (rnn.rs:67): Wb forget block
    # frame energies of real speech reach ~2e2 (zh.wav: 221): the conv weights are scaled so the LSTM pre-activations stay O(1)
    consts = {o_cw: 0.01 * rng.standard_normal(8 * nfr) / np.sqrt(nfr), o_w: rng.standard_normal(4 * H * 8) / np.sqrt(8.0),
              o_r: rng.standard_normal(4 * H * H) / np.sqrt(H), o_b: bias}
    return prog, m.synth_blob(prog, 21, consts)


def read_wav_s16(path):
    """fixtures/zh.wav as the reference's WavReader produces it (examples/silero/src/main.rs:70, examples/sensevoice/src/audio.rs:57):
    canonical 44-byte RIFF header, mono s16 little-endian -> f32 / 32768."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE" and raw[36:40] == b"data"
    n = int.from_bytes(raw[40:44], "little")
    return (np.frombuffer(raw, "<i2", count=n // 2, offset=44).astype(np.float32) / np.float32(32768.0))


def vad_chunk_direct(m, blob, chunk, state, hidden=None):
    """The same chunk step written as direct oracle calls."""
    H, nfr = (hidden or VH), 33
    o_cw = 0; o_cb = o_cw + 8 * nfr * 4; o_w = o_cb + 32; o_r = o_w + 4 * H * 8 * 4; o_b = o_r + 4 * H * H * 4; o_g = o_b + 8 * H * 4; o_gb = o_g + H * 4
    W = lambda off, ln, shp: m.weight_view(blob, "weight_f32", off, ln, shp)
    x = R.mul((chunk * np.float32(32768.0)).reshape(1, -1), np.array([0.000030517578], np.float32))
    sp = R.stft(x, 64, 64, 64, None)
    en = R.reduce(R.mul(sp, sp), [-1], False, "sum")
    cv = R.conv1d(R.transpose(en, [0, 2, 1]), W(o_cw, 8 * nfr * 4, [8, nfr, 1]), W(o_cb, 32, [8]), [1], 1, [0, 0], [1], True)
    y, yh, yc = R.lstm(R.transpose(cv, [2, 0, 1]), W(o_w, 4 * H * 8 * 4, [1, 4 * H, 8]), W(o_r, 4 * H * H * 4, [1, 4 * H, H]), W(o_b, 8 * H * 4, [1, 8 * H]), state[0:1], state[1:2])
    prob = R.sigmoid(R.gemm(yh.reshape(1, H), W(o_g, H * 4, [1, H]), W(o_gb, 4, [1]).reshape(-1), 1.0, 1.0, False, True))
    return float(prob.reshape(-1)[0]), np.concatenate([yh, yc], 0)


CONVINT_TEXT = """
pub struct T9Workspace { pub buf_0: Vec<f32>, }
pub struct T9<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T9Workspace, q: TensorView<'w, f32>, qz: TensorView<'w, f32>) -> TensorView<'static, f32> {
        let y = lele::kernels::conv_integer(&q, &self.weight_u8(0, 54, &[3, 2, 3, 3]), Some(&qz), Some(&self.weight_u8(54, 1, &[])), &[1, 1], 1, &[1, 1, 1, 1], &[2, 2], &mut ws.buf_0);
        y.to_owned()
    }
"""


def convint_forms(m):
    prog = m.parse_model_rs(CONVINT_TEXT)
    rng = np.random.default_rng(12)
    wq = rng.integers(0, 256, 54)
    blob = m.synth_blob(prog, 1, {0: wq, 54: [120]})
    q = rng.integers(0, 256, (1, 2, 5, 6)).astype(np.float32)
    return prog, blob, [q, np.array([7.0], np.float32)]


def convint_forms_direct(m, blob, xs):
    """Direct evaluation in float64: pad the raw tensor with zeros, shift by the zero points, 3x3 stride-2 windows (conv2d.rs:2025)."""
    q = xs[0]
    wq = m.weight_view(blob, "weight_u8", 0, 54, [3, 2, 3, 3]).astype(np.float64)
    xp = np.zeros((1, 2, 7, 8), np.float64); xp[:, :, 1:6, 1:7] = q; xp -= float(xs[1][0])
    w = wq - 120.0
    want = np.zeros((1, 3, 3, 3))
    for oc in range(3):
        for i in range(3):
            for j in range(3):
                want[0, oc, i, j] = (xp[0, :, 2 * i:2 * i + 3, 2 * j:2 * j + 3] * w[oc]).sum()
    return [want.astype(np.float32)]
