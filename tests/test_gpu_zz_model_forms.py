"""GPU replay of the generated-style statement forms in tests/model_forms.py over the C ABI (CudaOps) in lock-step with the CPU
oracle: every statement runs on both back-ends with the same inputs (the oracle's result is carried forward, so one rounding
difference cannot flip a quantiser code or a `where` condition further down).  Bars: exact for the integer stages (u8 codes, int32
sums, fused quantised linear), 1e-4 relative with a 1e-4 * max|ref| floor for f32, as in tests/test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import model_forms as MF            # noqa: E402
from tests.test_gpu_parity import close        # noqa: E402

from tests import gpu_backend as G                       # noqa: E402
from tests.kat_runner import KATS, kat_id, run_kat      # noqa: E402


@pytest.mark.parametrize("k", [k for k in KATS if "checks" in k], ids=kat_id)
def test_reference_kat_on_gpu_second_batch(k):
    """The KATs transcribed after the round's last GPU run (tests/golden/make_reference_kats.py, `kat2` entries)."""
    run_kat(G, k)


EXACT = {"dynamic_quantize_linear", "mat_mul_integer", "fused_quantized_linear", "pad", "expand", "where", "concat", "transpose", "slice", "less", "equal", "not"}


class Lockstep:
    def __init__(self, MR):
        self.gpu, self.cpu, self.calls = MR.CudaOps(), MR._NamespaceOps(MF.R), []

    def __getattr__(self, name):
        def f(*args):
            g, c = getattr(self.gpu, name)(*args), getattr(self.cpu, name)(*args)
            op = args[0] if name in ("binary", "unary") else name
            for a, b in zip(g if isinstance(g, tuple) else (g,), c if isinstance(c, tuple) else (c,)):
                if op in EXACT:
                    np.testing.assert_array_equal(np.asarray(a, np.float32), np.asarray(b, np.float32), err_msg=op)
                else:
                    close(np.asarray(a), np.asarray(b))
            self.calls.append(op)
            return c
        return f


@pytest.mark.parametrize("case", ["math", "recurrent", "quant", "shape", "const", "convint"])
def test_statement_forms_lockstep(case):
    from lele_b200 import model_rs as MR
    make, direct = {"math": (MF.math_forms, MF.math_forms_direct), "recurrent": (MF.recurrent_forms, MF.recurrent_forms_direct),
                    "quant": (MF.quant_forms, MF.quant_forms_direct), "shape": (MF.shape_forms, lambda m, blob, xs: MF.shape_forms_direct(*xs)), "const": (MF.const_forms, MF.const_forms_direct), "convint": (MF.convint_forms, MF.convint_forms_direct)}[case]
    prog, blob, x = make(MR)
    ops = Lockstep(MR)
    got = MR.run_program(prog, blob, x if isinstance(x, list) else [x], ops)
    for a, b in zip(got, direct(MR, blob, x)):
        np.testing.assert_array_equal(a, b)           # the carried values are the oracle's
    assert len(ops.calls) >= {"math": 10, "recurrent": 3, "quant": 8, "shape": 6, "const": 10, "convint": 1}[case]


def test_streaming_vad_on_device():
    """the Silero-style chunk loop (x32768, carried (h, c) state) over a replayed recurrent model -- first in
    lock-step with the oracle per statement, then free-running on the C ABI against the oracle's own stream."""
    from lele_b200 import model_rs as MR
    from lele_b200.vad import StreamingVad
    prog, blob = MF.vad_model(MR)
    audio = (0.1 * np.random.default_rng(33).standard_normal(512 * 3 + 200)).astype(np.float32)
    ops = Lockstep(MR)
    locked = StreamingVad(prog, blob, ops=ops, state_shape=(2, 1, MF.VH)).process(audio)
    assert locked.shape == (4,) and len(ops.calls) == 4 * 13
    ref = StreamingVad(prog, blob, ops=MF.R, state_shape=(2, 1, MF.VH)).process(audio)
    np.testing.assert_array_equal(locked, ref)
    free = StreamingVad(prog, blob, state_shape=(2, 1, MF.VH))            # default operator namespace: CudaOps
    np.testing.assert_allclose(free.process(audio), ref, rtol=1e-3, atol=1e-4)
    cpu = StreamingVad(prog, blob, ops=MF.R, state_shape=(2, 1, MF.VH)); cpu.process(audio)
    np.testing.assert_allclose(free.state, cpu.state, rtol=2e-3, atol=2e-3)   # 3e-6 relative noise per operator moves the carried state by < 2e-4 (measured on the CPU)


def test_c_caller_runs_an_operator(tmp_path):
    """examples/c/abi_tour.c: a plain C99 program drives layer_norm through the explicit-copy entry points and checks the reference's
    known answer (tests/verify_operators.rs:34)."""
    import os
    import subprocess
    from lele_b200 import SO_PATH
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_tour")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c", "abi_tour.c"),
                    "-L", os.path.dirname(SO_PATH), "-llele_b200", "-lm", "-Wl,-rpath," + os.path.dirname(SO_PATH), "-o", exe], check=True)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.stdout, run.stderr)
    import re
    assert "layer_norm([1,2,3]) = -1.2247" in run.stdout
    assert int(re.search(r"kernel launches issued by this context: (\d+)", run.stdout).group(1)) >= 1


def test_generated_model_prepares_int8_weights_once_on_device():
    """GeneratedModel keeps each quantised linear's packed weight resident (lele_b200_prepare_weights once per model, the role of
    B_WEIGHT_CACHE upstream): repeated forwards are bit-identical to each other and to the per-call path."""
    from lele_b200 import model_rs as MR
    prog, blob, x = MF.quant_forms(MR)
    model = MR.GeneratedModel(MF.QUANT_TEXT, blob)
    first, second = model(x), model(x)
    per_call = MR.run_program(prog, blob, [x], MR.CudaOps())
    for a, b, c in zip(first, second, per_call):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    assert sum(1 for k in model._cache if k[0] == "pw") == 2


@pytest.mark.parametrize("hid,isz,seq,n_seq", [(128, 128, 40, 24), (16, 8, 9, 5)])
def test_recurrent_streams_match_single_sequence_calls(hid, isz, seq, n_seq):
    """SURVEY 8f rank 4, the device half: many VAD streams advance in one launch (one CTA per stream, shared weights, per-stream
    carried state).  Each stream must be bit-identical to the batch-1 call the reference's `lstm` / `gru` correspond to, and
    within the f32 bar of the oracle."""
    from lele_b200 import kernels as K
    rng = np.random.default_rng(hid + n_seq)
    x = rng.standard_normal((n_seq, seq, isz)).astype(np.float32)
    w = (rng.standard_normal((1, 4 * hid, isz)) / np.sqrt(isz)).astype(np.float32); r = (rng.standard_normal((1, 4 * hid, hid)) / np.sqrt(hid)).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, 8 * hid))).astype(np.float32)
    h0 = rng.standard_normal((n_seq, hid)).astype(np.float32); c0 = rng.standard_normal((n_seq, hid)).astype(np.float32)
    y, h, c = K.lstm_streams(x, w, r, b, h0, c0)
    assert y.shape == (n_seq, seq, hid) and h.shape == c.shape == (n_seq, hid)
    for s in (0, 1, n_seq - 1):
        y1, h1, c1 = K.lstm(x[s][:, None, :], w, r, b, None, h0[s].reshape(1, 1, hid), c0[s].reshape(1, 1, hid))
        np.testing.assert_array_equal(y[s], y1.reshape(seq, hid)); np.testing.assert_array_equal(h[s], h1.reshape(hid)); np.testing.assert_array_equal(c[s], c1.reshape(hid))
        yr, hr, cr = MF.R.lstm(x[s][:, None, :], w, r, b, h0[s].reshape(1, 1, hid), c0[s].reshape(1, 1, hid))
        close(y[s], yr.reshape(seq, hid)); close(c[s], cr.reshape(hid))
    w3, r3, b3 = w[:, :3 * hid], r[:, :3 * hid], b[:, :6 * hid]
    yg, hg = K.gru_streams(x, w3, r3, b3, h0)
    for s in (0, n_seq - 1):
        y1, h1 = K.gru(x[s][:, None, :], w3, r3, b3, h0[s].reshape(1, 1, hid))
        np.testing.assert_array_equal(yg[s], y1.reshape(seq, hid)); np.testing.assert_array_equal(hg[s], h1.reshape(hid))
        yr, hr = MF.R.gru(x[s][:, None, :], w3, r3, b3, h0[s].reshape(1, 1, hid))
        close(yg[s], yr.reshape(seq, hid))
    y0, h0z, _ = K.lstm_streams(x, w, r, None)                      # no bias, zero initial state
    y1, _, _ = K.lstm(x[2][:, None, :], w, r, None)
    np.testing.assert_array_equal(y0[2], y1.reshape(seq, hid))


def test_stft_shape_rules_on_device():
    """math.rs:2313-2316, :2362-2367, :2381-2384: empty signal -> empty tensor, rank >= 2 input -> leading batch dim."""
    from lele_b200 import kernels as K
    sig = np.sin(np.arange(800, dtype=np.float32) * np.float32(0.01))
    for power in (False, True):
        assert K.stft(np.zeros(0, np.float32), 256, 64, 256, None, power).shape == MF.R.stft(np.zeros(0, np.float32), 256, 64, 256, None, power).shape
        g, r = K.stft(sig[None, :], 256, 128, 256, None, power), MF.R.stft(sig[None, :], 256, 128, 256, None, power)
        assert g.shape == r.shape == ((1, 5, 129) if power else (1, 5, 129, 2))
        close(g, r, atol_frac=1e-5)


def test_reduce_axes_edge_semantics_on_device():
    """math.rs:1611-1650: axes de-duplicated; an empty list reduces nothing (sum / mean -> 0 + x, max -> x, l2 -> |x|)."""
    x = np.array([[1.0, -2.0, 0.5], [3.0, -0.0, -7.0]], np.float32)
    for kind in ("sum", "mean", "max", "l2"):
        for keep in (True, False):
            np.testing.assert_array_equal(G.reduce(x, [], keep, kind), MF.R.reduce(x, [], keep, kind))
        np.testing.assert_array_equal(G.reduce(x, [1, 1, -1], False, kind).shape, (2,))
        close(G.reduce(x, [1, 1, -1], False, kind), MF.R.reduce(x, [1, 1, -1], False, kind), atol_frac=1e-6)


def test_gather_with_scalar_index_drops_the_axis():
    """manipulation.rs:604-611: the output shape splices in the INDEX shape, so a rank-0 index removes the gathered axis."""
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    for axis in (0, 1, -1):
        g, r = G.gather(x, np.array(1, np.int64), axis), MF.R.gather(x, np.array(1, np.int64), axis)
        assert g.shape == r.shape and g.ndim == 2
        np.testing.assert_array_equal(g, r)
    np.testing.assert_array_equal(G.gather(x, np.array([[2, -1]], np.int64), 1), MF.R.gather(x, np.array([[2, -1]], np.int64), 1))


def test_expand_zero_means_input_dim_on_device():
    """math.rs:2189: a 0 in the Expand target keeps the input's size at that position."""
    x = np.arange(6, dtype=np.float32).reshape(2, 1, 3)
    for tgt in ([0, 4, 0], [2, 4, 3], [5, 1, 1, 1], [0, 0, 0]):
        np.testing.assert_array_equal(G.expand(x, tgt), MF.R.expand(x, tgt))
