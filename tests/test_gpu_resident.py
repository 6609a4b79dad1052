"""GPU tests of the device-resident operator path: src/tensor.rs's buffer arena mapped to HBM through
lele_b200_arena_bind, a replayed model.rs whose values never leave the device between statements, the batched / graph-captured
replay (BASELINE configs 5 and 3 are built on it), and the C-ABI entries added with it (conv_integer, batched MatMulInteger,
NCCL comm, device-side gather range check).  The CPU box runs the same host logic against tests/fake_device.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lele_b200 import LeleB200Error                        # noqa: E402
from lele_b200 import _lib                                 # noqa: E402
from lele_b200 import kernels as K                         # noqa: E402
from lele_b200 import model_rs as MR                       # noqa: E402
from oracle import reference_api as R                      # noqa: E402  (checker only)
from tests import model_forms as MF                        # noqa: E402
from tests.test_abi_and_host import RESIDENT_TEXT          # noqa: E402
from tests.test_gpu_parity import close                    # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp, sz = C.c_void_p, C.c_size_t


class CallCounter:
    """Counts C-ABI calls by name while installed (the traffic assertions of the resident replay)."""

    def __init__(self, monkeypatch):
        self.n = {}
        real = _lib.call

        def call(name, *a):
            self.n[name] = self.n.get(name, 0) + 1
            return real(name, *a)
        for mod in (_lib, K):
            monkeypatch.setattr(mod, "call", call)

    def __getitem__(self, name):
        return self.n.get("lele_b200_" + name, 0)

    def clear(self):
        self.n.clear()


def test_arena_through_the_c_abi():
    """lele_b200_arena_bind / _release (csrc/ctx.cu): the device mirror of a host Vec, keyed by its address -- first bind allocates,
    a larger bind grows and KEEPS the contents (Vec::reserve, kernels/utils.rs:10), a smaller one returns the same storage, release
    frees, a re-bind after release starts empty-handed at a (possibly) new address."""
    ctx = K.default_context()
    key = (C.c_char * 8)()
    p1, p2, p3 = vp(), vp(), vp()
    _lib.call("lele_b200_arena_bind", ctx.h, vp(C.addressof(key)), sz(64), C.byref(p1))
    src = np.arange(16, dtype=np.float32)
    _lib.call("lele_b200_h2d", ctx.h, p1, src.ctypes.data_as(vp), sz(64)); ctx.sync()
    _lib.call("lele_b200_arena_bind", ctx.h, vp(C.addressof(key)), sz(1 << 20), C.byref(p2))     # grow
    assert p2.value != p1.value
    back = np.zeros(16, np.float32)
    _lib.call("lele_b200_d2h", ctx.h, back.ctypes.data_as(vp), p2, sz(64)); ctx.sync()
    np.testing.assert_array_equal(back, src)
    _lib.call("lele_b200_arena_bind", ctx.h, vp(C.addressof(key)), sz(128), C.byref(p3))         # fits: same storage
    assert p3.value == p2.value
    other = (C.c_char * 8)(); p4 = vp()
    _lib.call("lele_b200_arena_bind", ctx.h, vp(C.addressof(other)), sz(64), C.byref(p4))        # another Vec: its own mirror
    assert p4.value not in (p2.value, None)
    _lib.call("lele_b200_arena_release", ctx.h, vp(C.addressof(key)))
    _lib.call("lele_b200_arena_release", ctx.h, vp(C.addressof(other)))
    _lib.call("lele_b200_arena_release", ctx.h, vp(C.addressof(other)))                          # releasing twice is harmless
    with pytest.raises(LeleB200Error):
        _lib.call("lele_b200_arena_bind", ctx.h, vp(None), sz(64), C.byref(p4))


def test_device_tensor_operator_calls_stay_on_the_device(monkeypatch):
    """kernels.py with DeviceTensor operands: no copies, results are DeviceTensors (owned, or placed in the named workspace buffer);
    values equal the host form bit for bit (same kernels)."""
    ctx = K.default_context()
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal((4, 33, 64)).astype(np.float32), rng.standard_normal((64,)).astype(np.float32)
    want = K.softmax(K.add(K.matmul(a, a.transpose(0, 2, 1).copy()), np.float32(0.5)), -1)
    cnt = CallCounter(monkeypatch)
    da = ctx.to_device(a); dat = K.transpose(da, (0, 2, 1))
    ws = K.Workspace(ctx)
    ctx.out_slots([(ws, "ws.buf_0")])
    mm = K.matmul(da, dat)
    assert isinstance(mm, K.DeviceTensor) and mm.slot == "ws.buf_0" and mm.shape == (4, 33, 33)
    got = K.softmax(K.add(mm, np.array([0.5], np.float32)), -1)
    assert isinstance(got, K.DeviceTensor) and got.slot is None
    assert cnt["d2h"] == 0 and cnt["h2d"] == 2                     # `a` and the [1] constant; nothing came back yet
    np.testing.assert_array_equal(got.numpy(), want)
    v = K.reshape(mm, [4, -1]); assert v.ptr == mm.ptr and v.shape == (4, 33 * 33) and v.slot == "ws.buf_0"    # zero-copy view (shape.rs:2)
    assert K.unsqueeze(v, [0]).shape == (1, 4, 1089) and K.flatten(mm, 2).shape == (132, 33) and K.squeeze(K.unsqueeze(v, [0]), [0]).shape == (4, 1089)
    y, h, c = K.lstm(ctx.to_device(rng.standard_normal((5, 1, 8)).astype(np.float32)), rng.standard_normal((1, 16, 8)).astype(np.float32),
                     rng.standard_normal((1, 16, 4)).astype(np.float32))
    assert all(isinstance(t, K.DeviceTensor) for t in (y, h, c)) and y.shape == (5, 1, 1, 4)
    del b
    ws.release()


def test_resident_replay_matches_oracle_and_host_form(monkeypatch):
    prog = MR.parse_model_rs(RESIDENT_TEXT)
    blob = MR.synth_blob(prog, 11)
    x = np.random.default_rng(20).standard_normal((1, 3, 8, 8)).astype(np.float32)
    want = MR.run_program(prog, blob, [x], R)
    host_form = MR.run_program(prog, blob, [x], MR.CudaOps())
    cnt = CallCounter(monkeypatch)
    model = MR.GeneratedModel(prog, blob, resident=True)
    for rnd in range(2):
        cnt.clear()
        got = model.forward(x)
        for g, h, w in zip(got, host_form, want):
            np.testing.assert_array_equal(g, h)                     # same kernels, same operands: bit-identical to upload/launch/download
            close(g, w)
        assert cnt["d2h"] == 2 and cnt["h2d"] == (3 if rnd == 0 else 1)   # weights go up once per model; only the graph outputs come down
    assert {"ws.buf_0", "ws.buf_1", "ws.buf_2"} <= set(model.workspace().bytes)


def _yolo_case():
    prog = json.load(open(os.path.join(ROOT, "tests", "golden", "yolo26seg_program.json")))
    pts, strd = [], []
    for s_, g in ((8, 80), (16, 40), (32, 20)):
        ys, xs = np.meshgrid(np.arange(g) + 0.5, np.arange(g) + 0.5, indexing="ij")
        pts.append(np.stack([xs.reshape(-1), ys.reshape(-1)], 0)); strd.append(np.full(g * g, s_, np.float32))
    consts = {int(k): v for k, v in prog["constants"].items()}
    consts[prog["anchor_points_offset"]] = np.concatenate(pts, 1)[None]; consts[prog["anchor_strides_offset"]] = np.concatenate(strd)[None]
    return prog, MR.synth_blob(prog, 7, consts)


def test_yolo26seg_resident_replay_is_the_host_replay(monkeypatch):
    """The reference's committed lele_gen output (337 statements, 21 workspace buffers) with every value resident: identical, bit for
    bit, to the statement-by-statement upload / launch / download replay that test_gpu_parity.py holds to the oracle; the only
    transfers of a steady-state forward are the image up and the two graph outputs down."""
    prog, blob = _yolo_case()
    x = np.random.default_rng(7).random((1, 3, 640, 640), dtype=np.float32)
    host_form = MR.run_program(prog, blob, [x], MR.CudaOps())
    cnt = CallCounter(monkeypatch)
    model = MR.GeneratedModel(prog, blob, resident=True)
    model.forward(x)                                                # first forward: sizes the arena, uploads the weights
    cnt.clear()
    got = model.forward(x)
    assert cnt["h2d"] == 1 and cnt["d2h"] == 2 and cnt["malloc"] == 1 and cnt["free"] <= 1
    for g, h in zip(got, host_form):
        np.testing.assert_array_equal(g, h)
    named = {k for k in model.workspace().bytes if k.startswith("ws.buf_")}
    assert len(named) == 21                                         # yolo26seg.rs:14-36


def test_batch_runner_graph_replay_matches_single_forwards():
    """Config 5's execution form at a small batch: 5 images over 3 lanes; run 1 eager, run 2 captured, run 3 a pure graph replay --
    every image's outputs equal its own resident single forward bit for bit, in every round."""
    prog, blob = _yolo_case()
    rng = np.random.default_rng(9)
    xs = [rng.random((1, 3, 640, 640), dtype=np.float32) for _ in range(5)]
    single = MR.GeneratedModel(prog, blob, resident=True)
    want = [single.forward(x) for x in xs]
    model = MR.GeneratedModel(prog, blob, resident=True)
    br = model.batch_runner(5, lanes=3)
    try:
        for rnd in range(3):
            order = [(i + rnd) % 5 for i in range(5)]
            got = br.run([[xs[i]] for i in order])
            for it, i in zip(got, order):
                for g, w in zip(it, want[i]):
                    np.testing.assert_array_equal(g, w)
        assert br.graph is not None and br.launch_count() > 3 * 5 * 300
    finally:
        br.close()


def test_batch_folded_yolo_replay_matches_single_forwards():
    """Config 5's execution form: the 337-statement graph with the batch folded into the statements that allow it (every
    convolution sees nb = B: one implicit GEMM per layer for the whole batch) and the detection tail per item.  3 images, 2 lanes,
    eager / captured / replayed rounds: each image's outputs are those of its own single resident forward."""
    prog, blob = _yolo_case()
    rng = np.random.default_rng(10)
    xs = [rng.random((1, 3, 640, 640), dtype=np.float32) for _ in range(3)]
    single = MR.GeneratedModel(prog, blob, resident=True)
    want = [single.forward(x) for x in xs]
    br = MR.GeneratedModel(prog, blob, resident=True).batch_runner(3, lanes=2, fold=True)
    try:
        for rnd in range(3):
            got = br.run([[x] for x in xs])
            for it, w in zip(got, want):
                close(it[1], w[1], rtol=1e-5, atol_frac=1e-6)                  # prototype masks [1, 32, 160, 160]
                key = lambda r: (int(r[5]), tuple(np.round(r[:4], 1)))           # detections: same set (near-equal scores may reorder)
                kg, kw = {key(r) for r in it[0][0]}, {key(r) for r in w[0][0]}
                assert len(kg & kw) >= 0.97 * len(kw)
        rep = prog["_fold_report"][3]
        assert rep["folded_statements"] >= 300 and rep["stopped_at"] is not None, rep     # the backbone, neck and heads fold; the top-k tail does not
        assert br.graph is not None
    finally:
        br.close()


def test_conv_integer_entry_matches_the_pad_shift_convolve_composition():
    """lele_b200_conv_integer (conv2d.rs:2216): padded cells hold raw zeros, i.e. contribute (0 - x_zp)(w - w_zp)."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (2, 4, 9, 7)).astype(np.float32); w = rng.integers(0, 256, (6, 4, 3, 3)).astype(np.float32)
    for pads, strides, xz, wz in (([1, 1, 1, 1], [2, 2], 120.0, 128.0), ([0, 2, 1, 0], [1, 1], 0.0, 7.0), ([0, 0, 0, 0], [1, 2], 3.0, 0.0)):
        got = K.conv_integer(x, w, xz, wz, (1, 1), 1, pads, strides)
        xp = np.pad(x, ((0, 0), (0, 0), (pads[0], pads[2]), (pads[1], pads[3]))) - np.float32(xz)
        want = R.conv2d(xp, w - np.float32(wz), None, (1, 1), 1, (0, 0, 0, 0), strides, 0)
        assert got.shape == want.shape
        close(got, want, atol_frac=1e-6)                            # integer-valued sums below 2^24: exact up to the GEMM's order
    prog, blob, xin = MF.convint_forms(MR)
    for g, r in zip(MR.run_program(prog, blob, xin, MR.CudaOps()), MF.convint_forms_direct(MR, blob, xin)):
        close(g, r, atol_frac=1e-6)


def test_mat_mul_integer_batched_b_and_shape_checks():
    """quantization.rs:1157-1173: b may carry the batch; the side with batch 1 is broadcast.  A K mismatch is refused on the host."""
    rng = np.random.default_rng(8)
    a = rng.integers(0, 256, (3, 5, 12)).astype(np.float32); b = rng.integers(0, 256, (3, 12, 7)).astype(np.float32)
    want = np.einsum("bmk,bkn->bmn", a.astype(np.int64) - 9, b.astype(np.int64) - 130).astype(np.float32)
    np.testing.assert_array_equal(K.mat_mul_integer(a, b, 9.0, 130.0), want)
    np.testing.assert_array_equal(K.mat_mul_integer(a[:1], b, 9.0, 130.0), np.einsum("mk,bkn->bmn", a[0].astype(np.int64) - 9, b.astype(np.int64) - 130).astype(np.float32))
    np.testing.assert_array_equal(K.mat_mul_integer(a, b[0], 9.0, 130.0), np.einsum("bmk,kn->bmn", a.astype(np.int64) - 9, b[0].astype(np.int64) - 130).astype(np.float32))
    with pytest.raises(LeleB200Error):
        K.mat_mul_integer(a, b[:, :11], 0.0, 0.0)
    with pytest.raises(LeleB200Error):
        K.mat_mul_integer(a[:2], b, 0.0, 0.0)


def test_gather_index_out_of_range_is_reported_at_sync():
    """The reference panics on the slice bounds (manipulation.rs:589); the device clamps the access and raises the context's error
    word, which the next lele_b200_sync reports.  In-range negative indices still wrap."""
    x = np.arange(24, dtype=np.float32).reshape(4, 6)
    np.testing.assert_array_equal(K.gather(x, np.array([-1, 0], np.int64), 0), x[[3, 0]])
    with pytest.raises(LeleB200Error, match="out of range"):
        K.gather(x, np.array([1, 4], np.int64), 0)
    with pytest.raises(LeleB200Error, match="out of range"):
        K.gather_elements(x, np.full((4, 2), -7, np.float32), 1)
    np.testing.assert_array_equal(K.gather(x, np.array([2], np.int64), 0), x[[2]])    # the flag was cleared by the failing sync


def test_argmax_treats_both_zeros_as_equal():
    """partial_cmp: -0.0 == +0.0, so a tie between them goes to the LAST index (tokenizer.rs:55)."""
    x = np.array([[0.0, -1.0, -0.0, -2.0], [-0.0, -3.0, 0.0, -0.0]], np.float32)
    np.testing.assert_array_equal(K.argmax_last(x), [2, 3])


def test_comm_entries_single_rank():
    """lele_b200_comm_* with a world of one (the 2..8-rank form runs under torchrun in bench.py): NCCL loads, a communicator forms,
    broadcast is the identity and gather copies the rank's own contribution into slot 0."""
    from lele_b200 import distributed as D                 # (sets LELE_B200_NCCL_LIB to the NCCL torch itself loads)
    ctx = K.default_context()
    ver = D.nccl_version()
    assert ver >= 22000, ver
    uid = (C.c_char * 128)()
    _lib.call("lele_b200_comm_unique_id", uid)
    comm = vp()
    _lib.call("lele_b200_comm_create", ctx.h, uid, C.c_int(1), C.c_int(0), C.byref(comm))
    assert _lib.lib.lele_b200_comm_world(comm) == 1 and _lib.lib.lele_b200_comm_rank(comm) == 0
    src = ctx.to_device(np.arange(32, dtype=np.float32)); dst = ctx.to_device(np.zeros(32, np.float32))
    _lib.call("lele_b200_comm_broadcast", ctx.h, comm, vp(src.ptr), sz(128), C.c_int(0))
    _lib.call("lele_b200_comm_gather", ctx.h, comm, vp(src.ptr), vp(dst.ptr), sz(128), C.c_int(0))
    np.testing.assert_array_equal(dst.numpy(), np.arange(32, dtype=np.float32))
    with pytest.raises(LeleB200Error):
        _lib.call("lele_b200_comm_broadcast", ctx.h, comm, vp(src.ptr), sz(128), C.c_int(3))
    _lib.lib.lele_b200_comm_destroy(comm)


def test_config1_vad_on_the_reference_fixture_on_device():
    """BASELINE configs[0] on the device: the same 175-chunk walk over fixtures/zh.wav (state [2,1,128] carried) replayed through the
    C ABI; probabilities within the f32 bar of the CPU oracle's, identical segment lists."""
    from lele_b200.vad import StreamingVad, collect_segments, merge_segments
    audio = MF.read_wav_s16(os.path.join(ROOT, "tests", "golden", "zh.wav"))
    prog, blob = MF.vad_model(MR, hidden=128)
    cpu = StreamingVad(prog, blob, ops=R, state_shape=(2, 1, 128)); p_cpu = cpu.process(audio)
    gpu = StreamingVad(prog, blob, state_shape=(2, 1, 128)); p_gpu = gpu.process(audio)
    assert p_gpu.shape == (175,) and np.isfinite(p_gpu).all() and (p_gpu >= 0).all() and (p_gpu <= 1).all()
    np.testing.assert_allclose(p_gpu, p_cpu, rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(gpu.state, cpu.state, rtol=2e-3, atol=2e-3)
    seg = lambda p: merge_segments(collect_segments(p, audio.size))
    assert seg(p_gpu) == seg(p_cpu)


@pytest.mark.parametrize("outer,n,k", [(1, 24000, 300), (32, 8400, 300), (3, 1000, 1000), (2, 70000, 2048), (5, 300, 7)])
def test_topk_long_rows_select_and_sort(outer, n, k):
    """conv2d.rs:1385: stable descending sort, ties keep the lower index; the long-row kernel (radix select + ordered tie compaction +
    bitonic sort) against the oracle, on rows full of ties (values drawn from 50 levels), with +-0.0 and -inf present."""
    rng = np.random.default_rng(outer * 7 + k)
    x = rng.integers(0, 50, (outer, n)).astype(np.float32) / np.float32(7.0) - np.float32(3.0)
    x[0, ::97] = 0.0; x[0, 5::97] = -0.0; x[-1, 3] = -np.inf; x[-1, 11] = np.inf
    v, i = K.topk(x, k)
    rv, ri = R.topk(x, k)
    np.testing.assert_array_equal(i, ri)
    np.testing.assert_array_equal(v.view(np.uint32), rv.view(np.uint32))         # the original values, bit for bit (a -0.0 stays -0.0)
    y = rng.standard_normal((outer, n)).astype(np.float32)                       # distinct values
    v, i = K.topk(y, k); rv, ri = R.topk(y, k)
    np.testing.assert_array_equal(i, ri); np.testing.assert_array_equal(v, rv)
