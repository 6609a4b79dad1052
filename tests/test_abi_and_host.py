"""CPU-side tests: the C-ABI library loads and exports every symbol include/lele_b200.h declares
(no compute without a GPU), host logic (blob layout, synthetic PCM, shard plan), and the N>1
path (weights broadcast + ids gather) under gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def so_path():
    from lele_b200 import SO_PATH      # built by tests/conftest.py before the package is imported
    return SO_PATH


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lele_b200.h")).read()
    return sorted(set(re.findall(r"\b(lele_b200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(so_path):
    lib = ctypes.CDLL(so_path)
    names = declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lele_b200_[a-z0-9_]+)", out))
    assert exported == set(names), (sorted(exported - set(names)), sorted(set(names) - exported))   # nothing undeclared leaks either


def test_header_is_strict_c99_and_c_example_runs(so_path, tmp_path):
    """The boundary is a C ABI: include/lele_b200.h must compile as pedantic C99 (no C++-isms), and examples/c/abi_tour.c -- a plain
    C caller -- must build against it, link the library, run its host-only part, and stop cleanly at the device check on a CPU box."""
    exe = str(tmp_path / "abi_tour")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c", "abi_tour.c"), "-L", os.path.dirname(so_path), "-llele_b200", "-lm",
           "-Wl,-rpath," + os.path.dirname(so_path), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.stdout, run.stderr)
    assert "hann(4) = 0.00 0.75 0.75 0.00" in run.stdout and "1598 frames -> 267 LFR rows" in run.stdout
    assert ("no CUDA device" in run.stdout) or ("layer_norm([1,2,3]) = -1.2247" in run.stdout)


def test_library_sass_is_blackwell_native(so_path):
    """What B200_PROFILING.md calls the proof of a Blackwell-native kernel, checked on the built library: tcgen05.mma (UTCIMMA for the
    int8 GEMM, UTCHMMA for the 3xTF32 attention / f32 GEMM), tensor-memory loads, TMA loads and stores -- and no legacy mma.sync."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", so_path], capture_output=True, text=True, timeout=600).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per_kernel[cur] = set()
        elif cur:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                per_kernel[cur].add(m.group(1))
    find = lambda part: [ops for k, ops in per_kernel.items() if part in k]
    assert find("gemm_i8_tc_kernel") and all({"UTCIMMA", "LDTM", "UTMALDG", "UTCBAR"} <= ops for ops in find("gemm_i8_tc_kernel"))
    assert all({"UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"} <= ops for ops in find("attn_tc_kernel")) and find("attn_tc_kernel")
    assert all({"UTCHMMA", "LDTM"} <= ops for ops in find("gemm_tf32x3_nt_kernel")) and len(find("gemm_tf32x3_nt_kernel")) == 3
    assert not any({"HMMA", "IMMA", "HGMMA", "IGMMA"} & ops for ops in per_kernel.values())


def test_rust_ffi_matches_header():
    """bindings/rust/cuda_ffi.rs (the reference-side binding) is generated from include/lele_b200.h:
    the committed file must be what the generator produces now, with one declaration per symbol, the same argument count, and
    `*const` exactly where the C prototype says `const`."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    header = open(os.path.join(ROOT, "include", "lele_b200.h")).read()
    committed = open(os.path.join(ROOT, "bindings", "rust", "cuda_ffi.rs")).read()
    assert committed == g.generate(header), "run python tools/gen_rust_ffi.py"
    decls = {name: (ret, args) for ret, name, args in g.c_declarations(header)}
    assert sorted(decls) == declared_symbols()
    rust = dict(re.findall(r"pub fn (lele_b200_\w+)\((.*)\) -> ", committed))
    assert sorted(rust) == sorted(decls)
    for name, (ret, args) in decls.items():
        rargs = [a for a in rust[name].split(", ") if a]
        assert len(rargs) == len(args), name
        for (ctype, _), ra in zip(args, rargs):
            assert ctype.count("*") == ra.count("*"), (name, ctype, ra)
            if "*" in ctype:                                   # the pointee's constness is the innermost Rust pointer's
                innermost = ra.split(": ")[1].split(" ")[-2]
                assert innermost == ("*const" if ctype.startswith("const ") else "*mut"), (name, ctype, ra)
    assert g.rust_type("const float* const*") == "*const *const f32" and g.rust_type("lele_b200_ctx**") == "*mut *mut Ctx"
    assert g.rust_type("const int32_t*") == "*const i32" and g.rust_type("size_t") == "usize" and g.rust_type("unsigned long long") == "c_ulonglong"


def test_host_only_entry_points(so_path):
    """hann_window / mel_filterbank / frame counting are host functions of the ABI: callable without a GPU
    and identical to the oracle (same f32 libm arithmetic as the reference)."""
    from lele_b200 import features as F
    from oracle import reference_api as R
    np.testing.assert_array_equal(F.hann_window(400), R.hann_window(400))
    np.testing.assert_array_equal(F.mel_filterbank(16000.0, 512, 80, 20.0), R.mel_filterbank(16000.0, 512, 80, 20.0))
    assert F.hann_window(1)[0] == 1.0 and F.hann_window(0).size == 0
    lib = ctypes.CDLL(so_path)
    assert lib.lele_b200_frontend_num_frames(256000) == 1598 and lib.lele_b200_frontend_out_rows(256000) == 267
    assert lib.lele_b200_frontend_num_frames(399) == 0 and lib.lele_b200_frontend_num_frames(89472) == 557


def test_compute_fails_loudly_without_gpu(so_path):
    import lele_b200
    if lele_b200.lib.lele_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(lele_b200.LeleB200Error):
        lele_b200.kernels.add(np.ones(4, np.float32), np.ones(4, np.float32))
    with pytest.raises(lele_b200.LeleB200Error):
        lele_b200.features.SenseVoiceFrontend().compute(np.zeros(16000, np.float32))


def test_product_never_imports_oracle():
    """The product path may not route through the CPU oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "lele_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", "").replace("oracle/sensevoice_ref.c", ""), f"{f} references the oracle"


def test_blob_layout_and_synth():
    from lele_b200.sensevoice_weights import SenseVoiceConfig, blob_nbytes, build_blob, synth_pcm
    cfg = SenseVoiceConfig(n_layers=2, vocab=300, n_stage1=1, max_t=32)
    blob = build_blob(cfg, seed=1)
    assert blob.nbytes == blob_nbytes(cfg)
    assert blob_nbytes(SenseVoiceConfig()) > 230e6            # ~233 MB u8 + f32 vectors
    hdr = blob[:256].view(np.int32)
    assert hdr[0] == 0x454C454C and hdr[12] == 10 + 2 * 21
    table = blob[256:256 + 16 * int(hdr[12])].view(np.uint64).reshape(-1, 2)
    assert (table[:, 0] % 256 == 0).all() and int(table[-1, 0] + table[-1, 1]) <= blob.nbytes
    assert int(table[2 + 10, 1]) == 560 * 1536                 # layer-0 qkv weight is [560, 1536] u8
    np.testing.assert_array_equal(build_blob(cfg, seed=1), blob)
    a = synth_pcm(3, 4000)
    np.testing.assert_array_equal(a, synth_pcm(3, 4000))
    assert np.abs(a).max() < 0.111 and not np.array_equal(a, synth_pcm(4, 4000))
    np.testing.assert_array_equal(synth_pcm(3, 8000)[:4000], a)   # LCG jump-ahead is prefix-stable
    from oracle.binding import SenseVoiceRef
    assert SenseVoiceRef(blob).vocab == 300


def test_shard_range_partition():
    from lele_b200.distributed import shard_range
    for n, w in [(512, 8), (64, 1), (10, 4), (3, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from lele_b200.distributed import broadcast_blob, gather_ids, shard_range, max_over_ranks
from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, blob_nbytes
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = SenseVoiceConfig(n_layers=1, vocab=64, n_stage1=1, max_t=16)
nb = blob_nbytes(cfg)
blob = torch.from_numpy(build_blob(cfg, seed=5)) if rank == 0 else torch.zeros(nb, dtype=torch.uint8)
broadcast_blob(blob, 0)
assert np.array_equal(blob.numpy(), build_blob(cfg, seed=5)), "broadcast mismatch"
n_total, T = 7, 5                       # ragged: shards of 4 and 3 clips
s, e = shard_range(n_total, rank, world)
local = torch.arange(s * T, e * T, dtype=torch.int32).reshape(e - s, T)   # stands in for this rank's greedy ids
allids = gather_ids(local, n_total, 0)
if rank == 0:
    assert allids.shape == (n_total, T) and torch.equal(allids.reshape(-1), torch.arange(n_total * T, dtype=torch.int32))
else:
    assert allids is None
assert max_over_ranks(float(rank + 1)) == float(world)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_broadcast_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_bench_reference_arm_cli_contract():
    """bench.py --impl reference must parse and (without running the 10 s/clip workload here) expose the flags."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "--impl" in r.stdout and "--gpus" in r.stdout and "--steps" in r.stdout and "--warmup" in r.stdout


def test_tokenizer_host_logic(tmp_path):
    """Tokenizer.from_file / skip_mask / text follow examples/sensevoice/src/tokenizer.rs:10-82 (split at the LAST space,
    blank + "<|...|>" tokens skipped, sentencepiece underscore -> space, trim); checked against the oracle restatement."""
    from lele_b200.tokenizer import Tokenizer
    from oracle import np_ops as N
    toks = ["<blank>", "<|zh|>", "\u2581he", "llo", "\u2581wor ld", "<|NEUTRAL|>", "!", "<|", "|>"]
    f = tmp_path / "tokens.txt"
    f.write_text("".join(f"{t} {i}\n" for i, t in enumerate(toks)) + "malformed_line_without_id\n", encoding="utf-8")
    tk = Tokenizer.from_file(str(f))
    assert tk.id_to_token == toks and tk.vocab_size() == len(toks)
    assert tk.skip_mask().tolist() == [1, 1, 0, 0, 0, 1, 0, 0, 0]
    rng = np.random.default_rng(3)
    logits = rng.standard_normal((3, 17, len(toks))).astype(np.float32)
    logits[0, 5, 2] = logits[0, 5, 6] = 9.0                      # tie: the LAST maximum wins (tokenizer.rs:55)
    ids = logits.shape[2] - 1 - np.argmax(logits[:, :, ::-1], axis=2)
    kept = N.greedy_filter(ids, tk.skip_mask())
    assert [tk.text(k) for k in kept] == N.decode_greedy(logits, toks)
    assert tk.text([2, 3, 4, 6]) == "hello wor ld!"


def test_model_rs_parser_and_replay_on_oracle():
    """lele_b200/model_rs.py: the statement forms lele_gen emits (src/compiler/generate.rs:802-997) parse into a program and
    replay against the shared operator vocabulary; here on the CPU oracle with a tiny generated-style body."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("model_rs", os.path.join(ROOT, "lele_b200", "model_rs.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    from oracle import reference_api as R
    text = """
pub struct TinyWorkspace { pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, }
pub struct Tiny<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut TinyWorkspace, images: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {
        let a = lele::kernels::conv2d_silu(&images, &self.weight_f32(0, 432, &[4, 3, 3, 3]), Some(&self.weight_f32(432, 16, &[4])), &[1, 1], 1, &[1, 1, 1, 1], &[2, 2], &mut ws.buf_0);
        let splits_slice = &[2, 2];
        let mut split_results = lele::kernels::split_owned(&a, 1, splits_slice);
        let a1 = split_results.swap_remove(1);
        let a0 = split_results.swap_remove(0);
        let b = lele::kernels::add(&a0, &a1, &mut ws.buf_1);
        let c = lele::kernels::concat(&[&a0, &b], 1, &mut ws.buf_0);
        let d = lele::kernels::resize_nearest(&c, Some(&self.weight_f32(448, 16, &[4]).data), None, "asymmetric", &mut ws.buf_1);
        let e = lele::kernels::reshape(&d, &[1, 4, -1]);
        let mut buf_v = Vec::<f32>::new();
        let mut buf_i = Vec::<f32>::new();
        let (tv, ti) = lele::kernels::topk(&e, self.weight_i64(464, 8, &[1]).data[0] as usize, -1, true, true, &mut buf_v, &mut buf_i);
        let f = tv.clone(); // Cast f32->f32 is no-op
        (f.to_owned(), ti.to_owned())
    }
"""
    prog = m.parse_model_rs(text)
    assert prog["class"] == "Tiny" and prog["inputs"] == ["images"] and prog["outputs"] == ["f", "ti"] and prog["workspace_buffers"] == 2
    assert [st["op"] for st in prog["statements"]] == ["conv2d_silu", "split_take", "split_take", "add", "concat", "resize_nearest", "reshape", "topk", "identity"]
    blob = m.synth_blob(prog, 1, {448: [1, 1, 2, 2], 464: [5]})
    assert len(blob) == 472
    x = np.random.default_rng(0).random((1, 3, 8, 8), dtype=np.float32)
    f, ti = m.run_program(prog, blob, [x], R)
    w = m.weight_view(blob, "weight_f32", 0, 432, [4, 3, 3, 3]); b = m.weight_view(blob, "weight_f32", 432, 16, [4])
    a = R.conv2d(x, w, b, (1, 1), 1, (1, 1, 1, 1), (2, 2), 2)
    c = np.concatenate([a[:, :2], a[:, :2] + a[:, 2:]], 1)
    e = np.repeat(np.repeat(c, 2, 2), 2, 3).reshape(1, 4, -1)
    want_v, want_i = R.topk(e, 5)
    np.testing.assert_array_equal(f, want_v); np.testing.assert_array_equal(ti, want_i)
    with pytest.raises(ValueError):
        m.parse_model_rs(text.replace("let b = lele::kernels::add(&a0, &a1, &mut ws.buf_1);", "let b = unsafe { transmute(a0) };"))


def _model_rs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("model_rs", os.path.join(ROOT, "lele_b200", "model_rs.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_model_rs_more_statement_forms_on_oracle():
    """More of the forms src/compiler/ops/*.rs emits (layer_norm, gemm, conv1d, unary / binary math, reductions, pad, expand,
    squeeze, where_op): parsed and replayed on the oracle, checked against direct oracle calls."""
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob, x = MF.math_forms(m)
    s_out, e_out = m.run_program(prog, blob, [x], MF.R)
    s_ref, e_ref = MF.math_forms_direct(m, blob, x)
    np.testing.assert_array_equal(e_out, e_ref)
    np.testing.assert_array_equal(s_out, s_ref)
    assert s_out.shape == (3, 9)


def test_model_rs_recurrent_and_tuple_forms_on_oracle():
    """Multi-output statements of ops/nn.rs: LSTM (Y, H, C), GRU with and without the `_` placeholder, LayerNormalization
    with unused extra outputs, and the empty-TensorView scale that the x86 kernel cannot take."""
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob, x = MF.recurrent_forms(m)
    assert [s["outs"] for s in prog["statements"]] == [["y", "yh", "yc"], ["y2"], ["g", "_"], ["n"]]
    got = m.run_program(prog, blob, [x], MF.R)
    for a, b in zip(got, MF.recurrent_forms_direct(m, blob, x)):
        np.testing.assert_array_equal(a, b)
    bad = MF.recurrent_text().replace(f"&self.weight_f32(4000, {MF.H*4}, &[{MF.H}])", "&lele::tensor::TensorView::empty()")
    with pytest.raises(ValueError, match="scale and bias are required"):
        m.run_program(m.parse_model_rs(bad), blob, [x], MF.R)


def test_model_rs_quantised_forms_on_oracle():
    """The int8 path of a generated model: cfg(aarch64) / cfg(not(aarch64)) statement pairs, self.linear_quantized[_relu],
    self.layer_norm, DynamicQuantizeLinear tuple + to_owned, MatMulInteger with scalar zero-point tensors, clip, self.linear."""
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob, x = MF.quant_forms(m)
    ops_seen = [s["op"] for s in prog["statements"]]
    assert "self.linear_quantized_relu_arm" not in ops_seen and "self.mat_mul_integer_arm" not in ops_seen
    assert ops_seen.count("identity") == 3 and prog["workspace_buffers"] == 2
    got = m.run_program(prog, blob, [x], MF.R)
    ref = MF.quant_forms_direct(m, blob, x)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)
    assert got[0].shape == (2, 5, 8) and np.abs(got[1]).max() > 0


def test_model_rs_constants_stft_and_libm_math_on_oracle():
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob, x = MF.const_forms(m)
    assert [s["op"] for s in prog["statements"]][:3] == ["constant", "literal", "literal"]
    got = m.run_program(prog, blob, [x], MF.R)
    for a, b in zip(got, MF.const_forms_direct(m, blob, x)):
        np.testing.assert_array_equal(a, b)
    assert got[0].shape == (22, 33) and not got[1].any()          # (400 - 64) / 16 + 1 frames x 33 bins; not(x) never equals x


def test_model_rs_multi_chunk_model():
    """A graph split into run_chunk_0 / run_chunk_1 (generate.rs:704): statements are the chunks in order, graph inputs and outputs
    come from forward_with_workspace -- here the second chunk consumes a graph input the first never sees, and a chunk with a
    single live output returns it without a tuple."""
    from tests import model_forms as MF
    m = _model_rs()
    text = """
pub struct T7Workspace { pub buf_0: Vec<f32>, }
pub struct T7<'a> { data: &'a [u8] }
    #[inline(never)]
    fn run_chunk_0<'w>(&self, ws: &'w mut T7Workspace, x: TensorView<'w, f32>) -> TensorView<'static, f32> {
        let a = lele::kernels::relu(&x, &mut ws.buf_0);
        a.to_owned()
    }

    #[inline(never)]
    fn run_chunk_1<'w>(&self, ws: &'w mut T7Workspace, a: TensorView<'w, f32>, y: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {
        let b = lele::kernels::add(&a, &y, &mut ws.buf_0);
        let c = lele::kernels::mul(&b, &self.weight_f32(0, 12, &[3]), &mut ws.buf_0);
        (b.to_owned(), c.to_owned())
    }

    pub fn forward_with_workspace<'w>(&self, ws: &'w mut T7Workspace, x: TensorView<'w>, y: TensorView<'w>) -> (TensorView<'w>, TensorView<'w>) {
        let (a) = self.run_chunk_0(ws, x);
        let (b, c) = self.run_chunk_1(ws, a, y);
        (c, b)
    }
}
"""
    prog = m.parse_model_rs(text)
    assert prog["inputs"] == ["x", "y"] and prog["outputs"] == ["c", "b"] and [s["op"] for s in prog["statements"]] == ["relu", "add", "mul"]
    blob = m.synth_blob(prog, 1)
    x = np.array([[-1.0, 2.0, 3.0]], np.float32); y = np.array([[10.0, 20.0, 30.0]], np.float32)
    c, b = m.run_program(prog, blob, [x, y], MF.R)
    np.testing.assert_array_equal(b, np.maximum(x, 0) + y)
    np.testing.assert_array_equal(c, b * m.weight_view(blob, "weight_f32", 0, 12, [3]))


def test_vad_segment_logic_known_answers():
    """Segment state machine and merge pass of examples/silero/src/main.rs:155-229, against hand-computed answers
    (16 kHz, 512-sample chunks: min silence 3200, min speech 6400, pad 1920, merge gap 3200 samples)."""
    from lele_b200.vad import VadConfig, collect_segments, merge_segments, ms_to_samples
    assert [ms_to_samples(ms, 16000) for ms in (200.0, 400.0, 120.0)] == [3200, 6400, 1920]
    assert ms_to_samples(0.03125, 16000) == 1 and ms_to_samples(0.03, 16000) == 0          # 0.5 rounds away from zero
    assert ms_to_samples(120.0, 8000) == 960
    hi, lo = 0.9, 0.1
    # speech in chunks 10..29, then silence: released once 7 silent chunks (3584 >= 3200) have passed, at frame_end = 37 * 512
    probs = [lo] * 10 + [hi] * 20 + [lo] * 20
    n = 50 * 512 - 100
    assert collect_segments(probs, n) == [(10 * 512 - 1920, 37 * 512 + 1920)]
    # a burst shorter than min_speech is dropped: 3 chunks + 7 release chunks + pads = 1920 + 10*512 + 1920 = 8960 >= 6400 -> kept;
    # with pad 0 it is 10 * 512 = 5120 < 6400 -> dropped
    probs = [lo] * 5 + [hi] * 3 + [lo] * 30
    assert collect_segments(probs, 38 * 512) == [(5 * 512 - 1920, 15 * 512 + 1920)]
    assert collect_segments(probs, 38 * 512, config=VadConfig(speech_pad_ms=0.0)) == []
    # the threshold is inclusive, a short dip (6 chunks < min silence) does not release, start saturates at 0
    probs = [0.3] * 4 + [lo] * 6 + [0.3] * 10 + [lo] * 20
    assert collect_segments(probs, 40 * 512) == [(0, 27 * 512 + 1920)]
    # still triggered at the end of the clip: the segment ends at the unpadded length; the release pad is clamped to it too
    probs = [lo] * 5 + [hi] * 20
    assert collect_segments(probs, 25 * 512 - 7) == [(5 * 512 - 1920, 25 * 512 - 7)]
    probs = [lo] * 3 + [hi] * 20 + [lo] * 7
    assert collect_segments(probs, 30 * 512 - 300) == [(0, 30 * 512 - 300)]            # 3 * 512 - 1920 saturates at 0
    # merge: overlap, gap <= 3200 joined, larger gap kept apart, input order irrelevant
    assert merge_segments([(20000, 30000), (0, 10000), (9000, 12000), (15200, 16000), (40000, 41000)]) == [(0, 16000), (20000, 30000), (40000, 41000)]
    assert merge_segments([(0, 10), (3211, 3300)]) == [(0, 10), (3211, 3300)] and merge_segments([(0, 10), (3210, 3300)]) == [(0, 3300)]
    assert merge_segments([]) == []


def test_streaming_vad_carries_state_across_chunks():
    """StreamingVad over a replayed recurrent model (synthetic weights, Silero's calling convention): chunk loop, x32768 scaling, zero
    padding of the last chunk and the (h, c) hand-over, against the same steps written as direct oracle calls."""
    from lele_b200 import model_rs as MR
    from lele_b200.vad import StreamingVad
    from tests import model_forms as MF
    prog, blob = MF.vad_model(MR)
    assert prog["inputs"] == ["input", "state", "sr"] and prog["outputs"] == ["output", "stateN"]
    rng = np.random.default_rng(33)
    audio = (0.1 * rng.standard_normal(512 * 3 + 200)).astype(np.float32)
    vad = StreamingVad(prog, blob, ops=MF.R, state_shape=(2, 1, MF.VH))
    probs = vad.process(audio)
    assert probs.shape == (4,) and np.all((probs > 0) & (probs < 1))
    padded = np.zeros(4 * 512, np.float32); padded[:audio.size] = audio
    state = np.zeros((2, 1, MF.VH), np.float32); want = []
    for i in range(4):
        p, state = MF.vad_chunk_direct(MR, blob, padded[i * 512:(i + 1) * 512], state)
        want.append(p)
    np.testing.assert_array_equal(probs, np.asarray(want, np.float32))
    np.testing.assert_array_equal(vad.state, state)
    assert len(set(np.round(probs, 6))) > 1                        # the state matters: chunks are not scored independently
    fresh = StreamingVad(prog, blob, ops=MF.R, state_shape=(2, 1, MF.VH))
    assert fresh.push(padded[512:1024]) != probs[1]
    with pytest.raises(ValueError, match="expected 512"):
        vad.push(np.zeros(100, np.float32))
    segs = vad.segments(audio)
    assert isinstance(segs, list) and len(vad.probs) == 4          # segments() restarts the stream


def test_model_rs_if_blocks_and_embedding_concat():
    """ONNX If as emitted by ops/control_flow.rs (Silero's sample-rate switch is one): the first element of the condition picks the
    branch, branches may nest and may yield stored tensors (`self.weight(...)`); plus the ConstantOfShape + Concat helper."""
    from tests import model_forms as MF
    m = _model_rs()
    text = """
pub struct T8Workspace { pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, }
pub struct T8<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T8Workspace, sr: TensorView<'w, i64>, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>, TensorView<'static, f32>) {
        let mut buf_is16 = Vec::<i64>::new();
        let is16 = lele::kernels::equal_i64(&sr, &self.weight_i64(0, 8, &[]), &mut buf_is16);
        let (y, z) = if is16.data.get(0).map(|v| *v != 0).unwrap_or(false) {
            let a = lele::kernels::mul(&x, &self.weight_f32(8, 4, &[1]), &mut ws.buf_0);
            let (p) = if a.data.get(0).map(|v| *v != 0.0).unwrap_or(false) {
                let q = lele::kernels::relu(&a, &mut ws.buf_1);
                (q.to_owned())
            } else {
                (a.to_owned())
            };
            (p.to_owned(), self.weight(12, 12, &[3]).to_owned())
        } else {
            let b = lele::kernels::neg(&x, &mut ws.buf_0);
            (b.to_owned(), x.to_owned())
        };
        let e = self.embedding_concat(&self.weight_i64(24, 16, &[2]), 0.5, self.weight_f32(40, 24, &[2, 3]), &mut ws.buf_1);
        (y.to_owned(), z.to_owned(), e.to_owned())
    }

    pub fn forward_with_workspace<'w>(&self, ws: &'w mut T8Workspace, x: TensorView<'w>, sr: TensorView<'w, i64>) -> (TensorView<'w>, TensorView<'w>, TensorView<'w>) {
        let (y, z, e) = self.run_chunk_0(ws, sr, x);
        (y, z, e)
    }
}
"""
    prog = m.parse_model_rs(text)
    st_if = prog["statements"][1]
    assert st_if["op"] == "if" and st_if["outs"] == ["y", "z"] and st_if["then"]["statements"][1]["op"] == "if"
    assert st_if["then"]["outputs"][1] == {"weight": ["weight_f32", 12, 12, [3]]}
    blob = m.synth_blob(prog, 2, {0: [16000], 8: [2.0], 12: [7.0, 8.0, 9.0], 24: [1, 3], 40: [1, 2, 3, 4, 5, 6]})
    x = np.array([-1.0, 2.0, -3.0], np.float32)
    y, z, e = m.run_program(prog, blob, [x, np.array([16000], np.int64)], MF.R)
    np.testing.assert_array_equal(y, [0.0, 4.0, 0.0]); np.testing.assert_array_equal(z, [7.0, 8.0, 9.0])       # then / then
    np.testing.assert_array_equal(e, [[1, 2, 3], [4, 5, 6], [0.5, 0.5, 0.5]])
    y, z, _ = m.run_program(prog, blob, [np.array([0.0, 2.0, -3.0], np.float32), np.array([16000], np.int64)], MF.R)
    np.testing.assert_array_equal(y, [0.0, 4.0, -6.0])                                                           # then / else (a[0] == 0)
    y, z, _ = m.run_program(prog, blob, [x, np.array([8000], np.int64)], MF.R)
    np.testing.assert_array_equal(y, [1.0, -2.0, 3.0]); np.testing.assert_array_equal(z, x)                      # else
    with pytest.raises(ValueError, match="one value per output"):
        m.parse_model_rs(text.replace("(b.to_owned(), x.to_owned())", "(b.to_owned())"))


def test_e2e_golden_checks_skip_and_compare(tmp_path):
    """lele_b200.e2e mirrors examples/*/tests/e2e_test.rs: SKIP (None) while a weights / .npy file is missing, the reference's
    tolerances once they exist.  The goldens here are synthetic (written by the oracle replay, then perturbed)."""
    from lele_b200 import e2e, model_rs as MR
    from tests import model_forms as MF
    d = str(tmp_path)
    # --- silero-shaped
    prog, blob = MF.vad_model(MR)
    open(os.path.join(d, "silerovad.rs"), "w").write(MF.vad_text()); open(os.path.join(d, "silerovad_weights.bin"), "wb").write(blob)
    args = (os.path.join(d, "silerovad.rs"), os.path.join(d, "silerovad_weights.bin"), [os.path.join(d, "missing"), d])
    assert e2e.check_silero(*args, ops=MF.R) is None                                       # fixtures absent -> SKIP
    x = (1000 * np.random.default_rng(4).standard_normal((1, 512))).astype(np.float32); st = np.zeros((2, 1, MF.VH), np.float32)
    out, st_out = MR.run_program(prog, blob, [x, st, np.array([16000], np.int64)], MF.R)
    for n, v in (("silero_input", x), ("silero_state_in", st), ("silero_sr", np.array([16000], np.int64)), ("silero_output", out + 5e-5), ("silero_state_out", st_out)):
        np.save(os.path.join(d, n + ".npy"), v)
    rep = e2e.check_silero(*args, ops=MF.R)
    assert rep["output_diff"] < 1e-4 and rep["state_max_diff"] == 0.0
    np.save(os.path.join(d, "silero_output.npy"), out + 1e-3)
    with pytest.raises(AssertionError, match="silero output diff"):
        e2e.check_silero(*args, ops=MF.R)
    # --- sensevoice-shaped: logits MAE <= 1.0 and at least one arg-max frame in common
    text = """
pub struct SvWorkspace { pub buf_0: Vec<f32>, }
pub struct Sv<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut SvWorkspace, x: TensorView<'w, f32>) -> TensorView<'static, f32> {
        let logits = lele::kernels::matmul(&x, &self.weight_f32(0, 15680, &[560, 7]), &mut ws.buf_0);
        logits.to_owned()
    }

    pub fn forward_with_workspace<'w>(&self, ws: &'w mut SvWorkspace, x: TensorView<'w>, x_length: TensorView<'w, i64>, language: TensorView<'w, i64>, text_norm: TensorView<'w, i64>) -> TensorView<'w> {
        let (logits) = self.run_chunk_0(ws, x);
        logits
    }
}
"""
    prog = MR.parse_model_rs(text); blob = MR.synth_blob(prog, 3)
    assert prog["inputs"] == ["x", "x_length", "language", "text_norm"] and prog["outputs"] == ["logits"]
    open(os.path.join(d, "sensevoice.rs"), "w").write(text); open(os.path.join(d, "sensevoice_weights.bin"), "wb").write(blob)
    x = np.random.default_rng(5).standard_normal((1, 10, 560)).astype(np.float32)
    logits = MR.run_program(prog, blob, [x, np.array([10]), np.array([0]), np.array([15])], MF.R)[0]
    np.save(os.path.join(d, "sensevoice_input_x.npy"), x)
    for n, v in (("x_length", 10), ("language", 0), ("text_norm", 15)):
        np.save(os.path.join(d, f"sensevoice_input_{n}.npy"), np.array([v], np.int32))
    args = (os.path.join(d, "sensevoice.rs"), os.path.join(d, "sensevoice_weights.bin"), [d])
    assert e2e.check_sensevoice(*args, ops=MF.R, vocab=7) is None                          # logits golden still missing
    np.save(os.path.join(d, "sensevoice_logits.npy"), logits + 0.25)
    rep = e2e.check_sensevoice(*args, ops=MF.R, vocab=7)
    assert abs(rep["mae"] - 0.25) < 1e-4 and rep["argmax_match"] == rep["frames"] == 10
    np.save(os.path.join(d, "sensevoice_logits.npy"), logits + 1.5)
    with pytest.raises(AssertionError, match="mae"):
        e2e.check_sensevoice(*args, ops=MF.R, vocab=7)
    assert e2e.check_yolo26(os.path.join(d, "yolo26.rs"), os.path.join(d, "yolo26_weights.bin"), [d]) is None


def test_model_rs_conv_integer_pads_with_raw_zeros():
    """ConvInteger as the reference computes it (conv2d.rs:1507-2025): f32 convolution of (x - x_zp) and (w - w_zp) where the padding
    is applied to the raw tensor, i.e. a padded position contributes (0 - x_zp).  Checked against a direct float64 evaluation."""
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob, xs = MF.convint_forms(m)
    y = m.run_program(prog, blob, xs, MF.R)[0]
    assert y.shape == (1, 3, 3, 3)
    np.testing.assert_array_equal(y, MF.convint_forms_direct(m, blob, xs)[0])          # integer-valued and far below 2^24: exact in f32


def test_generated_model_object(tmp_path):
    """GeneratedModel = the generated struct's `new(&bin)` + `forward`: file naming of the compiler, input arity, blob size check."""
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob = MF.vad_model(m)
    rs = tmp_path / "synthvad.rs"; rs.write_text(MF.vad_text()); (tmp_path / "synthvad_weights.bin").write_bytes(blob)
    model = m.GeneratedModel.from_files(str(rs), ops=MF.R)
    assert model.class_name == "SynthVad" and model.input_names == ["input", "state", "sr"] and model.output_names == ["output", "stateN"]
    x = (500 * np.random.default_rng(1).standard_normal((1, 512))).astype(np.float32); st = np.zeros((2, 1, MF.VH), np.float32)
    out, st2 = model(x, st, np.array([16000], np.int64))
    want = m.run_program(prog, blob, [x, st, np.array([16000], np.int64)], MF.R)
    np.testing.assert_array_equal(out, want[0]); np.testing.assert_array_equal(st2, want[1])
    with pytest.raises(ValueError, match="takes 3 tensors"):
        model.forward(x)
    with pytest.raises(ValueError, match="reads up to byte"):
        m.GeneratedModel(MF.vad_text(), blob[:-4], MF.R)
    assert m._blob_extent(prog) == len(blob)


def test_generated_model_prepares_int8_weights_once():
    """The per-model cache: weight views are decoded once, and a namespace that offers prepare_weights (CudaOps does) is asked to
    pack each quantised linear's weight exactly once across forwards -- the B_WEIGHT_CACHE behaviour (avx/quantization.rs:47-95)."""
    import inspect
    from lele_b200 import kernels as K, model_rs as MR
    from tests import model_forms as MF
    prog, blob, x = MF.quant_forms(MR)
    made = []

    class PackingOps(MR._NamespaceOps):
        def prepare_weights(self, w, ws, wz, bias):
            made.append((np.asarray(w).shape, int(wz)))
            return ("packed", np.asarray(w), ws, wz, bias)

        def fused_quantized_linear(self, x_, w, ws, wz, bias, relu):
            if isinstance(w, tuple):
                _, w, ws, wz, bias = w
            return self.ns.fused_quantized_linear(x_, w, ws, wz, bias, relu)

    model = MR.GeneratedModel(MF.QUANT_TEXT, blob, PackingOps(MF.R))
    first = model(x)
    for _ in range(2):
        again = model(x)
        for a, b in zip(first, again):
            np.testing.assert_array_equal(a, b)
    for a, b in zip(first, MF.quant_forms_direct(MR, blob, x)):
        np.testing.assert_array_equal(a, b)
    assert made == [((16, 24), 128), ((24, 16), 121)]                     # two linears, three forwards, two packings
    assert sum(1 for k in model._cache if k[0] == "w") >= 8
    # the real namespace builds lele_b200.kernels.PreparedWeights with arguments its constructor accepts
    sig = inspect.signature(K.PreparedWeights.__init__)
    sig.bind(None, np.zeros((4, 4), np.uint8), np.ones(1, np.float32), 3, None, None)
    assert "prepare_weights" in MR.CudaOps.__dict__


RESHAPE_CASES = [  # (input shape, target, expected) -- shape.rs:2-93; None = the reference panics
    ([2, 2], [4], [4]), ([2, 2], [1, -1], [1, 4]),                      # src/kernels/shape.rs:194-203
    ([2, 3], [-1, 2], [3, 2]), ([2, 2, 3], [2, -1, 2], [2, 3, 2]),      # tests/regression_kernels.rs:934-945
    ([2, 3, 4], [0, 12], [2, 12]), ([2, 3, 4], [0, 0, -1], [2, 3, 4]),  # pass 1: 0 copies the input dim
    ([4, 6], [0, 3], [8, 3]),                                           # pass 2: [4, 3] has the wrong count, so 0 is re-read as -1
    ([1, 64, 601], [1, 96000, 64, 601], [1, 1, 64, 601]),               # pass 3: the example in the reference's comment (shape.rs:26)
    ([6], [2, 5, 3], [2, 1, 3]),                                        # pass 3 on a rank-1 input: [first, -1] + last 0 dims ... = [2, -1] fails, then panic?
    ([6], [4], None), ([2, 3], [-1, -1], None), ([2, 3], [0, 0, 0], None),
]


def test_reshape_and_squeeze_follow_the_reference_rules(so_path):
    """Host shape logic in three places (oracle, lele_b200.kernels, model_rs replay) against the rules of shape.rs."""
    from lele_b200 import LeleB200Error, kernels as K, model_rs as MR
    from oracle import reference_api as R
    for ish, tgt, want in RESHAPE_CASES:
        x = np.arange(int(np.prod(ish)), dtype=np.float32).reshape(ish)
        if ish == [6] and tgt == [2, 5, 3]:
            want = None                                              # rank-1 input: the collapsed target is [2, -1] -> [2, 3], see below
            assert MR.resolve_reshape(ish, tgt) == [2, 3]
            assert list(R.reshape(x, tgt).shape) == [2, 3] and list(K.reshape(x, tgt).shape) == [2, 3]
            continue
        if want is None:
            with pytest.raises(ValueError, match="element count mismatch"):
                MR.resolve_reshape(ish, tgt)
            with pytest.raises(ValueError, match="element count mismatch"):
                R.reshape(x, tgt)
            with pytest.raises(LeleB200Error, match="element count mismatch"):
                K.reshape(x, tgt)
            continue
        assert MR.resolve_reshape(ish, tgt) == want, (ish, tgt)
        for got in (R.reshape(x, tgt), K.reshape(x, tgt), MR._reshape(x, tgt)):
            assert list(got.shape) == want
            np.testing.assert_array_equal(got.reshape(-1), x.reshape(-1))
    x = np.zeros((1, 3, 1, 2), np.float32)
    for axes, want in ((None, [3, 2]), ([], [1, 3, 1, 2]), ([0], [3, 1, 2]), ([0, 1], [3, 1, 2]), ([-2, 0], [3, 2]), ([1, 3], [1, 3, 1, 2])):
        assert MR.squeeze_shape(x.shape, axes) == want, axes
        assert list(R.squeeze(x, axes).shape) == want and list(K.squeeze(x, axes).shape) == want
    v = np.zeros(3, np.float32)
    for axes, want in (([0], [1, 3]), ([-1], [3, 1]), ([-1, -2], [3, 1, 1]), ([0, -1], [1, 3, 1]), ([1, 2], [3, 1, 1]), ([0, 1], [1, 1, 3]), ([5], [3, 1])):
        assert MR.unsqueeze_shape(v.shape, axes) == want, axes
        assert list(R.unsqueeze(v, axes).shape) == want and list(K.unsqueeze(v, axes).shape) == want
        if max(axes) < 3:
            assert np.expand_dims(v, tuple(axes)).shape == tuple(want)     # agrees with numpy wherever numpy accepts the axes
    assert list(K.flatten(np.zeros((2, 3, 4)), -1).shape) == [6, 4] and list(R.flatten(np.zeros((2, 3, 4)), -1).shape) == [6, 4]
    for ish, tgt, want in (([3, 1], [3, 4], [3, 4]), ([1, 1, 3], [1, 2, 3], [1, 2, 3]), ([3], [2, 1], [2, 3]), ([2, 3], [0, 0], [2, 3]), ([2, 1], [0, 5], [2, 5]),
                           ([4], [1], [4]), ([2, 3], [3, 1, 1], [3, 2, 3])):                                           # math.rs:2175-2204
        assert MR.expand_shape(ish, tgt) == want, (ish, tgt)
        assert list(R.expand(np.zeros(ish, np.float32), tgt).shape) == want
    with pytest.raises(ValueError, match="Expand: incompatible dimensions at dim index 1"):
        MR.expand_shape([2, 3], [2, 4])
    with pytest.raises(ValueError, match="Expand: incompatible dimensions"):
        R.expand(np.zeros((2, 3), np.float32), [2, 4])
    with pytest.raises(LeleB200Error, match="Expand: incompatible dimensions"):
        K.expand(np.zeros((2, 3), np.float32), [2, 4], ctx=object())
    one = np.ones((1, 1), np.float32)                                   # src/kernels/shape.rs:214-223
    assert R.squeeze(one, None).shape == () and K.squeeze(one).shape == () and K.unsqueeze(K.squeeze(one), [0]).shape == (1,)


def test_precondition_failures_raise_before_any_device_work(so_path):
    """Where the reference panics on a violated precondition, the mirror raises LeleB200Error with the same message -- from host code,
    before a context or a device buffer is touched (so this runs on the CPU box; `ctx=object()` proves no context is used)."""
    from lele_b200 import LeleB200Error, kernels as K
    from oracle import reference_api as R
    z = lambda *s: np.zeros(s, np.float32)
    nothing = object()
    cases = [
        ("splits sum mismatch", lambda: K.split(z(2, 6), 1, [2, 2], ctx=nothing)),                        # manipulation.rs:1173
        ("axis out of bounds", lambda: K.split(z(2, 6), 2, [6], ctx=nothing)),                             # manipulation.rs:1169
        ("element count mismatch", lambda: K.reshape(z(2, 3), [4])),                                       # shape.rs:48
        ("repeats length must match input rank", lambda: K.tile(z(2, 3), [2], ctx=nothing)),              # math.rs:2256
        ("sizes H and W must be positive", lambda: K.resize_nearest(z(1, 1, 1, 1), None, [1, 1, -1, 10], ctx=nothing)),   # conv2d.rs:3619 (should_panic test)
        ("scales must be positive", lambda: K.resize_nearest(z(1, 1, 2, 2), [1, 1, 0.0, 2.0], ctx=nothing)),             # conv2d.rs:1312
        ("either scales or sizes", lambda: K.resize_nearest(z(1, 1, 2, 2), ctx=nothing)),                                 # conv2d.rs:1318
        ("output dimensions must be positive", lambda: K.resize_nearest(z(1, 1, 2, 2), [1, 1, 0.25, 1.0], ctx=nothing)),  # conv2d.rs:1323
        ("conv2d: output dimensions must be positive", lambda: K.conv2d(z(1, 1, 2, 2), z(1, 1, 3, 3), ctx=nothing)),       # conv2d.rs:288
        ("Concat: ranks mismatch", lambda: K.concat([z(2, 3), z(3)], 0, ctx=nothing)),                     # manipulation.rs:159
        ("Concat: inner dim mismatch", lambda: K.concat([z(2, 3), z(2, 4), z(0)], 0, ctx=nothing)),        # manipulation.rs:162
        ("Pad: Rank 5 not fully implemented", lambda: K.pad(z(1, 1, 1, 1, 2), [0] * 10, ctx=nothing)),    # manipulation.rs:485
        ("MatMul K dim mismatch: 3 vs 4", lambda: K.matmul(z(2, 3), z(4, 5), ctx=nothing)),                # gemm.rs:129
        ("MatMul broadcast not fully supported yet", lambda: K.matmul(z(2, 3), z(4, 3, 5), ctx=nothing)),  # gemm.rs:134
        ("rank >= 2", lambda: K.matmul(z(3), z(3, 5), ctx=nothing)),                                       # gemm.rs:122
        ("Gemm K dim mismatch", lambda: K.gemm(z(2, 3), z(4, 5), ctx=nothing)),                            # gemm.rs:465
        ("only the last axis", lambda: K.softmax(z(2, 3), 0, ctx=nothing)),                                # norm.rs:218
        ("Only batch_size=1", lambda: K.lstm(z(2, 2, 4), z(1, 8, 4), z(1, 8, 2), ctx=nothing)),            # rnn.rs:88
        ("Only num_directions=1", lambda: K.gru(z(2, 1, 4), z(2, 6, 4), z(2, 6, 2), ctx=nothing)),         # rnn.rs:262
        ("expected rank-4", lambda: K.conv_transpose(z(1, 2, 3), z(2, 2, 1, 1), ctx=nothing)),             # conv2d.rs:2989
        ("group > 1 not supported", lambda: K.conv_transpose(z(1, 2, 3, 3), z(2, 2, 1, 1), group=2, ctx=nothing)),   # conv2d.rs:3042
    ]
    for msg, fn in cases:
        with pytest.raises(LeleB200Error, match=msg):
            fn()
    np.testing.assert_array_equal(R.pad(z(2), [1, 1], 3.0, "wrap"), [3.0, 0.0, 0.0, 3.0])                      # unknown mode = constant fill
    assert R.concat([z(0), z(0)], 0).shape == (0,) and K.concat([z(0), z(0)], 0, ctx=nothing).shape == (0,)       # manipulation.rs:131-144
    assert R.concat([z(0), z(2, 3), z(1, 3)], -2).shape == (3, 3)
    for msg, fn in (("splits sum mismatch", lambda: R.split(z(2, 6), 1, [2, 2])), ("element count mismatch", lambda: R.reshape(z(2, 3), [4])),
                    ("repeats length must match", lambda: R.tile(z(2, 3), [2])), ("Concat: ranks mismatch", lambda: R.concat([z(2, 3), z(3)], 0)),
                    ("Concat: inner dim mismatch", lambda: R.concat([z(2, 3), z(2, 4)], 0)),
                    ("sizes H and W must be positive", lambda: R.resize_nearest(z(1, 1, 1, 1), None, [1, 1, -1, 10])),
                    ("conv2d: output dimensions must be positive", lambda: R.conv2d(z(1, 1, 2, 2), z(1, 1, 3, 3))),
                    ("MatMul K dim mismatch: 3 vs 4", lambda: R.matmul(z(2, 3), z(4, 5))), ("broadcast not fully supported", lambda: R.matmul(z(2, 3), z(4, 3, 5))), ("Rank 5 not fully implemented", lambda: R.pad(z(1, 1, 1, 1, 2), [0] * 10))):
        with pytest.raises(ValueError, match=msg):
            fn()


def test_model_rs_topk_smallest():
    """TopK with largest = false (conv2d.rs:1421): ascending, ties in original order; k clamps to the axis length."""
    from tests import model_forms as MF
    m = _model_rs()
    text = """
pub struct TkWorkspace { pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, }
pub struct Tk<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut TkWorkspace, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>, TensorView<'static, f32>) {
        let (lo, lo_i) = lele::kernels::topk(&x, self.weight_i64(0, 8, &[1]).data[0] as usize, -1, false, true, &mut ws.buf_0, &mut ws.buf_1);
        let (hi, hi_i) = lele::kernels::topk(&x, self.weight_i64(8, 8, &[1]).data[0] as usize, -1, true, true, &mut ws.buf_0, &mut ws.buf_1);
        (lo.to_owned(), lo_i.to_owned(), hi.to_owned())
    }
"""
    prog = m.parse_model_rs(text)
    blob = m.synth_blob(prog, 1, {0: [3], 8: [9]})
    x = np.array([[2.0, -1.0, 2.0, 0.5, -1.0], [0.0, -0.0, 5.0, -3.0, 5.0]], np.float32)
    lo, lo_i, hi = m.run_program(prog, blob, [x], MF.R)
    np.testing.assert_array_equal(lo, [[-1.0, -1.0, 0.5], [-3.0, 0.0, 0.0]])
    np.testing.assert_array_equal(lo_i, [[1, 4, 3], [3, 0, 1]])
    np.testing.assert_array_equal(hi, [[2.0, 2.0, 0.5, -1.0, -1.0], [5.0, 5.0, 0.0, 0.0, -3.0]])       # k = 9 clamps to 5


def test_host_wrappers_marshal_and_shape_like_the_oracle(so_path, monkeypatch):
    """Every lele_b200.kernels wrapper is run with a stand-in context and a recording `call` (no device): the Python side must get
    through its shape logic, hand the C entry point exactly as many arguments as its prototype declares, size its output buffer for
    what it downloads, and return the shape the oracle returns for the same inputs."""
    import importlib.util
    from lele_b200 import kernels as K
    from oracle import reference_api as R
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    nargs = {name: len(args) for _, name, args in g.c_declarations(open(os.path.join(ROOT, "include", "lele_b200.h")).read())}
    seen = []

    def fake_call(name, *args):
        assert len(args) == nargs[name], (name, len(args), nargs[name])
        seen.append(name)

    class Buf:
        def __init__(self, n): self.ptr, self.n = 4096, n
        def free(self): pass

    class Ctx:
        h = None
        _slots = ()                                # host form: no output placement pending (kernels._resident)
        def upload(self, a, dtype=np.float32): return Buf(np.asarray(a).size)
        def empty(self, n, itemsize=4): return Buf(int(n))
        def download(self, b, shape, dtype=np.float32):
            assert int(np.prod(shape, dtype=np.int64)) <= max(b.n, 1), (shape, b.n)
            return np.zeros(shape, dtype)
        def sync(self): pass

    monkeypatch.setattr(K, "call", fake_call)
    ctx = Ctx()
    z = lambda *s: np.zeros(s, np.float32)
    o = lambda *s: np.ones(s, np.float32)
    cases = [   # (device-wrapper call, oracle call)
        (lambda: K.matmul(z(3, 4, 5), z(5, 6), ctx=ctx), lambda: R.matmul(z(3, 4, 5), z(5, 6))),
        (lambda: K.matmul_fused_add(z(4, 5), z(5, 6), z(6), ctx=ctx), lambda: R.matmul_fused_add(z(4, 5), z(5, 6), z(6))),
        (lambda: K.gemm(z(5, 4), z(6, 5), z(6), 1.0, 1.0, True, True, ctx=ctx), lambda: R.gemm(z(5, 4), z(6, 5), z(6), 1.0, 1.0, True, True)),
        (lambda: K.layer_norm(z(2, 3, 8), o(8), z(8), -1, 1e-5, ctx=ctx), lambda: R.layer_norm(z(2, 3, 8), o(8), z(8), -1, 1e-5)),
        (lambda: K.softmax(z(2, 7), -1, ctx=ctx), lambda: R.softmax(z(2, 7), -1)),
        (lambda: K.batch_norm(z(2, 3, 4), o(3), z(3), z(3), o(3), ctx=ctx), lambda: R.batch_norm(z(2, 3, 4), o(3), z(3), z(3), o(3))),
        (lambda: K.dynamic_quantize_linear(o(2, 5), ctx=ctx)[0], lambda: R.dynamic_quantize_linear(o(2, 5))[0]),
        (lambda: K.mat_mul_integer(z(2, 3, 4), z(4, 5), 1.0, 2.0, ctx=ctx), lambda: R.mat_mul_integer(z(2, 3, 4), z(4, 5), 1.0, 2.0)),
        (lambda: K.fused_quantized_linear(z(2, 5, 16), z(16, 24), o(24), np.array([128.0], np.float32), z(24), True, ctx=ctx),
         lambda: R.fused_quantized_linear(z(2, 5, 16), z(16, 24), o(24), 128, z(24), True)),
        (lambda: K.conv1d_fused(z(1, 2, 9), z(4, 2, 3), z(4), [1], 1, [1, 1], [2], True, ctx=ctx), lambda: R.conv1d(z(1, 2, 9), z(4, 2, 3), z(4), [1], 1, [1, 1], [2], True)),
        (lambda: K.conv2d(z(1, 2, 7, 8), z(3, 2, 3, 3), None, (1, 1), 1, (1, 1, 1, 1), (2, 2), ctx=ctx), lambda: R.conv2d(z(1, 2, 7, 8), z(3, 2, 3, 3), None, (1, 1), 1, (1, 1, 1, 1), (2, 2))),
        (lambda: K.conv_transpose(z(1, 2, 4, 4), z(2, 3, 3, 3), None, (1, 1), (1, 1, 1, 1), (2, 2), ctx=ctx), lambda: R.conv_transpose(z(1, 2, 4, 4), z(2, 3, 3, 3), None, (1, 1), (1, 1, 1, 1), (2, 2))),
        (lambda: K.max_pool2d(z(1, 2, 7, 8), (3, 2), (1, 0, 1, 0), (2, 2), (1, 1), True, ctx=ctx), lambda: R.max_pool2d(z(1, 2, 7, 8), (3, 2), (1, 0, 1, 0), (2, 2), (1, 1), True)),
        (lambda: K.resize_nearest(z(1, 2, 3, 4), [1, 1, 2.0, 1.5], None, "asymmetric", ctx=ctx), lambda: R.resize_nearest(z(1, 2, 3, 4), [1, 1, 2.0, 1.5], None, "asymmetric")),
        (lambda: K.lstm(z(5, 1, 3), z(1, 16, 3), z(1, 16, 4), z(1, 32), None, z(1, 1, 4), z(1, 1, 4), ctx=ctx)[0], lambda: R.lstm(z(5, 1, 3), z(1, 16, 3), z(1, 16, 4), z(1, 32), z(1, 1, 4), z(1, 1, 4))[0]),
        (lambda: K.gru(z(5, 1, 3), z(1, 12, 3), z(1, 12, 4), None, None, ctx=ctx)[1], lambda: R.gru(z(5, 1, 3), z(1, 12, 3), z(1, 12, 4), None, None)[1]),
        (lambda: K.add(z(3, 1, 5), z(4, 5), ctx=ctx), lambda: R.add(z(3, 1, 5), z(4, 5))),
        (lambda: K.prelu(z(2, 3), o(1, 1, 1), ctx=ctx), lambda: R.prelu(z(2, 3), o(1, 1, 1))),
        (lambda: K.tanh_kernel(z(7), ctx=ctx), lambda: R.tanh(z(7))),
        (lambda: K.not_(z(2, 2), ctx=ctx), lambda: R.not_(z(2, 2))),
        (lambda: K.clip(z(4), -1.0, float("inf"), ctx=ctx), lambda: R.clip(z(4), -1.0, float("inf"))),
        (lambda: K.reduce_sum(z(2, 3, 4), [0, 2], False, ctx=ctx), lambda: R.reduce(z(2, 3, 4), [0, 2], False, "sum")),
        (lambda: K.reduce_l2(z(2, 3), [], True, ctx=ctx), lambda: R.reduce(z(2, 3), [], True, "l2")),
        (lambda: K.reduce_max(z(2, 3), [1, -1], True, ctx=ctx), lambda: R.reduce(z(2, 3), [1, -1], True, "max")),
        (lambda: K.where_op(o(3, 1), z(1, 4), z(3, 4), ctx=ctx), lambda: R.where(o(3, 1), z(1, 4), z(3, 4))),
        (lambda: K.stft(z(1, 300), 64, 32, 64, None, False, ctx=ctx), lambda: R.stft(z(1, 300), 64, 32, 64, None, False)),
        (lambda: K.stft(z(300), 64, 32, 64, o(64), True, ctx=ctx), lambda: R.stft(z(300), 64, 32, 64, o(64), True)),
        (lambda: K.transpose(z(2, 3, 4), (2, 0, 1), ctx=ctx), lambda: R.transpose(z(2, 3, 4), (2, 0, 1))),
        (lambda: K.slice(z(4, 6), [1], [2**63 - 1], [1], [2], ctx=ctx), lambda: R.slice(z(4, 6), [1], [2**63 - 1], [1], [2])),
        (lambda: K.expand(z(2, 1, 3), [0, 4, 0], ctx=ctx), lambda: R.expand(z(2, 1, 3), [0, 4, 0])),
        (lambda: K.split(z(2, 6), 1, [1, 5], ctx=ctx)[1], lambda: R.split(z(2, 6), 1, [1, 5])[1]),
        (lambda: K.concat([z(2, 3), z(0), z(2, 1)], -1, ctx=ctx), lambda: R.concat([z(2, 3), z(0), z(2, 1)], -1)),
        (lambda: K.pad(z(2, 3), [1, 2], 0.0, "edge", ctx=ctx), lambda: R.pad(z(2, 3), [1, 2], 0.0, "edge")),
        (lambda: K.gather(z(2, 3, 4), np.array(1), 1, ctx=ctx), lambda: R.gather(z(2, 3, 4), np.array(1), 1)),
        (lambda: K.gather_elements(z(2, 3), z(2, 2), 1, ctx=ctx), lambda: R.gather_elements(z(2, 3), z(2, 2), 1)),
        (lambda: K.tile(z(2, 3), [2, 1], ctx=ctx), lambda: R.tile(z(2, 3), [2, 1])),
        (lambda: K.topk(z(2, 9), 4, ctx=ctx)[1], lambda: R.topk(z(2, 9), 4)[1]),
        (lambda: K.lstm_streams(z(3, 5, 2), z(1, 16, 2), z(1, 16, 4), None, z(3, 4), z(3, 4), ctx=ctx)[0], lambda: z(3, 5, 4)),
        (lambda: K.gru_streams(z(3, 5, 2), z(1, 12, 2), z(1, 12, 4), z(1, 24), None, ctx=ctx)[1], lambda: z(3, 4)),
    ]
    for dev, ref in cases:
        got, want = dev(), ref()
        assert np.asarray(got).shape == np.asarray(want).shape, (seen[-1:], np.asarray(got).shape, np.asarray(want).shape)
    assert len(set(seen)) >= 29, sorted(set(seen))


def test_model_rs_shape_arithmetic_stays_on_the_host():
    """i64 shape tensors (Shape / Gather / Concat / Range / Less / Cast / ConstantOfShape / Size, `&t.data[..]`, temp_i64 vectors,
    inline to_i64_vec) are evaluated as host int64 values; only f32 tensor work reaches the operator namespace."""
    from tests import model_forms as MF
    m = _model_rs()
    prog, blob, inputs = MF.shape_forms(m)

    class Spy(m._NamespaceOps):
        calls = []

        def __getattr__(self, name):
            f = getattr(self.ns, name)

            def g(*a, **k):
                Spy.calls.append(name)
                for v in a:
                    assert not (isinstance(v, np.ndarray) and v.dtype == np.int64 and name != "transpose"), (name, "i64 tensor reached the device namespace")
                return f(*a, **k)
            return g

        def binary(self, op, a, b):
            Spy.calls.append(op); assert np.asarray(a).dtype == np.float32 and np.asarray(b).dtype == np.float32
            return super().binary(op, a, b)

        def unary(self, op, x):
            Spy.calls.append(op); assert np.asarray(x).dtype == np.float32
            return super().unary(op, x)

    got = m.run_program(prog, blob, inputs, Spy(MF.R))
    ref = MF.shape_forms_direct(*inputs)
    for a, b in zip(got, ref):
        assert np.asarray(a).dtype == np.asarray(b).dtype and np.asarray(a).shape == np.asarray(b).shape
        np.testing.assert_array_equal(a, b)
    assert Spy.calls == ["transpose", "sqrt", "div", "mul", "slice", "reduce"]
    # integer division truncates toward zero, like Rust's i64 `/`
    assert m._host_i64_op("div", [np.array([-7, 7, -7], np.int64), np.array([2, -2, -2], np.int64)]).tolist() == [-3, -3, 3]
    assert m._host_i64_op("div", [np.ones(2, np.float32), np.ones(2, np.float32)]) is None


def test_cuda_ops_glue_binds_to_kernel_signatures(so_path):
    """CudaOps is the only untested-on-CPU piece of the replay: here every call it makes is bound against the real
    lele_b200.kernels signature (inspect.signature(...).bind) and then answered by the oracle, so an argument-order or
    keyword mistake in the glue fails on the CPU box, not on the GPU one."""
    import inspect
    from lele_b200 import kernels as K, model_rs as MR
    from tests import model_forms as MF
    R = MF.R
    seen = []

    class BoundK:
        def __getattr__(self, name):
            real = getattr(K, name)
            sig = inspect.signature(real)

            def f(*args, **kw):
                b = sig.bind(*args, **kw); b.apply_defaults()
                assert "ctx" in b.arguments
                seen.append(name)
                v = {k: b.arguments[k] for k in b.arguments if k != "ctx"}
                vals = list(v.values())
                if name == "layer_norm": return R.layer_norm(v["x"], v["scale"], v["bias"], v["axis"], v["epsilon"])
                if name == "gemm": return R.gemm(v["a"], v["b"], None if v["c"] is None else np.asarray(v["c"]).reshape(-1), v["alpha"], v["beta"], v["trans_a"], v["trans_b"])
                if name == "conv1d_fused": return R.conv1d(vals[0], vals[1], vals[2], tuple(vals[3]), vals[4], tuple(vals[5]), tuple(vals[6]), vals[7])
                if name == "lstm":
                    assert v["sequence_lens"] is None
                    return R.lstm(v["input"], v["w"], v["r"], v["bias"], v["initial_h"], v["initial_c"])
                if name == "gru":
                    assert v["linear_before_reset"] is False
                    return R.gru(v["input"], v["w"], v["r"], v["bias"], v["initial_h"])
                if name.startswith("reduce_"): return R.reduce(vals[0], list(vals[1]), vals[2], name[len("reduce_"):])
                if name == "fused_quantized_linear": return R.fused_quantized_linear(v["input"], v["weight_int8"], v["weight_scale"], v["weight_zero"], v["bias"], v["apply_relu"])
                if name == "mat_mul_integer":
                    assert v["scale"] is None and v["bias"] is None and v["relu"] is False
                    return R.mat_mul_integer(v["a"], v["b"], v["a_zero_point"], v["b_zero_point"])
                if name == "stft":
                    assert v["power"] is False
                    return R.stft(v["input"], v["n_fft"], v["hop_length"], v["win_length"], v["window"])
                return getattr(R, {"tanh_kernel": "tanh", "max": "maximum", "where_op": "where"}.get(name, name))(*vals)
            return f

    ops = MR.CudaOps.__new__(MR.CudaOps); ops.K, ops.ctx = BoundK(), None
    prog, blob, x = MF.math_forms(MR)
    for a, b in zip(MR.run_program(prog, blob, [x], ops), MF.math_forms_direct(MR, blob, x)):
        np.testing.assert_array_equal(a, b)
    prog, blob, x = MF.recurrent_forms(MR)
    for a, b in zip(MR.run_program(prog, blob, [x], ops), MF.recurrent_forms_direct(MR, blob, x)):
        np.testing.assert_array_equal(a, b)
    prog, blob, x = MF.quant_forms(MR)
    for a, b in zip(MR.run_program(prog, blob, [x], ops), MF.quant_forms_direct(MR, blob, x)):
        np.testing.assert_array_equal(a, b)
    prog, blob, x = MF.const_forms(MR)
    for a, b in zip(MR.run_program(prog, blob, [x], ops), MF.const_forms_direct(MR, blob, x)):
        np.testing.assert_array_equal(a, b)
    assert {"stft", "pow", "log", "sin", "cos", "less", "not_", "equal"} <= set(seen)
    assert {"fused_quantized_linear", "dynamic_quantize_linear", "mat_mul_integer", "clip", "matmul_fused_add", "mul"} <= set(seen)
    assert {"layer_norm", "gemm", "tanh_kernel", "max", "reduce_mean", "exp", "expand", "where_op", "pad", "conv1d_fused", "lstm", "gru"} <= set(seen)


def test_generated_model_fixture_is_consistent():
    """tests/golden/yolo26seg_program.json (from the reference's committed lele_gen output, see make_model_program.py):
    337 statements, 21 workspace buffers, a 10 993 208-byte weights.bin."""
    import importlib.util, json
    spec = importlib.util.spec_from_file_location("model_rs", os.path.join(ROOT, "lele_b200", "model_rs.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    prog = json.load(open(os.path.join(ROOT, "tests", "golden", "yolo26seg_program.json")))
    assert prog["class"] == "Yolo26Seg" and prog["workspace_buffers"] == 21 and len(prog["statements"]) == 337
    ops = [st["op"] for st in prog["statements"]]
    assert ops.count("conv2d") + ops.count("conv2d_silu") == 117 and ops.count("conv_transpose") == 1
    blob = m.synth_blob(prog, 7, {int(k): v for k, v in prog["constants"].items()})
    assert len(blob) == 10993208


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract keys, the
    oracle port on the host cores, the GPU arm's own 64-clip step, bounded by its time budget (here squeezed so the clips shorten
    to 1 s); the process never maps the product library."""
    import json
    env = dict(os.environ, LELE_B200_REF_BUDGET_S="0.1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["value"] > 0 and line["steps"] == 1 and line["vs_baseline"] is None and line["scaling"] == "weak"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == line["value"] and "64 clips x 1 s per step" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["clip_seconds"] == 1 and line["config"]["clips_per_gpu"] == 64 and "SenseVoiceSmall" in line["config"]["workload"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    ref_body = src[src.index("def run_reference(args):"):src.index("def run_ours(args):")]
    assert "import lele_b200" not in ref_body and "from lele_b200" not in ref_body      # the arm loads sensevoice_weights.py by path (no product .so)
    # under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without work or output
    out1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=120)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


# ------------------------------------------------------------------------------------------------------------------
# Resident replay: src/tensor.rs's buffer arena mapped to device memory.  Host logic, exercised here
# against a host-memory stand-in for the library (tests/fake_device.py); the same flows run on the B200 in
# tests/test_gpu_resident.py.
# ------------------------------------------------------------------------------------------------------------------
import json as _json_resident  # noqa: E402
RESIDENT_TEXT = """
pub struct T9Workspace { pub buf_0: Vec<f32>, pub buf_1: Vec<f32>, pub buf_2: Vec<f32>, }
pub struct T9<'a> { data: &'a [u8] }
    fn run_chunk_0<'w>(&self, ws: &'w mut T9Workspace, x: TensorView<'w, f32>) -> (TensorView<'static, f32>, TensorView<'static, f32>) {
        let a = lele::kernels::conv2d_silu(&x, &self.weight_f32(0, 432, &[4, 3, 3, 3]), Some(&self.weight_f32(432, 16, &[4])), &[1, 1], 1, &[1, 1, 1, 1], &[1, 1], &mut ws.buf_0);
        let b = lele::kernels::sigmoid(&a, &mut ws.buf_1);
        let c = lele::kernels::mul(&a, &b, &mut ws.buf_2);
        let splits_slice = &[2, 2];
        let mut split_results = lele::kernels::split_owned(&c, 1, splits_slice);
        let s1 = split_results.swap_remove(1);
        let s0 = split_results.swap_remove(0);
        let d = lele::kernels::add(&s0, &s1, &mut ws.buf_0);
        let e = lele::kernels::concat(&[&d, &s1], 1, &mut ws.buf_1);
        let p = lele::kernels::max_pool2d(&e, &[2, 2], &[2, 2], &[0, 0, 0, 0], &[1, 1], false, &mut ws.buf_2);
        let f = lele::kernels::reshape(&p, &[1, 4, -1]);
        let g = lele::kernels::transpose(&f, &[0, 2, 1], &mut ws.buf_0);
        (g.to_owned(), e.to_owned())
    }
"""


@pytest.fixture
def fake_dev(monkeypatch):
    """tests/fake_device.py installed for one test; objects created under it are collected before the real library comes back."""
    import gc
    from tests import fake_device
    dev = fake_device.install(monkeypatch)
    yield dev
    from lele_b200 import kernels
    kernels._default = None                 # the default context made under the stand-in must die under it as well
    gc.collect()


def _resident_case():
    from lele_b200 import model_rs as MR
    from oracle import reference_api as R
    prog = MR.parse_model_rs(RESIDENT_TEXT)
    blob = MR.synth_blob(prog, 11)
    xs = [np.random.default_rng(20 + i).standard_normal((1, 3, 8, 8)).astype(np.float32) for i in range(3)]
    want = [MR.run_program(prog, blob, [x], R) for x in xs]
    return MR, prog, blob, xs, want


def test_parser_keeps_the_workspace_buffer_of_every_statement():
    from lele_b200 import model_rs as MR
    prog = MR.parse_model_rs(RESIDENT_TEXT)
    bufs = {st["outs"][0]: st.get("bufs", []) for st in prog["statements"]}
    assert bufs["a"] == ["ws.buf_0"] and bufs["c"] == ["ws.buf_2"] and bufs["g"] == ["ws.buf_0"] and bufs["f"] == []
    assert prog["workspace_buffers"] == 3
    yolo = _json_resident.load(open(os.path.join(ROOT, "tests", "golden", "yolo26seg_program.json")))
    named = {b for st in yolo["statements"] for b in st.get("bufs", []) if b.startswith("ws.buf_")}
    assert len(named) == 21                                            # yolo26seg.rs:14-36: 21 workspace buffers


def test_resident_replay_keeps_values_on_the_device(fake_dev):
    MR, prog, blob, xs, want = _resident_case()
    dev = fake_dev
    model = MR.GeneratedModel(prog, blob, resident=True)
    got = model.forward(xs[0])
    for g, w in zip(got, want[0]):
        np.testing.assert_allclose(g, w, rtol=1e-6, atol=1e-6)
    first = list(dev.log)
    # uploads: the graph input, the two weight views (once per model); downloads: the two graph outputs only
    assert first.count("lele_b200_h2d") == 3 and first.count("lele_b200_d2h") == 2
    ws = model.workspace()
    assert {"ws.buf_0", "ws.buf_1", "ws.buf_2"} <= set(ws.bytes) and ws.bytes["ws.buf_0"] == 4 * 8 * 8 * 4
    dev.log.clear()
    got2 = model.forward(xs[1])                                        # second forward: sizes settled, weights resident
    for g, w in zip(got2, want[1]):
        np.testing.assert_allclose(g, w, rtol=1e-6, atol=1e-6)
    assert dev.log.count("lele_b200_h2d") == 1 and dev.log.count("lele_b200_d2h") == 2 and "arena_grow" not in dev.log
    assert dev.log.count("lele_b200_malloc") == 1                      # the input's staging buffer; every other value lives in the arena


def test_arena_grow_keeps_contents_and_release_frees(fake_dev):
    from lele_b200 import kernels as K
    dev = fake_dev
    ctx = K.Context(0)
    ws = K.Workspace(ctx)
    t = ws.tensor("ws.buf_0", (4,))
    dev.arr(t.ptr, (4,))[...] = [1, 2, 3, 4]
    t2 = ws.tensor("ws.buf_0", (2, 8))                                 # grows: ensure_capacity (utils.rs:10) keeps the old prefix
    assert t2.ptr != t.ptr and list(dev.arr(t2.ptr, (4,))) == [1, 2, 3, 4]
    t3 = ws.tensor("ws.buf_0", (3,))                                   # smaller request: same storage
    assert t3.ptr == t2.ptr and ws.bytes["ws.buf_0"] == 64
    ws.release()
    assert not dev.arena


def test_batch_runner_captures_the_step_once(fake_dev):
    MR, prog, blob, xs, want = _resident_case()
    dev = fake_dev
    model = MR.GeneratedModel(prog, blob, resident=True)
    br = model.batch_runner(3, lanes=2)
    assert len(br.lanes) == 2 and len(br.ws) == 3
    for rnd in range(3):                                               # eager, capture + launch, launch
        dev.log.clear()
        order = [(i + rnd) % 3 for i in range(3)]
        got = br.run([[xs[i]] for i in order])
        for it, i in zip(got, order):
            for g, w in zip(it, want[i]):
                np.testing.assert_allclose(g, w, rtol=1e-6, atol=1e-6)
        if rnd == 1:
            assert "lele_b200_capture_begin" in dev.log and "lele_b200_capture_end" in dev.log
        if rnd == 2:                                                   # steady state: inputs in, one graph launch, outputs out
            assert [n for n in dev.log if n not in ("lele_b200_h2d", "lele_b200_d2h", "lele_b200_sync")] == ["lele_b200_graph_launch"]
            assert dev.log.count("lele_b200_h2d") == 3 and dev.log.count("lele_b200_d2h") == 6
    br.close()


def test_config1_silero_shaped_vad_on_the_reference_fixture():
    """BASELINE configs[0] (plumbing): fixtures/zh.wav -> 89 472 samples -> 175 chunks of 512 (x32768), state [2,1,128] carried from
    chunk to chunk (examples/silero/src/main.rs:70-137), through a Silero-shaped recurrent model (STFT -> Conv1d -> LSTM H=128 ->
    Gemm -> Sigmoid; synthetic weights: no model file exists) replayed on the CPU oracle.  Pass = it runs, 175 finite probabilities
    in [0, 1], the state evolves, the segment pass terminates inside the clip."""
    from lele_b200 import model_rs as MR
    from lele_b200.vad import StreamingVad, collect_segments, merge_segments
    from oracle import reference_api as R
    from tests import model_forms as MF
    audio = MF.read_wav_s16(os.path.join(ROOT, "tests", "golden", "zh.wav"))
    assert audio.size == 89472 and np.abs(audio).max() <= 1.0
    prog, blob = MF.vad_model(MR, hidden=128)
    vad = StreamingVad(prog, blob, ops=R, state_shape=(2, 1, 128))
    probs = vad.process(audio)
    assert probs.shape == (175,) and np.isfinite(probs).all() and (probs >= 0).all() and (probs <= 1).all()
    assert vad.state.shape == (2, 1, 128) and np.isfinite(vad.state).all() and np.abs(vad.state).max() > 0
    assert len(np.unique(probs)) > 10                                   # the recurrent state makes every chunk's answer its own
    segs = merge_segments(collect_segments(probs, audio.size))
    assert all(0 <= s < e <= audio.size for s, e in segs)


def test_batch_folded_replay_runs_the_independent_prefix_once(fake_dev):
    """run_program(lift=B) / BatchRunner(fold=True): statements that are independent along the leading dimension run ONCE on the
    stacked [B, ...] values (reshape targets get their leading 1 replaced by B); the first statement that is not stops the folding
    and the rest runs per item on zero-copy slices.  Same results as B separate replays."""
    MR, prog, blob, xs, want = _resident_case()
    dev = fake_dev
    model = MR.GeneratedModel(prog, blob, resident=True)
    br = model.batch_runner(3, lanes=2, fold=True)
    for rnd in range(3):
        dev.log.clear()
        got = br.run([[x] for x in xs])
        for it, w in zip(got, want):
            for g, ww in zip(it, w):
                np.testing.assert_allclose(g, ww, rtol=1e-6, atol=1e-6)
        if rnd == 0:
            assert dev.log.count("lele_b200_conv2d") == 1 and dev.log.count("lele_b200_max_pool2d") == 1      # one launch for the whole batch
    rep = prog["_fold_report"][3]
    assert rep["first_per_item_statement"] == rep["statements"] and rep["stopped_at"] is None              # this graph folds completely
    br.close()
    # a statement that is not independent along the leading dimension ends the folded prefix: flatten(axis=2) moves the batch inward
    text = RESIDENT_TEXT.replace("let g = lele::kernels::transpose(&f, &[0, 2, 1], &mut ws.buf_0);",
                                 "let f2 = lele::kernels::flatten(&f, 2);\n        let g = lele::kernels::transpose(&f2, &[1, 0], &mut ws.buf_0);")
    prog2 = MR.parse_model_rs(text)
    from oracle import reference_api as R
    want2 = [MR.run_program(prog2, blob, [x], R) for x in xs]
    br2 = MR.GeneratedModel(prog2, blob, resident=True).batch_runner(3, lanes=2, fold=True)
    got2 = br2.run([[x] for x in xs])
    for it, w in zip(got2, want2):
        for g, ww in zip(it, w):
            np.testing.assert_allclose(g, ww, rtol=1e-6, atol=1e-6)
    assert prog2["_fold_report"][3]["stopped_at"] == "flatten" and prog2["_fold_report"][3]["folded_statements"] >= 8
    br2.close()


def test_quantiser_rounding_without_a_conversion_is_exact():
    """csrc/common.cuh lb_q8: clamp(rint(y), 0, 255) (round-half-even, the reference's `_mm256_cvtps_epi32` after the clamp-free fma,
    avx/quantization.rs:158-161, then the saturating pack) computed as  low byte of  f32(clamp(y, 0, 255) + 1.5 * 2^23).  The identity is
    checked here in numpy float32 on every half-way point, the neighbours of every integer, out-of-range / infinite / NaN inputs and a
    random sweep (the GPU kernels are compared bit for bit with the oracle's quantiser in test_gpu_parity.py)."""
    f32 = np.float32
    ks = np.arange(-3, 260, dtype=np.float64)
    pts = [ks, ks + 0.5, ks - 0.5]
    for d in (1e-3, 2.0 ** -16, 2.0 ** -20):
        pts += [ks + d, ks - d, ks + 0.5 + d, ks + 0.5 - d]
    rng = np.random.default_rng(5)
    pts.append(rng.uniform(-10.0, 270.0, 200000))
    pts.append(np.array([0.0, -0.0, 255.0, 255.49999, 255.5, 256.0, 1e9, -1e9, np.inf, -np.inf, np.nan, 1e-45, -1e-45]))
    y = np.concatenate(pts).astype(f32)
    with np.errstate(invalid="ignore"):
        want = np.clip(np.rint(np.nan_to_num(y.astype(np.float64), nan=0.0, posinf=1e30, neginf=-1e30)), 0, 255).astype(np.uint32)
        c = np.minimum(np.fmax(y, f32(0.0)), f32(255.0)).astype(f32)         # fmaxf(NaN, 0) = 0, as CUDA's fmaxf
        got = (c + f32(12582912.0)).astype(f32).view(np.uint32) & np.uint32(0xFF)
    np.testing.assert_array_equal(got, want)
