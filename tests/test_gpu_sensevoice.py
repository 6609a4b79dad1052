"""GPU parity of the SenseVoice-shaped graph runner (csrc/sensevoice.cu) against the CPU oracle
executing the reference's per-clip op-by-op sequence (oracle/sensevoice_ref.c), plus the
size-independent properties used at BASELINE.json's full size (64 clips x 16 s)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lele_b200 import SenseVoice                               # noqa: E402
from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, synth_batch  # noqa: E402
from oracle import reference_api as R                            # noqa: E402  (checker only)

SMALL = SenseVoiceConfig(n_layers=3, vocab=1000, n_stage1=2, max_t=128)


def rel_err(got, ref):
    return float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30))


def _make(blob, attn, **kw):
    """attn = "simt": CUDA-core f32 attention with the oracle's exact summation order (bit-exact encoder);
    attn = "tc": the product path, fused tcgen05 3xTF32 attention (f32-grade, not bit-identical)."""
    import os
    os.environ["LELE_B200_ATTN_SIMT"] = "1" if attn == "simt" else "0"
    try:
        return SenseVoice(blob, **kw)
    finally:
        os.environ.pop("LELE_B200_ATTN_SIMT", None)


@pytest.fixture(scope="module")
def small_model():
    blob = build_blob(SMALL, seed=7)
    m = _make(blob, "simt", max_clips=4, max_samples=89472)
    yield blob, m
    m.close()


@pytest.fixture(scope="module")
def small_model_tc():
    blob = build_blob(SMALL, seed=7)
    m = _make(blob, "tc", max_clips=4, max_samples=89472)
    yield blob, m
    m.close()


def _qkv_with_v(m, B, T, H=4, dk=128):
    """The projection's [q | k | v] rows of the last forward.  The product path keeps v only transposed (V^T [B][H][dk][Tp], the
    attention kernel's P.V operand, which the FSMN block reads too) and never writes the row-major v columns, so they are rebuilt
    from the "vt" workspace buffer here."""
    d = H * dk
    qkv = m.workspace("qkv", (B * T, 3 * d)).copy()
    Tp = (T + 3) // 4 * 4
    vt = m.workspace("vt", (B, H, dk, Tp))
    qkv[:, 2 * d:] = vt[..., :T].transpose(0, 3, 1, 2).reshape(B * T, d)
    return qkv


def _oracle_attention(qkv, T, H=4, dk=128):
    d = H * dk
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    qh = (q.reshape(T, H, dk).transpose(1, 0, 2) * np.float32(1.0 / np.sqrt(np.float32(dk)))).astype(np.float32)
    kh = k.reshape(T, H, dk).transpose(1, 2, 0); vh = v.reshape(T, H, dk).transpose(1, 0, 2)
    p = R.softmax(R.matmul(qh, kh))
    return R.matmul(p, vh).transpose(1, 0, 2).reshape(T, d)


@pytest.mark.parametrize("t", [40, 93, 29])
def test_tcgen05_attention_stage(small_model_tc, t):
    """The fused tensor-core attention against the reference op sequence (mul, matmul, softmax, matmul)
    on the GPU's own qkv (identical input): 3xTF32 keeps it at f32 accuracy (1e-5 bar here, 1e-4 required)."""
    blob, m = small_model_tc
    rng = np.random.default_rng(t)
    B = 3
    feats = (rng.standard_normal((B, t, 560)) * np.array([1.0, 2.5, 0.3])[:, None, None]).astype(np.float32)
    m.forward(feats, 3, 0, n_layers=1)
    T = t + 4
    qkv = _qkv_with_v(m, B, T); att = m.workspace("att", (B * T, 512))
    keys = m.workspace("keys", (SMALL.n_layers * 4 + 1, B, 8, 2), np.uint32)    # [site][clip][slot][min,max], sharded atomics
    for c in range(B):
        want = _oracle_attention(qkv[c * T:(c + 1) * T], T)
        assert rel_err(att[c * T:(c + 1) * T], want) < 1e-5, (t, c)
        k = np.array([keys[1, c, :, 0].min(), keys[1, c, :, 1].max()], np.uint32)   # fused per-clip min/max of the attention output
        dec = np.where(k & 0x80000000, k & 0x7fffffff, ~k).astype(np.uint32).view(np.float32)
        got = att[c * T:(c + 1) * T]
        assert dec[0] == got.min() and dec[1] == got.max()


@pytest.mark.parametrize("lanes", [2, 3])
def test_clip_lanes_bit_identical(lanes, monkeypatch):
    """The runner splits a batch into clip lanes that execute on concurrent streams (own workspace slices, own
    min/max keys).  Clips are independent, so any lane count must give bit-identical logits and ids
    (7 clips -> uneven lanes; product configuration: tcgen05 attention, two-pass FFN1, CUDA-graph replay)."""
    blob = build_blob(SMALL, seed=11)
    pcm = synth_batch(3, 7, 16000 * 3)
    outs = []
    for n in (1, lanes):
        monkeypatch.setenv("LELE_B200_LANES", str(n))
        m = SenseVoice(blob, max_clips=7, max_samples=pcm.shape[1])
        ids, logits = m.transcribe(pcm, want_logits=True)
        for _ in range(2):                                # eager + capture, then a replay of the captured graph
            np.testing.assert_array_equal(ids, m.transcribe(pcm))
        outs.append((ids, logits))
        m.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


_SWITCHED = {"LELE_B200_FFN_FUSED": "0", "LELE_B200_W_UNSIGNED": "1", "LELE_B200_LNQ_STREAM": "1",
             # round 2c: in-place residual added by the L2 (TMA reduce-add) vs added in registers; the FSMN residual tile fetched by TMA vs
             # per-thread loads; the register-window FSMN kernel vs the tap-gathering one; B tiles multicast over 2-CTA clusters; the early
             # look-up of the fused quantiser's parameters -- all re-schedulings of the same IEEE operations
             "LELE_B200_GEMM_RED": "0", "LELE_B200_GEMM_R1_TMA": "0", "LELE_B200_FSMN_V2": "0", "LELE_B200_GEMM_MC": "1",
             "LELE_B200_FFN_PREFETCH": "1", "LELE_B200_GEMM_CG2": "1",      # + FFN2 as cta_group::2 pair MMAs (256 x 256 per cluster)
             "LELE_B200_GEMM_AFUSE": "1"}    # the attention output quantised inside the out-projection instead of by its own kernel


@pytest.mark.parametrize("switch", sorted(_SWITCHED))
def test_one_pass_ffn1_and_s8_weights_bit_identical(switch, monkeypatch):
    """Round-2 GEMM changes are pure re-schedulings of exact arithmetic: (a) FFN1 as ONE pass (the dequantised tile waits in TMEM for
    the clip's max, gemm_i8_fused_q_kernel) vs the max-only + quantising passes; (b) the weight operand as s8 (w - 128, no per-row
    zero-point term) vs u8.  Logits and ids must be bit-identical either way -- uneven batch (7 clips of 3 s: groups of clips, a partial
    last m-block per group, warps that straddle two clips), eager + capture + replay."""
    blob = build_blob(SMALL, seed=12)
    pcm = synth_batch(5, 7, 16000 * 3)
    outs = []
    for v in ("default", "switched"):
        if v == "switched":
            monkeypatch.setenv(switch, _SWITCHED[switch])
        m = SenseVoice(blob, max_clips=7, max_samples=pcm.shape[1])
        ids, logits = m.transcribe(pcm, want_logits=True)
        for _ in range(2):
            np.testing.assert_array_equal(ids, m.transcribe(pcm))
        outs.append((ids, logits))
        m.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_pipelined_host_entry_matches_blocking(small_model_tc):
    """transcribe_host_async / transcribe_wait (two batches in flight, copies on their own streams) returns exactly the ids
    of the blocking entry for every batch, in submission order; a busy slot is refused."""
    import torch
    from lele_b200 import LeleB200Error
    blob, m = small_model_tc
    batches = [synth_batch(10 * i, 3, 16000 * 2) for i in range(5)]
    want = [m.transcribe(b) for b in batches]
    T = want[0].shape[1]
    pins = [torch.from_numpy(b).pin_memory() for b in batches]
    outs = [torch.zeros((3, T), dtype=torch.int32).pin_memory() for _ in batches]
    for i, p in enumerate(pins):
        if i >= 2:
            m.transcribe_wait(i % 2)
        m.transcribe_host_async(p.data_ptr(), 3, p.shape[1], outs[i].data_ptr(), i % 2)
    with pytest.raises(LeleB200Error):
        m.transcribe_host_async(pins[0].data_ptr(), 3, pins[0].shape[1], outs[0].data_ptr(), (len(pins) - 1) % 2)   # still in flight
    m.transcribe_wait(0); m.transcribe_wait(1)
    for w, o in zip(want, outs):
        np.testing.assert_array_equal(w, o.numpy())


def test_tc_network_close_to_oracle(small_model_tc):
    """Whole small network with the tensor-core attention: f32-grade attention differences can flip
    individual u8 roundings of the next dynamic quantiser (1 LSB), so the bar is normwise."""
    blob, m = small_model_tc
    ref = R.SenseVoiceRef(blob)
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((2, 40, 560)).astype(np.float32)
    got = m.forward(feats, 3, 0)
    for c in range(2):
        want = ref.forward(feats[c], 3, 0)
        err = np.abs(got[c] - want)
        assert err.max() / np.abs(want).max() < 0.15 and err.mean() / np.abs(want).mean() < 0.03
        assert (np.argmax(got[c], 1) == np.argmax(want, 1)).mean() > 0.8


def test_forward_features_matches_oracle_per_layer(small_model):
    """Hidden state after 0,1,2 layers and final logits; per-clip dynamic quantisation means each
    clip of the batch must match its own single-clip oracle run."""
    blob, m = small_model
    ref = R.SenseVoiceRef(blob)
    rng = np.random.default_rng(0)
    feats = (rng.standard_normal((3, 40, 560)) * np.array([1.0, 2.5, 0.3])[:, None, None]).astype(np.float32)
    for nl, tol in [(0, 1e-6), (1, 1e-5), (2, 1e-5), (-1, 1e-5)]:
        got = m.forward(feats, 3, 0, n_layers=nl)
        for c in range(3):
            want = ref.forward(feats[c], 3, 0, n_layers=nl)
            assert got[c].shape == want.shape
            e = rel_err(got[c], want)
            # identical features in => integer core exact and identical f32 op order (LayerNorm / softmax in
            # the AVX2 accumulator order, sequential-k FMA GEMMs): the encoder is bit-exact up to ~1 ulp
            assert e < tol, (nl, c, e)


def test_pcm_to_ids_matches_oracle(small_model):
    """End to end from PCM.  Dynamic int8 quantisation turns 1e-6-level feature differences into
    1-LSB rounding flips, so PCM-level logits are compared the way the reference's own e2e test
    does (examples/sensevoice/tests/e2e_test.rs:69: MAE <= 1.0, argmax agreement), while the strict
    bars are applied per stage: front-end 1e-4, encoder on identical features ~bit-exact."""
    blob, m = small_model
    ref = R.SenseVoiceRef(blob)
    pcm = synth_batch(0, 2, 89472)
    ids, logits = m.transcribe(pcm, want_logits=True)
    lfr_gpu = m.workspace("lfr", (2, 93, 560))
    feats_gpu = m.workspace("feats", (2, 93, 560))
    ids_host_path = m.transcribe(pcm)                          # host-buffer entry, fused-argmax epilogue (no logits written)
    np.testing.assert_array_equal(ids, ids_host_path)
    for c in range(2):
        assert rel_err(lfr_gpu[c], R.frontend(pcm[c])) < 1e-4               # stage 1a: fbank + LFR
        assert rel_err(feats_gpu[c], R.cmvn(lfr_gpu[c])) < 1e-4             # stage 1b: CMVN on identical input (1/std amplifies, so not chained)
        rlog = ref.forward(feats_gpu[c], 3, 0)                              # stage 2: encoder on identical features
        assert rel_err(logits[c], rlog) < 1e-5
        np.testing.assert_array_equal(ids[c], rlog.shape[1] - 1 - np.argmax(rlog[:, ::-1], axis=1))
        rids, rlog_pcm = ref.pcm_to_ids(pcm[c], want_logits=True)          # whole path on the CPU
        assert float(np.abs(logits[c] - rlog_pcm).mean()) < 0.05           # reference's own bar is MAE <= 1.0
        assert (ids[c] == rids).mean() > 0.8
        # the GPU ids are the LAST-max argmax of the GPU logits (tokenizer.rs:55)
        np.testing.assert_array_equal(ids[c], logits[c].shape[1] - 1 - np.argmax(logits[c][:, ::-1], axis=1))


def test_prompt_ids_and_bounds(small_model):
    blob, m = small_model
    rng = np.random.default_rng(1)
    f = rng.standard_normal((1, 10, 560)).astype(np.float32)
    a = m.forward(f, 3, 0, n_layers=0); b = m.forward(f, 4, 1, n_layers=0)
    assert not np.array_equal(a[0, 0], b[0, 0]) and np.array_equal(a[0, 1:3], b[0, 1:3]) and np.array_equal(a[0, 4:], b[0, 4:])
    from lele_b200 import LeleB200Error
    with pytest.raises(LeleB200Error):
        m.forward(f, 99, 0)
    with pytest.raises(LeleB200Error):
        m.forward(np.zeros((5, 10, 560), np.float32))           # more clips than the workspace holds


@pytest.fixture(scope="module")
def full_model():
    blob = build_blob(SenseVoiceConfig(), seed=1234)
    m = SenseVoice(blob, max_clips=64, max_samples=256000)      # product configuration (tcgen05 attention)
    yield blob, m
    m.close()


def test_full_size_first_layers_vs_oracle(full_model):
    """BASELINE config shapes (T'=271, d=512, ffn 2048): 2 layers of one 16 s clip vs the oracle."""
    blob, m = full_model
    ref = R.SenseVoiceRef(blob)
    pcm = synth_batch(5, 1, 256000)
    feats = R.cmvn(R.frontend(pcm[0]))
    got = m.forward(feats[None], 3, 0, n_layers=2)[0]
    want = ref.forward(feats, 3, 0, n_layers=2)
    assert got.shape == (271, 512)
    err = np.abs(got - want)
    assert err.max() / np.abs(want).max() < 0.1 and err.mean() / np.abs(want).mean() < 0.02   # normwise: tensor-core attention (see test_tc_network_close_to_oracle)
    qkv = _qkv_with_v(m, 1, 271); att = m.workspace("att", (271, 512))                            # layer-1 buffers of the last forward
    assert rel_err(att, _oracle_attention(qkv, 271)) < 1e-5                                       # the attention stage itself at T'=271


def test_full_size_batch_properties(full_model):
    """64 clips x 16 s: (1) deterministic, (2) batch-invariant bit-for-bit (clip c inside a batch of
    64 == the same clip in a batch of 2: everything per-tensor is per clip), (3) ids are the
    argmax of the logits, (4) host-buffer path == device path."""
    blob, m = full_model
    pcm = synth_batch(0, 64, 256000)
    ids64 = m.transcribe(pcm)
    assert ids64.shape == (64, 271)
    np.testing.assert_array_equal(ids64, m.transcribe(pcm))
    ids2, logits2 = m.transcribe(pcm[[7, 63]], want_logits=True)
    np.testing.assert_array_equal(ids2[0], ids64[7]); np.testing.assert_array_equal(ids2[1], ids64[63])
    np.testing.assert_array_equal(ids2, logits2.shape[2] - 1 - np.argmax(logits2[:, :, ::-1], axis=2))
    assert np.isfinite(logits2).all() and len(np.unique(ids64)) > 10


# ------------------------------------------------------------------------------------------------------------------
# Parity of the BENCHED configuration (BASELINE configs[1]: 64 clips x 16 s, 70 layers, seed-1234 blob, product path =
# tcgen05 3xTF32 attention + CUDA-graph replay) against the oracle.
#
# What can be asked of it.  Stage by stage the path is exact or f32-faithful (integer core bit-exact, LayerNorm / softmax in the
# reference's accumulator order, attention 1e-5): those bars are the tests above.  End to end, the network contains 281 dynamic
# quantisers: a value that sits within an ulp of a .5 rounding boundary flips its u8 code, the flip moves later min/max, and
# the two runs decorrelate until the difference saturates at the quantisation-noise level.  That is a property of the NETWORK,
# not of an implementation, and it is measured here on the oracle itself: the CPU oracle run twice, the second time on features
# multiplied by (1 + 1e-7 * N(0,1)) -- about one ulp -- disagrees with itself by ~1.5 % (layer 10), ~2 % (35), ~3 % (70) of the
# mean magnitude, shares ~93 % of the greedy ids and has a logits MAE of ~0.025 (profiles/r02_oracle_self_sensitivity.json).
# No implementation that differs from the reference by a single rounding anywhere can be closer than that, so that is the bar:
# the product path must deviate from the oracle by no more than the oracle deviates from its own 1-ulp twin (x1.5 margin), on
# top of the reference's own end-to-end bar for this model (logits MAE <= 1.0 and shared arg-max tokens,
# examples/sensevoice/tests/e2e_test.rs:125-185 -- lele vs ONNX Runtime, two f32 implementations with different summation orders).
# The numbers are written to gpurun_out/parity_full_size.json (committed as profiles/r02_parity_full_size.json).
# ------------------------------------------------------------------------------------------------------------------
PARITY_CLIPS = [0, 21, 42, 63]
PARITY_DEPTHS = [1, 10, 35, 70]


def _last_argmax(logits):
    return logits.shape[-1] - 1 - np.argmax(logits[..., ::-1], axis=-1)


def _drift_rows(run_a, run_b, n_clips):
    """run_x(depth, clip) -> hidden state / logits; rows of relative max / mean deviation (and arg-max agreement at depth 70)."""
    rows = []
    for depth in PARITY_DEPTHS:
        row = {"layers": depth, "rel_max": [], "rel_mean": [], "argmax_agreement": [], "mae": []}
        for j in range(n_clips):
            a, b = run_a(depth, j), run_b(depth, j)
            e = np.abs(a - b)
            row["rel_max"].append(float(e.max() / np.abs(b).max())); row["rel_mean"].append(float(e.mean() / np.abs(b).mean()))
            if depth == 70:
                row["argmax_agreement"].append(float((_last_argmax(a) == _last_argmax(b)).mean())); row["mae"].append(float(e.mean()))
        rows.append(row)
    return rows


def _oracle_self_sensitivity(ref, feats):
    """The oracle against itself on features perturbed by ~1 ulp (relative 1e-7 Gaussian noise, seed 0): the network's own floor."""
    rng = np.random.default_rng(0)
    twin = (feats * (1 + 1e-7 * rng.standard_normal(feats.shape))).astype(np.float32)
    nl = lambda d: -1 if d == 70 else d
    return _drift_rows(lambda d, j: ref.forward(twin[j], 3, 0, n_layers=nl(d)), lambda d, j: ref.forward(feats[j], 3, 0, n_layers=nl(d)), len(feats))


def test_benched_configuration_vs_oracle(full_model):
    import json
    import os
    blob, m = full_model
    ref = R.SenseVoiceRef(blob)
    pcm = synth_batch(0, 64, 256000)
    for _ in range(3):                                             # eager pass, capture, replay: the ids compared are a graph replay's
        ids64 = m.transcribe(pcm)
    sub = pcm[PARITY_CLIPS]
    ids4, logits4 = m.transcribe(sub, want_logits=True)
    np.testing.assert_array_equal(ids4, ids64[PARITY_CLIPS])       # batch-invariant: the 4-clip logits are those of the benched batch
    report = {"clips": PARITY_CLIPS, "rows_per_clip": int(ids64.shape[1]), "per_clip_from_pcm": []}
    for j, c in enumerate(PARITY_CLIPS):
        rids, rlog = ref.pcm_to_ids(pcm[c], want_logits=True)     # the whole path on the CPU: front-end, CMVN, 70 layers, CTC head
        err = np.abs(logits4[j] - rlog)
        report["per_clip_from_pcm"].append({"clip": c, "ids_agreement": float((ids64[c] == rids).mean()), "logits_mae": float(err.mean()),
                                            "logits_max_abs": float(err.max()), "logits_ref_mean_abs": float(np.abs(rlog).mean()),
                                            "logits_ref_max_abs": float(np.abs(rlog).max())})
    # encoder only, identical (oracle-computed) features: hidden state after 1, 10, 35 layers and the logits after all 70
    feats = np.stack([R.cmvn(R.frontend(pcm[c])) for c in PARITY_CLIPS])
    nl = lambda d: -1 if d == 70 else d
    gpu = {d: m.forward(feats, 3, 0, n_layers=nl(d)) for d in PARITY_DEPTHS}
    report["gpu_vs_oracle"] = _drift_rows(lambda d, j: gpu[d][j], lambda d, j: ref.forward(feats[j], 3, 0, n_layers=nl(d)), len(PARITY_CLIPS))
    report["oracle_vs_its_one_ulp_twin"] = _oracle_self_sensitivity(ref, feats)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "parity_full_size.json"), "w") as fh:
        json.dump(report, fh, indent=1)
    print("PARITY", json.dumps(report))
    floor = {r["layers"]: r for r in report["oracle_vs_its_one_ulp_twin"]}
    for row in report["gpu_vs_oracle"]:
        f = floor[row["layers"]]
        if row["layers"] >= 10:                                    # (at depth 1 the twin has hardly flipped a code yet; the stage bars above cover it)
            assert max(row["rel_mean"]) <= 1.5 * max(f["rel_mean"]), (row, f)
        if row["layers"] == 70:
            assert min(row["argmax_agreement"]) >= min(f["argmax_agreement"]) - 0.05, (row, f)
            assert max(row["mae"]) <= 1.5 * max(f["mae"]), (row, f)
    assert max(report["gpu_vs_oracle"][0]["rel_mean"]) < 1e-3      # one layer in: 1e-4-class (a handful of flipped codes)
    for r in report["per_clip_from_pcm"]:
        assert r["logits_mae"] <= 1.0, r                           # the reference's own bar (e2e_test.rs:143)
        assert r["logits_mae"] <= 0.1 * r["logits_ref_mean_abs"], r   # measured 0.03-0.08 of mean |logit| (front-end 1e-4 differences flip layer-0 codes)
        assert r["ids_agreement"] >= 0.75, r                       # measured 0.83-0.93


def test_full_size_simt_attention_is_bit_identical():
    """Same clips, same seed-1234 70-layer blob, CUDA-core attention in the oracle's summation order (LELE_B200_ATTN_SIMT=1): on
    identical features the whole encoder + CTC head IS the oracle's arithmetic -- integer core exact, LayerNorm / softmax in the AVX2
    accumulator order, the polynomial exp in the SIMD bodies and a correctly rounded exp in the scalar tails (lb_libm_expf) -- and the
    result is bit-identical at full size: hidden state after 1, 10, 35 layers, the 271 x 25055 logits after 70, and the ids.  (Before
    the scalar-tail exp was made libm-exact, 1-ulp differences in 7 of the 271 softmax columns were enough to flip quantiser codes
    from layer ~20 on and put this path on the same floor as the tensor-core one -- which is what shows the floor is the network's.)"""
    import json
    import os
    blob = build_blob(SenseVoiceConfig(), seed=1234)
    m = _make(blob, "simt", max_clips=2, max_samples=256000)
    try:
        ref = R.SenseVoiceRef(blob)
        pcm = synth_batch(0, 64, 256000)[[0, 63]]
        feats = np.stack([R.cmvn(R.frontend(p)) for p in pcm])
        nl = lambda d: -1 if d == 70 else d
        gpu = {d: m.forward(feats, 3, 0, n_layers=nl(d)) for d in PARITY_DEPTHS}
        rows = _drift_rows(lambda d, j: gpu[d][j], lambda d, j: ref.forward(feats[j], 3, 0, n_layers=nl(d)), 2)
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", "parity_full_size_simt.json"), "w") as fh:
            json.dump({"gpu_simt_vs_oracle": rows}, fh, indent=1)
        print("PARITY_SIMT", json.dumps(rows))
        for d in PARITY_DEPTHS:
            for j in range(2):
                np.testing.assert_array_equal(gpu[d][j], ref.forward(feats[j], 3, 0, n_layers=nl(d)), err_msg=f"depth {d} clip {j}")
        got, ids = m.forward(feats, 3, 0, want_ids=True)
        np.testing.assert_array_equal(ids, _last_argmax(got))       # the fused arg-max epilogue is the last-max arg-max of the logits it writes
        for j in range(2):
            np.testing.assert_array_equal(ids[j], _last_argmax(ref.forward(feats[j], 3, 0)))
    finally:
        m.close()
