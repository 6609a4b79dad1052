"""Synthetic SenseVoiceSmall-shaped weight blob and synthetic PCM (numpy only).

No SenseVoice model file exists in the reference checkout (`.gitignore:7-10`), so the
headline workload runs a random-weight network of the documented architecture
(`src/bin/wasm_bench.rs:758,891,1105`, `examples/sensevoice/tests/e2e_test.rs:150`):
70 SANM layers (1 + 49 + 20), d=512, 4 heads, FFN 2048, FSMN k=11, vocab 25055,
int8 (u8 weights + u8 zero point + per-channel f32 scale) linears.

The blob plays the role of lele's `<class>_weights.bin` (`src/compiler/mod.rs:1381-1434`):
raw little-endian tensors at aligned offsets; here a table of (offset, nbytes) stands in
for the literals lele_gen bakes into `model.rs`.  Both the CPU oracle
(`oracle/sensevoice_ref.c`) and the CUDA runner (`csrc/sensevoice.cu`) read this layout.
"""
from __future__ import annotations

import dataclasses
import numpy as np

MAGIC = 0x454C454C  # b"LELE"
NUM_GLOBAL = 10
NUM_LAYER = 21
ALIGN = 256


@dataclasses.dataclass(frozen=True)
class SenseVoiceConfig:
    n_layers: int = 70
    d_model: int = 512
    d_in: int = 560
    ffn: int = 2048
    heads: int = 4
    fsmn_k: int = 11
    vocab: int = 25055
    n_embed: int = 16
    max_t: int = 512
    n_stage1: int = 50  # after_norm follows this many layers; the rest are tp_encoders


def sinusoidal_positions(max_t: int, depth: int) -> np.ndarray:
    """FunASR SinusoidalPositionEncoder: positions 1..T, cat(sin, cos)."""
    pos = np.arange(1, max_t + 1, dtype=np.float64)[:, None]
    inc = np.log(10000.0) / (depth / 2 - 1)
    inv = np.exp(np.arange(depth // 2, dtype=np.float64) * -inc)[None, :]
    st = pos * inv
    return np.concatenate([np.sin(st), np.cos(st)], axis=1).astype(np.float32)


def _linear(rng: np.random.Generator, k: int, n: int):
    w = rng.integers(0, 256, size=(k, n), dtype=np.uint8)
    # (w-128) has std ~74; scale so the dequantised weight has std ~ 1/sqrt(k)
    scale = ((1.0 / 74.0) / np.sqrt(k) * (1.0 + 0.1 * rng.uniform(-1, 1, size=n))).astype(np.float32)
    bias = (0.02 * rng.standard_normal(n)).astype(np.float32)
    zp = np.array([128], dtype=np.uint8)
    return w, scale, bias, zp


def build_blob(cfg: SenseVoiceConfig = SenseVoiceConfig(), seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    tensors: list[np.ndarray] = []
    d, din = cfg.d_model, cfg.d_in

    def ln(n):
        return ((1.0 + 0.02 * rng.standard_normal(n)).astype(np.float32),
                (0.02 * rng.standard_normal(n)).astype(np.float32))

    embed = rng.standard_normal((cfg.n_embed, din)).astype(np.float32)
    pos = sinusoidal_positions(cfg.max_t, din)
    ag, ab = ln(d)
    tg, tb = ln(d)
    cw, cs, cb, cz = _linear(rng, d, cfg.vocab)
    tensors += [embed, pos, ag, ab, tg, tb, cw, cs, cb, cz]
    assert len(tensors) == NUM_GLOBAL
    for l in range(cfg.n_layers):
        cur = din if l == 0 else d
        g1, b1 = ln(cur)
        qw, qs, qb, qz = _linear(rng, cur, 3 * d)
        fs = (0.3 * rng.standard_normal((d, 1, cfg.fsmn_k))).astype(np.float32)
        ow, os_, ob, oz = _linear(rng, d, d)
        g2, b2 = ln(d)
        w1, s1, bb1, z1 = _linear(rng, d, cfg.ffn)
        w2, s2, bb2, z2 = _linear(rng, cfg.ffn, d)
        tensors += [g1, b1, qw, qs, qb, qz, fs, ow, os_, ob, oz, g2, b2, w1, s1, bb1, z1, w2, s2, bb2, z2]
    n_t = len(tensors)
    assert n_t == NUM_GLOBAL + cfg.n_layers * NUM_LAYER
    table_off = 256
    off = table_off + 16 * n_t
    off = (off + ALIGN - 1) // ALIGN * ALIGN
    table = np.zeros((n_t, 2), dtype=np.uint64)
    for i, t in enumerate(tensors):
        table[i] = (off, t.nbytes)
        off = (off + t.nbytes + ALIGN - 1) // ALIGN * ALIGN
    blob = np.zeros(off, dtype=np.uint8)
    hdr = np.zeros(64, dtype=np.int32)
    hdr[:13] = [MAGIC, 1, cfg.n_layers, d, din, cfg.ffn, cfg.heads, cfg.fsmn_k, cfg.vocab,
                cfg.n_embed, cfg.max_t, cfg.n_stage1, n_t]
    blob[:256] = hdr.view(np.uint8)
    blob[table_off:table_off + 16 * n_t] = table.reshape(-1).view(np.uint8)
    for i, t in enumerate(tensors):
        o = int(table[i, 0])
        blob[o:o + t.nbytes] = np.ascontiguousarray(t).reshape(-1).view(np.uint8)
    return blob


def _tensor_nbytes(cfg: SenseVoiceConfig) -> list[int]:
    d, din, v, f = cfg.d_model, cfg.d_in, cfg.vocab, cfg.ffn
    lin = lambda k, n: [k * n, 4 * n, 4 * n, 1]            # w u8, scale f32, bias f32, zp u8
    out = [4 * cfg.n_embed * din, 4 * cfg.max_t * din, 4 * d, 4 * d, 4 * d, 4 * d] + lin(d, v)
    for l in range(cfg.n_layers):
        cur = din if l == 0 else d
        out += [4 * cur, 4 * cur] + lin(cur, 3 * d) + [4 * d * cfg.fsmn_k] + lin(d, d) + [4 * d, 4 * d] + lin(d, f) + lin(f, d)
    return out


def blob_nbytes(cfg: SenseVoiceConfig = SenseVoiceConfig()) -> int:
    """Size of build_blob(cfg) without generating it (ranks that receive the NCCL broadcast)."""
    sizes = _tensor_nbytes(cfg)
    off = (256 + 16 * len(sizes) + ALIGN - 1) // ALIGN * ALIGN
    for s in sizes:
        off = (off + s + ALIGN - 1) // ALIGN * ALIGN
    return off


def synth_pcm(clip_id: int, n_samples: int = 256000) -> np.ndarray:
    """0.1*sin(2*pi*f_c*n/16000) + 0.01*u[n], f_c = 200+37*clip_id,
    u ~ uniform(-1,1) from a 64-bit LCG seeded 0x5EED0000+clip_id."""
    n = np.arange(n_samples, dtype=np.float64)
    fc = 200.0 + 37.0 * clip_id
    # vectorised 64-bit LCG (Knuth MMIX constants): x_{i+1} = a*x_i + c mod 2^64
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    x = np.empty(n_samples, dtype=np.uint64)
    s = np.uint64(0x5EED0000 + clip_id)
    # jump-ahead by doubling so the generator stays O(n) in numpy
    with np.errstate(over="ignore"):
        x[0] = a * s + c
        filled = 1
        am, cm = a, c
        while filled < n_samples:
            take = min(filled, n_samples - filled)
            x[filled:filled + take] = am * x[:take] + cm
            cm = am * cm + cm
            am = am * am
            filled += take
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24) * 2.0 - 1.0
    return (0.1 * np.sin(2 * np.pi * fc * n / 16000.0) + 0.01 * u).astype(np.float32)


def synth_batch(first_clip: int, n_clips: int, n_samples: int = 256000) -> np.ndarray:
    return np.stack([synth_pcm(first_clip + i, n_samples) for i in range(n_clips)])
