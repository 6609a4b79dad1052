"""Host-side mirror of `lele::features` (src/features/*.rs) over the C ABI.

`SenseVoiceFrontend`, `FeatureConfig`, `Cmvn`, `Lfr`, `hann_window`, `mel_filterbank`, `RealFft`
keep the reference's names and semantics; compute runs in the CUDA kernels of csrc/frontend.cu.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from ._lib import LeleB200Error, call, f32, i32, i64, lib, vp
from .kernels import _f, _run, default_context


@dataclasses.dataclass
class FeatureConfig:  # pipeline.rs:7-27
    sample_rate: int = 16000
    n_mels: int = 80
    frame_length_ms: float = 25.0
    frame_shift_ms: float = 10.0
    lfr_m: int = 7
    lfr_n: int = 6


def hann_window(size: int) -> np.ndarray:  # window.rs:2
    out = np.empty(size, np.float32)
    call("lele_b200_hann_window", i32(size), out.ctypes.data_as(vp))
    return out


def mel_filterbank(sample_rate, n_fft, n_mels, f_min, f_max=None) -> np.ndarray:  # mel.rs:7
    f_max = sample_rate / 2.0 if f_max is None else f_max
    out = np.empty((n_mels, n_fft // 2 + 1), np.float32)
    call("lele_b200_mel_filterbank", f32(sample_rate), i32(n_fft), i32(n_mels), f32(f_min), f32(f_max), out.ctypes.data_as(vp))
    return out


def hz_to_mel_htk(hz: float) -> float:  # mel.rs:1 (host scalar helper)
    return float(np.float32(2595.0) * np.log10(np.float32(1.0) + np.float32(hz) / np.float32(700.0), dtype=np.float32))


class RealFft:  # features/fft.rs:1-50
    def __init__(self, length: int):
        self.n = length

    def process(self, x, ctx=None):
        """-> (re, im) of the n/2+1 non-redundant bins, one row per input row."""
        ctx = ctx or default_context()
        x = _f(x).reshape(-1, self.n)
        half = self.n // 2 + 1
        bx = ctx.upload(x); re = ctx.empty(x.shape[0] * half); im = ctx.empty(x.shape[0] * half)
        call("lele_b200_rfft", ctx.h, vp(bx.ptr), i32(x.shape[0]), i32(self.n), vp(re.ptr), vp(im.ptr))
        out = ctx.download(re, (x.shape[0], half)), ctx.download(im, (x.shape[0], half))
        for b in (bx, re, im):
            b.free()
        return out


class Lfr:  # lfr.rs
    def __init__(self, m: int = 7, n: int = 6):
        self.m, self.n = m, n

    def compute(self, x, ctx=None):
        x = _f(x)
        if x.ndim == 3 and x.shape[0] == 1:
            x = x[0]
        if x.ndim != 2:
            raise LeleB200Error(f"LFR expects [T, D] or [1, T, D] input, got {list(x.shape)} (lfr.rs:24)")
        t, d = x.shape
        t_lfr = (t + self.n - 1) // self.n
        if t == 0:
            return np.zeros((0, d * self.m), np.float32)
        return _run((t_lfr, d * self.m), lambda c, o, px: call("lele_b200_lfr", c.h, px, i32(1), i32(t), i32(d), i32(self.m), i32(self.n), o), x, ctx=ctx)


class Cmvn:  # cmvn.rs
    def __init__(self, eps: float = 1e-5):
        self.eps = eps

    def compute(self, x, ctx=None):
        x = _f(x)
        shp = x.shape
        if x.ndim == 3 and x.shape[0] == 1:
            x2 = x[0]
        elif x.ndim == 2:
            x2 = x
        else:
            raise LeleB200Error(f"CMVN expects [T, D] or [1, T, D] input, got {list(shp)} (cmvn.rs:22)")
        t, d = x2.shape
        if t == 0:
            return x.copy()
        return _run(shp, lambda c, o, px: call("lele_b200_cmvn", c.h, px, i32(1), i32(t), i32(d), f32(self.eps), o), x2, ctx=ctx)


class SenseVoiceFrontend:  # pipeline.rs:28-193
    def __init__(self, config: FeatureConfig = FeatureConfig()):
        self.config = config
        frame_len = int(np.float32(config.sample_rate) * np.float32(config.frame_length_ms) / np.float32(1000.0))
        hop = int(np.float32(config.sample_rate) * np.float32(config.frame_shift_ms) / np.float32(1000.0))
        if (config.sample_rate, config.n_mels, frame_len, hop, config.lfr_m, config.lfr_n) != (16000, 80, 400, 160, 7, 6):
            raise LeleB200Error("SenseVoiceFrontend: only the SenseVoice configuration (16 kHz, 80 mel, 25/10 ms, LFR 7/6) is built")

    def compute(self, pcm, want_mel=False, ctx=None):
        """pcm [n] or [n_clips, n] -> [T_lfr, 560] or [n_clips, T_lfr, 560] (empty when shorter than a frame)."""
        ctx = ctx or default_context()
        pcm = _f(pcm)
        single = pcm.ndim == 1
        p2 = pcm.reshape(1, -1) if single else pcm
        nclips, n = p2.shape
        frames = lib.lele_b200_frontend_num_frames(i32(n))
        if frames == 0:
            return np.zeros((0, 560) if single else (nclips, 0, 560), np.float32)
        t = (frames + 5) // 6
        bp = ctx.upload(p2); out = ctx.empty(nclips * t * 560)
        mel = ctx.empty(nclips * frames * 80) if want_mel else None
        call("lele_b200_frontend_compute", ctx.h, vp(bp.ptr), i32(nclips), i32(n), i64(n), vp(None if mel is None else mel.ptr), vp(out.ptr))
        res = ctx.download(out, (nclips, t, 560))
        melh = ctx.download(mel, (nclips, frames, 80)) if want_mel else None
        for b in (bp, out, mel):
            if b is not None:
                b.free()
        if single:
            res = res[0]; melh = None if melh is None else melh[0]
        return (melh, res) if want_mel else res
