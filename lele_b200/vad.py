"""Streaming voice-activity front end: the caller of examples/silero/src/main.rs over a replayed `model.rs`.

SURVEY.md 8f rank 4: the step in front of the ASR path.  The reference's loop (main.rs:70-137) pads the clip to a multiple of
512 samples, scales each chunk by 32768, and calls `forward_with_workspace(ws, input [1,512], state [2,1,128], sr [1] i64)`
once per chunk, carrying the returned state into the next call; the per-chunk speech probabilities are then turned into
sample-accurate segments (main.rs:155-229).  This module mirrors both halves:

* `StreamingVad` drives any parsed generated model with that signature through `model_rs.run_program` (the operators run
  on the B200 through the C ABI; the loop and the state hand-over are host code, as upstream);
* `collect_segments` / `merge_segments` are the segment state machine and the merge pass, same integer arithmetic.

The Silero graph itself is not in the reference checkout (no model files), so the tests drive a synthetic
recurrent model with the same calling convention.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import model_rs

__all__ = ["VadConfig", "StreamingVad", "collect_segments", "merge_segments", "ms_to_samples"]


@dataclass(frozen=True)
class VadConfig:
    """main.rs:18-28 defaults"""
    threshold: float = 0.3
    min_silence_ms: float = 200.0
    min_speech_ms: float = 400.0
    speech_pad_ms: float = 120.0
    merge_gap_ms: float = 200.0


def ms_to_samples(ms: float, sample_rate: int) -> int:
    """main.rs:157: `((sr as f32) * (ms / 1000.0)).round() as usize` -- f32 arithmetic, ties away from zero."""
    v = np.float32(sample_rate) * (np.float32(ms) / np.float32(1000.0))
    return int(np.floor(np.float64(v) + 0.5)) if v > 0 else 0


def collect_segments(probs, n_samples: int, sample_rate: int = 16000, chunk_size: int = 512, config: VadConfig = VadConfig()):
    """The trigger / release state machine of main.rs:164-205.  `probs[i]` is the speech probability of chunk i of the padded clip,
    `n_samples` the unpadded clip length.  Returns [(start, end)] in samples, in detection order."""
    padded_len = -(-n_samples // chunk_size) * chunk_size
    min_silence = max(ms_to_samples(config.min_silence_ms, sample_rate), 1)
    min_speech = max(ms_to_samples(config.min_speech_ms, sample_rate), 1)
    pad = ms_to_samples(config.speech_pad_ms, sample_rate)
    thr = np.float32(config.threshold)
    segments, triggered, start, silence = [], False, 0, 0
    for i, p in enumerate(probs):
        offset = i * chunk_size
        frame_end = min(offset + chunk_size, padded_len)
        if np.float32(p) >= thr:
            if not triggered:
                triggered, start = True, max(offset - pad, 0)      # saturating_sub
            silence = 0
        elif triggered:
            silence += frame_end - offset
            if silence >= min_silence:
                end = min(frame_end + pad, n_samples)
                if end > start and end - start >= min_speech:
                    segments.append((start, end))
                triggered, silence = False, 0
    if triggered:
        end = n_samples
        if end > start and end - start >= min_speech:
            segments.append((start, end))
    return segments


def merge_segments(segments, sample_rate: int = 16000, config: VadConfig = VadConfig()):
    """main.rs:208-229: sort by start; overlapping segments and gaps of at most merge_gap_ms are joined."""
    gap_max = ms_to_samples(config.merge_gap_ms, sample_rate)
    merged: list[list[int]] = []
    for s, e in sorted(segments, key=lambda se: se[0]):
        if merged and (s <= merged[-1][1] or s - merged[-1][1] <= gap_max):
            merged[-1][1] = max(merged[-1][1], e)
            continue
        merged.append([s, e])
    return [(s, e) for s, e in merged]


class StreamingVad:
    """One audio stream.  `program` = `model_rs.parse_model_rs(...)` of a generated model whose forward takes
    (input [1, chunk], state, sr [1] i64) and returns (probability, new state), in that order (main.rs:121)."""

    def __init__(self, program: dict, blob, ops=None, sample_rate: int = 16000, chunk_size: int = 512, state_shape=(2, 1, 128)):
        if len(program["inputs"]) != 3 or len(program["outputs"]) != 2:
            raise ValueError("StreamingVad: the model must take (input, state, sr) and return (output, state)")
        self.program, self.blob = program, blob
        self.ops = ops if ops is not None else model_rs.CudaOps()
        self.sample_rate, self.chunk_size, self.state_shape = int(sample_rate), int(chunk_size), tuple(state_shape)
        self._cache = {}                                           # decoded weights / prepared int8 weights: once per stream object
        self.reset()

    def reset(self):
        self.state = np.zeros(self.state_shape, np.float32)       # main.rs:88
        self.probs: list[float] = []

    def push(self, chunk) -> float:
        """One chunk of [-1, 1] samples -> speech probability; the recurrent state moves on."""
        chunk = np.asarray(chunk, np.float32).reshape(-1)
        if chunk.size != self.chunk_size:
            raise ValueError(f"StreamingVad: chunk of {chunk.size} samples, expected {self.chunk_size}")
        x = (chunk * np.float32(32768.0)).reshape(1, self.chunk_size)                                   # main.rs:115
        out, new_state = model_rs.run_program(self.program, self.blob, [x, self.state, np.array([self.sample_rate], np.int64)], self.ops,
                                                  cache=self._cache)
        self.state = np.asarray(new_state, np.float32).reshape(self.state_shape)                       # main.rs:129
        out = np.asarray(out).reshape(-1)
        if out.size:                                                                                    # main.rs:124
            self.probs.append(float(out[0]))
        return self.probs[-1] if self.probs else float("nan")

    def process(self, audio):
        """Whole clip: zero-pad to a multiple of the chunk size (main.rs:72-80) and push every chunk.  Returns the probabilities."""
        audio = np.asarray(audio, np.float32).reshape(-1)
        n_chunks = -(-audio.size // self.chunk_size)
        padded = np.zeros(n_chunks * self.chunk_size, np.float32)
        padded[:audio.size] = audio
        for i in range(n_chunks):
            self.push(padded[i * self.chunk_size:(i + 1) * self.chunk_size])
        return np.asarray(self.probs, np.float32)

    def segments(self, audio, config: VadConfig = VadConfig()):
        """Clip -> merged speech segments [(start, end)] in samples."""
        self.reset()
        audio = np.asarray(audio, np.float32).reshape(-1)
        probs = self.process(audio)
        return merge_segments(collect_segments(probs, audio.size, self.sample_rate, self.chunk_size, config), self.sample_rate, config)
