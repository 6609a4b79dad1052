"""Host-side mirror of the SenseVoice example's `Tokenizer` (examples/sensevoice/src/tokenizer.rs).

`from_file` / `decode_greedy` keep the reference's names and semantics (greedy arg-max per frame with the
last-maximum tie rule, blank id 0 and "<|...|>" special tokens skipped, "▁" -> space, trim).  The id filter
itself runs on the device (`lele_b200_greedy_filter`), so only the kept ids of each clip leave the GPU --
SURVEY.md 8(f) rank 3, the step immediately after the hot path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import LeleB200Error, call, i32, vp
from .kernels import default_context


class Tokenizer:
    def __init__(self, id_to_token):
        self.id_to_token = list(id_to_token)
        self._mask_dev = None   # (ctx, DevBuf)

    @classmethod
    def from_file(cls, path):  # tokenizer.rs:10-34: "<token> <id>" per line, split at the LAST space
        id_to_token = []
        with open(path, encoding="utf-8") as fh:
            for line in fh:
                line = line.rstrip("\n")
                parts = line.rsplit(" ", 1)
                if len(parts) != 2:
                    continue
                try:
                    idx = int(parts[1])
                except ValueError:
                    idx = 0
                if idx < 0:
                    idx = 0
                if idx >= len(id_to_token):
                    id_to_token.extend([""] * (idx + 1 - len(id_to_token)))
                id_to_token[idx] = parts[0]
        return cls(id_to_token)

    def vocab_size(self) -> int:
        return len(self.id_to_token)

    def skip_mask(self) -> np.ndarray:
        """1 for ids the greedy decode drops: blank (0) and "<|...|>" tokens (tokenizer.rs:64)."""
        m = np.zeros(max(len(self.id_to_token), 1), np.uint8)
        m[0] = 1
        for i, tok in enumerate(self.id_to_token):
            if tok.startswith("<|") and tok.endswith("|>"):
                m[i] = 1
        return m

    def text(self, kept_ids) -> str:  # tokenizer.rs:71-79
        toks = [self.id_to_token[i] for i in kept_ids if 0 <= i < len(self.id_to_token)]
        return "".join(toks).replace("▁", " ").strip()

    def filter_ids_device(self, ids_dev_ptr: int, n_clips: int, t: int, ctx=None):
        """ids [n_clips, t] int32 already in HBM (the runner's output) -> list of kept-id arrays, one per clip."""
        ctx = ctx or default_context()
        if self._mask_dev is None or self._mask_dev[0] is not ctx:
            self._mask_dev = (ctx, ctx.upload(self.skip_mask(), np.uint8))
        out = ctx.empty(max(n_clips * t, 1)); ln = ctx.empty(max(n_clips, 1))
        call("lele_b200_greedy_filter", ctx.h, vp(ids_dev_ptr), i32(n_clips), i32(t), vp(self._mask_dev[1].ptr), i32(len(self.id_to_token)),
             vp(out.ptr), vp(ln.ptr))
        lens = ctx.download(ln, (n_clips,), np.int32)
        kept = ctx.download(out, (n_clips, t), np.int32)
        out.free(); ln.free()
        return [kept[c, :lens[c]].copy() for c in range(n_clips)]

    def decode_ids(self, ids, ctx=None):
        """Host ids [n_clips, t] -> texts (device filter)."""
        ctx = ctx or default_context()
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if ids.ndim != 2:
            raise LeleB200Error("decode_ids expects [n_clips, t] ids")
        b = ctx.upload(ids, np.int32)
        kept = self.filter_ids_device(b.ptr, ids.shape[0], ids.shape[1], ctx)
        b.free()
        return [self.text(k) for k in kept]

    def decode_greedy(self, logits, batch_size: int, time_steps: int, vocab_size: int, ctx=None):
        """tokenizer.rs:37 -- logits [batch, time, vocab] -> texts; arg-max (last maximum wins) and the id filter on the device."""
        ctx = ctx or default_context()
        lg = np.ascontiguousarray(logits, dtype=np.float32).reshape(-1)
        if lg.size != batch_size * time_steps * vocab_size:
            raise LeleB200Error("decode_greedy: logits length mismatch (tokenizer.rs:44)")
        bl = ctx.upload(lg); ids = ctx.empty(max(batch_size * time_steps, 1))
        call("lele_b200_argmax_last", ctx.h, vp(bl.ptr), C.c_longlong(batch_size * time_steps), i32(vocab_size), vp(ids.ptr))
        kept = self.filter_ids_device(ids.ptr, batch_size, time_steps, ctx)
        bl.free(); ids.free()
        return [self.text(k) for k in kept]
