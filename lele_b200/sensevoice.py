"""SenseVoice-shaped model object over the C-ABI graph runner (csrc/sensevoice.cu).

Mirrors what the reference app does with its generated model (examples/sensevoice/src/main.rs):
`SenseVoice.new(weights_blob)` keeps the blob resident (here: in HBM), `forward(speech, ...)`
takes CMVN'd features, `transcribe(pcm)` is the end-to-end host-buffer call
(front-end + CMVN + encoder + greedy ids).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import LeleB200Error, call, i32, lib, sz, vp
from .kernels import Context, DevBuf, default_context

lib.lele_b200_sensevoice_rows.argtypes = [vp, i32]
lib.lele_b200_sensevoice_vocab.argtypes = [vp]


class SenseVoice:
    def __init__(self, blob: np.ndarray, max_clips: int = 64, max_samples: int = 256000, ctx: Context | None = None,
                 blob_dev_ptr: int | None = None):
        """blob: the weights blob (lele_b200.sensevoice_weights.build_blob).  blob_dev_ptr: optional device copy that
        already exists (e.g. a torch tensor that received the NCCL weight broadcast)."""
        self.ctx = ctx or default_context()
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        hdr = blob[:256].view(np.int32)
        n_t = int(hdr[12])
        self.header = np.ascontiguousarray(blob[:256 + 16 * n_t])
        self.n_layers, self.d_model, self.d_in, self.vocab = int(hdr[2]), int(hdr[3]), int(hdr[4]), int(hdr[8])
        self._own = None
        if blob_dev_ptr is None:
            self._own = self.ctx.upload(blob, np.uint8)
            blob_dev_ptr = self._own.ptr
        h = vp()
        call("lele_b200_sensevoice_create", self.ctx.h, vp(blob_dev_ptr), sz(blob.size), self.header.ctypes.data_as(vp),
             sz(self.header.size), i32(max_clips), i32(max_samples), C.byref(h))
        self.h = h
        self.max_clips, self.max_samples = max_clips, max_samples

    def rows(self, n_samples: int) -> int:
        return int(lib.lele_b200_sensevoice_rows(self.h, i32(n_samples)))

    # ---- device-pointer forms (bench uses these with torch tensors) ----
    def forward_pcm_dev(self, pcm_ptr: int, n_clips: int, n_samples: int, ids_ptr: int | None, logits_ptr: int | None = None,
                        lang: int = 3, textnorm: int = 0, n_layers: int = -1):
        call("lele_b200_sensevoice_forward", self.ctx.h, self.h, vp(pcm_ptr), i32(n_clips), i32(n_samples), i32(lang), i32(textnorm),
             i32(n_layers), vp(ids_ptr), vp(logits_ptr))

    def transcribe_host_ptr(self, pcm_host_ptr: int, n_clips: int, n_samples: int, ids_host_ptr: int, lang: int = 3, textnorm: int = 0):
        call("lele_b200_sensevoice_transcribe_host", self.ctx.h, self.h, vp(pcm_host_ptr), i32(n_clips), i32(n_samples), i32(lang),
             i32(textnorm), vp(ids_host_ptr))

    def transcribe_host_async(self, pcm_host_ptr: int, n_clips: int, n_samples: int, ids_host_ptr: int, slot: int, lang: int = 3, textnorm: int = 0):
        """Pipelined serving form: submit into slot 0/1, collect with transcribe_wait(slot); copies overlap the forward."""
        call("lele_b200_sensevoice_transcribe_host_async", self.ctx.h, self.h, vp(pcm_host_ptr), i32(n_clips), i32(n_samples), i32(lang),
             i32(textnorm), vp(ids_host_ptr), i32(slot))

    def set_comm(self, comm, root: int = 0):
        """Attach a lele_b200.distributed.Comm: transcribe_host_async then gathers all ranks' ids on the device and only `root`
        receives them ([world, n_clips, T'] in its ids buffer); pass ids_host_ptr = None on the other ranks."""
        call("lele_b200_sensevoice_set_comm", self.h, None if comm is None else comm.h, i32(root))

    def transcribe_wait(self, slot: int):
        call("lele_b200_sensevoice_transcribe_wait", self.ctx.h, self.h, i32(slot))

    # ---- numpy forms ----
    def forward(self, speech, language: int = 3, text_norm: int = 0, n_layers: int = -1, want_ids: bool = False):
        """model.forward(speech [B,T,560] or [T,560], speech_lengths, language, text_norm) -> logits [B,T+4,vocab]
        (main.rs:140).  With n_layers < total the hidden state after that many layers is returned instead."""
        x = np.ascontiguousarray(speech, dtype=np.float32)
        single = x.ndim == 2
        if single:
            x = x[None]
        b, t, din = x.shape
        if din != self.d_in:
            raise LeleB200Error(f"forward: feature dim {din} != {self.d_in}")
        full = n_layers < 0 or n_layers >= self.n_layers
        width = self.vocab if full else (self.d_in if n_layers == 0 else self.d_model)
        T = t + 4
        bx = self.ctx.upload(x); out = self.ctx.empty(b * T * width); ids = self.ctx.empty(b * T)
        call("lele_b200_sensevoice_forward_features", self.ctx.h, self.h, vp(bx.ptr), i32(b), i32(t), i32(language), i32(text_norm),
             i32(n_layers), vp(ids.ptr if (want_ids and full) else None), vp(out.ptr))
        res = self.ctx.download(out, (b, T, width))
        idh = self.ctx.download(ids, (b, T), np.int32) if (want_ids and full) else None
        for q in (bx, out, ids):
            q.free()
        if single:
            res = res[0]; idh = None if idh is None else idh[0]
        return (res, idh) if want_ids else res

    def transcribe(self, pcm, language: int = 3, text_norm: int = 0, want_logits: bool = False):
        """pcm [B, n] host floats in [-1,1) -> greedy ids [B, T'] (+ logits).  Host buffers in,
        host buffers out: the H2D copy, every kernel and the D2H copy run on the context stream."""
        p = np.ascontiguousarray(pcm, dtype=np.float32)
        single = p.ndim == 1
        if single:
            p = p[None]
        b, n = p.shape
        T = self.rows(n)
        if T == 0:
            raise LeleB200Error("transcribe: clip shorter than one frame (400 samples)")
        if not want_logits:
            ids = np.empty((b, T), np.int32)
            self.transcribe_host_ptr(p.ctypes.data, b, n, ids.ctypes.data, language, text_norm)
            return ids[0] if single else ids
        bp = self.ctx.upload(p); ids = self.ctx.empty(b * T); lg = self.ctx.empty(b * T * self.vocab)
        self.forward_pcm_dev(bp.ptr, b, n, ids.ptr, lg.ptr, language, text_norm)
        idh = self.ctx.download(ids, (b, T), np.int32); lgh = self.ctx.download(lg, (b, T, self.vocab))
        for q in (bp, ids, lg):
            q.free()
        return (idh[0], lgh[0]) if single else (idh, lgh)

    def transcribe_text(self, pcm, tokenizer, language: int = 3, text_norm: int = 0):
        """pcm [B, n] -> one string per clip: the whole example pipeline (main.rs:73-150) with only the kept token ids
        of each clip leaving the GPU (device arg-max in the CTC epilogue + device greedy filter, tokenizer.rs:37)."""
        p = np.ascontiguousarray(pcm, dtype=np.float32)
        if p.ndim == 1:
            p = p[None]
        b, n = p.shape
        T = self.rows(n)
        if T == 0:
            raise LeleB200Error("transcribe_text: clip shorter than one frame (400 samples)")
        bp = self.ctx.upload(p); ids = self.ctx.empty(b * T)
        self.forward_pcm_dev(bp.ptr, b, n, ids.ptr, None, language, text_norm)
        kept = tokenizer.filter_ids_device(ids.ptr, b, T, self.ctx)
        bp.free(); ids.free()
        return [tokenizer.text(k) for k in kept]

    def workspace(self, name: str, shape, dtype=np.float32) -> np.ndarray:
        """Host copy of the leading `shape` elements of a workspace buffer after a forward
        (forward_with_workspace-style borrow; used by tests to localise divergences)."""
        p = vp(); nb = sz(0)
        call("lele_b200_sensevoice_workspace", self.h, name.encode(), C.byref(p), C.byref(nb))
        out = np.empty(shape, dtype=dtype)
        if out.nbytes > nb.value:
            raise LeleB200Error(f"workspace {name}: {out.nbytes} bytes requested, buffer holds {nb.value}")
        self.ctx.sync()
        call("lele_b200_d2h", self.ctx.h, out.ctypes.data_as(vp), p, sz(out.nbytes))
        self.ctx.sync()
        return out

    # ---- profiling (kernels/timing.rs analogue) ----
    def set_profiling(self, on: bool):
        call("lele_b200_sensevoice_set_profiling", self.h, i32(int(on)))

    def last_profile(self) -> dict:
        cap = 32
        names = (C.c_char_p * cap)(); ms = (C.c_float * cap)(); calls = (C.c_int * cap)(); n = C.c_int(0)
        call("lele_b200_sensevoice_last_profile", self.h, names, ms, calls, i32(cap), C.byref(n))
        return {names[i].decode(): {"ms": float(ms[i]), "calls": int(calls[i])} for i in range(n.value)}

    def close(self):
        if getattr(self, "h", None):
            lib.lele_b200_sensevoice_destroy(self.ctx.h, self.h)
            self.h = None
        if self._own is not None:
            self._own.free(); self._own = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
