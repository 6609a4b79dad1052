"""Host-side mirror of `lele::kernels::*` over the C ABI (include/lele_b200.h).

Same names, argument order and meaning as the reference operators (src/kernels/mod.rs:23-39);
two tensor types stand in for `TensorView` (src/tensor.rs:5):
  * a numpy array (host memory): the call uploads its operands, launches on the context stream and downloads the result --
    the form the parity tests use;
  * a `DeviceTensor` (the view's `data` lives in HBM): when any operand is one, nothing is copied -- operands are passed as
    device pointers, the result is a `DeviceTensor` placed in the workspace buffer the caller named (`Context.out_slots`, the
    `&mut ws.buf_N` argument of the generated code, backed by `lele_b200_arena_bind`) or in a fresh allocation it owns.  This is
    the drop-in execution form: a replayed `model.rs` keeps every value resident (lele_b200/model_rs.py).
Precondition failures raise `LeleB200Error` where the reference panics.  No CPU fallback: without a CUDA device every call fails.
"""
from __future__ import annotations

import builtins as _b
import ctypes as C

import numpy as np

from ._lib import LeleB200Error, call, f32, i32, i64, lib, sz, vp

__all__ = ["Context", "DeviceTensor", "Workspace", "Graph", "default_context", "LeleB200Error"]


class DevBuf:
    """A device allocation owned by a Context (freed with it or explicitly)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = vp()
        call("lele_b200_malloc", ctx.h, sz(_b.max(self.nbytes, 16)), C.byref(p))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            call("lele_b200_free", self.ctx.h, vp(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceTensor:
    """`TensorView<f32>` whose data lives in HBM: (device pointer, shape) plus whatever keeps the memory alive (a `DevBuf` it
    owns, a `Workspace` slot, or another tensor it is a zero-copy view of -- reshape / flatten / squeeze / unsqueeze, shape.rs)."""
    __slots__ = ("ptr", "shape", "ctx", "owner", "dtype", "slot")

    def __init__(self, ptr, shape, ctx, owner=None, dtype=np.float32, slot=None):
        self.ptr, self.shape, self.ctx, self.owner, self.dtype = int(ptr), tuple(int(d) for d in shape), ctx, owner, np.dtype(dtype)
        self.slot = slot              # name of the workspace buffer holding the data (None: an allocation of its own)

    @property
    def ndim(self): return len(self.shape)

    @property
    def size(self): return _prod(self.shape)

    @property
    def nbytes(self): return self.size * self.dtype.itemsize

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)) else shape
        shape = [int(d) for d in (shape if hasattr(shape, "__len__") else [shape])]
        if -1 in shape:
            known = _prod([d for d in shape if d != -1])
            shape[shape.index(-1)] = self.size // _b.max(known, 1)
        if _prod(shape) != self.size:
            raise LeleB200Error(f"DeviceTensor.reshape: {self.shape} -> {tuple(shape)} changes the element count")
        return DeviceTensor(self.ptr, shape, self.ctx, self, self.dtype, self.slot)

    def numpy(self) -> np.ndarray:
        """`.data` read on the host (the one place a resident value crosses PCIe): joins the stream first."""
        out = np.empty(self.shape, dtype=self.dtype)
        if out.nbytes:
            call("lele_b200_d2h", self.ctx.h, out.ctypes.data_as(vp), vp(self.ptr), sz(out.nbytes))
        self.ctx.sync()
        return out

    def _own(self, buf: "DevBuf"):
        self.ptr, self.owner = buf.ptr, buf
        return self

    def __repr__(self):
        return f"DeviceTensor(shape={self.shape}, ptr=0x{self.ptr:x})"


class Workspace:
    """The generated `<Model>Workspace` (src/compiler/mod.rs:1057-1070: one grow-only `Vec<f32>` per allocator colour, `buf_N`)
    mapped to HBM: every named buffer is a grow-only device mirror obtained from `lele_b200_arena_bind`, keyed by a stable host
    address that stands in for the `Vec`'s own (`&ws.buf_N`).  Sizes settle after the first forward, as upstream
    (kernels/utils.rs:10 `ensure_capacity`)."""

    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        self._keys: dict[str, C.Array] = {}
        self.bytes: dict[str, int] = {}

    def _key(self, name: str):
        k = self._keys.get(name)
        if k is None:
            k = self._keys[name] = (C.c_char * 8)()          # its address is the arena key (the Vec object's address upstream)
        return C.addressof(k)

    def bind(self, name: str, nbytes: int) -> int:
        """Device pointer of buffer `name`, grown (contents kept) to hold at least nbytes."""
        p = vp()
        call("lele_b200_arena_bind", self.ctx.h, vp(self._key(name)), sz(_b.max(int(nbytes), 16)), C.byref(p))
        self.bytes[name] = _b.max(self.bytes.get(name, 0), int(nbytes))
        return p.value

    def tensor(self, name: str, shape, dtype=np.float32) -> DeviceTensor:
        n = _prod(shape) * np.dtype(dtype).itemsize
        return DeviceTensor(self.bind(name, n), shape, self.ctx, self, dtype, name)

    def release(self):
        for name in list(self._keys):
            call("lele_b200_arena_release", self.ctx.h, vp(self._key(name)))
        self._keys.clear(); self.bytes.clear()

    def total_bytes(self) -> int:
        return sum(self.bytes.values())


class Context:
    def __init__(self, device: int = 0, stream: int | None = None):
        h = vp()
        call("lele_b200_ctx_create", i32(device), vp(stream), C.byref(h))
        self.h = h
        self.device = device
        self._slots: list = []          # output placements for the next resident call(s): (Workspace, name) pairs, consumed in order
        self._consts: dict = {}         # host operands of resident calls, uploaded once: key -> DeviceTensor
        self._keep: list = []           # host arrays whose address is a _consts key (kept alive so the address stays theirs)
        self._pinned: list = []         # page-locked host allocations handed out by pinned_empty

    # -- memory --
    def upload(self, a: np.ndarray, dtype=np.float32) -> DevBuf:
        a = np.ascontiguousarray(a, dtype=dtype)
        b = DevBuf(self, a.nbytes)
        if a.nbytes:
            call("lele_b200_h2d", self.h, vp(b.ptr), a.ctypes.data_as(vp), sz(a.nbytes))
            self.sync()  # `a` may be a temporary
        return b

    def empty(self, n_elems: int, itemsize: int = 4) -> DevBuf:
        return DevBuf(self, int(n_elems) * itemsize)

    def download(self, b: DevBuf, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            call("lele_b200_d2h", self.h, out.ctypes.data_as(vp), vp(b.ptr), sz(out.nbytes))
        self.sync()
        return out

    # -- resident values --
    def to_device(self, a, dtype=np.float32) -> DeviceTensor:
        """A host array as a DeviceTensor that owns its allocation (a graph input, `TensorView::from_slice` upstream)."""
        if isinstance(a, DeviceTensor):
            return a
        a = np.asarray(a, dtype=dtype, order="C")
        b = self.upload(a, dtype)
        return DeviceTensor(b.ptr, a.shape, self, b, dtype)

    def persist(self, a: np.ndarray) -> DeviceTensor:
        """Uploads a long-lived host array (a view into weights.bin) once; later resident calls that receive this array -- or a
        reshape of it -- as an operand use the device copy (keyed by the array's data address, (blob_base, offset) upstream)."""
        a = np.asarray(a)
        if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a, dtype=np.float32)
        key = ("p", a.__array_interface__["data"][0], a.nbytes)
        t = self._consts.get(key)
        if t is None:
            t = self._consts[key] = self.to_device(a)
            self._keep.append(a)
        return t

    def _operand(self, a, dtype=np.float32) -> DeviceTensor:
        """Device view of one operand of a resident call.  Host operands are uploaded once and cached: by data address when the
        array was registered with `persist` (weights), else by content (shape constants, literals: small by construction)."""
        if isinstance(a, DeviceTensor):
            return a
        a = np.ascontiguousarray(a, dtype=dtype)
        t = self._consts.get(("p", a.__array_interface__["data"][0], a.nbytes))
        if t is not None:
            return t
        import hashlib
        key = ("c", a.dtype.str, a.nbytes, hashlib.blake2b(a.tobytes(), digest_size=16).digest())
        t = self._consts.get(key)
        if t is None:
            t = self._consts[key] = self.to_device(a, dtype)
        return t

    def out_slots(self, slots):
        """Names the workspace buffers the next resident call writes its outputs to, in output order (the `&mut ws.buf_N` arguments
        of a generated statement): a list of (Workspace, name).  Consumed by that call; unused entries are dropped by the caller."""
        self._slots = list(slots)

    def _out(self, shape, dtype=np.float32) -> DeviceTensor:
        if self._slots:
            ws, name = self._slots.pop(0)
            return ws.tensor(name, shape, dtype)
        n = _b.max(_prod(shape), 1) * np.dtype(dtype).itemsize
        return DeviceTensor(0, shape, self, None, dtype)._own(DevBuf(self, n))

    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """A numpy array over page-locked host memory (lele_b200_malloc_host), freed with the context: the destination of d2h / source
        of h2d copies that should run asynchronously at PCIe rate."""
        n = _b.max(_prod(shape) * np.dtype(dtype).itemsize, 16)
        p = vp()
        call("lele_b200_malloc_host", self.h, sz(n), C.byref(p))
        self._pinned.append(p.value)
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=_prod(shape)).reshape(shape)

    def sync(self):
        call("lele_b200_sync", self.h)

    def launch_count(self) -> int:
        return int(lib.lele_b200_launch_count(self.h))

    # -- CUDA-graph capture of a sequence of C-ABI calls (a replayed model after its first, arena-sizing forward) --
    def capture_begin(self):
        call("lele_b200_capture_begin", self.h)

    def capture_end(self, lane_launches: int = 0) -> "Graph":
        g = vp()
        call("lele_b200_capture_end", self.h, C.c_ulonglong(int(lane_launches)), C.byref(g))
        return Graph(self, g)

    def fork(self, lane: "Context"):
        """`lane`'s stream waits for everything enqueued on this context so far (inside a capture it joins the capture)."""
        call("lele_b200_stream_fork", self.h, lane.h)

    def join(self, lane: "Context"):
        """this context's stream waits for everything enqueued on `lane` so far."""
        call("lele_b200_stream_join", self.h, lane.h)

    def close(self):
        if self.h:
            self._consts.clear(); self._keep.clear()
            for p in self._pinned:
                lib.lele_b200_free_host(self.h, vp(p))
            self._pinned = []
            lib.lele_b200_ctx_destroy(self.h)
            self.h = None


class Graph:
    """An instantiated CUDA graph of captured C-ABI calls; `launch()` replays it on the capturing context's stream."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    def launch(self):
        call("lele_b200_graph_launch", self.ctx.h, self.h)

    def close(self):
        if self.h:
            call("lele_b200_graph_destroy", self.ctx.h, self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: Context | None = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(0)
    return _default


def _f(a):
    return a if isinstance(a, DeviceTensor) else np.ascontiguousarray(a, dtype=np.float32)


def _ll(v):
    return (i64 * _b.max(len(v), 1))(*[int(x) for x in v])


def _ints(v):
    return (i32 * _b.max(len(v), 1))(*[int(x) for x in v])


def _prod(s):
    return int(np.prod(s, dtype=np.int64)) if len(s) else 1


def _resident(ctx, *inputs):
    return bool(ctx._slots) or any(isinstance(a, DeviceTensor) for a in inputs)


def _run(out_shape, fn, *inputs, ctx=None, out_dtype=np.float32):
    """fn(ctx, out_ptr, *in_ptrs) with the operands marshalled (None passes NULL).
    Host form (all operands numpy): upload -> launch -> download.  Resident form (any operand a DeviceTensor, or an output
    slot named): no copies -- device pointers in, a DeviceTensor out."""
    ctx = ctx or default_context()
    if _resident(ctx, *inputs):
        ops = [None if a is None else ctx._operand(a) for a in inputs]
        out = ctx._out(out_shape, out_dtype)
        fn(ctx, vp(out.ptr), *[vp(None if o is None else o.ptr) for o in ops])
        return out
    bufs = [None if a is None else ctx.upload(a) for a in inputs]
    out = ctx.empty(_b.max(_prod(out_shape), 1), np.dtype(out_dtype).itemsize)
    fn(ctx, vp(out.ptr), *[vp(None if b is None else b.ptr) for b in bufs])
    res = ctx.download(out, out_shape, out_dtype)
    for b in bufs:
        if b is not None:
            b.free()
    out.free()
    return res


def _run_multi(out_shapes, fn, *inputs, ctx=None):
    """Several outputs (lstm, gru, topk, dynamic_quantize_linear): fn(ctx, [out_ptrs], *in_ptrs) -> tuple, host or resident as _run."""
    ctx = ctx or default_context()
    if _resident(ctx, *inputs):
        ops = [None if a is None else ctx._operand(a) for a in inputs]
        outs = [ctx._out(s) for s in out_shapes]
        fn(ctx, [vp(o.ptr) for o in outs], *[vp(None if o is None else o.ptr) for o in ops])
        return tuple(outs)
    bufs = [None if a is None else ctx.upload(a) for a in inputs]
    outs = [ctx.empty(_b.max(_prod(s), 1)) for s in out_shapes]
    fn(ctx, [vp(o.ptr) for o in outs], *[vp(None if b is None else b.ptr) for b in bufs])
    res = tuple(ctx.download(o, s) for o, s in zip(outs, out_shapes))
    for b in bufs + outs:
        if b is not None:
            b.free()
    return res


# ---------------------------------------------------------------- gemm.rs
def _lead(a, b):
    ba = _prod(a.shape[:-2]); bb = _prod(b.shape[:-2])
    return ba, bb, (a.shape[:-2] if ba >= bb else b.shape[:-2])


def matmul(a, b, ctx=None):
    """gemm.rs:112"""
    a, b = _f(a), _f(b)
    if a.ndim < 2 or b.ndim < 2:
        raise LeleB200Error("MatMul: both operands need rank >= 2 (gemm.rs:122-123)")
    if a.shape[-1] != b.shape[-2]:
        raise LeleB200Error(f"MatMul K dim mismatch: {a.shape[-1]} vs {b.shape[-2]} (gemm.rs:129)")
    m, k = a.shape[-2:]; n = b.shape[-1]
    ba, bb, lead = _lead(a, b)
    if not (bb == 1 or bb == ba):
        raise LeleB200Error("MatMul broadcast not fully supported yet (gemm.rs:134)")
    return _run(tuple(lead) + (m, n), lambda c, o, pa, pb: call("lele_b200_matmul", c.h, pa, pb, i32(ba), i32(bb), i32(m), i32(k), i32(n), o), a, b, ctx=ctx)


def matmul_fused_add(a, b, bias, ctx=None):
    """gemm.rs:223"""
    a, b, bias = _f(a), _f(b), _f(bias).reshape(-1)
    m, k = a.shape[-2:]; n = b.shape[-1]
    ba, bb, lead = _lead(a, b)
    return _run(tuple(lead) + (m, n), lambda c, o, pa, pb, pc: call("lele_b200_matmul_fused_add", c.h, pa, pb, pc, i32(bias.size), i32(ba), i32(bb), i32(m), i32(k), i32(n), o), a, b, bias, ctx=ctx)


def gemm(a, b, c=None, alpha=1.0, beta=1.0, trans_a=False, trans_b=False, ctx=None):
    """gemm.rs:433 (output always [M, N])"""
    a, b = _f(a), _f(b)
    m = a.shape[-1] if trans_a else a.shape[-2]; k = a.shape[-2] if trans_a else a.shape[-1]
    n = b.shape[-2] if trans_b else b.shape[-1]; k2 = b.shape[-1] if trans_b else b.shape[-2]
    if k != k2:
        raise LeleB200Error("Gemm K dim mismatch (gemm.rs:465)")
    cc = None if c is None else _f(c).reshape(-1)
    return _run((m, n), lambda cx, o, pa, pb, pc: call("lele_b200_gemm", cx.h, pa, pb, pc, i32(0 if cc is None else cc.size), f32(alpha), f32(beta), i32(int(trans_a)), i32(int(trans_b)), i32(m), i32(k), i32(n), o), a, b, cc, ctx=ctx)


# ---------------------------------------------------------------- norm.rs
def layer_norm(x, scale, bias, axis=-1, epsilon=1e-5, ctx=None):
    """norm.rs:226"""
    x = _f(x); ax = axis % x.ndim
    n = _prod(x.shape[ax:]); outer = x.size // _b.max(n, 1)
    return _run(x.shape, lambda c, o, px, pg, pb: call("lele_b200_layer_norm", c.h, px, pg, pb, i64(outer), i32(n), f32(epsilon), o), x, _f(scale).reshape(-1), _f(bias).reshape(-1), ctx=ctx)


def softmax(x, axis=-1, ctx=None):
    """norm.rs:8 -- last axis only; other axes are `unimplemented!` upstream (norm.rs:218)."""
    x = _f(x)
    if axis % x.ndim != x.ndim - 1:
        raise LeleB200Error("softmax: only the last axis is implemented (norm.rs:218)")
    n = x.shape[-1]
    return _run(x.shape, lambda c, o, px: call("lele_b200_softmax", c.h, px, i64(x.size // n), i32(n), o), x, ctx=ctx)


def batch_norm(x, scale, bias, mean, var, epsilon=1e-5, ctx=None):
    x = _f(x); nb, ch = x.shape[:2]; inner = x.size // (nb * ch)
    return _run(x.shape, lambda c, o, px, ps, pb, pm, pv: call("lele_b200_batch_norm", c.h, px, ps, pb, pm, pv, i32(nb), i32(ch), i64(inner), f32(epsilon), o), x, _f(scale), _f(bias), _f(mean), _f(var), ctx=ctx)


def rms_norm(x, w, epsilon=1e-5, ctx=None):
    x = _f(x); n = x.shape[-1]
    return _run(x.shape, lambda c, o, px, pw: call("lele_b200_rms_norm", c.h, px, pw, i64(x.size // n), i32(n), f32(epsilon), o), x, _f(w), ctx=ctx)


# ---------------------------------------------------------------- quantization.rs
def dynamic_quantize_linear(x, ctx=None):
    """quantization.rs:1628 -> (q as f32, scale, zero_point); scale / zero_point are [1] tensors (host floats in the host form)"""
    x = _f(x)
    q, s_, z = _run_multi([x.shape, (1,), (1,)], lambda c, o, px: call("lele_b200_dynamic_quantize_linear", c.h, px, i32(1), i64(x.size), o[0], o[1], o[2]), x, ctx=ctx)
    return (q, s_, z) if isinstance(q, DeviceTensor) else (q, s_[0], z[0])


def mat_mul_integer(a, b, a_zero_point=0.0, b_zero_point=0.0, scale=None, bias=None, relu=False, ctx=None):
    """mat_mul_integer / _with_scale_bias / _with_scale_bias_relu (quantization.rs:8-72).  a [.., M, K], b [.., K, N]: equal batch, or
    the side whose batch is 1 is broadcast (quantization.rs:1157-1173)."""
    a, b = _f(a), _f(b)
    if a.ndim < 2 or b.ndim < 2:
        raise LeleB200Error("MatMulInteger: both operands need rank >= 2")
    m, k = a.shape[-2:]; n = b.shape[-1]; ba = _prod(a.shape[:-2]); bb = _prod(b.shape[:-2])
    if b.shape[-2] != k:
        raise LeleB200Error(f"MatMulInteger: K mismatch {k} vs {b.shape[-2]} (the reference panics on the slice bounds, quantization.rs:1176)")
    if not (ba == bb or ba == 1 or bb == 1):
        raise LeleB200Error(f"MatMulInteger: batch {ba} vs {bb} (equal, or one side 1; quantization.rs:1157-1173)")
    lead = a.shape[:-2] if ba >= bb else b.shape[:-2]
    sc = None if scale is None else _f(scale).reshape(-1); bi = None if bias is None else _f(bias).reshape(-1)
    return _run(tuple(lead) + (m, n), lambda c, o, pa, pb, ps, pbi: call("lele_b200_mat_mul_integer_batched", c.h, pa, pb, i32(ba), i32(bb), i32(m), i32(k), i32(n), f32(a_zero_point), f32(b_zero_point), ps, i32(0 if sc is None else sc.size), pbi, i32(int(relu)), o), a, b, sc, bi, ctx=ctx)


class PreparedWeights:
    """prepare_weights (quantization.rs:221): K-major u8 weight + column sums resident in HBM."""

    def __init__(self, weight_u8, weight_scale, weight_zero, bias=None, ctx=None):
        self.ctx = ctx or default_context()
        w = np.ascontiguousarray(weight_u8, dtype=np.uint8)
        self.k, self.n = w.shape
        ws = _f(weight_scale).reshape(-1)
        bw = self.ctx.upload(w, np.uint8); bs = self.ctx.upload(ws)
        bb = None if bias is None else self.ctx.upload(_f(bias).reshape(-1))
        h = vp()
        call("lele_b200_prepare_weights", self.ctx.h, vp(bw.ptr), i32(self.k), i32(self.n), vp(bs.ptr), i32(ws.size), i32(int(weight_zero)), vp(None if bb is None else bb.ptr), C.byref(h))
        self.ctx.sync()
        self.h = h
        for b in (bw, bs, bb):
            if b is not None:
                b.free()

    def __del__(self):
        try:
            if self.h:
                lib.lele_b200_qweights_destroy(self.ctx.h, self.h); self.h = None
        except Exception:
            pass


def fused_quantized_linear(input, weight_int8, weight_scale, weight_zero, bias, apply_relu=False, ctx=None):
    """quantization.rs:77.  `weight_int8` is the u8 weight [K,N] (f32-coded or u8) or a PreparedWeights."""
    ctx = ctx or default_context()
    x = _f(input)
    pw = weight_int8 if isinstance(weight_int8, PreparedWeights) else PreparedWeights(
        np.clip(np.asarray(weight_int8), 0, 255).astype(np.uint8), weight_scale, int(np.asarray(weight_zero).reshape(-1)[0]) if np.size(weight_zero) else 0,
        None if bias is None or np.size(bias) == 0 else bias, ctx)
    m, k = x.shape[-2:]
    if k != pw.k:
        raise LeleB200Error(f"fused_quantized_linear: K mismatch {k} vs {pw.k}")
    batch = _prod(x.shape[:-2])
    return _run(x.shape[:-1] + (pw.n,), lambda c, o, px: call("lele_b200_fused_quantized_linear", c.h, px, i32(batch), i32(m), pw.h, i32(int(apply_relu)), o), x, ctx=ctx)


# ---------------------------------------------------------------- conv
def _pads4(p):
    p = list(p)
    if len(p) == 0: return [0, 0, 0, 0]
    if len(p) == 2: return [p[0], p[1], p[0], p[1]]
    return p


def conv1d_fused(input, weights, bias, dilations, group, pads, strides, relu, ctx=None):
    """conv1d.rs:853"""
    x, w = _f(input), _f(weights)
    if x.ndim == 2: x = x.reshape((x.shape[0], 1, x.shape[1]))
    if x.ndim != 3: raise LeleB200Error(f"Conv1d: Unsupported input rank {x.ndim} (conv1d.rs:872)")
    nb, ic, l = x.shape; oc, _, k = w.shape
    dil = dilations[0] if len(dilations) else 1; st = strides[0] if len(strides) else 1
    pl = pads[0] if len(pads) else 0; pr = pads[1] if len(pads) > 1 else 0
    ol = (l + pl + pr - dil * (k - 1) - 1) // st + 1
    bi = None if bias is None else _f(bias)
    return _run((nb, oc, ol), lambda c, o, px, pw, pb: call("lele_b200_conv1d", c.h, px, pw, pb, i32(nb), i32(ic), i32(l), i32(oc), i32(k), i32(group), i32(pl), i32(pr), i32(st), i32(dil), i32(int(relu)), o), x, w, bi, ctx=ctx)


def conv1d(input, weights, bias=None, dilations=(1,), group=1, pads=(0, 0), strides=(1,), ctx=None):
    """conv1d.rs:837"""
    return conv1d_fused(input, weights, bias, dilations, group, pads, strides, False, ctx)


def _conv2d(x, w, bias, dilations, group, pads, strides, act, ctx):
    x, w = _f(x), _f(w); nb, ic, h, wd = x.shape; oc, _, kh, kw = w.shape
    p = _pads4(pads); s = list(strides) or [1, 1]; d = list(dilations) or [1, 1]
    eh = h + p[0] + p[2] - d[0] * (kh - 1) - 1; ew = wd + p[1] + p[3] - d[1] * (kw - 1) - 1
    oh = eh // s[0] + 1; ow = ew // s[1] + 1
    if eh < 0 or ew < 0 or oh <= 0 or ow <= 0:   # upstream the unsigned subtraction overflows or the assertion fires (conv2d.rs:274-291)
        raise LeleB200Error(f"conv2d: output dimensions must be positive, got out_h={oh} out_w={ow} (conv2d.rs:288)")
    bi = None if bias is None else _f(bias)
    return _run((nb, oc, oh, ow), lambda c, o, px, pw, pb: call("lele_b200_conv2d", c.h, px, pw, pb, i32(nb), i32(ic), i32(h), i32(wd), i32(oc), i32(kh), i32(kw), i32(group), _ints(p), _ints(s), _ints(d), i32(act), o), x, w, bi, ctx=ctx)


def conv2d(input, weights, bias=None, dilations=(1, 1), group=1, pads=(0, 0, 0, 0), strides=(1, 1), act=0, ctx=None):
    """conv2d.rs:107 (act=0), conv2d_fused :155 (act=1 ReLU), conv2d_silu :124 (act=2)"""
    return _conv2d(input, weights, bias, dilations, group, pads, strides, act, ctx)


def conv2d_fused(input, weights, bias, dilations, group, pads, strides, relu, ctx=None):
    return _conv2d(input, weights, bias, dilations, group, pads, strides, 1 if relu else 0, ctx)


def conv2d_silu(input, weights, bias, dilations, group, pads, strides, ctx=None):
    return _conv2d(input, weights, bias, dilations, group, pads, strides, 2, ctx)


def conv_integer(input, weights, x_zero_point=0.0, w_zero_point=0.0, dilations=(1, 1), group=1, pads=(0, 0, 0, 0), strides=(1, 1), ctx=None):
    """conv2d.rs:2216 (emitted by ops/nn.rs:328): (x - x_zp) (*) (w - w_zp); padded positions hold raw zeros (conv2d.rs:2025).
    Zero points: the scalar the reference reads from the zero-point tensor (`zp.data[0]`, 0 when absent or empty)."""
    x, w = _f(input), _f(weights); nb, ic, h, wd = x.shape; oc, _, kh, kw = w.shape
    p = _pads4(pads); s = list(strides) or [1, 1]; d = list(dilations) or [1, 1]
    eh = h + p[0] + p[2] - d[0] * (kh - 1) - 1; ew = wd + p[1] + p[3] - d[1] * (kw - 1) - 1
    oh = eh // s[0] + 1; ow = ew // s[1] + 1
    if eh < 0 or ew < 0 or oh <= 0 or ow <= 0:
        raise LeleB200Error(f"conv_integer: output dimensions must be positive, got out_h={oh} out_w={ow} (conv2d.rs:288)")
    return _run((nb, oc, oh, ow), lambda c, o, px, pw: call("lele_b200_conv_integer", c.h, px, pw, f32(x_zero_point), f32(w_zero_point), i32(nb), i32(ic), i32(h), i32(wd), i32(oc), i32(kh), i32(kw), i32(group), _ints(p), _ints(s), _ints(d), o), x, w, ctx=ctx)


def conv_transpose(input, weights, bias=None, dilations=(1, 1), pads=(0, 0, 0, 0), strides=(1, 1), group=1, ctx=None):
    """conv2d.rs:2952 (rank-4, group 1)"""
    x, w = _f(input), _f(weights)
    if x.ndim != 4 or w.ndim != 4: raise LeleB200Error("ConvTranspose: expected rank-4 input/weight (conv2d.rs:2989)")
    if group != 1: raise LeleB200Error("ConvTranspose: group > 1 not supported yet (conv2d.rs:3042)")
    nb, ic, h, wd = x.shape; _, oc, kh, kw = w.shape
    p = _pads4(pads); s = (list(strides) + [1, 1])[:2] if len(strides) < 2 else list(strides); d = (list(dilations) + [1, 1])[:2]
    oh = (h - 1) * s[0] - (p[0] + p[2]) + d[0] * (kh - 1) + 1; ow = (wd - 1) * s[1] - (p[1] + p[3]) + d[1] * (kw - 1) + 1
    bi = None if bias is None else _f(bias)
    return _run((nb, oc, oh, ow), lambda c, o, px, pw, pb: call("lele_b200_conv_transpose", c.h, px, pw, pb, i32(nb), i32(ic), i32(h), i32(wd), i32(oc), i32(kh), i32(kw), _ints(p), _ints(s), _ints(d), o), x, w, bi, ctx=ctx)


def max_pool2d(input, kernel_shape, pads=(0, 0, 0, 0), strides=(1, 1), dilations=(1, 1), ceil_mode=False, ctx=None):
    """conv2d.rs:1051"""
    x = _f(input); nb, c, h, w = x.shape; p = _pads4(pads); s = list(strides); d = list(dilations); kh, kw = kernel_shape
    nh = h + p[0] + p[2] - d[0] * (kh - 1) - 1; nw = w + p[1] + p[3] - d[1] * (kw - 1) - 1
    oh = (-(-nh // s[0]) if ceil_mode else nh // s[0]) + 1; ow = (-(-nw // s[1]) if ceil_mode else nw // s[1]) + 1
    return _run((nb, c, oh, ow), lambda cx, o, px: call("lele_b200_max_pool2d", cx.h, px, i32(nb), i32(c), i32(h), i32(w), i32(kh), i32(kw), _ints(p), _ints(s), _ints(d), i32(int(ceil_mode)), o), x, ctx=ctx)


def resize_nearest(input, scales=None, sizes=None, coordinate_transform_mode="asymmetric", ctx=None):
    """conv2d.rs:1261"""
    x = _f(input); nb, c, h, w = x.shape
    if sizes is not None:
        if len(sizes) < 4: raise LeleB200Error("Resize: sizes must have at least 4 elements (conv2d.rs:1301)")
        if not (sizes[2] > 0 and sizes[3] > 0): raise LeleB200Error("Resize: sizes H and W must be positive (conv2d.rs:1305)")
        oh, ow = int(sizes[2]), int(sizes[3])
    elif scales is not None:
        sh = np.float32(scales[2]) if len(scales) >= 3 else np.float32(1); sw = np.float32(scales[3]) if len(scales) >= 4 else np.float32(1)
        if not (sh > 0 and sw > 0): raise LeleB200Error("Resize: scales must be positive (conv2d.rs:1312)")
        oh, ow = int(np.float64(h) * np.float64(sh)), int(np.float64(w) * np.float64(sw))
    else:
        raise LeleB200Error("Resize: either scales or sizes must be provided (conv2d.rs:1318)")
    if not (oh > 0 and ow > 0): raise LeleB200Error(f"Resize: output dimensions must be positive, got out_h={oh} out_w={ow} (conv2d.rs:1323)")
    mode = 0 if coordinate_transform_mode == "asymmetric" else 1
    return _run((nb, c, oh, ow), lambda cx, o, px: call("lele_b200_resize_nearest", cx.h, px, i32(nb), i32(c), i32(h), i32(w), i32(oh), i32(ow), i32(mode), o), x, ctx=ctx)


# ---------------------------------------------------------------- rnn.rs
def lstm(input, w, r, bias=None, sequence_lens=None, initial_h=None, initial_c=None, ctx=None):
    """rnn.rs:67 -> (Y [S,1,1,H], H [1,1,H], C [1,1,H])"""
    x, w, r = _f(input), _f(w), _f(r)
    seq, bs, isz = x.shape
    if w.shape[0] != 1: raise LeleB200Error("LSTM: Only num_directions=1 supported (rnn.rs:85)")
    if bs != 1: raise LeleB200Error("LSTM: Only batch_size=1 supported (rnn.rs:88)")
    hid = w.shape[1] // 4
    opt = [None if a is None else _f(a) for a in (bias, initial_h, initial_c)]
    return _run_multi([(seq, 1, 1, hid), (1, 1, hid), (1, 1, hid)],
                      lambda c, o, px, pw, pr, pb, ph, pc: call("lele_b200_lstm", c.h, px, pw, pr, pb, ph, pc, i32(1), i32(seq), i32(isz), i32(hid), o[0], o[1], o[2]),
                      x, w, r, *opt, ctx=ctx)


def gru(input, w, r, bias=None, initial_h=None, linear_before_reset=False, ctx=None):
    """rnn.rs:246 -> (Y [S,1,1,H], H [1,1,H])"""
    x, w, r = _f(input), _f(w), _f(r)
    seq, bs, isz = x.shape
    if w.shape[0] != 1: raise LeleB200Error("GRU: Only num_directions=1 supported (rnn.rs:262)")
    if bs != 1: raise LeleB200Error("GRU: Only batch_size=1 supported (rnn.rs:265)")
    hid = w.shape[1] // 3
    opt = [None if a is None else _f(a) for a in (bias, initial_h)]
    return _run_multi([(seq, 1, 1, hid), (1, 1, hid)],
                      lambda c, o, px, pw, pr, pb, ph: call("lele_b200_gru", c.h, px, pw, pr, pb, ph, i32(1), i32(seq), i32(isz), i32(hid), o[0], o[1]),
                      x, w, r, *opt, ctx=ctx)


def _rnn_streams(gates, input, w, r, bias, initial_h, initial_c, ctx):
    """n_seq independent batch-1 sequences that share one set of weights, in ONE launch (the `n_seq` argument of
    lele_b200_lstm / lele_b200_gru; one CTA per stream).  Every stream gets exactly what the single-sequence call returns for it:
    the hoisted W.x GEMM and the recurrence sum in the same order whatever n_seq is."""
    x, w, r = _f(input), _f(w), _f(r)
    if x.ndim != 3: raise LeleB200Error("rnn streams: input must be [n_seq, seq, input_size]")
    n_seq, seq, isz = x.shape
    if w.shape[0] != 1: raise LeleB200Error("RNN: Only num_directions=1 supported (rnn.rs:85)")
    hid = w.shape[1] // gates
    if w.shape[2] != isz or tuple(r.shape[1:]) != (gates * hid, hid): raise LeleB200Error("rnn streams: W / R shape mismatch")
    states = [None if a is None else _f(a).reshape(n_seq, hid) for a in ((initial_h, initial_c) if gates == 4 else (initial_h,))]
    dims = (i32(n_seq), i32(seq), i32(isz), i32(hid))
    shapes = [(n_seq, seq, hid)] + [(n_seq, hid)] * len(states)
    if gates == 4:
        fn = lambda c, o, px, pw, pr, pb, ph, pc: call("lele_b200_lstm", c.h, px, pw, pr, pb, ph, pc, *dims, o[0], o[1], o[2])
    else:
        fn = lambda c, o, px, pw, pr, pb, ph: call("lele_b200_gru", c.h, px, pw, pr, pb, ph, *dims, o[0], o[1])
    return _run_multi(shapes, fn, x, w, r, None if bias is None else _f(bias), *states, ctx=ctx)


def lstm_streams(input, w, r, bias=None, initial_h=None, initial_c=None, ctx=None):
    """Many streams per launch: input [n_seq, S, I], states [n_seq, H] -> (Y [n_seq, S, H], H [n_seq, H], C [n_seq, H]);
    stream s equals `lstm(input[s][:, None, :], ..., initial_h[s], initial_c[s])` (rnn.rs:67 is batch-1 only)."""
    return _rnn_streams(4, input, w, r, bias, initial_h, initial_c, ctx)


def gru_streams(input, w, r, bias=None, initial_h=None, ctx=None):
    """input [n_seq, S, I], initial_h [n_seq, H] -> (Y [n_seq, S, H], H [n_seq, H]); stream s equals `gru` on that stream alone (rnn.rs:246)."""
    return _rnn_streams(3, input, w, r, bias, initial_h, None, ctx)


# ---------------------------------------------------------------- math.rs
_BIN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "max": 4, "pow": 5, "mod_f32": 6, "prelu": 7, "equal": 8, "less": 9}
_UN = {"relu": 0, "sigmoid": 1, "tanh_kernel": 2, "silu": 3, "erf": 4, "gelu": 5, "exp": 6, "softplus": 7, "log": 8, "sqrt": 9,
       "neg": 10, "reciprocal": 11, "sin": 12, "cos": 13, "not": 14, "fast_gelu": 15}


def _binary(op, a, b, ctx=None):
    a, b = _f(a), _f(b)
    try:
        shp = np.broadcast_shapes(a.shape, b.shape)
    except ValueError as e:
        raise LeleB200Error(f"broadcast: {e}")
    return _run(shp, lambda c, o, pa, pb: call("lele_b200_binary", c.h, i32(_BIN[op]), pa, _ll(a.shape), i32(a.ndim), pb, _ll(b.shape), i32(b.ndim), o), a, b, ctx=ctx)


def _unary(op, x, ctx=None):
    x = _f(x)
    return _run(x.shape, lambda c, o, px: call("lele_b200_unary", c.h, i32(_UN[op]), px, i64(x.size), o), x, ctx=ctx)


def add(a, b, ctx=None): return _binary("add", a, b, ctx)
def sub(a, b, ctx=None): return _binary("sub", a, b, ctx)
def mul(a, b, ctx=None): return _binary("mul", a, b, ctx)
def div(a, b, ctx=None): return _binary("div", a, b, ctx)
def max(a, b, ctx=None): return _binary("max", a, b, ctx)  # noqa: A001 (lele::kernels::max)
def pow(a, b, ctx=None): return _binary("pow", a, b, ctx)  # noqa: A001
def mod_f32(a, b, ctx=None): return _binary("mod_f32", a, b, ctx)
def prelu(a, slope, ctx=None):
    """math.rs:2012: a one-element slope keeps the input's shape whatever the slope's rank; otherwise NumPy broadcasting"""
    s = _f(slope)
    return _binary("prelu", a, s.reshape(1) if s.size == 1 else s, ctx)
def equal(a, b, ctx=None): return _binary("equal", a, b, ctx)
def less(a, b, ctx=None): return _binary("less", a, b, ctx)
def relu(x, ctx=None): return _unary("relu", x, ctx)
def sigmoid(x, ctx=None): return _unary("sigmoid", x, ctx)
def tanh_kernel(x, ctx=None): return _unary("tanh_kernel", x, ctx)
def silu(x, ctx=None): return _unary("silu", x, ctx)
def erf(x, ctx=None): return _unary("erf", x, ctx)
def gelu(x, ctx=None): return _unary("gelu", x, ctx)
def fast_gelu(x, ctx=None): return _unary("fast_gelu", x, ctx)
def exp(x, ctx=None): return _unary("exp", x, ctx)
def softplus(x, ctx=None): return _unary("softplus", x, ctx)
def log(x, ctx=None): return _unary("log", x, ctx)
def sqrt(x, ctx=None): return _unary("sqrt", x, ctx)
def neg(x, ctx=None): return _unary("neg", x, ctx)
def reciprocal(x, ctx=None): return _unary("reciprocal", x, ctx)
def sin(x, ctx=None): return _unary("sin", x, ctx)
def cos(x, ctx=None): return _unary("cos", x, ctx)
def not_(x, ctx=None): return _unary("not", x, ctx)  # lele::kernels::not (math.rs:1508); `not` is a Python keyword


def clip(x, lo, hi, ctx=None):
    x = _f(x)
    return _run(x.shape, lambda c, o, px: call("lele_b200_clip", c.h, px, i64(x.size), f32(lo), f32(hi), o), x, ctx=ctx)


def _reduce(kind, x, axes, keepdims, ctx=None):
    """math.rs:1527-1921.  Arbitrary axes: transposed to a single middle axis first."""
    x = _f(x)
    axes = sorted({a % x.ndim for a in axes})   # deduplicated (math.rs:1629); an empty list reduces nothing upstream (reduce_mask stays false, :1631)
    keep = [i for i in range(x.ndim) if i not in axes]
    trailing = axes == list(range(x.ndim - len(axes), x.ndim))
    xt = x if trailing else (transpose(x, keep + axes, ctx=ctx) if isinstance(x, DeviceTensor) else np.ascontiguousarray(np.transpose(x, keep + axes)))
    outer = _prod([x.shape[i] for i in keep]); alen = _prod([x.shape[i] for i in axes])
    out = _run((outer,), lambda c, o, px: call("lele_b200_reduce", c.h, i32(kind), px, i64(outer), i32(alen), i64(1), o), xt, ctx=ctx)
    shp = [1 if i in axes else x.shape[i] for i in range(x.ndim)] if keepdims else [x.shape[i] for i in keep]
    return out.reshape(shp)


def reduce_sum(x, axes=(), keepdims=True, ctx=None): return _reduce(0, x, axes, keepdims, ctx)
def reduce_mean(x, axes=(), keepdims=True, ctx=None): return _reduce(1, x, axes, keepdims, ctx)
def reduce_max(x, axes=(), keepdims=True, ctx=None): return _reduce(2, x, axes, keepdims, ctx)
def reduce_l2(x, axes=(), keepdims=True, ctx=None): return _reduce(3, x, axes, keepdims, ctx)


def where_op(condition, x, y, ctx=None):
    """manipulation.rs:1215"""
    c_, x, y = _f(condition), _f(x), _f(y)
    shp = np.broadcast_shapes(c_.shape, x.shape, y.shape)
    return _run(shp, lambda c, o, pc, px, py: call("lele_b200_where", c.h, pc, _ll(c_.shape), i32(c_.ndim), px, _ll(x.shape), i32(x.ndim), py, _ll(y.shape), i32(y.ndim), o), c_, x, y, ctx=ctx)


def stft(input, n_fft, hop_length, win_length, window=None, power=False, ctx=None):
    """math.rs:2304 (power=False: [.., frames, n_fft/2+1, 2]) / stft_power_spectrum :2372"""
    x = _f(input); sig = x.reshape(-1); L = sig.size
    nfr = n_fft // 2 + 1
    if L == 0:
        return np.zeros((0, 0, nfr) if power else (0, 0, nfr, 2), np.float32)
    frames = 1 if L < win_length else (L - win_length) // hop_length + 1
    shp = (frames, nfr) if power else (frames, nfr, 2)
    w = None if window is None else _f(window)
    fo = i32(0)
    out = _run(shp, lambda c, o, px, pw: call("lele_b200_stft", c.h, px, i32(L), i32(n_fft), i32(hop_length), i32(win_length), pw, i32(int(power)), o, C.byref(fo)), sig, w, ctx=ctx)
    if x.ndim > 1:  # batch dim lives only in the shape (math.rs:2364)
        out = out.reshape((_prod(x.shape[:-1]),) + shp) if _prod(x.shape[:-1]) == 1 else out
    return out


def stft_power_spectrum(input, n_fft, hop_length, win_length, window=None, ctx=None):
    return stft(input, n_fft, hop_length, win_length, window, True, ctx)


# ---------------------------------------------------------------- manipulation.rs / shape.rs
def _strided(x, out_shape, in_strides, offset=0, ctx=None):
    x = _f(x)
    return _run(tuple(out_shape), lambda c, o, px: call("lele_b200_strided_copy", c.h, px, i64(offset), _ll(out_shape), _ll(in_strides), i32(len(out_shape)), o), x, ctx=ctx)


def _estrides(shape):
    s, acc = [], 1
    for d in reversed(shape):
        s.append(acc); acc *= d
    return list(reversed(s))


def transpose(x, perm=(), ctx=None):
    """manipulation.rs:644 (empty perm = reverse)"""
    x = _f(x); perm = list(perm) if len(perm) else list(reversed(range(x.ndim)))
    st = _estrides(x.shape)
    return _strided(x, [x.shape[p] for p in perm], [st[p] for p in perm], 0, ctx)


def slice(x, starts, ends, axes=(), steps=(), ctx=None):  # noqa: A001
    """manipulation.rs:209 (ONNX clamp rules incl. i64 sentinels, negative steps)"""
    x = _f(x); st = _estrides(x.shape)
    shape, strides, off = list(x.shape), list(st), 0
    I64MAX, I64MIN = 2**63 - 1, -2**63
    for i in range(len(starts)):
        ax = i if len(axes) == 0 else (axes[i] + x.ndim if axes[i] < 0 else axes[i])
        dim = x.shape[ax]; step = steps[i] if i < len(steps) else 1
        s0, e0 = int(starts[i]), int(ends[i])
        emax, emin = e0 > I64MAX // 2, e0 < I64MIN // 2
        s = dim if s0 > dim else (-dim if s0 < -dim else s0)
        e = dim if emax else (-dim if emin else (dim if e0 > dim else (-dim if e0 < -dim else e0)))
        ns = s + dim if s < 0 else s
        ne = (dim if step > 0 else -1) if emax else ((0 if step > 0 else -1) if emin else (e + dim if e < 0 else e))
        if step > 0:
            s_, e_ = min(_b.max(ns, 0), dim), min(_b.max(ne, 0), dim)
            cnt = 0 if s_ >= e_ else (e_ - s_ + step - 1) // step
        else:
            s_, e_ = min(_b.max(ns, 0), dim - 1), min(_b.max(ne, -1), dim - 1)
            cnt = 0 if s_ <= e_ else (s_ - e_ + (-step) - 1) // (-step)
        shape[ax] = cnt; off += s_ * st[ax] if cnt else 0; strides[ax] = st[ax] * step
    if _prod(shape) == 0:
        return np.zeros(shape, np.float32)
    return _strided(x, shape, strides, off, ctx)




def expand(x, shape, ctx=None):
    """math.rs:2168"""
    from .model_rs import expand_shape
    x = _f(x)
    try:
        tgt = tuple(expand_shape(x.shape, shape))       # 0 = the input's size, two-way broadcasting (math.rs:2189-2204)
    except ValueError as e:
        raise LeleB200Error(str(e))
    xs = [1] * (len(tgt) - x.ndim) + list(x.shape); st = _estrides(xs)
    return _strided(x, list(tgt), [0 if xs[i] == 1 else st[i] for i in range(len(tgt))], 0, ctx)


def split(x, axis, splits, ctx=None):
    """manipulation.rs:1091 / split_owned :1153"""
    x = _f(x)
    if not -x.ndim <= axis < x.ndim: raise LeleB200Error("Split: axis out of bounds (manipulation.rs:1169)")
    ax = axis % x.ndim; st = _estrides(x.shape); outs, o = [], 0
    if sum(int(s) for s in splits) != x.shape[ax]: raise LeleB200Error("Split: splits sum mismatch (manipulation.rs:1173)")
    for s in splits:
        shp = list(x.shape); shp[ax] = int(s)
        outs.append(_strided(x, shp, st, o * st[ax], ctx)); o += int(s)
    return outs


split_owned = split


def concat(inputs, axis, ctx=None):
    """manipulation.rs:108"""
    ctx = ctx or default_context()
    xs = [_f(a) for a in inputs]
    ne = [a for a in xs if a.size > 0]
    if not ne:
        return np.zeros((0,), np.float32)
    ax = axis % ne[0].ndim
    for a in ne:
        if a.ndim != ne[0].ndim: raise LeleB200Error("Concat: ranks mismatch (manipulation.rs:159)")
        if any(a.shape[i] != ne[0].shape[i] for i in range(a.ndim) if i != ax): raise LeleB200Error("Concat: inner dim mismatch (manipulation.rs:162)")
    outer = _prod(ne[0].shape[:ax]); inner = _prod(ne[0].shape[ax + 1:])
    shp = list(ne[0].shape); shp[ax] = sum(a.shape[ax] for a in ne)
    if len(ne) > 16:   # the ABI takes up to 16 inputs per call
        raise LeleB200Error("concat: more than 16 inputs per call not supported by this wrapper")

    def fn(c, o, *ptrs):
        call("lele_b200_concat", c.h, (vp * len(ne))(*ptrs), _ll([a.shape[ax] for a in ne]), i32(len(ne)), i64(outer), i64(inner), o)
    return _run(tuple(shp), fn, *ne, ctx=ctx)


def pad(x, pads, constant_value=0.0, mode="constant", ctx=None):
    """manipulation.rs:382"""
    x = _f(x); r = x.ndim
    if r > 4: raise LeleB200Error(f"Pad: Rank {r} not fully implemented (manipulation.rs:485)")
    p = [_b.max(int(v), 0) for v in pads]
    if len(p) < 2 * r:
        half = len(p) // 2; miss = r - half; full = [0] * (2 * r)
        for i in range(half):
            full[miss + i] = p[i]; full[r + miss + i] = p[half + i]
        p = full
    shp = [x.shape[i] + p[i] + p[i + r] for i in range(r)]
    m = {"edge": 1, "reflect": 2}.get(mode, 0)     # any other mode string leaves the constant fill in place (manipulation.rs:487-491)
    return _run(tuple(shp), lambda c, o, px: call("lele_b200_pad", c.h, px, _ll(x.shape), i32(r), _ll(p), i32(m), f32(constant_value), o), x, ctx=ctx)


def gather(data, indices, axis=0, ctx=None):
    """manipulation.rs:589"""
    x = _f(data); idx = _f(indices); ax = axis % x.ndim
    outer = _prod(x.shape[:ax]); inner = _prod(x.shape[ax + 1:])
    ishape = indices.shape if isinstance(indices, DeviceTensor) else np.shape(indices)
    shp = tuple(x.shape[:ax]) + tuple(ishape) + tuple(x.shape[ax + 1:])   # a rank-0 index removes the axis (manipulation.rs:609)
    return _run(shp, lambda c, o, px, pi: call("lele_b200_gather", c.h, px, i64(outer), i32(x.shape[ax]), i64(inner), pi, i64(idx.size), o), x, idx, ctx=ctx)


def gather_elements(data, indices, axis, ctx=None):
    """conv2d.rs:1438"""
    x = _f(data); idx = _f(indices); ax = axis % x.ndim
    outer = _prod(x.shape[:ax]); inner = _prod(x.shape[ax + 1:])
    return _run(idx.shape, lambda c, o, px, pi: call("lele_b200_gather_elements", c.h, px, pi, i64(outer), i32(x.shape[ax]), i32(idx.shape[ax]), i64(inner), o), x, idx, ctx=ctx)


def tile(x, repeats, ctx=None):
    """math.rs:2249"""
    x = _f(x); rep = [int(r) for r in repeats]
    if len(rep) != x.ndim: raise LeleB200Error("Tile: repeats length must match input rank (math.rs:2256)")
    shp = [x.shape[i] * rep[i] for i in range(x.ndim)]
    return _run(tuple(shp), lambda c, o, px: call("lele_b200_tile", c.h, px, _ll(x.shape), _ll(rep), i32(x.ndim), o), x, ctx=ctx)


def topk(x, k, axis=-1, ctx=None):
    """conv2d.rs:1385 (last axis; indices as f32; `_axis` ignored upstream)"""
    x = _f(x); n = x.shape[-1]; k = min(int(k), n); outer = x.size // n
    shp = tuple(x.shape[:-1]) + (k,)
    return _run_multi([shp, shp], lambda c, o, px: call("lele_b200_topk", c.h, px, i64(outer), i32(n), i32(k), o[0], o[1]), x, ctx=ctx)


def argmax_last(x, ctx=None):
    x = _f(x); n = x.shape[-1]
    return _run(x.shape[:-1], lambda c, o, px: call("lele_b200_argmax_last", c.h, px, i64(x.size // n), i32(n), o), x, ctx=ctx, out_dtype=np.int32)


# zero-copy shape ops stay on the host (shape.rs:2-223)
def _fh(a):   # like _f, but a rank-0 value keeps its rank (np.ascontiguousarray would make it rank 1)
    return a if isinstance(a, DeviceTensor) else np.asarray(a, dtype=np.float32, order="C")


def reshape(x, shape):
    """shape.rs:2 (three-pass target resolution)"""
    from .model_rs import resolve_reshape
    x = _fh(x)
    try:
        return x.reshape(resolve_reshape(x.shape, shape))
    except ValueError as e:
        raise LeleB200Error(str(e))


def flatten(x, axis=1):
    """shape.rs:105: [prod(shape[:axis]), prod(shape[axis:])], negative axis counted from the end"""
    x = _fh(x); axis = axis + x.ndim if axis < 0 else axis
    return x.reshape((_prod(x.shape[:axis]), _prod(x.shape[axis:])))


def unsqueeze(x, axes):
    """shape.rs:133 (raw axes sorted, resolved against the output rank, inserted in turn)"""
    from .model_rs import unsqueeze_shape
    x = _fh(x); return x.reshape(unsqueeze_shape(x.shape, axes))


def squeeze(x, axes=None):
    """shape.rs:157: axes None = every dim of 1; with axes only listed dims that ARE 1 go (an empty list removes nothing)"""
    from .model_rs import squeeze_shape
    x = _fh(x); return x.reshape(squeeze_shape(x.shape, axes))
