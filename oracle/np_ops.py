"""numpy restatement of lele's indexing / shape / element-wise operators.
TEST INFRASTRUCTURE ONLY (see oracle/lele_oracle.h).  Citations: /root/reference paths.

These ops are bit-exact copies or single IEEE operations, so numpy float32 arithmetic
reproduces the reference exactly (no accumulation order is involved except reductions).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _a(x):
    return np.ascontiguousarray(x, dtype=np.float32)


# ---- manipulation.rs ----
def concat(xs, axis):  # manipulation.rs:108-207 (empty inputs skipped)
    xs = [_a(x) for x in xs if np.asarray(x).size > 0]
    if not xs:
        return np.zeros((0,), np.float32)                       # TensorView::empty(), manipulation.rs:131-144
    ax = axis + xs[0].ndim if axis < 0 else axis
    for x in xs:
        if x.ndim != xs[0].ndim:
            raise ValueError("Concat: ranks mismatch")          # manipulation.rs:159
        if any(x.shape[d] != xs[0].shape[d] for d in range(x.ndim) if d != ax):
            raise ValueError("Concat: inner dim mismatch")      # manipulation.rs:162
    return np.concatenate(xs, axis=ax)


def _slice_bounds(dim, start, end, step):  # manipulation.rs:268-330
    I64MAX, I64MIN = 2**63 - 1, -2**63
    end_max = end > I64MAX // 2
    end_min = end < I64MIN // 2
    s = dim if start > dim else (-dim if start < -dim else start)
    if end_max: e = dim
    elif end_min: e = -dim
    elif end > dim: e = dim
    elif end < -dim: e = -dim
    else: e = end
    ns = s + dim if s < 0 else s
    if end_max: ne = dim if step > 0 else -1
    elif end_min: ne = 0 if step > 0 else -1
    elif e < 0: ne = e + dim
    else: ne = e
    if step > 0:
        return min(max(ns, 0), dim), min(max(ne, 0), dim)
    return min(max(ns, 0), dim - 1), min(max(ne, -1), dim - 1)


def slice_(x, starts, ends, axes=(), steps=()):  # manipulation.rs:209-381
    x = _a(x)
    idx = [np.arange(d) for d in x.shape]
    for i in range(len(starts)):
        ax = i if len(axes) == 0 else (axes[i] + x.ndim if axes[i] < 0 else axes[i])
        step = steps[i] if i < len(steps) else 1
        s, e = _slice_bounds(x.shape[ax], int(starts[i]), int(ends[i]), step)
        idx[ax] = np.arange(s, e, step)
    return x[np.ix_(*idx)]


def pad(x, pads, value=0.0, mode="constant"):  # manipulation.rs:382-588
    x = _a(x)
    r = x.ndim
    if r > 4:
        raise ValueError(f"Pad: Rank {r} not fully implemented")   # manipulation.rs:485
    p = [max(int(v), 0) for v in pads]
    if len(p) < 2 * r:
        half = len(p) // 2
        miss = r - half
        full = [0] * (2 * r)
        for i in range(half):
            full[miss + i] = p[i]
            full[r + miss + i] = p[half + i]
        p = full
    widths = [(p[i], p[i + r]) for i in range(r)]
    if mode not in ("edge", "reflect"):                             # only these two are special-cased (manipulation.rs:487-491)
        return np.pad(x, widths, mode="constant", constant_values=f32(value))
    return np.pad(x, widths, mode=mode)  # "edge" / "reflect" match numpy's definitions


def gather(x, indices, axis=0):  # manipulation.rs:589-640
    x = _a(x)
    idx = np.asarray(indices).astype(np.int64)
    ax = axis + x.ndim if axis < 0 else axis
    idx = np.where(idx < 0, idx + x.shape[ax], idx)
    return np.take(x, idx, axis=ax)


def transpose(x, perm=()):  # manipulation.rs:644-1080 (empty perm = reverse)
    x = _a(x)
    return np.ascontiguousarray(np.transpose(x, perm if len(perm) else None))


def split(x, axis, splits):  # manipulation.rs:1091-1213
    x = _a(x)
    ax = axis + x.ndim if axis < 0 else axis
    if not 0 <= ax < x.ndim:
        raise ValueError("Split: axis out of bounds")          # manipulation.rs:1169
    if sum(int(s) for s in splits) != x.shape[ax]:
        raise ValueError("Split: splits sum mismatch")         # manipulation.rs:1173
    outs, o = [], 0
    for s in splits:
        sl = [slice(None)] * x.ndim
        sl[ax] = slice(o, o + int(s))
        outs.append(np.ascontiguousarray(x[tuple(sl)]))
        o += int(s)
    return outs


def where_op(cond, x, y):  # manipulation.rs:1215-1399 (non-zero = true)
    return np.where(np.asarray(cond) != 0, _a(x), _a(y)).astype(np.float32)


def expand(x, shape):  # math.rs:2168-2248
    x = _a(x); shape = [int(s) for s in shape]; n = max(x.ndim, len(shape)); tgt = []
    for i in range(n):                                   # right-aligned; 0 = "the input's size here" (math.rs:2189); two-way broadcast
        d_in = x.shape[i - (n - x.ndim)] if i >= n - x.ndim else 1
        d_t = shape[i - (n - len(shape))] if i >= n - len(shape) else 1
        d_t = d_in if d_t == 0 else d_t
        if d_in != d_t and d_in != 1 and d_t != 1:
            raise ValueError(f"Expand: incompatible dimensions at dim index {i} (from left): in={d_in} target={d_t}")
        tgt.append(d_in if d_t == 1 else d_t)
    return np.ascontiguousarray(np.broadcast_to(x, tgt))


def tile(x, repeats):  # math.rs:2249-2300
    x = _a(x)
    if len(repeats) != x.ndim:
        raise ValueError("Tile: repeats length must match input rank")   # math.rs:2256
    return np.tile(x, tuple(int(r) for r in repeats))


def flatten(x, axis=1):  # shape.rs:105-120
    x = _a(x); axis = axis + x.ndim if axis < 0 else axis
    return x.reshape(int(np.prod(x.shape[:axis], dtype=np.int64)), int(np.prod(x.shape[axis:], dtype=np.int64)))


def _try_reshape(in_shape, tgt, total):  # try_reshape_with_zeros, shape.rs:54-93
    new, known, infer = [], 1, None
    for i, d in enumerate(tgt):
        if d == -1:
            if infer is not None:
                return None
            infer = i
        elif d == 0:
            if i >= len(in_shape):
                return None
            new.append(in_shape[i]); known *= in_shape[i]
        else:
            new.append(int(d)); known *= int(d)
    if infer is not None:
        if known == 0 or total % known:
            return None
        new.insert(infer, total // known)
    return new if int(np.prod(new, dtype=np.int64)) == total else None


def reshape(x, shape):  # shape.rs:2-52: ONNX 0 / -1 rules, then 0 read as -1, then rank collapse [first, -1, last rank-1 dims]
    x = _a(x); tgt = [int(s) for s in shape]; ish = list(x.shape)
    tries = [tgt, [-1 if d == 0 else d for d in tgt]]
    if len(tgt) > x.ndim > 0:
        tries.append([tgt[0] if tgt[0] > 0 else -1, -1] + [d if d > 0 else -1 for d in (tgt[len(tgt) - (x.ndim - 1):] if x.ndim > 1 else [])])
    for t in tries:
        shp = _try_reshape(ish, t, x.size)
        if shp is not None:
            return x.reshape(shp)
    raise ValueError(f"Reshape: element count mismatch (input={ish} target={tgt})")


def unsqueeze(x, axes):  # shape.rs:133-156: raw axes sorted, resolved against the OUTPUT rank, inserted one after the other
    x = np.asarray(x, np.float32); new = list(x.shape); rank = x.ndim + len(axes)
    for a in sorted(int(a) for a in axes):
        idx = rank + a if a < 0 else a
        new.insert(idx, 1) if idx <= len(new) else new.append(1)
    return x.reshape(new)


def squeeze(x, axes=None):  # shape.rs:157-183
    x = np.asarray(x, np.float32)
    if axes is None:
        return x.reshape([d for d in x.shape if d != 1])
    pick = {a + x.ndim if a < 0 else a for a in axes}
    return x.reshape([d for i, d in enumerate(x.shape) if not (d == 1 and i in pick)])


def topk(x, k):  # conv2d.rs:1385-1437: last axis, stable, indices as f32
    x = _a(x)
    k = min(int(k), x.shape[-1])
    order = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(x, order, axis=-1), order.astype(np.float32)


def gather_elements(x, indices, axis):  # conv2d.rs:1438-1506: f32 indices, negative wrap
    x = _a(x)
    ax = axis + x.ndim if axis < 0 else axis
    idx = np.asarray(indices).astype(np.int64)
    idx = np.where(idx < 0, idx + x.shape[ax], idx)
    return np.take_along_axis(x, idx, axis=ax)


def resize_nearest(x, scales=None, sizes=None, mode="asymmetric"):  # conv2d.rs:1261-1384
    x = _a(x)
    n, c, h, w = x.shape
    if sizes is not None:
        if len(sizes) < 4:
            raise ValueError("Resize: sizes must have at least 4 elements")        # conv2d.rs:1301
        if not (sizes[2] > 0 and sizes[3] > 0):
            raise ValueError("Resize: sizes H and W must be positive")             # conv2d.rs:1305 (test at :3619)
        oh, ow = int(sizes[2]), int(sizes[3])
    elif scales is not None:
        sh = scales[2] if len(scales) >= 3 else 1.0
        sw = scales[3] if len(scales) >= 4 else 1.0
        if not (sh > 0 and sw > 0):
            raise ValueError("Resize: scales must be positive")                    # conv2d.rs:1312
        oh, ow = int(np.float64(h) * np.float64(f32(sh))), int(np.float64(w) * np.float64(f32(sw)))
    else:
        raise ValueError("Resize: either scales or sizes must be provided")        # conv2d.rs:1318
    if not (oh > 0 and ow > 0):
        raise ValueError(f"Resize: output dimensions must be positive, got out_h={oh} out_w={ow}")   # conv2d.rs:1323
    hs, ws = f32(h) / f32(oh), f32(w) / f32(ow)

    def src(o, scale, lim):
        o = np.arange(o, dtype=np.float32)
        if mode == "asymmetric":
            v = np.minimum(np.floor(o * scale), f32(lim - 1))
        else:
            t = (o + f32(0.5)) * scale - f32(0.5)
            v = np.sign(t) * np.floor(np.abs(t) + f32(0.5))  # f32::round = half away from zero
            v = np.minimum(np.maximum(v, f32(0.0)), f32(lim - 1))
        return v.astype(np.int64)

    ih, iw = src(oh, hs, h), src(ow, ws, w)
    return np.ascontiguousarray(x[:, :, ih][:, :, :, iw])


# ---- math.rs element-wise (NumPy broadcasting, utils.rs:107) ----
def add(a, b): return (_a(a) + _a(b)).astype(np.float32)           # math.rs:414
def sub(a, b): return (_a(a) - _a(b)).astype(np.float32)           # math.rs:838
def mul(a, b): return (_a(a) * _a(b)).astype(np.float32)           # math.rs:611
def div(a, b): return (_a(a) / _a(b)).astype(np.float32)           # math.rs:1106
def maximum(a, b): return np.maximum(_a(a), _a(b))                 # math.rs:1922
def neg(x): return -_a(x)                                          # math.rs:2142
def sqrt(x): return np.sqrt(_a(x))                                 # math.rs:1460
def reciprocal(x): return (f32(1.0) / _a(x)).astype(np.float32)    # math.rs:893
def clip(x, lo, hi): return np.minimum(np.maximum(_a(x), f32(lo)), f32(hi))  # math.rs:1984
# libm-backed f32 functions (Rust's f32::powf / ln / sin / cos call the platform libm; numpy's float32 loops do the same):
# parity bar 1e-5 relative with an absolute floor, not bit-exact
def pow(a, b): return np.power(_a(a), _a(b)).astype(np.float32)    # noqa: A001  math.rs:1481
def log(x):                                                          # math.rs:2130
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(_a(x)).astype(np.float32)
def sin(x): return np.sin(_a(x)).astype(np.float32)                # math.rs:2084
def cos(x): return np.cos(_a(x)).astype(np.float32)                # math.rs:2096
def equal(a, b): return (_a(a) == _a(b)).astype(np.float32)        # math.rs:1193: 1.0 / 0.0
def less(a, b): return (_a(a) < _a(b)).astype(np.float32)          # math.rs:2154
def not_(x): return (_a(x) == 0).astype(np.float32)                # math.rs:1508


def mod_f32(a, b):  # math.rs:1163-1192 : a - b*floor(a/b), 0 when b == 0
    a, b = np.broadcast_arrays(_a(a), _a(b))
    with np.errstate(divide="ignore", invalid="ignore"):
        r = a - b * np.floor(a / b)
    return np.where(b == 0, f32(0), r).astype(np.float32)


def prelu(x, slope):  # math.rs:2012: a one-element slope keeps the input's shape (:2017-2028), otherwise broadcasting
    x = _a(x); s = _a(slope)
    if s.size == 1:
        s = s.reshape(1)
    return np.where(x < 0, x * s, x).astype(np.float32)


def reduce(x, axes, keepdims, kind):  # math.rs:1527-1921
    x = _a(x)
    axes = tuple(sorted({a + x.ndim if a < 0 else a for a in axes}))   # sorted + dedup (math.rs:1628); an EMPTY list reduces nothing (reduce_mask stays false, math.rs:1631): sum -> 0 + x, l2 -> |x|
    if kind == "sum":
        return np.add.reduce(x, axis=axes, keepdims=keepdims, dtype=np.float32)
    if kind == "mean":
        cnt = int(np.prod([x.shape[a] for a in axes]))
        return (np.add.reduce(x, axis=axes, keepdims=keepdims, dtype=np.float32) * (f32(1.0) / f32(cnt))).astype(np.float32)
    if kind == "max":
        return np.max(x, axis=axes, keepdims=keepdims)
    if kind == "l2":
        return np.sqrt(np.add.reduce(x * x, axis=axes, keepdims=keepdims, dtype=np.float32)).astype(np.float32)
    raise ValueError(kind)


# ---- greedy decode (examples/sensevoice/src/tokenizer.rs:37-82): test infrastructure only ----
def greedy_filter(ids, skip_mask):
    """Per clip: frame ids that are not blank (0) and not flagged in skip_mask, in frame order (tokenizer.rs:62-68)."""
    out = []
    for row in np.asarray(ids):
        out.append(np.array([int(i) for i in row if i != 0 and 0 < i < len(skip_mask) and not skip_mask[i]], np.int32))
    return out


def decode_greedy(logits, id_to_token):
    """logits [B, T, V] -> texts: arg-max with the LAST maximum winning (Iterator::max_by, tokenizer.rs:55), skip id 0 and
    "<|...|>" tokens, join, replace the sentencepiece underscore with a space, trim (tokenizer.rs:71-79)."""
    lg = np.asarray(logits, np.float32)
    texts = []
    for b in range(lg.shape[0]):
        toks = []
        for t in range(lg.shape[1]):
            row = lg[b, t]
            tid = int(row.shape[0] - 1 - np.argmax(row[::-1]))
            if tid < len(id_to_token):
                tok = id_to_token[tid]
                if tid == 0 or (tok.startswith("<|") and tok.endswith("|>")):
                    continue
                toks.append(tok)
        texts.append("".join(toks).replace("\u2581", " ").strip())
    return texts
