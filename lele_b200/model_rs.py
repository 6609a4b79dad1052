"""Runs a `lele_gen`-generated `model.rs` on the B200 back-end without a Rust toolchain.

The AOT compiler emits every model as a list of `lele::kernels::<op>(...)` calls (straight-line except for ONNX `If` blocks,
spread over one or more `run_chunk_N` functions) whose weights are
`(offset, len, shape)` views into one `weights.bin` blob (src/compiler/mod.rs:1053-1092, :1381-1505; a committed
sample is examples/yolo26n-seg/src/yolo26seg.rs).  `parse_model_rs` turns such a file into a small JSON-able
"program" (statement list + weight literals); `run_program` replays it against an operator namespace -- by default
the CUDA product (`lele_b200.kernels` over the C ABI), in tests also the CPU oracle -- so a compiled model is a
drop-in: same operator names, argument order and `weights.bin` layout.

Only the statement forms the code generator emits are understood (src/compiler/generate.rs:802-997,
src/compiler/ops/*.rs, patterns.rs, snippets/default_methods.rs; table in INTEGRATION.md); anything else raises.  Tensor work
goes to the operator namespace; i64 shape arithmetic (Shape / Gather / Concat / Range ... on a handful of elements) is
evaluated on the host, as upstream.
"""
from __future__ import annotations

import re

import numpy as np

__all__ = ["parse_model_rs", "run_program", "CudaOps", "weight_view", "GeneratedModel", "BatchRunner", "resolve_reshape", "squeeze_shape", "unsqueeze_shape", "expand_shape"]

_DTYPES = {"weight_f32": ("<f4", None), "weight_i64": ("<i8", None), "weight_i64_f32": ("<i8", np.float32), "weight_i32": ("<i4", None),
           "weight_i32_i64": ("<i4", np.int64), "weight_i32_f32": ("<i4", np.float32), "weight_u8": ("u1", np.float32), "weight_i8": ("i1", np.float32),
           "weight_f16": ("<f2", np.float32)}


def weight_view(blob, kind: str, offset: int, length: int, shape):
    """TensorView::from_bytes_* (src/tensor.rs:131-257): typed view of blob[offset : offset+len] with the literal shape."""
    src, cast = _DTYPES[kind]
    a = np.frombuffer(blob, dtype=src, count=length // np.dtype(src).itemsize, offset=offset)
    if cast is not None:
        a = a.astype(cast)
    return a.reshape(tuple(shape)) if len(shape) else a.reshape(())


# ------------------------------------------------------------------------------------------------ parsing
def _split_top(s: str):
    """Split on top-level commas (parentheses / brackets / string literals respected)."""
    out, depth, cur, in_str = [], 0, [], False
    for ch in s:
        if in_str:
            cur.append(ch)
            if ch == '"':
                in_str = False
            continue
        if ch == '"':
            in_str = True; cur.append(ch); continue
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip()); cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


_W = re.compile(r"^&?self\.(weight_[a-z0-9_]+)\((\d+),\s*(\d+),\s*&\[([^\]]*)\]\)(.*)$")


def _parse_arg(a: str):
    a = a.strip()
    if a in ("None", "&lele::tensor::TensorView::empty()"):
        return None
    if a.startswith("Some(") and a.endswith(")"):
        return _parse_arg(a[5:-1])
    if a in ("true", "false"):
        return a == "true"
    if a.startswith('"') and a.endswith('"'):
        return {"str": a[1:-1]}
    m = _W.match(a)
    if m:
        kind, off, ln, shp, tail = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5).strip()
        shape = [int(x) for x in shp.split(",") if x.strip()]
        w = {"weight": [kind, off, ln, shape]}
        if tail in ("", ".data"):
            return w
        if tail == ".data[0] as usize":
            return {"weight_scalar": w["weight"]}
        if tail == ".as_i64_vec()":
            return {"weight_list": w["weight"]}
        raise ValueError(f"model.rs: unsupported weight expression tail {tail!r}")
    m = re.fullmatch(r"&(\w+)\.data\[\.\.\]", a)                       # an i64 shape tensor used as a slice (ops/tensor.rs:27)
    if m:
        return {"i64vec": m.group(1)}
    m = re.fullmatch(r"&lele::kernels::to_i64_vec\((.+)\)", a)           # inline conversion of a tensor or a weight literal (ops/math.rs:232)
    if m:
        return {"i64vec_of": _parse_arg(m.group(1))}
    if a.startswith("&[") and a.endswith("]"):
        items = _split_top(a[2:-1])
        if items and items[0].startswith("&"):
            if all(re.fullmatch(r"&\w+", i) for i in items):
                return {"vars": [i.lstrip("&").strip() for i in items]}
            return {"items": [_parse_arg(i) for i in items]}            # tensors and weight literals mixed (Concat of a shape with constants)
        return {"list": [_num(i) for i in items]}
    if re.fullmatch(r"-?\d+i64", a):
        return {"i64": int(a[:-3])}
    if a.startswith("&mut "):
        return {"out": a[5:].strip()}
    if re.fullmatch(r"&\s*[A-Za-z_][A-Za-z0-9_]*", a):
        return {"var": a[1:].strip()}
    if re.fullmatch(r"-?\d+(\.\d+)?([eE][-+]?\d+)?(f32|i64|usize)?", a):
        return _num(a)
    if re.fullmatch(r"[A-Za-z_][A-Za-z0-9_]*", a):
        return {"var": a}
    raise ValueError(f"model.rs: unsupported argument {a!r}")


def _num(t: str):
    t = re.sub(r"(f32|i64|usize|i32)$", "", t.strip())
    return float(t) if any(c in t for c in ".eE") and not t.lstrip("-").isdigit() else int(t)


_IF = re.compile(r"^let \(([\w, ]*)\) = if (\w+)\.data\.get\(0\)\.map\(\|v\| \*v != 0(?:\.0)?\)\.unwrap_or\(false\) \{$")


def _parse_tail(expr: str):
    """`(a.to_owned(), self.weight(0, 4, &[1]).to_owned())` or `a.to_owned()` -> [arg]: tensor names or stored tensors."""
    expr = expr.strip()
    items = _split_top(expr[1:-1]) if expr.startswith("(") else [expr]
    out = []
    for it in items:
        it = re.sub(r"\.to_owned\(\)$", "", it.strip())
        it = re.sub(r"^self\.weight\(", "self.weight_f32(", it)            # control_flow.rs:84 names the f32 loader `weight`
        out.append(_parse_arg(it))
    return out


def _parse_block(lines, i):
    """Statements from lines[i] up to the line that closes the block (`}`, `} else {` or `};`).  Returns (statements, tail, index of
    the closing line); `tail` = the block's value expression as a list of args, or None."""
    stmts, tail = [], None
    splits, split_src, skip_next = None, None, False
    while i < len(lines):
        line = lines[i]
        if line in ("}", "} else {", "};"):
            return stmts, tail, i
        i += 1
        if line.startswith("//") or not line:
            continue
        if line == '#[cfg(target_arch = "aarch64")]':          # the pre-packed NEON arm of a statement pair (patterns.rs:383, ops/math.rs:60):
            skip_next = True                                     # this back-end takes the portable arm that follows
            continue
        if line.startswith("#[cfg(not(target_arch"):
            continue
        if skip_next:
            skip_next = False
            continue
        m = _IF.match(line)                                      # ONNX If (ops/control_flow.rs:17-152): both branches are blocks with a tail
        if m:
            then_s, then_t, i = _parse_block(lines, i)
            if lines[i] != "} else {":
                raise ValueError("model.rs: malformed if / else block")
            else_s, else_t, i = _parse_block(lines, i + 1)
            if lines[i] != "};":
                raise ValueError("model.rs: malformed if / else block")
            i += 1
            outs = [o.strip() for o in m.group(1).split(",") if o.strip()]
            if then_t is None or else_t is None or len(then_t) != len(outs) or len(else_t) != len(outs):
                raise ValueError("model.rs: if / else branches must yield one value per output")
            stmts.append({"outs": outs, "op": "if", "args": [{"var": m.group(2)}],
                          "then": {"statements": then_s, "outputs": then_t}, "else": {"statements": else_s, "outputs": else_t}})
            continue
        if re.match(r"^\(.*\)$", line) or re.match(r"^\w+\.to_owned\(\)$", line):   # tail expression: (a.to_owned(), b.to_owned()) or a.to_owned()
            tail = _parse_tail(line)
            continue
        m = re.match(r"^let splits_slice = &\[(.*)\];$", line)
        if m:
            splits = [int(x) for x in m.group(1).split(",") if x.strip()]
            continue
        m = re.match(r"^let mut split_results = lele::kernels::split_owned\(&(\w+), (-?\d+), splits_slice\);$", line)
        if m:
            split_src = (m.group(1), int(m.group(2)), splits)
            continue
        m = re.match(r"^let (\w+) = split_results\.swap_remove\((\d+)\);$", line)
        if m:
            stmts.append({"outs": [m.group(1)], "op": "split_take", "args": [{"var": split_src[0]}, split_src[1], {"list": split_src[2]}, int(m.group(2))]})
            continue
        if re.match(r"^let mut (buf|temp_cast_buf)_\w+ = Vec::<(f32|i64)>::new\(\);$", line):
            continue
        m = re.match(r"^let (\w+) = (self\.weight_\w+\(.*\));$", line)     # Identity / Constant of a stored tensor (ops/tensor.rs:443)
        if m:
            stmts.append({"outs": [m.group(1)], "op": "constant", "args": [_parse_arg(m.group(2))]})
            continue
        m = re.match(r"^let (\w+) = lele::tensor::TensorView::from_owned\(vec!\[(.*)\], vec!\[(.*)\]\);$", line)   # small inline constant (:459, :489)
        if m:
            vals = [float(v) for v in m.group(2).split(",") if v.strip()]
            stmts.append({"outs": [m.group(1)], "op": "literal", "args": [{"list": vals}, {"list": [int(v) for v in m.group(3).split(",") if v.strip()]}]})
            continue
        m = re.match(r"^let (\w+) = lele::tensor::TensorView::empty\(\);", line)           # absent / oversized constant (:483, :514)
        if m:
            stmts.append({"outs": [m.group(1)], "op": "literal", "args": [{"list": []}, {"list": [0]}]})
            continue
        m = re.match(r"^let (\w+) = self\.(\w+)\((.*)\);$", line)   # helper methods of src/compiler/snippets/default_methods.rs
        if m:
            args = [_parse_arg(a) for a in _split_top(m.group(3))]
            bufs = [a["out"] for a in args if isinstance(a, dict) and "out" in a]
            args = [a for a in args if not (isinstance(a, dict) and "out" in a)]
            stmts.append({"outs": [m.group(1)], "op": "self." + m.group(2), "args": args, "bufs": bufs})
            continue
        m = re.match(r"^let (\w+) = (\w+)\.(?:clone|to_owned)\(\);", line)                 # incl. "// Cast f32->f32 is no-op" (ops/tensor.rs:236)
        if m:
            stmts.append({"outs": [m.group(1)], "op": "identity", "args": [{"var": m.group(2)}]})
            continue
        m = re.match(r"^let \((\w+), ([\w, ]+)\) = \(lele::kernels::(layer_norm)\((.*)\), lele::tensor::TensorView::empty\(\)[^;]*\);$", line)
        if m:                                                    # LayerNormalization with unused Mean / InvStdDev outputs (ops/nn.rs:261)
            line = f"let {m.group(1)} = lele::kernels::layer_norm({m.group(4)});"
        m = re.match(r"^let \(([\w, ]+)\) = lele::kernels::(\w+)\((.*)\);$", line) or re.match(r"^let (\w+) = lele::kernels::(?:utils::)?(\w+)\((.*)\);$", line)
        if m:
            outs = [o.strip() for o in m.group(1).split(",")]   # "_" = an output the graph never reads
            args = [_parse_arg(a) for a in _split_top(m.group(3))]
            bufs = [a["out"] for a in args if isinstance(a, dict) and "out" in a]     # `&mut ws.buf_N` / `&mut buf_<name>`: where the outputs live
            args = [a for a in args if not (isinstance(a, dict) and "out" in a)]      # (the resident replay binds them to HBM, see run_program)
            stmts.append({"outs": outs, "op": m.group(2), "args": args, "bufs": bufs})
            continue
        raise ValueError(f"model.rs: unsupported statement: {line[:160]}")
    raise ValueError("model.rs: unterminated block")


def parse_model_rs(text: str) -> dict:
    """-> {"class", "workspace_buffers", "inputs": [names], "statements": [{"outs", "op", "args"}], "outputs": [names]}
    (an `if` statement carries its two branches as nested {"statements", "outputs"} blocks)."""
    text = text.replace(".data.iter().map(|&v| v as i64).collect::<Vec<_>>()", ".as_i64_vec()")
    cls = re.search(r"pub struct (\w+)<'a>", text)
    n_bufs = len(re.findall(r"pub buf_\d+: Vec<f32>", text))
    lines = [raw.strip() for raw in text.splitlines()]
    stmts, inputs, outputs, i = [], None, None, 0
    while i < len(lines):
        m = re.match(r"fn run_chunk_\d+<'w>\(&self, ws: [^,]+, (.*)\) -> ", lines[i])
        i += 1
        if not m:
            continue
        if inputs is None:
            inputs = re.findall(r"(\w+): TensorView", m.group(1))
        body, tail, i = _parse_block(lines, i)
        if lines[i] != "}":
            raise ValueError("model.rs: malformed run_chunk body")
        stmts += body
        if tail is not None:
            if not all(isinstance(t, dict) and "var" in t for t in tail):
                raise ValueError("model.rs: a chunk returns tensors by name (generate.rs:772)")
            outputs = [t["var"] for t in tail]
    if inputs is None or outputs is None:
        raise ValueError("model.rs: no run_chunk body found")
    # A model split into several run_chunk_N functions shares one set of tensor names (generate.rs:704-790), so the statement list
    # is simply the chunks in order; the graph's own inputs and outputs are those of forward_with_workspace (mod.rs:1306-1350).
    fw = re.search(r"pub fn forward_with_workspace<'w>\(&self, ws: [^,]+, (.*?)\) -> [^{]*\{\n(.*?)\n    \}", text, re.S)
    if fw:
        inputs = re.findall(r"(\w+): TensorView", fw.group(1))
        tail = fw.group(2).strip().splitlines()[-1].strip()
        outputs = [p.strip() for p in _split_top(tail[1:-1] if tail.startswith("(") else tail)]
        if not all(re.fullmatch(r"\w+", o) for o in outputs):
            raise ValueError(f"model.rs: graph output that is a stored tensor is not supported: {tail[:120]}")
    return {"class": cls.group(1) if cls else "Model", "workspace_buffers": n_bufs, "inputs": inputs, "statements": stmts, "outputs": outputs}


# ------------------------------------------------------------------------------------------------ execution
class CudaOps:
    """The operator namespace `run_program` drives, over `lele_b200.kernels` (every call is a C-ABI launch)."""

    def __init__(self, ctx=None):
        from . import kernels as K
        self.K, self.ctx = K, ctx

    def conv2d(self, x, w, bias, dilations, group, pads, strides, act): return self.K.conv2d(x, w, bias, dilations, group, pads, strides, act, ctx=self.ctx)
    def conv_transpose(self, x, w, bias, dilations, pads, strides): return self.K.conv_transpose(x, w, bias, dilations, pads, strides, ctx=self.ctx)
    def conv_integer(self, x, w, x_zp, w_zp, dilations, group, pads, strides): return self.K.conv_integer(x, w, x_zp, w_zp, dilations, group, pads, strides, ctx=self.ctx)
    def matmul(self, a, b): return self.K.matmul(a, b, ctx=self.ctx)
    def softmax(self, x, axis): return self.K.softmax(x, axis, ctx=self.ctx)
    def max_pool2d(self, x, kernel, pads, strides, dilations, ceil_mode): return self.K.max_pool2d(x, kernel, pads, strides, dilations, ceil_mode, ctx=self.ctx)
    def concat(self, xs, axis): return self.K.concat(xs, axis, ctx=self.ctx)
    def slice(self, x, starts, ends, axes, steps): return self.K.slice(x, starts, ends, axes, steps, ctx=self.ctx)
    def gather(self, x, idx, axis): return self.K.gather(x, idx, axis, ctx=self.ctx)
    def gather_elements(self, x, idx, axis): return self.K.gather_elements(x, idx, axis, ctx=self.ctx)
    def transpose(self, x, perm): return self.K.transpose(x, perm, ctx=self.ctx)
    def split(self, x, axis, splits): return self.K.split(x, axis, splits, ctx=self.ctx)
    def tile(self, x, repeats): return self.K.tile(x, repeats, ctx=self.ctx)
    def topk(self, x, k): return self.K.topk(x, k, ctx=self.ctx)
    def resize_nearest(self, x, scales, sizes, mode): return self.K.resize_nearest(x, scales, sizes, mode, ctx=self.ctx)
    def reduce(self, x, axes, keepdims, kind):
        return {"sum": self.K.reduce_sum, "mean": self.K.reduce_mean, "max": self.K.reduce_max, "l2": self.K.reduce_l2}[kind](x, axes, keepdims, ctx=self.ctx)
    def binary(self, op, a, b): return getattr(self.K, op)(a, b, ctx=self.ctx)
    def unary(self, op, x): return getattr(self.K, "not_" if op == "not" else op)(x, ctx=self.ctx)
    def layer_norm(self, x, g, b, axis, eps): return self.K.layer_norm(x, g, b, axis, eps, ctx=self.ctx)
    def gemm(self, a, b, c, alpha, beta, ta, tb): return self.K.gemm(a, b, c, alpha, beta, ta, tb, ctx=self.ctx)
    def matmul_fused_add(self, a, b, bias): return self.K.matmul_fused_add(a, b, bias, ctx=self.ctx)
    def conv1d(self, x, w, bias, dilations, group, pads, strides, relu): return self.K.conv1d_fused(x, w, bias, dilations, group, pads, strides, relu, ctx=self.ctx)
    def pad(self, x, pads, value, mode): return self.K.pad(x, pads, value, mode, ctx=self.ctx)
    def expand(self, x, shape): return self.K.expand(x, shape, ctx=self.ctx)
    def where(self, c, x, y): return self.K.where_op(c, x, y, ctx=self.ctx)
    def fused_quantized_linear(self, x, w, ws, wz, bias, relu): return self.K.fused_quantized_linear(x, w, ws, wz, bias, relu, ctx=self.ctx)
    def prepare_weights(self, w, ws, wz, bias):     # prepare_weights (quantization.rs:221): packed once, resident in HBM
        return self.K.PreparedWeights(np.clip(np.asarray(w), 0, 255).astype(np.uint8), ws, wz, None if bias is None or np.size(bias) == 0 else bias, self.ctx)
    def dynamic_quantize_linear(self, x): return self.K.dynamic_quantize_linear(x, ctx=self.ctx)
    def mat_mul_integer(self, a, b, zp_a, zp_b): return self.K.mat_mul_integer(a, b, zp_a, zp_b, ctx=self.ctx)
    def clip(self, x, lo, hi): return self.K.clip(x, lo, hi, ctx=self.ctx)
    def stft(self, x, n_fft, hop, win, window): return self.K.stft(x, n_fft, hop, win, window, ctx=self.ctx)
    def batch_norm(self, x, scale, bias, mean, var, eps): return self.K.batch_norm(x, scale, bias, mean, var, eps, ctx=self.ctx)
    def lstm(self, x, w, r, bias, h0, c0): return self.K.lstm(x, w, r, bias, None, h0, c0, ctx=self.ctx)
    def gru(self, x, w, r, bias, h0): return self.K.gru(x, w, r, bias, h0, ctx=self.ctx)


class _NamespaceOps:
    """Adapter for a module exposing the shared operator vocabulary of the test suite (same names as CudaOps)."""

    def __init__(self, ns):
        self.ns = ns

    def __getattr__(self, name):
        return getattr(self.ns, name)

    _ALIAS = {"tanh_kernel": "tanh", "max": "maximum", "not": "not_"}

    def binary(self, op, a, b): return getattr(self.ns, self._ALIAS.get(op, op))(a, b)
    def unary(self, op, x): return getattr(self.ns, self._ALIAS.get(op, op))(x)
    def resize_nearest(self, x, scales, sizes, mode): return self.ns.resize_nearest(x, scales=scales, sizes=sizes, mode=mode)


def _is_dev(x):
    return hasattr(x, "ptr") and hasattr(x, "numpy")           # lele_b200.kernels.DeviceTensor (not imported: the parser half of this
                                                               # module is also loaded stand-alone, without the built library)


def _host(x):
    """`.data` read: a resident value crosses to the host only where the generated code itself reads the data (ops/nn.rs:434 Resize
    scales, to_i64_vec, If conditions, scalar attributes carried by tensors)."""
    return x.numpy() if _is_dev(x) else x


def _c(x, dtype=None):  # C-contiguous without np.ascontiguousarray's promotion of rank-0 values to rank 1
    if _is_dev(x):
        return x
    return np.asarray(x, dtype=dtype, order="C")


def resolve_reshape(in_shape, target):
    """The three passes of `reshape` (shape.rs:2-52): (1) ONNX rules -- 0 copies the input dim at that position, one -1 is inferred;
    (2) every 0 re-read as -1; (3) a target of higher rank than the input collapsed to [first, -1, last rank-1 dims].  The first
    pass whose element count matches wins; otherwise the reference panics."""
    in_shape = [int(d) for d in in_shape]; target = [int(d) for d in target]
    total = 1
    for d in in_shape:
        total *= d

    def attempt(tgt):                                  # try_reshape_with_zeros, shape.rs:54-93
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
                new.append(d); known *= d
        if infer is not None:
            if known == 0 or total % known != 0:
                return None
            new.insert(infer, total // known)
        prod = 1
        for d in new:
            prod *= d
        return new if prod == total else None

    for tgt in (target, [-1 if d == 0 else d for d in target]):
        got = attempt(tgt)
        if got is not None:
            return got
    if len(target) > len(in_shape) > 0:
        tail = target[len(target) - (len(in_shape) - 1):] if len(in_shape) > 1 else []
        got = attempt([target[0] if target[0] > 0 else -1, -1] + [d if d > 0 else -1 for d in tail])
        if got is not None:
            return got
    raise ValueError(f"Reshape: element count mismatch (input={in_shape} target={target})")


def squeeze_shape(in_shape, axes):
    """`squeeze` (shape.rs:157-183): with axes, a dim goes only if it is listed AND equals 1 (anything else listed is silently kept;
    an empty list removes nothing); without axes (None), every dim equal to 1 goes."""
    if axes is None:
        return [int(d) for d in in_shape if d != 1]
    pick = {int(a) + len(in_shape) if a < 0 else int(a) for a in axes}
    return [int(d) for i, d in enumerate(in_shape) if not (d == 1 and i in pick)]


def unsqueeze_shape(in_shape, axes):
    """`unsqueeze` (shape.rs:133-156): the output rank is fixed first, the RAW axes are sorted (negative ones come first), each is
    resolved against the output rank and inserted in turn; a position past the current end appends."""
    new, rank = [int(d) for d in in_shape], len(in_shape) + len(axes)
    for a in sorted(int(a) for a in axes):
        idx = rank + a if a < 0 else a
        if idx <= len(new):
            new.insert(idx, 1)
        else:
            new.append(1)
    return new


def expand_shape(in_shape, target):
    """`expand` (math.rs:2175-2210): shapes right-aligned, a target entry of 0 means "the input's size here", then two-way
    broadcasting (equal, or either side 1); anything else is the reference's panic."""
    n = max(len(in_shape), len(target)); out = []
    for i in range(n):
        d_in = int(in_shape[i - (n - len(in_shape))]) if i >= n - len(in_shape) else 1
        d_t = int(target[i - (n - len(target))]) if i >= n - len(target) else 1
        d_t = d_in if d_t == 0 else d_t
        if d_in == d_t or d_in == 1:
            out.append(d_t)
        elif d_t == 1:
            out.append(d_in)
        else:
            raise ValueError(f"Expand: incompatible dimensions at dim index {i} (from left): in={d_in} target={d_t}. Full shapes: in={list(in_shape)} target={list(target)}")
    return out


def _reshape(x, shape):
    x = _c(x)
    return x.reshape(tuple(resolve_reshape(x.shape, shape)))


def _to_i64_list(x):  # to_i64_vec (manipulation.rs:1082): `as i64` truncation of every element
    if isinstance(x, (list, tuple)):
        return [int(v) for v in x]
    return [int(v) for v in np.asarray(_host(x)).reshape(-1)]


def _is_i64(x):
    return isinstance(x, (np.ndarray, np.generic)) and np.asarray(x).dtype == np.int64


def _scalar0(x, dflt):
    x = np.asarray(_host(x)).reshape(-1)
    return x[0] if x.size else dflt


def _host_i64_op(op, a):
    """Shape arithmetic.  The code generator types shape-carrying values as i64 tensors (generate.rs var_types) and the reference
    computes them with the same generic kernels on the host; they are a few elements each and decide tensor SHAPES, so they stay
    host values here as well (numpy int64) and never reach the device.  Returns None when the statement is not of this class."""
    if op == "shape":                                 # shape.rs:100 (a resident value carries its shape on the host)
        return np.array(a[0].shape if _is_dev(a[0]) else np.asarray(a[0]).shape, np.int64)
    if op == "size":                                  # shape.rs:95: rank-0 tensor holding the element count
        return np.array(a[0].size if _is_dev(a[0]) else np.asarray(a[0]).size, np.int64)
    if op == "to_i64_vec":
        return _to_i64_list(a[0])
    if op in ("cast_to_i64", "cast_to_f32", "equal_i64", "equal_i64_f32_r", "equal_i64_f32_r_i64", "equal_i64_f32_lhs", "less_i64"):
        a = [_host(v) for v in a]                     # element types change here: host values from this point (they feed shape logic)
    if op == "cast_to_i64":                           # utils.rs:85
        return np.trunc(np.asarray(a[0])).astype(np.int64) if np.asarray(a[0]).dtype.kind == "f" else np.asarray(a[0]).astype(np.int64)
    if op == "cast_to_f32":                           # utils.rs:71
        return np.asarray(a[0]).astype(np.float32)
    if op == "constant_of_shape":                     # shape.rs:122: the value literal decides the element type
        shp = tuple(_to_i64_list(a[0]))
        return np.full(shp, a[1], np.int64 if isinstance(a[1], np.int64) else np.float32)
    if op == "range_i64":                             # math.rs:2057
        s0, lim, d = int(_scalar0(a[0], 0)), int(_scalar0(a[1], 0)), int(_scalar0(a[2], 1))
        n = max(int(np.ceil((lim - s0) / d)), 0) if d != 0 else 0
        return s0 + np.arange(n, dtype=np.int64) * d
    if op == "range":                                 # math.rs:2033: start + i * delta in f32
        s0, lim, d = np.float32(_scalar0(a[0], 0.0)), np.float32(_scalar0(a[1], 0.0)), np.float32(_scalar0(a[2], 1.0))
        n = int(max(np.ceil(np.float32(lim - s0) / d), 0.0)) if d != 0 else 0
        return (s0 + np.arange(n, dtype=np.float32) * d).astype(np.float32)
    if op in ("equal_i64", "equal_i64_f32_r", "equal_i64_f32_r_i64", "equal_i64_f32_lhs"):   # math.rs:1201-1236: operands compared as i64
        x, y = (np.trunc(np.asarray(v)).astype(np.int64) if np.asarray(v).dtype.kind == "f" else np.asarray(v, np.int64) for v in a[:2])
        return (x == y).astype(np.int64)
    if op == "less_i64":                              # math.rs:2161
        return (np.asarray(a[0]) < np.asarray(a[1])).astype(np.int64)
    tens = [v for v in a if isinstance(v, (np.ndarray, np.generic))]
    if op == "concat":
        tens = list(a[0])
    if not tens or not all(_is_i64(v) for v in tens):
        return None
    if op in ("add", "sub", "mul"):
        return {"add": np.add, "sub": np.subtract, "mul": np.multiply}[op](a[0], a[1]).astype(np.int64)
    if op == "div":                                   # Rust i64 division truncates toward zero
        x, y = np.broadcast_arrays(np.asarray(a[0]), np.asarray(a[1]))
        return (np.sign(x) * np.sign(y) * (np.abs(x) // np.abs(y))).astype(np.int64)
    if op == "not":                                   # math.rs:1508
        return (np.asarray(a[0]) == 0).astype(np.int64)
    if op == "gather":                                # manipulation.rs:589 on an i64 tensor (picking dims out of a shape)
        x, idx = np.asarray(a[0]), np.asarray(a[1])
        return np.take(x, np.where(idx < 0, idx + x.shape[a[2]], idx), axis=a[2]).astype(np.int64)
    if op == "concat":
        return np.concatenate([np.atleast_1d(v) for v in a[0]], axis=a[1])
    if op == "slice":
        x = np.asarray(a[0]); sl = [np.s_[:]] * x.ndim
        starts, ends, axes, steps = a[1], a[2], (a[3] or list(range(len(a[1])))), (a[4] or [1] * len(a[1]))
        for st_, en, ax, sp in zip(starts, ends, axes, steps):
            sl[ax] = np.s_[int(np.clip(st_, -2**62, 2**62)):int(np.clip(en, -2**62, 2**62)):int(sp)]
        return _c(x[tuple(sl)])
    if op == "where_op":
        return np.where(np.asarray(a[0]) != 0, a[1], a[2]).astype(np.int64)
    if op == "expand":
        x = np.asarray(a[0])
        return _c(np.broadcast_to(x, tuple(expand_shape(x.shape, a[1]))))
    return None                                       # reshape / unsqueeze / squeeze / flatten / identity keep the dtype below


def _item_view(t, i, B):
    """Item i of a folded value: the i-th [1, ...] slice of a contiguous [B, ...] DeviceTensor (zero-copy)."""
    from .kernels import DeviceTensor
    per = t.size // B
    return DeviceTensor(t.ptr + i * per * t.dtype.itemsize, (1,) + tuple(t.shape[1:]), t.ctx, t, t.dtype, t.slot)


_LIFT_UNARY = {"sigmoid", "silu", "relu", "tanh_kernel", "erf", "exp", "sqrt", "neg", "reciprocal", "softplus", "log", "sin", "cos", "clip", "identity"}
_LIFT_BINARY = {"add", "sub", "mul", "div", "max", "pow", "prelu", "mod_f32"}
_LIFT_NCHW = {"conv2d", "conv2d_silu", "conv2d_fused", "conv_transpose", "max_pool2d", "resize_nearest", "batch_norm"}


def _liftable(st, a, lifted, env, B):
    """Can this statement run ONCE on folded operands ([B, ...] standing for B independent [1, ...] values)?
    -> None: no (the folded replay stops here); False: it has no folded operand (run it as it is); else the statement to execute
    (the original, or a copy whose reshape target has its leading 1 replaced by B).  Conservative: anything not listed stops."""
    if st["op"] == "if":
        return None
    def names(x):
        if isinstance(x, dict):
            if "var" in x: return [x["var"]]
            if "vars" in x: return list(x["vars"])
            if "items" in x: return [n for it in x["items"] for n in names(it)]
            if "i64vec" in x: return [x["i64vec"]]
            if "i64vec_of" in x: return names(x["i64vec_of"])
        return []
    arg_names = [names(x) for x in st["args"]]
    used = [n for ns in arg_names for n in ns]
    if not any(n in lifted for n in used):
        return False
    op = st["op"]
    folded = lambda k: any(n in lifted for n in arg_names[k])
    rank = lambda k: len(a[k].shape)
    norm = lambda ax, r: ax + r if ax < 0 else ax
    try:
        if op in _LIFT_NCHW:
            return st if folded(0) and not any(folded(k) for k in range(1, len(a))) and rank(0) >= 3 else None
        if op in _LIFT_UNARY:
            return st if not any(folded(k) for k in range(1, len(a))) else None
        if op in _LIFT_BINARY:
            r = max(len(np.shape(v)) if not _is_dev(v) else len(v.shape) for v in a[:2])
            for k in (0, 1):
                if folded(k) and rank(k) != r: return None                      # a folded operand must carry the leading (batch) dim of the result
                if not folded(k) and len(np.shape(a[k]) if not _is_dev(a[k]) else a[k].shape) == r and (np.shape(a[k]) if not _is_dev(a[k]) else a[k].shape)[0] != 1: return None
            return st
        if op == "concat":
            if not all(n in lifted for n in arg_names[0]): return None
            return st if norm(a[1], len(a[0][0].shape)) != 0 else None
        if op == "split_take":
            return st if folded(0) and norm(a[1], rank(0)) != 0 else None
        if op == "reshape":
            tgt = list(a[1])
            if not folded(0) or not tgt or tgt[0] not in (0, 1) or a[0].shape[0] != B: return None
            return dict(st, args=[st["args"][0], {"list": [B] + tgt[1:]}])
        if op == "flatten":
            return st if folded(0) and a[1] == 1 else None
        if op == "unsqueeze":
            r_out = rank(0) + len(a[1])
            return st if folded(0) and all(norm(int(x), r_out) != 0 for x in a[1]) else None
        if op == "squeeze":
            return st if folded(0) and a[1] is not None and all(norm(int(x), rank(0)) != 0 for x in a[1]) else None
        if op == "transpose":
            perm = list(a[1]) if len(a[1]) else list(reversed(range(rank(0))))
            return st if folded(0) and perm[0] == 0 else None
        if op == "matmul":
            if not folded(0) or rank(0) < 3: return None
            if folded(1): return st if rank(1) == rank(0) else None
            return st if len(np.shape(a[1])) == 2 else None
        if op == "softmax":
            return st if folded(0) and rank(0) >= 2 and norm(a[1], rank(0)) != 0 else None
        if op == "layer_norm":
            return st if folded(0) and not folded(1) and not folded(2) and norm(a[3], rank(0)) != 0 else None
        if op == "slice":
            axes = list(a[3]) if a[3] else list(range(len(a[1])))
            return st if folded(0) and all(norm(int(x), rank(0)) != 0 for x in axes) else None
        if op == "topk":
            return st if folded(0) and rank(0) >= 2 else None
        if op in ("reduce_max", "reduce_sum", "reduce_mean", "reduce_l2"):
            return st if folded(0) and len(a[1]) > 0 and all(norm(int(x), rank(0)) != 0 for x in a[1]) else None
        if op == "tile":
            return st if folded(0) and len(a[1]) == rank(0) and int(a[1][0]) == 1 else None
        if op == "gather_elements":
            return st if folded(0) and folded(1) and norm(a[2], rank(0)) != 0 else None
    except Exception:
        return None
    return None


def run_program(program: dict, blob, inputs, ops=None, trace=None, cache=None, workspace=None, download=True, env0=None, lift=None, items=None):
    """Replays the statement list.  `blob` = the model's weights.bin bytes, `inputs` = arrays in `program["inputs"]` order.
    `ops`: CudaOps (default) or any module with the shared operator vocabulary.  Returns the outputs as numpy arrays.
    `cache`: a dict the caller keeps across calls of one model -- decoded weight views and, where the namespace offers
    `prepare_weights`, the device-resident packed int8 weights of each quantised linear are made once (the role of
    B_WEIGHT_CACHE upstream, avx/quantization.rs:47-95: keyed by the weight's place in the blob).
    `workspace` (a `lele_b200.kernels.Workspace`, CudaOps only): the RESIDENT replay -- src/tensor.rs's buffer arena mapped to
    HBM.  Inputs are uploaded once, every statement's outputs are placed in the workspace buffer the generated code names
    (`&mut ws.buf_N` -> `lele_b200_arena_bind`; values the generated code owns -- split_owned pieces, intermediates of composed
    statements -- get a buffer keyed by their statement), weights are uploaded once per model, and nothing returns to the host
    except where the generated code itself reads `.data` and, with `download`, the graph outputs (`.to_owned()`, generate.rs:748).
    `lift` = B (resident replay only): the BATCH-FOLDED replay.  Generated code bakes batch 1 into its shapes, but almost every
    statement of a vision graph is independent along the leading dimension, so B inputs stacked as [B, ...] run through the SAME
    statement as one launch (the convolutions see nb = B: one implicit GEMM over the whole batch) -- a value whose logical shape is
    [1, ...] is carried as [B, ...].  A statement is folded only if `_liftable` can prove it independent per leading index (reshape
    targets get their leading 1 replaced by B); at the first top-level statement that is not (the detection tail: flatten to a
    row-major list, gather on axis 0 ...) the folded values are sliced per item (zero-copy views) and the remaining statements run
    per item: `items` = [(ops_i, workspace_i)] -- item i's operator namespace (its lane) and its own workspace.  Returns one output
    list per item.  `env0`: a ready environment (internal: the per-item tail)."""
    if ops is None:
        ops = CudaOps()
    elif not hasattr(ops, "binary"):
        ops = _NamespaceOps(ops)
    resident = workspace is not None
    if resident and not isinstance(ops, CudaOps):
        raise ValueError("run_program: a workspace (resident replay) needs the CUDA operator namespace")
    rctx = workspace.ctx if resident else None
    if lift is not None and not resident:
        raise ValueError("run_program: the batch-folded replay (lift) needs a workspace")
    env = dict(env0) if env0 is not None else {}
    lifted = set(program["inputs"]) if lift is not None else set()     # names of values carried as [B, ...] for logical [1, ...]
    for n, a in zip(program["inputs"], inputs if env0 is None else []):
        if _is_dev(a):
            env[n] = a
        elif np.asarray(a).dtype.kind in "iu":
            env[n] = _c(a, np.int64)
        else:
            env[n] = rctx.to_device(np.asarray(a, np.float32)) if resident else _c(a, np.float32)
    split_cache = {}
    stmt_no = [0 if env0 is None else 100000]         # (a per-item tail never shares statement-keyed buffer names with the folded prefix)

    def val(a):
        if isinstance(a, dict):
            if "var" in a: return env[a["var"]]
            if "vars" in a: return [env[v] for v in a["vars"]]
            if "items" in a: return [val(i) for i in a["items"]]
            if "i64vec" in a: return _to_i64_list(env[a["i64vec"]])
            if "i64vec_of" in a: return _to_i64_list(val(a["i64vec_of"]))
            if "i64" in a: return np.int64(a["i64"])
            if "weight" in a:
                if cache is None:
                    return weight_view(blob, *a["weight"])
                key = ("w", a["weight"][0], a["weight"][1], a["weight"][2], tuple(a["weight"][3]))
                if key not in cache:
                    cache[key] = weight_view(blob, *a["weight"])
                    if resident and a["weight"][0] == "weight_f32" and cache[key].size >= 64:
                        rctx.persist(cache[key])          # device copy keyed by the view's address ((blob_base, offset) upstream)
                return cache[key]
            if "weight_scalar" in a: return int(weight_view(blob, *a["weight_scalar"]).reshape(-1)[0])
            if "weight_list" in a: return [int(v) for v in weight_view(blob, *a["weight_list"]).reshape(-1)]
            if "list" in a: return list(a["list"])
            if "str" in a: return a["str"]
        return a

    def exec_block(statements):
        for st in statements:
            exec_statement(st)

    def exec_statement(st):
        if st["op"] == "if":                          # ops/control_flow.rs:41: the first element of the condition decides; names are
            cond = np.asarray(_host(env[st["args"][0]["var"]])).reshape(-1)   # graph-wide, so the taken branch runs in the same environment
            branch = st["then"] if (cond.size and cond[0] != 0) else st["else"]
            exec_block(branch["statements"])
            for n, o in zip(st["outs"], branch["outputs"]):
                env[n] = val(o)
            if trace is not None:
                trace.append((st["outs"][0] if st["outs"] else "", "if", env[st["outs"][0]] if st["outs"] else None))
            return
        op, a = st["op"], [val(x) for x in st["args"]]
        if resident:
            # the statement's own buffers first, then buffers keyed by the statement for whatever else it materialises
            # (the allocator never hands a statement one of its own input buffers, mod.rs:234-253; a hand-written statement that does
            # gets a statement-keyed buffer instead of an in-place launch)
            stmt_no[0] += 1
            held = {v.slot for x in a for v in (x if isinstance(x, (list, tuple)) else [x]) if _is_dev(v)}
            rctx.out_slots([(workspace, b) for b in st.get("bufs", []) if b not in held] + [(workspace, f"stmt{stmt_no[0]}_{k}") for k in range(8)])
        r = _host_i64_op(op, a)
        if r is not None:
            pass
        elif op == "constant":
            r = a[0]
        elif op == "literal":
            r = np.asarray(a[0], np.float32).reshape(tuple(a[1]))
        elif op == "stft":                            # (signal, n_fft, frame_step, n_fft, window)  ops/math.rs:480
            r = ops.stft(a[0], a[1], a[2], a[3], a[4])
        elif op in ("pow", "equal", "less"):
            r = ops.binary(op, a[0], a[1])
        elif op in ("log", "sin", "cos", "not"):
            r = ops.unary(op, a[0])
        elif op in ("conv2d", "conv2d_silu", "conv2d_fused"):
            act = 2 if op == "conv2d_silu" else (1 if (op == "conv2d_fused" and a[7]) else 0)
            r = ops.conv2d(a[0], a[1], a[2], a[3], a[4], a[5], a[6], act)
        elif op == "conv_integer":                    # (x, w, x_zero_point, w_zero_point, dilations, group, pads, strides)  ops/nn.rs:328
            # conv2d.rs:1507-2000: an f32 convolution of (x - x_zp) with (w - w_zp); the im2col pads with RAW zeros, so a padded
            # position contributes (0 - x_zp) (conv2d.rs:2025) -- hence pad first, shift second, convolve without padding
            xz, wz = (0.0 if z is None or np.size(_host(z)) == 0 else float(np.asarray(_host(z)).reshape(-1)[0]) for z in (a[2], a[3]))
            p4 = list(a[6]) if len(a[6]) >= 4 else (list(a[6]) * 2 if len(a[6]) == 2 else [0, 0, 0, 0])
            r = ops.conv_integer(a[0], a[1], xz, wz, a[4], a[5], p4, a[7])   # one C-ABI entry on the device (lele_b200_conv_integer)
        elif op == "conv_transpose":
            if a[4] != 1:
                raise ValueError("ConvTranspose: group > 1 not supported yet (conv2d.rs:3042)")
            r = ops.conv_transpose(a[0], a[1], a[2], a[3], a[5], a[6])
        elif op in ("add", "sub", "mul", "div", "mod_f32", "max", "prelu"):
            r = ops.binary(op, a[0], a[1])
        elif op in ("sigmoid", "silu", "relu", "tanh_kernel", "erf", "exp", "sqrt", "neg", "reciprocal", "softplus"):
            r = ops.unary(op, a[0])
        elif op == "layer_norm":                      # (x, scale, bias, axis, epsilon)  ops/nn.rs:282
            if a[1] is None or a[2] is None:          # the x86 kernel reads scale and bias unconditionally (norm.rs:244)
                raise ValueError("layer_norm: scale and bias are required (norm.rs:244)")
            r = ops.layer_norm(a[0], a[1], a[2], a[3], a[4])
        elif op == "gemm":                            # (a, b, c, alpha, beta, trans_a, trans_b)  ops/nn.rs:109
            r = ops.gemm(a[0], a[1], a[2], a[3], a[4], a[5], a[6])
        elif op == "lstm":                            # (x, w, r, bias, sequence_lens, initial_h, initial_c) -> (Y, H, C)  ops/nn.rs:146
            r = ops.lstm(a[0], a[1], a[2], a[3], a[5], a[6])
        elif op == "gru":                             # (x, w, r, bias, initial_h, linear_before_reset) -> (Y, H)  ops/nn.rs:195
            r = ops.gru(a[0], a[1], a[2], a[3], a[4])
        elif op in ("self.linear_quantized", "self.linear_quantized_relu"):   # (x, weight_u8 [K,N], weight_scale, weight_zero, bias)  default_methods.rs:36-62
            zero = int(np.asarray(_host(a[3])).reshape(-1)[0]) if np.size(_host(a[3])) else 0
            w = a[1]
            if cache is not None and hasattr(ops, "prepare_weights") and all(isinstance(x, dict) and "weight" in x for x in st["args"][1:5]):
                key = ("pw",) + tuple(x["weight"][1] for x in st["args"][1:5])
                if key not in cache:
                    cache[key] = ops.prepare_weights(a[1], a[2], zero, a[4])
                w = cache[key]
            r = ops.fused_quantized_linear(a[0], w, a[2], zero, a[4], op.endswith("_relu"))
        elif op == "self.layer_norm":                 # (x, scale, bias, epsilon tensor, two)  default_methods.rs:24: axis -1, eps = epsilon[0] or 1e-5
            eps = np.asarray(_host(a[3])).reshape(-1)
            r = ops.layer_norm(a[0], a[1], a[2], -1, float(eps[0]) if eps.size else 1e-5)
        elif op == "self.linear":
            r = ops.matmul_fused_add(a[0], a[1], a[2])
        elif op == "self.conv1d_relu":                # (x, w, bias, stride, dilation, groups, padding)  default_methods.rs:1
            r = ops.conv1d(a[0], a[1], a[2], [a[4]], a[5], [a[6], a[6]], [a[3]], True)
        elif op == "dynamic_quantize_linear":         # -> (q as f32, scale, zero_point)  ops/tensor.rs:407
            r = ops.dynamic_quantize_linear(a[0])
        elif op == "mat_mul_integer":                 # (a, b, a_zero_point, b_zero_point)  ops/math.rs:43; zero points are scalar tensors
            zp = [0.0 if z is None or np.size(_host(z)) == 0 else float(np.asarray(_host(z)).reshape(-1)[0]) for z in (a[2], a[3])]
            r = ops.mat_mul_integer(a[0], a[1], zp[0], zp[1])
        elif op == "clip":                            # (x, min, max): scalar tensors; a missing bound is -inf / +inf (math.rs:15-20, :1990-1997)
            lim = [d if z is None or np.size(_host(z)) == 0 else float(np.asarray(_host(z)).reshape(-1)[0]) for z, d in ((a[1], float("-inf")), (a[2], float("inf")))]
            r = ops.clip(a[0], lim[0], lim[1])
        elif op == "batch_norm":                      # (x, scale, bias, mean, var, epsilon)  ops/nn.rs:352
            r = ops.batch_norm(a[0], a[1], a[2], a[3], a[4], a[5])
        elif op in ("self.embedding_concat", "self.embedding_concat_i64"):   # (shape, value, weight)  default_methods.rs:182-232: ConstantOfShape + Concat
            shp = _to_i64_list(a[0]); w = np.asarray(a[2])          # on axis 0 as one flat append: a table with a constant tail, built on the host
            if op.endswith("_i64"):
                w = w.astype(np.int64)
            tail = np.full(int(np.prod(shp, dtype=np.int64)) if shp else 1, a[1], w.dtype)
            r = np.concatenate([w.reshape(-1), tail]).reshape((w.shape[0] + (shp[0] if shp else 1),) + tuple(w.shape[1:]))
        elif op == "matmul_fused_add":
            r = ops.matmul_fused_add(a[0], a[1], a[2])
        elif op in ("conv1d", "conv1d_fused"):        # same argument form as conv2d (ops/nn.rs:57); _fused appends the ReLU flag
            r = ops.conv1d(a[0], a[1], a[2], a[3], a[4], a[5], a[6], bool(a[7]) if (op == "conv1d_fused" and len(a) > 7) else False)
        elif op in ("reduce_sum", "reduce_mean", "reduce_l2"):
            r = ops.reduce(a[0], a[1], a[2], op[len("reduce_"):])
        elif op == "pad":                             # (x, pads, constant_value, mode)  ops/tensor.rs:403
            r = ops.pad(a[0], a[1], float(a[2]) if a[2] is not None else 0.0, a[3])
        elif op == "expand":
            r = ops.expand(a[0], a[1])
        elif op == "squeeze":
            x = _c(a[0])
            r = x.reshape(tuple(squeeze_shape(x.shape, a[1])))
        elif op == "where_op":
            r = ops.where(a[0], a[1], a[2])
        elif op == "concat":
            r = ops.concat(a[0], a[1])
        elif op == "split_take":
            key = (st["args"][0]["var"], a[1], tuple(a[2]))
            if key not in split_cache:
                split_cache[key] = ops.split(a[0], a[1], a[2])
            r = split_cache[key][a[3]]
        elif op == "identity":
            r = a[0]
        elif op == "reshape":
            r = _reshape(a[0], a[1])
        elif op == "flatten":
            x = _c(a[0]); ax = a[1] + x.ndim if a[1] < 0 else a[1]                 # shape.rs:109: negative axis counted from the end
            r = x.reshape((int(np.prod(x.shape[:ax], dtype=np.int64)), int(np.prod(x.shape[ax:], dtype=np.int64))))
        elif op == "unsqueeze":
            r = _c(a[0])
            r = r.reshape(tuple(unsqueeze_shape(r.shape, a[1])))
        elif op == "transpose":
            r = ops.transpose(a[0], a[1])
        elif op == "resize_nearest":
            scales = None if a[1] is None else [float(v) for v in np.asarray(_host(a[1])).reshape(-1)]
            r = ops.resize_nearest(a[0], scales, a[2], a[3])
        elif op == "matmul":
            r = ops.matmul(a[0], a[1])
        elif op == "softmax":
            r = ops.softmax(a[0], a[1])
        elif op == "max_pool2d":                     # (x, kernel, strides, pads, dilations, ceil_mode)
            r = ops.max_pool2d(a[0], a[1], a[3], a[2], a[4], a[5])
        elif op == "slice":
            r = ops.slice(a[0], a[1], a[2], a[3], a[4])
        elif op == "reduce_max":
            r = ops.reduce(a[0], a[1], a[2], "max")
        elif op == "topk":                            # (x, k, axis, largest, sorted): last axis, descending, stable (conv2d.rs:1385)
            if len(a) > 3 and a[3] is False:          # largest = false: ascending stable order = the descending stable order of -x
                v, i = ops.topk(ops.unary("neg", a[0]), a[1])
                r = (ops.unary("neg", v), i)
            else:
                r = ops.topk(a[0], a[1])
        elif op == "tile":
            r = ops.tile(a[0], a[1])
        elif op == "gather_elements":
            r = ops.gather_elements(a[0], a[1], a[2])
        elif op == "gather":
            r = ops.gather(a[0], a[1], a[2])
        else:
            raise ValueError(f"model.rs: operator {op} is not wired into run_program")
        if len(st["outs"]) == 1:
            env[st["outs"][0]] = r
        else:
            for n, v in zip(st["outs"], r):
                if n != "_":
                    env[n] = v
        if trace is not None:
            trace.append((st["outs"][0], op, env[st["outs"][0]]))

    if lift is None:
        exec_block(program["statements"])
        if resident:
            rctx.out_slots([])
        outs = [env[n] for n in program["outputs"]]
        return [_host(o) for o in outs] if download else outs

    # ---- batch-folded replay ----
    B, stmts, cut = int(lift), program["statements"], None
    folded_statements = 0
    for idx, st in enumerate(stmts):
        patched = _liftable(st, [val(x) for x in st["args"]] if st["op"] != "if" else None, lifted, env, B)
        if patched is None:
            cut = idx
            break
        if patched is False:                          # no folded operand: an ordinary (shared) statement
            exec_statement(st)
            continue
        exec_statement(patched)
        folded_statements += 1
        lifted.update(n for n in st["outs"] if n != "_" and _is_dev(env.get(n)))
    rctx.out_slots([])
    per_item = []
    if cut is None:
        cut = len(stmts)
    program.setdefault("_fold_report", {})[B] = {"folded_statements": folded_statements, "first_per_item_statement": cut, "statements": len(stmts),
                                                 "stopped_at": None if cut == len(stmts) else stmts[cut]["op"]}
    tail = dict(program, statements=stmts[cut:], inputs=[])
    lanes = []
    for ops_i, ws_i in (items or []):
        if ws_i.ctx is not rctx and ws_i.ctx not in lanes:
            lanes.append(ws_i.ctx)
    for c in lanes:
        rctx.fork(c)                                  # the per-item tails start when the folded prefix is done
    for i in range(B):
        ops_i, ws_i = items[i] if items else (ops, workspace)
        env_i = {n: (_item_view(v, i, B) if n in lifted and _is_dev(v) else v) for n, v in env.items()}
        per_item.append(run_program(tail, blob, [], ops_i, cache=cache, workspace=ws_i, download=False, env0=env_i))
    for c in lanes:
        rctx.join(c)
    if download:
        per_item = [[_host(o) for o in outs] for outs in per_item]
    return per_item


def synth_blob(program: dict, seed: int = 7, constants=None) -> bytes:
    """A weights.bin stand-in for a parsed program when the real blob is not available:
    convolution / matmul weights ~ N(0, 1/sqrt(fan_in)), biases ~ N(0, 0.1), every other f32 view ~ N(0, 1);
    `constants` = {offset: array} overrides (shape constants, anchors, k ...), written with the view's own dtype."""
    size, views = 0, {}

    def walk(statements):
        for st in statements:
            yield st
            for br in ("then", "else"):
                if br in st:
                    yield from walk(st[br]["statements"])
                    yield {"op": "tail", "args": list(st[br]["outputs"])}

    for st in walk(program["statements"]):
        flat = [(i, a) for i, a in enumerate(st["args"])] + [(i, b) for i, a in enumerate(st["args"]) if isinstance(a, dict) for b in (a.get("items", []) + ([a["i64vec_of"]] if "i64vec_of" in a else []))]
        for i, a in flat:
            if isinstance(a, dict):
                for key in ("weight", "weight_scalar", "weight_list"):
                    if key in a:
                        kind, off, ln, shape = a[key]
                        size = max(size, off + ln)
                        role = "w" if (st["op"] in ("conv2d", "conv2d_silu", "conv_transpose") and i == 1) else ("b" if st["op"].startswith("conv") and i == 2 else "x")
                        views[off] = (kind, ln, shape, role)
    blob = np.zeros(size, np.uint8)
    rng = np.random.default_rng(seed)
    for off in sorted(views):
        kind, ln, shape, role = views[off]
        src = np.dtype(_DTYPES[kind][0])
        n = ln // src.itemsize
        if constants is not None and off in constants:
            v = np.asarray(constants[off]).astype(src).reshape(-1)
        elif src.kind == "f":
            fan_in = int(np.prod(shape[1:])) if (role == "w" and len(shape) > 1) else 1
            sd = 1.0 / np.sqrt(fan_in) if role == "w" else (0.1 if role == "b" else 1.0)
            v = (rng.standard_normal(n) * sd).astype(src)
        else:
            v = np.ones(n, src)
        if v.size != n:
            raise ValueError(f"synth_blob: constant at offset {off} has {v.size} elements, the view holds {n}")
        blob[off:off + ln] = v.view(np.uint8)
    return blob.tobytes()


class GeneratedModel:
    """The Python stand-in for the struct `lele_gen` emits (`pub struct <Class><'a> { data: &'a [u8] }`, mod.rs:1100-1135):
    `GeneratedModel(open("yolo26seg.rs").read(), weights_bytes)` is `Yolo26Seg::new(&bin)`, `forward(*inputs)` is `forward`
    (inputs in the order of `forward_with_workspace`; one array back for a single-output graph, a tuple otherwise).
    `resident=True` (CUDA namespace): the replay keeps every value in HBM -- `forward_with_workspace` over a `Workspace` whose
    buffers are device mirrors (`lele_b200_arena_bind`); only the graph outputs are downloaded."""

    def __init__(self, model_rs, weights: bytes, ops=None, resident: bool = False):
        self.program = model_rs if isinstance(model_rs, dict) else parse_model_rs(model_rs)   # a parsed program (tests/golden/*.json) or the source text
        self.class_name = self.program["class"]
        need = _blob_extent(self.program)
        if len(weights) < need:
            raise ValueError(f"{self.class_name}: weights blob has {len(weights)} bytes, the generated code reads up to byte {need}")
        self.weights = weights
        self.ops = ops
        self.resident = resident
        self._cache = {}
        self._ws = None

    @classmethod
    def from_files(cls, rs_path: str, weights_path: str | None = None, ops=None, resident: bool = False):
        """`weights_path` defaults to `<stem>_weights.bin` next to the source, the name the compiler writes (mod.rs:1372)."""
        import os
        if weights_path is None:
            weights_path = os.path.splitext(rs_path)[0] + "_weights.bin"
        return cls(open(rs_path).read(), open(weights_path, "rb").read(), ops, resident)

    @property
    def input_names(self):
        return list(self.program["inputs"])

    @property
    def output_names(self):
        return list(self.program["outputs"])

    def workspace(self):
        """`<Model>Workspace::new()` (mod.rs:1066): the device arena of this model instance, created on first use."""
        if self._ws is None:
            from . import kernels as K
            if self.ops is None:
                self.ops = CudaOps()
            self._ws = K.Workspace(self.ops.ctx or K.default_context())
        return self._ws

    def forward(self, *inputs, trace=None):
        if len(inputs) != len(self.program["inputs"]):
            raise ValueError(f"{self.class_name}.forward takes {len(self.program['inputs'])} tensors ({', '.join(self.program['inputs'])})")
        if self.ops is None:
            self.ops = CudaOps()
        out = run_program(self.program, self.weights, list(inputs), self.ops, trace=trace, cache=self._cache,
                          workspace=self.workspace() if self.resident else None)
        return out[0] if len(out) == 1 else tuple(out)

    __call__ = forward

    def batch_runner(self, n_items: int, lanes: int = 8, ctx=None, graph: bool = True, fold: bool = False):
        return BatchRunner(self, n_items, lanes, ctx, graph, fold)


class BatchRunner:
    """`n_items` independent inputs through one generated model, resident.  Generated code bakes batch = 1 into its reshape /
    gather constants, so the batch lives below the boundary exactly as in the SenseVoice runner: item i is its own
    replay of the call list, with its own `Workspace` (its outputs stay valid until collected), on lane i % lanes -- a lane is a
    `Context` (stream, scratch) of the same device, so the short launches of different items overlap on the GPU.  The first
    `run` executes eagerly (sizes the arenas, uploads the weights once, shared by all lanes); the second captures the whole
    step -- every item, every lane -- into one CUDA graph (`lele_b200_capture_begin/_end`, lanes joined by `stream_fork/_join`)
    and every later `run` is: H2D of the inputs, one graph launch, D2H of the outputs."""

    def __init__(self, model: GeneratedModel, n_items: int, lanes: int = 8, ctx=None, graph: bool = True, fold: bool = False):
        from . import kernels as K
        self.K, self.model, self.n = K, model, int(n_items)
        self.fold = bool(fold)      # batch-folded replay (run_program lift=): the statements that are independent along the leading dimension
                                    # run ONCE on the stacked [n_items, ...] values; only the tail that is not runs per item on the lanes
        origin = ctx or K.default_context()
        self.lanes = [origin] + [K.Context(origin.device) for _ in range(max(1, min(int(lanes), self.n)) - 1)]
        for c in self.lanes[1:]:
            c._consts, c._keep = origin._consts, origin._keep          # one device copy of the weights for every lane
        self.ops = [CudaOps(c) for c in self.lanes]
        self.ws = [K.Workspace(self.lanes[i % len(self.lanes)]) for i in range(self.n)]
        self.ws_folded = K.Workspace(origin) if self.fold else None
        self._host_out = {}
        self.use_graph, self.graph, self.runs = graph, None, 0
        self.inputs_dev, self.outs_dev = None, None

    def _enqueue(self):
        L = len(self.lanes)
        origin = self.lanes[0]
        if self.fold:
            return run_program(self.model.program, self.model.weights, self.inputs_dev, self.ops[0], cache=self.model._cache,
                               workspace=self.ws_folded, download=False, lift=self.n, items=[(self.ops[i % L], self.ws[i]) for i in range(self.n)])
        for c in self.lanes[1:]:
            origin.fork(c)
        outs = []
        for i in range(self.n):
            outs.append(run_program(self.model.program, self.model.weights, self.inputs_dev[i], self.ops[i % L], cache=self.model._cache,
                                    workspace=self.ws[i], download=False))
        for c in self.lanes[1:]:
            origin.join(c)
        return outs

    def upload(self, items):
        """items[i] = list of host arrays (program input order).  Device staging buffers are allocated once; later calls copy into them
        on the origin stream (asynchronously when the host arrays are pinned)."""
        K, origin = self.K, self.lanes[0]
        if len(items) != self.n:
            raise ValueError(f"BatchRunner: {len(items)} items given, built for {self.n}")
        from ._lib import call, sz, vp
        if self.fold:                                                      # one stacked [n_items, ...] device tensor per program input
            if self.inputs_dev is None:
                self.inputs_dev = []
                for k in range(len(items[0])):
                    first = np.asarray(items[0][k])
                    if first.dtype.kind != "f":
                        self.inputs_dev.append(np.asarray(first, np.int64)); continue
                    if first.shape[0] != 1:
                        raise ValueError("BatchRunner(fold=True): every float input needs a leading dimension of 1 (the folded batch axis)")
                    self.inputs_dev.append(origin.to_device(np.zeros((self.n,) + first.shape[1:], np.float32)))
            for k, d in enumerate(self.inputs_dev):
                if not _is_dev(d):
                    continue
                per = d.nbytes // self.n
                for i, it in enumerate(items):
                    a = np.ascontiguousarray(it[k], np.float32)
                    if a.nbytes != per:
                        raise ValueError("BatchRunner: input shapes are fixed after the first run (the captured graph bakes them)")
                    call("lele_b200_h2d", origin.h, vp(d.ptr + i * per), a.ctypes.data_as(vp), sz(per))
            return
        if self.inputs_dev is None:
            self.inputs_dev = [[origin.to_device(np.asarray(a, np.float32)) if np.asarray(a).dtype.kind == "f" else np.asarray(a, np.int64) for a in it] for it in items]
            return
        for it, devs in zip(items, self.inputs_dev):
            for a, d in zip(it, devs):
                if _is_dev(d):
                    a = np.ascontiguousarray(a, np.float32)
                    if a.shape != d.shape:
                        raise ValueError("BatchRunner: input shapes are fixed after the first run (the captured graph bakes them)")
                    call("lele_b200_h2d", origin.h, vp(d.ptr), a.ctypes.data_as(vp), sz(a.nbytes))

    def launch(self):
        """Enqueues one step on the device (no host synchronisation once the graph exists)."""
        origin = self.lanes[0]
        self.runs += 1
        if self.graph is not None:
            self.graph.launch()
            return
        if self.use_graph and self.runs == 2:
            import gc
            l0 = [c.launch_count() for c in self.lanes[1:]]
            gc.collect()                                    # finalisers free device buffers, and a free joins the stream: none may run
            gc_was = gc.isenabled(); gc.disable()           # inside the capture (the library also defers such frees, csrc/ctx.cu)
            origin.capture_begin()
            try:
                self.outs_dev = self._enqueue()
            except Exception:
                try:
                    origin.capture_end(0).close()
                except Exception:
                    pass
                raise
            finally:
                if gc_was:
                    gc.enable()
            self.graph = origin.capture_end(sum(c.launch_count() - a for c, a in zip(self.lanes[1:], l0)))
            self.graph.launch()
            return
        self.outs_dev = self._enqueue()

    def collect(self):
        """Host copies of every item's graph outputs (`.to_owned()`): all device-to-host copies are enqueued on the origin stream
        (which the lanes were joined to), then the stream is joined once."""
        from ._lib import call, sz, vp
        origin = self.lanes[0]
        res = []
        for i, outs in enumerate(self.outs_dev):
            row = []
            for k, o in enumerate(outs):
                if _is_dev(o):
                    h = self._host_out.get((i, k))
                    if h is None or h.shape != tuple(o.shape):          # page-locked, allocated once: the copies run at PCIe rate
                        h = self._host_out[(i, k)] = origin.pinned_empty(o.shape, o.dtype)
                    if h.nbytes:
                        call("lele_b200_d2h", origin.h, h.ctypes.data_as(vp), vp(o.ptr), sz(h.nbytes))
                    row.append(h)
                else:
                    row.append(o)
            res.append(row)
        origin.sync()
        return res                                                       # (views of the runner's page-locked buffers: valid until the next collect)

    def run(self, items):
        self.upload(items)
        self.launch()
        return self.collect()

    def launch_count(self) -> int:
        return sum(c.launch_count() for c in self.lanes)

    def close(self):
        if self.graph is not None:
            self.graph.close(); self.graph = None
        self.lanes[0].sync()
        for w in self.ws + ([self.ws_folded] if self.ws_folded is not None else []):
            w.release()
        for c in self.lanes[1:]:
            c._consts, c._keep = {}, []
            c.close()


def _blob_extent(program: dict) -> int:
    """Highest byte any weight literal of the program touches."""
    end = 0

    def visit(a):
        nonlocal end
        if isinstance(a, dict):
            for key in ("weight", "weight_scalar", "weight_list"):
                if key in a:
                    end = max(end, a[key][1] + a[key][2])
            for b in a.get("items", []):
                visit(b)
            if "i64vec_of" in a:
                visit(a["i64vec_of"])

    def walk(statements):
        for st in statements:
            for a in st["args"]:
                visit(a)
            for br in ("then", "else"):
                if br in st:
                    walk(st[br]["statements"])
                    for o in st[br]["outputs"]:
                        visit(o)

    walk(program["statements"])
    return end
