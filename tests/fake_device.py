"""Host-memory stand-in for liblele_b200's entry points -- TEST INFRASTRUCTURE for the CPU box only.

The resident replay (lele_b200.kernels.DeviceTensor / Workspace / Context.out_slots, lele_b200.model_rs.run_program with a
workspace, BatchRunner) is host logic: which buffer a statement writes, which operands are uploaded once, what is downloaded.
That logic must be testable where there is no GPU (`-m "not gpu"`), so this module replaces `lele_b200._lib.call` with a
dispatcher in which "device memory" is host memory (ctypes buffers, arena keyed like lele_b200_arena_bind) and a handful of
compute entries are carried out by the oracle (tests may use oracle/; the product never does).  It records every call so tests
can assert on the traffic (no h2d / d2h between statements, arena growth keeps contents, ...).  The real library is what the
`-m gpu` tests and every product path run; nothing in lele_b200/ knows this file exists.
"""
import ctypes as C

import numpy as np

from oracle import reference_api as R


def _v(a):
    return a.value if hasattr(a, "value") else a


class FakeDevice:
    def __init__(self):
        self.mem = {}            # base address -> ctypes buffer
        self.arena = {}          # (ctx, host key) -> (address, nbytes)
        self.log = []            # (entry name, nbytes or None)
        self.capturing = False
        self.captured = []       # entries recorded while capturing (replayed by graph_launch)
        self.graphs = {}
        self.next_ctx = 1

    # -- memory helpers --
    def _alloc(self, n):
        buf = (C.c_char * max(int(n), 16))()
        self.mem[C.addressof(buf)] = buf
        return C.addressof(buf)

    def arr(self, ptr, shape, dtype=np.float32):
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        if n == 0:
            return np.zeros(shape, dtype)
        return np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(ptr)), dtype=dtype).reshape(shape)

    def _shape(self, p, rank):
        return [int(x) for x in p[:int(rank)]]

    # -- dispatcher --
    def call(self, name, *args):
        fn = getattr(self, "e_" + name[len("lele_b200_"):], None)
        if fn is None:
            raise NotImplementedError(f"fake device: {name} is not emulated")
        if self.capturing and name not in ("lele_b200_capture_end", "lele_b200_arena_bind", "lele_b200_stream_fork", "lele_b200_stream_join", "lele_b200_malloc_host"):
            if name in ("lele_b200_sync", "lele_b200_d2h", "lele_b200_free"):
                raise RuntimeError(f"fake device: {name} during graph capture (would invalidate a real capture)")
            self.captured.append((fn, args))
        self.log.append(name)
        return fn(*args)

    # -- context / memory --
    def e_ctx_create(self, device, stream, out):
        out._obj.value = self.next_ctx; self.next_ctx += 1

    def e_ctx_destroy(self, ctx): pass
    def e_sync(self, ctx): pass

    def e_malloc(self, ctx, nbytes, out):
        out._obj.value = self._alloc(_v(nbytes))

    def e_free(self, ctx, p):
        self.mem.pop(_v(p), None)

    def e_malloc_host(self, ctx, nbytes, out):
        out._obj.value = self._alloc(_v(nbytes))

    def e_free_host(self, ctx, p):
        self.mem.pop(_v(p), None)

    def e_h2d(self, ctx, dst, src, n):
        C.memmove(_v(dst), _v(src), _v(n))

    def e_d2h(self, ctx, dst, src, n):
        C.memmove(_v(dst), _v(src), _v(n))

    def e_d2d(self, ctx, dst, src, n):
        C.memmove(_v(dst), _v(src), _v(n))

    def e_arena_bind(self, ctx, key, nbytes, out):
        k = (_v(ctx), _v(key)); n = _v(nbytes)
        cur = self.arena.get(k)
        if cur is None or cur[1] < n:
            p = self._alloc(n)
            if cur is not None:
                C.memmove(p, cur[0], cur[1])                      # grow keeps contents (Vec::reserve)
                self.mem.pop(cur[0], None)
            self.arena[k] = cur = (p, n)
            self.log.append("arena_grow")
        out._obj.value = cur[0]

    def e_arena_release(self, ctx, key):
        cur = self.arena.pop((_v(ctx), _v(key)), None)
        if cur:
            self.mem.pop(cur[0], None)

    # -- streams / graphs --
    def e_stream_fork(self, ctx, lane): pass
    def e_stream_join(self, ctx, lane): pass

    def e_capture_begin(self, ctx):
        self.capturing, self.captured = True, []

    def e_capture_end(self, ctx, lane_launches, out):
        self.capturing = False
        gid = len(self.graphs) + 1
        self.graphs[gid] = list(self.captured)
        out._obj.value = gid

    def e_graph_launch(self, ctx, g):
        for fn, args in self.graphs[_v(g)]:
            fn(*args)

    def e_graph_destroy(self, ctx, g):
        self.graphs.pop(_v(g), None)

    # -- a few operators, carried out by the oracle --
    def e_binary(self, ctx, op, a, ash, ar, b, bsh, br, out):
        sa, sb = self._shape(ash, _v(ar)), self._shape(bsh, _v(br))
        x, y = self.arr(_v(a), sa), self.arr(_v(b), sb)
        r = [R.add, R.sub, R.mul, R.div, R.maximum][_v(op)](x, y)
        self.arr(_v(out), r.shape)[...] = r

    def e_unary(self, ctx, op, x, n, out):
        a = self.arr(_v(x), (_v(n),))
        r = {0: R.relu, 1: R.sigmoid, 3: R.silu}[_v(op)](a)
        self.arr(_v(out), r.shape)[...] = r

    def e_conv2d(self, ctx, x, w, bias, nb, ic, h, wd, oc, kh, kw, group, pads, strides, dils, act, out):
        nb, ic, h, wd, oc, kh, kw, group = (_v(t) for t in (nb, ic, h, wd, oc, kh, kw, group))
        X = self.arr(_v(x), (nb, ic, h, wd)); W = self.arr(_v(w), (oc, ic // group, kh, kw))
        B = None if not _v(bias) else self.arr(_v(bias), (oc,))
        r = R.conv2d(X, W, B, list(dils[:2]), group, list(pads[:4]), list(strides[:2]), _v(act))
        self.arr(_v(out), r.shape)[...] = r

    def e_concat(self, ctx, ptrs, axis_lens, n, outer, inner, out):
        n, outer, inner = _v(n), _v(outer), _v(inner)
        parts = [self.arr(_v(ptrs[i]), (outer, int(axis_lens[i]), inner)) for i in range(n)]
        r = np.concatenate(parts, axis=1)
        self.arr(_v(out), r.shape)[...] = r

    def e_strided_copy(self, ctx, x, off, oshape, istrides, rank, out):
        rank = _v(rank); shp = self._shape(oshape, rank); st = self._shape(istrides, rank)
        total = int(np.prod(shp, dtype=np.int64)) if shp else 1
        idx = np.zeros(shp, np.int64) + _v(off)
        for d in range(rank):
            view = [1] * rank; view[d] = shp[d]
            idx = idx + (np.arange(shp[d]) * st[d]).reshape(view)
        src = self.arr(_v(x), (int(idx.max()) + 1,)) if total else np.zeros(0, np.float32)
        self.arr(_v(out), shp)[...] = src[idx]

    def e_max_pool2d(self, ctx, x, nb, c, h, w, kh, kw, pads, strides, dils, ceil_mode, out):
        nb, c, h, w, kh, kw = (_v(t) for t in (nb, c, h, w, kh, kw))
        r = R.max_pool2d(self.arr(_v(x), (nb, c, h, w)), (kh, kw), list(pads[:4]), list(strides[:2]), list(dils[:2]), bool(_v(ceil_mode)))
        self.arr(_v(out), r.shape)[...] = r


def install(monkeypatch):
    """Routes every C-ABI call of lele_b200.kernels / model_rs through a FakeDevice for the duration of one test."""
    from lele_b200 import _lib, kernels, model_rs
    dev = FakeDevice()

    def call(name, *args):
        dev.call(name, *args)

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(kernels, "call", call)
    monkeypatch.setattr(kernels, "_default", None)
    fake_lib = type("FakeLib", (), {"lele_b200_launch_count": staticmethod(lambda h: 0), "lele_b200_ctx_destroy": staticmethod(lambda h: 0),
                                    "lele_b200_free_host": staticmethod(lambda h, p: 0)})()
    monkeypatch.setattr(kernels, "lib", fake_lib)
    return dev
