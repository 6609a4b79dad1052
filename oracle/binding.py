"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (lele_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _host_tag() -> str:
    """-march=native binaries are host specific: key the .so by the CPU's ISA flags so a box with a
    different CPU (the GPU box) rebuilds instead of executing foreign instructions."""
    import hashlib
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


_SO = os.path.join(_HERE, f"liblele_oracle_{_host_tag()}.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("lele_oracle.c", "sensevoice_ref.c", "lele_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", f"OUT={os.path.basename(_SO)}"])
    return _SO


_lib = None
F = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.lo_hz_to_mel_htk.restype = C.c_float
        _lib.lo_hz_to_mel_htk.argtypes = [C.c_float]
        _lib.lo_mel_to_hz_htk.restype = C.c_float
        _lib.lo_mel_to_hz_htk.argtypes = [C.c_float]
        _lib.lo_cephes_expf.restype = C.c_float
        _lib.lo_cephes_expf.argtypes = [C.c_float]
        _lib.lo_sv_create.restype = C.c_void_p
        _lib.lo_sv_create.argtypes = [C.c_void_p, C.c_size_t]
        _lib.lo_sv_destroy.argtypes = [C.c_void_p]
        _lib.lo_sv_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.lo_sv_pcm_to_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.lo_sv_vocab.argtypes = [C.c_void_p]
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ints(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


# ---- front-end ----
def hann_window(n):
    o = np.empty(n, np.float32); lib().lo_hann_window(C.c_int(n), _p(o)); return o


def rfft(x):
    x = _f(x); n = x.size
    re = np.empty(n // 2 + 1, np.float32); im = np.empty_like(re)
    lib().lo_rfft_forward(_p(x), C.c_int(n), _p(re), _p(im)); return re, im


def mel_filterbank(sr, n_fft, n_mels, f_min, f_max=None):
    f_max = sr / 2.0 if f_max is None else f_max
    w = np.empty((n_mels, n_fft // 2 + 1), np.float32)
    lib().lo_mel_filterbank(C.c_float(sr), C.c_int(n_fft), C.c_int(n_mels), C.c_float(f_min), C.c_float(f_max), _p(w))
    return w


def frontend(pcm, want_mel=False):
    pcm = _f(pcm); n = pcm.size
    frames = lib().lo_frontend_num_frames(C.c_int(n))
    if frames == 0:
        return (np.zeros((0, 80), np.float32), np.zeros((0, 560), np.float32)) if want_mel else np.zeros((0, 560), np.float32)
    t = (frames + 5) // 6
    mel = np.empty((frames, 80), np.float32); out = np.empty((t, 560), np.float32)
    lib().lo_frontend_compute(_p(pcm), C.c_int(n), _p(mel), _p(out))
    return (mel, out) if want_mel else out


def lfr(x, m=7, n=6):
    x = _f(x); t, d = x.shape
    out = np.empty(((t + n - 1) // n, d * m), np.float32)
    lib().lo_lfr(_p(x), C.c_int(t), C.c_int(d), C.c_int(m), C.c_int(n), _p(out)); return out


def cmvn(x, eps=1e-5):
    x = _f(x); t, d = x.shape; out = np.empty_like(x)
    lib().lo_cmvn(_p(x), C.c_int(t), C.c_int(d), C.c_float(eps), _p(out)); return out


def stft(sig, n_fft, hop, win, window=None, power=False):
    sig_in = sig
    sig = _f(sig).reshape(-1); L = sig.size
    if L == 0:   # math.rs:2313-2316 / :2381-2384: an empty signal gives an empty tensor, not one zero frame
        return np.zeros((0, 0, n_fft // 2 + 1) if power else (0, 0, n_fft // 2 + 1, 2), np.float32)
    frames = 1 if L < win else (L - win) // hop + 1
    nfr = n_fft // 2 + 1
    out = np.empty((frames, nfr) if power else (frames, nfr, 2), np.float32)
    w = None if window is None else _f(window)
    lib().lo_stft(_p(sig), C.c_int(L), C.c_int(n_fft), C.c_int(hop), C.c_int(win), _p(w), C.c_int(int(power)), _p(out))
    lead = np.shape(sig_in)[:-1]
    if len(lead) >= 1:   # math.rs:2362-2367: an input of rank >= 2 gets a leading batch dim; the data is still one flat signal
        if int(np.prod(lead)) != 1:
            raise ValueError("stft: batch > 1 gives a shape that does not match the data upstream (math.rs:2364)")
        out = out.reshape((1,) + out.shape)
    return out


# ---- quantised ----
def dynamic_quantize_linear(x):
    x = _f(x); q = np.empty_like(x); s = C.c_float(); z = C.c_float()
    lib().lo_dynamic_quantize_linear(_p(x), C.c_size_t(x.size), _p(q), C.byref(s), C.byref(z))
    return q, np.float32(s.value), np.float32(z.value)


def mat_mul_integer(a, b, a_zp=0.0, b_zp=0.0, scale=None, bias=None, relu=False):
    a = _f(a); b = _f(b)
    m, k = a.shape[-2:]; n = b.shape[-1]; batch = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    out = np.empty(a.shape[:-1] + (n,), np.float32)
    sc = None if scale is None else _f(scale).reshape(-1); bi = None if bias is None else _f(bias).reshape(-1)
    lib().lo_mat_mul_integer(_p(a), _p(b), C.c_int(batch), C.c_int(m), C.c_int(k), C.c_int(n), C.c_float(a_zp),
                             C.c_float(b_zp), _p(sc), C.c_int(0 if sc is None else sc.size), _p(bi), C.c_int(int(relu)), _p(out))
    return out


def fused_quantized_linear(x, w_u8, w_scale, w_zp, bias, relu=False):
    x = _f(x); w = np.ascontiguousarray(w_u8, dtype=np.uint8)
    m, k = x.shape[-2:]; n = w.shape[-1]; batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    ws = _f(w_scale).reshape(-1); bi = None if bias is None else _f(bias).reshape(-1)
    out = np.empty(x.shape[:-1] + (n,), np.float32)
    lib().lo_fused_quantized_linear(_p(x), C.c_int(batch), C.c_int(m), C.c_int(k), C.c_int(n), _p(w), _p(ws),
                                    C.c_int(ws.size), C.c_int(int(w_zp)), _p(bi), C.c_int(int(relu)), _p(out))
    return out


# ---- gemm ----
def matmul(a, b):
    a = _f(a); b = _f(b)
    if a.ndim < 2 or b.ndim < 2:
        raise ValueError("MatMul: both operands need rank >= 2")                               # gemm.rs:122-123
    m, k = a.shape[-2:]; n = b.shape[-1]
    if k != b.shape[-2]:
        raise ValueError(f"MatMul K dim mismatch: {k} vs {b.shape[-2]}")                        # gemm.rs:129
    ba = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    bb = int(np.prod(b.shape[:-2])) if b.ndim > 2 else 1
    if not (bb == 1 or bb == ba):
        raise ValueError("MatMul broadcast not fully supported yet")                           # gemm.rs:134
    lead = a.shape[:-2] if ba >= bb else b.shape[:-2]
    out = np.empty(tuple(lead) + (m, n), np.float32)
    lib().lo_matmul(_p(a), _p(b), C.c_int(ba), C.c_int(bb), C.c_int(m), C.c_int(k), C.c_int(n), _p(out))
    return out


def matmul_fused_add(a, b, bias):
    a = _f(a); b = _f(b); bias = _f(bias).reshape(-1)
    m, k = a.shape[-2:]; n = b.shape[-1]
    ba = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    bb = int(np.prod(b.shape[:-2])) if b.ndim > 2 else 1
    lead = a.shape[:-2] if ba >= bb else b.shape[:-2]
    out = np.empty(tuple(lead) + (m, n), np.float32)
    lib().lo_matmul_fused_add(_p(a), _p(b), _p(bias), C.c_int(bias.size), C.c_int(ba), C.c_int(bb), C.c_int(m), C.c_int(k), C.c_int(n), _p(out))
    return out


def gemm(a, b, c=None, alpha=1.0, beta=1.0, trans_a=False, trans_b=False):
    a = _f(a); b = _f(b)
    m = a.shape[-1] if trans_a else a.shape[-2]; k = a.shape[-2] if trans_a else a.shape[-1]
    n = b.shape[-2] if trans_b else b.shape[-1]
    cc = None if c is None else _f(c).reshape(-1)
    out = np.empty((m, n), np.float32)
    lib().lo_gemm(_p(a), _p(b), _p(cc), C.c_int(0 if cc is None else cc.size), C.c_float(alpha), C.c_float(beta),
                  C.c_int(int(trans_a)), C.c_int(int(trans_b)), C.c_int(m), C.c_int(k), C.c_int(n), _p(out))
    return out


# ---- norms / activations ----
def layer_norm(x, gamma, beta, axis=-1, eps=1e-5):
    x = _f(x); ax = axis % x.ndim
    n = int(np.prod(x.shape[ax:])); outer = x.size // n
    out = np.empty_like(x)
    lib().lo_layer_norm(_p(x), _p(_f(gamma).reshape(-1)), _p(_f(beta).reshape(-1)), C.c_int(outer), C.c_int(n), C.c_float(eps), _p(out))
    return out


def softmax(x):
    x = _f(x); n = x.shape[-1]; out = np.empty_like(x)
    lib().lo_softmax(_p(x), C.c_int(x.size // n), C.c_int(n), _p(out)); return out


def batch_norm(x, scale, bias, mean, var, eps=1e-5):
    x = _f(x); nb, c = x.shape[:2]; inner = x.size // (nb * c); out = np.empty_like(x)
    lib().lo_batch_norm(_p(x), _p(_f(scale)), _p(_f(bias)), _p(_f(mean)), _p(_f(var)), C.c_int(nb), C.c_int(c), C.c_int(inner), C.c_float(eps), _p(out))
    return out


def rms_norm(x, w, eps=1e-5):
    x = _f(x); n = x.shape[-1]; out = np.empty_like(x)
    lib().lo_rms_norm(_p(x), _p(_f(w)), C.c_int(x.size // n), C.c_int(n), C.c_float(eps), _p(out)); return out


UNARY = {"relu": 0, "sigmoid": 1, "tanh": 2, "silu": 3, "erf": 4, "gelu": 5, "exp": 6, "softplus": 7}


def unary(name, x):
    x = _f(x); out = np.empty_like(x)
    lib().lo_unary(C.c_int(UNARY[name]), _p(x), C.c_size_t(x.size), _p(out)); return out


# ---- conv ----
def conv1d(x, w, bias=None, dilations=(1,), group=1, pads=(0, 0), strides=(1,), relu=False):
    x = _f(x); w = _f(w)
    if x.ndim == 2:
        x = x[:, None, :]
    nb, ic, l = x.shape; oc, _, k = w.shape
    dil = dilations[0] if len(dilations) else 1; st = strides[0] if len(strides) else 1
    pl = pads[0] if len(pads) else 0; pr = pads[1] if len(pads) > 1 else 0
    ol = (l + pl + pr - dil * (k - 1) - 1) // st + 1
    out = np.empty((nb, oc, ol), np.float32)
    bi = None if bias is None else _f(bias)
    lib().lo_conv1d(_p(x), _p(w), _p(bi), C.c_int(nb), C.c_int(ic), C.c_int(l), C.c_int(oc), C.c_int(k), C.c_int(group),
                    C.c_int(pl), C.c_int(pr), C.c_int(st), C.c_int(dil), C.c_int(int(relu)), _p(out))
    return out


def _pads4(pads):
    p = list(pads)
    if len(p) == 0: p = [0, 0, 0, 0]
    if len(p) == 2: p = [p[0], p[1], p[0], p[1]]
    return p


def conv2d(x, w, bias=None, dilations=(1, 1), group=1, pads=(0, 0, 0, 0), strides=(1, 1), act=0):
    x = _f(x); w = _f(w); nb, ic, h, wd = x.shape; oc, _, kh, kw = w.shape
    p = _ints(_pads4(pads)); s = _ints(strides or (1, 1)); d = _ints(dilations or (1, 1))
    oh = C.c_int(); ow = C.c_int()
    lib().lo_conv2d(_p(x), _p(w), None, C.c_int(nb), C.c_int(ic), C.c_int(h), C.c_int(wd), C.c_int(oc), C.c_int(kh), C.c_int(kw),
                    C.c_int(group), p, s, d, C.c_int(act), None, C.byref(oh), C.byref(ow))
    if oh.value <= 0 or ow.value <= 0 or h + p[0] + p[2] - d[0] * (kh - 1) - 1 < 0 or wd + p[1] + p[3] - d[1] * (kw - 1) - 1 < 0:
        raise ValueError(f"conv2d: output dimensions must be positive, got out_h={oh.value} out_w={ow.value}")   # conv2d.rs:274-291
    out = np.empty((nb, oc, oh.value, ow.value), np.float32)
    bi = None if bias is None else _f(bias)
    lib().lo_conv2d(_p(x), _p(w), _p(bi), C.c_int(nb), C.c_int(ic), C.c_int(h), C.c_int(wd), C.c_int(oc), C.c_int(kh), C.c_int(kw),
                    C.c_int(group), p, s, d, C.c_int(act), _p(out), C.byref(oh), C.byref(ow))
    return out


def conv_transpose(x, w, bias=None, dilations=(1, 1), pads=(0, 0, 0, 0), strides=(1, 1)):
    x = _f(x); w = _f(w); nb, ic, h, wd = x.shape; _, oc, kh, kw = w.shape
    p = _ints(_pads4(pads)); s = _ints(strides or (1, 1)); d = _ints(dilations or (1, 1))
    oh = C.c_int(); ow = C.c_int()
    lib().lo_conv_transpose(_p(x), _p(w), None, C.c_int(nb), C.c_int(ic), C.c_int(h), C.c_int(wd), C.c_int(oc), C.c_int(kh), C.c_int(kw),
                            p, s, d, None, C.byref(oh), C.byref(ow))
    if oh.value <= 0 or ow.value <= 0 or h + p[0] + p[2] - d[0] * (kh - 1) - 1 < 0 or wd + p[1] + p[3] - d[1] * (kw - 1) - 1 < 0:
        raise ValueError(f"conv2d: output dimensions must be positive, got out_h={oh.value} out_w={ow.value}")   # conv2d.rs:274-291
    out = np.empty((nb, oc, oh.value, ow.value), np.float32)
    bi = None if bias is None else _f(bias)
    lib().lo_conv_transpose(_p(x), _p(w), _p(bi), C.c_int(nb), C.c_int(ic), C.c_int(h), C.c_int(wd), C.c_int(oc), C.c_int(kh), C.c_int(kw),
                            p, s, d, _p(out), C.byref(oh), C.byref(ow))
    return out


def max_pool2d(x, kernel, pads=(0, 0, 0, 0), strides=(1, 1), dilations=(1, 1), ceil_mode=False):
    x = _f(x); nb, c, h, w = x.shape
    p = _ints(_pads4(pads)); s = _ints(strides); d = _ints(dilations)
    oh = C.c_int(); ow = C.c_int()
    lib().lo_max_pool2d(_p(x), C.c_int(nb), C.c_int(c), C.c_int(h), C.c_int(w), C.c_int(kernel[0]), C.c_int(kernel[1]), p, s, d,
                        C.c_int(int(ceil_mode)), None, C.byref(oh), C.byref(ow))
    out = np.empty((nb, c, oh.value, ow.value), np.float32)
    lib().lo_max_pool2d(_p(x), C.c_int(nb), C.c_int(c), C.c_int(h), C.c_int(w), C.c_int(kernel[0]), C.c_int(kernel[1]), p, s, d,
                        C.c_int(int(ceil_mode)), _p(out), C.byref(oh), C.byref(ow))
    return out


# ---- rnn ----
def lstm(x, w, r, bias=None, h0=None, c0=None):
    x = _f(x); w = _f(w); r = _f(r)
    seq, _, isz = x.shape; hid = w.shape[1] // 4
    y = np.empty((seq, 1, 1, hid), np.float32); h = np.empty((1, 1, hid), np.float32); c = np.empty((1, 1, hid), np.float32)
    lib().lo_lstm(_p(x), _p(w), _p(r), _p(None if bias is None else _f(bias)), _p(None if h0 is None else _f(h0)),
                  _p(None if c0 is None else _f(c0)), C.c_int(seq), C.c_int(isz), C.c_int(hid), _p(y), _p(h), _p(c))
    return y, h, c


def gru(x, w, r, bias=None, h0=None):
    x = _f(x); w = _f(w); r = _f(r)
    seq, _, isz = x.shape; hid = w.shape[1] // 3
    y = np.empty((seq, 1, 1, hid), np.float32); h = np.empty((1, 1, hid), np.float32)
    lib().lo_gru(_p(x), _p(w), _p(r), _p(None if bias is None else _f(bias)), _p(None if h0 is None else _f(h0)),
                 C.c_int(seq), C.c_int(isz), C.c_int(hid), _p(y), _p(h))
    return y, h


# ---- SenseVoice-shaped network ----
class SenseVoiceRef:
    def __init__(self, blob: np.ndarray):
        self.blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self.h = lib().lo_sv_create(_p(self.blob), C.c_size_t(self.blob.size))
        if not self.h:
            raise ValueError("bad SenseVoice blob")
        self.vocab = lib().lo_sv_vocab(C.c_void_p(self.h))
        hdr = self.blob[:256].view(np.int32)
        self.d_model = int(hdr[3])

    def forward(self, feats, lang=3, textnorm=0, n_layers=-1):
        feats = _f(feats); t = feats.shape[0]
        hdr = self.blob[:256].view(np.int32)
        full = n_layers < 0 or n_layers >= int(hdr[2])
        width = self.vocab if full else (int(hdr[4]) if n_layers == 0 else self.d_model)
        out = np.empty((t + 4, width), np.float32)
        lib().lo_sv_forward(C.c_void_p(self.h), _p(feats), C.c_int(t), C.c_int(lang), C.c_int(textnorm), C.c_int(n_layers), _p(out))
        return out

    def pcm_to_ids(self, pcm, lang=3, textnorm=0, want_logits=False):
        pcm = _f(pcm); frames = (pcm.size - 400) // 160 + 1; T = (frames + 5) // 6 + 4
        ids = np.empty(T, np.int32)
        logits = np.empty((T, self.vocab), np.float32) if want_logits else None
        lib().lo_sv_pcm_to_ids(C.c_void_p(self.h), _p(pcm), C.c_int(pcm.size), C.c_int(lang), C.c_int(textnorm), _p(ids), _p(logits))
        return (ids, logits) if want_logits else ids

    def __del__(self):
        try:
            lib().lo_sv_destroy(C.c_void_p(self.h))
        except Exception:
            pass
