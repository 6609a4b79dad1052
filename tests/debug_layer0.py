"""Diagnostic (not a pytest): localise a runner-vs-oracle divergence inside layer 0 by comparing the
runner's workspace buffers with the same intermediates computed by oracle operators."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lele_b200 import SenseVoice
from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob
from oracle import reference_api as R

cfg = SenseVoiceConfig(n_layers=3, vocab=1000, n_stage1=2, max_t=128)
blob = build_blob(cfg, seed=7)
hdr = blob[:256].view(np.int32); nt = int(hdr[12])
table = blob[256:256 + 16 * nt].view(np.uint64).reshape(-1, 2)
def T_(i, dt=np.float32): o, n = int(table[i, 0]), int(table[i, 1]); return blob[o:o + n].view(dt)
G0 = 10
def L(l, w, dt=np.float32): return T_(G0 + l * 21 + w, dt)

rng = np.random.default_rng(0)
B, t = 2, 40
feats = (rng.standard_normal((B, t, 560)) * np.array([1.0, 2.5])[:, None, None]).astype(np.float32)
m = SenseVoice(blob, max_clips=4, max_samples=89472)
Tm = m.max_clips  # noqa
got = m.forward(feats, 3, 0, n_layers=1)
T = t + 4; d = 512; M = B * T
def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
def keyinv(k):
    k = k.astype(np.uint32); b = np.where(k & 0x80000000, k & 0x7fffffff, ~k).astype(np.uint32); return b.view(np.float32)
keys8 = m.workspace("keys", (3 * 4 + 1, B, 8, 2), np.uint32)   # sharded: [site][clip][slot][min,max]
keys = np.stack([keys8[..., 0].min(-1), keys8[..., 1].max(-1)], -1)
qkv_g = m.workspace("qkv", (M, 1536)); fsmn_g = m.workspace("fsmn", (M, 512)); att_g = m.workspace("att", (M, 512)); f1_g = m.workspace("f1", (M, 2048))
x0_g = m.workspace("x0", (M, 560))
for c in range(B):
    x0 = got_x0 = x0_g[c * T:(c + 1) * T]
    embed = T_(0).reshape(16, 560); pos = T_(1).reshape(-1, 560)
    x0r = np.concatenate([embed[[3, 1, 2, 0]], feats[c]], 0) * np.float32(np.sqrt(np.float32(512))) + pos[:T]
    print(f"clip {c}: x0 rel {rel(x0, x0r):.2e}")
    h = R.layer_norm(x0r, L(0, 0), L(0, 1), -1, 1e-5)
    print("  site0 keys (min,max)", keyinv(keys[0, c]), "oracle", h.min(), h.max())
    qkv = R.fused_quantized_linear(h[None], L(0, 2, np.uint8).reshape(560, 1536), L(0, 3), int(L(0, 5, np.uint8)[0]), L(0, 4))[0]
    print(f"  qkv rel {rel(qkv_g[c*T:(c+1)*T], qkv):.2e}")
    q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
    ft = R.conv1d(v.T[None], L(0, 6).reshape(512, 1, 11), None, (1,), 512, (5, 5), (1,))[0].T
    fs = ft + v
    print(f"  fsmn rel {rel(fsmn_g[c*T:(c+1)*T], fs):.2e}")
    qh = (q.reshape(T, 4, 128).transpose(1, 0, 2) * np.float32(1.0 / np.sqrt(np.float32(128)))).astype(np.float32)
    kh = k.reshape(T, 4, 128).transpose(1, 2, 0); vh = v.reshape(T, 4, 128).transpose(1, 0, 2)
    p = R.softmax(R.matmul(qh, kh)); o = R.matmul(p, vh).transpose(1, 0, 2).reshape(T, 512)
    print(f"  att rel {rel(att_g[c*T:(c+1)*T], o):.2e}")
    print("  site1 keys", keyinv(keys[1, c]), "oracle", o.min(), o.max())
    a = R.fused_quantized_linear(o[None], L(0, 7, np.uint8).reshape(512, 512), L(0, 8), int(L(0, 10, np.uint8)[0]), L(0, 9))[0] + fs
    h2 = R.layer_norm(a, L(0, 11), L(0, 12), -1, 1e-5)
    print("  site2 keys", keyinv(keys[2, c]), "oracle", h2.min(), h2.max())
    f1 = R.fused_quantized_linear(h2[None], L(0, 13, np.uint8).reshape(512, 2048), L(0, 14), int(L(0, 16, np.uint8)[0]), L(0, 15), True)[0]
    print(f"  f1 rel {rel(f1_g[c*T:(c+1)*T], f1):.2e}")
    print("  site3 keys", keyinv(keys[3, c]), "oracle", f1.min(), f1.max())
    f2 = R.fused_quantized_linear(f1[None], L(0, 17, np.uint8).reshape(2048, 512), L(0, 18), int(L(0, 20, np.uint8)[0]), L(0, 19))[0]
    xr = a + f2
    print(f"  x(after layer0) rel {rel(got[c], xr):.2e}   vs oracle C network {rel(got[c], R.SenseVoiceRef(blob).forward(feats[c], 3, 0, n_layers=1)):.2e}")
