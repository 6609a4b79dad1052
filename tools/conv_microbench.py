"""Device time of lele_b200_conv2d (resident operands) over a few layer shapes: which term scales with what."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from lele_b200 import Context, kernels as K
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = Context(0, stream.cuda_stream)
rng = np.random.default_rng(0)
def run(nb, ic, hw, oc, k, stride, act, reps=20):
    x = ctx.to_device(rng.standard_normal((nb, ic, hw, hw)).astype(np.float32))
    w = ctx.to_device((rng.standard_normal((oc, ic, k, k)) / np.sqrt(ic * k * k)).astype(np.float32))
    b = ctx.to_device(rng.standard_normal(oc).astype(np.float32))
    ws = K.Workspace(ctx)
    p = k // 2
    def f():
        ctx.out_slots([(ws, "o")]); return K.conv2d(x, w, b, (1, 1), 1, (p, p, p, p), (stride, stride), act, ctx=ctx)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): y = f()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    oh = (hw + 2 * p - k) // stride + 1
    fl = 2.0 * nb * oc * ic * k * k * oh * oh
    by = (x.nbytes + nb * oc * oh * oh * 4)
    tiles = (nb * oh * oh + 127) // 128
    print(f"nb{nb} ic{ic} hw{hw} oc{oc} k{k} s{stride} act{act}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {by/ms/1e6:7.0f} GB/s  tiles {tiles} ({ms*1e3/ (tiles/148):.2f} us per tile-wave) chunks/tile {(ic*k*k+31)//32}")
    ws.release()
for cfg in [(32,48,160,64,1,1,2),(32,48,160,64,1,1,0),(32,48,160,16,1,1,2),(32,48,160,16,1,1,0),(32,64,160,64,3,1,2),(32,64,160,64,3,1,0),(32,16,160,64,3,1,0),
            (32,3,640,16,3,2,2),(32,3,640,16,3,2,0),(32,256,20,256,3,1,2),(32,128,40,128,3,1,2),(8,48,160,64,1,1,0),(32,32,160,32,1,1,0)]:
    run(*cfg)
