"""Per-operator device time of the batch-folded Yolo26n-seg step (eager replay, CUDA events around every CudaOps call)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
sys.argv = ["bench.py"]
import bench
from lele_b200 import Context, model_rs as MR

B = int(os.environ.get("B", "32"))
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = Context(0, stream.cuda_stream)
prog, blob = bench._yolo_program()
rng = np.random.default_rng(7)
items = [[rng.random((1, 3, 640, 640), dtype=np.float32)] for _ in range(B)]
model = MR.GeneratedModel(prog, blob, ops=MR.CudaOps(ctx), resident=True)
br = model.batch_runner(B, lanes=8, ctx=ctx, fold=True, graph=False)
br.run(items); br.run(items)
ev = []
ops = br.ops[0]
for name in dir(MR.CudaOps):
    if name.startswith("_") or name in ("prepare_weights",):
        continue
    real = getattr(ops, name)
    if not callable(real):
        continue
    def mk(real, name):
        def f(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); r = real(*a); e1.record(stream)
            tag = name if name not in ("binary", "unary") else f"{name}:{a[0]}"
            shp = next((tuple(x.shape) for x in a if hasattr(x, "shape") and hasattr(x, "ptr")), None)
            ev.append((tag, shp, e0, e1)); return r
        return f
    setattr(ops, name, mk(real, name))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(stream); br.launch(); t1.record(stream); torch.cuda.synchronize()
tot = {}
rows = []
for tag, shp, a, b in ev:
    ms = a.elapsed_time(b); tot.setdefault(tag, [0, 0.0]); tot[tag][0] += 1; tot[tag][1] += ms; rows.append((ms, tag, shp))
print("step ms", t0.elapsed_time(t1))
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:24s} {n:4d} calls {ms:8.3f} ms")
print("top single calls:")
for ms, tag, shp in sorted(rows, key=lambda r: -r[0])[:25]:
    print(f"  {ms:7.3f} ms {tag} {shp}")
