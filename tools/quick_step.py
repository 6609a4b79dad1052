"""Quick experiment loop (GPU box): a short SenseVoice-shaped stack (QS_LAYERS layers, 64 clips x 16 s),
graph-replay time per step and the per-class CUDA-event breakdown of one eager pass.  Not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lele_b200 import Context, SenseVoice
from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, synth_batch

L = int(os.environ.get("QS_LAYERS", "8"))
B = int(os.environ.get("QS_CLIPS", "64"))
cfg = SenseVoiceConfig(n_layers=L, n_stage1=max(1, L // 2), vocab=int(os.environ.get("QS_VOCAB", "512")))
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = Context(0, st.cuda_stream)
m = SenseVoice(build_blob(cfg, seed=1234), max_clips=B, max_samples=256000, ctx=ctx)
pcm = torch.from_numpy(synth_batch(0, B)).cuda()
ids = torch.empty((B, m.rows(256000)), dtype=torch.int32, device="cuda")
for _ in range(3):
    m.forward_pcm_dev(pcm.data_ptr(), B, 256000, ids.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 10
e0.record(st)
for _ in range(N):
    m.forward_pcm_dev(pcm.data_ptr(), B, 256000, ids.data_ptr())
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
m.set_profiling(True)
m.forward_pcm_dev(pcm.data_ptr(), B, 256000, ids.data_ptr())
prof = m.last_profile()
m.set_profiling(False)
fixed = sum(prof[k]["ms"] for k in ("frontend_fbank_lfr", "cmvn", "prompt_scale_pos") if k in prof)
print(f"QS layers={L} clips={B}: {ms:.3f} ms/step  (~{(ms - fixed) / L * 1000:.1f} us/layer incl. head; 70-layer estimate {fixed + (ms - fixed) / L * 70:.2f} ms)")
print("   ".join(f"{k}={v['ms'] * 1000 / max(v['calls'], 1):.1f}us x{v['calls']}" for k, v in prof.items() if v["calls"]))
print("ids checksum", int(ids.to(torch.int64).sum().item()))
