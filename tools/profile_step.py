"""Profiling driver (run under ncu on the GPU box): one warm forward, then one forward inside the
NVTX range "profiled" of the BASELINE config (64 clips x 16 s, 70 layers).  Not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lele_b200 import Context, SenseVoice
from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, synth_batch

B = int(os.environ.get("PROF_CLIPS", "64"))
cfg = SenseVoiceConfig()
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = Context(0, st.cuda_stream)
m = SenseVoice(build_blob(cfg, seed=1234), max_clips=B, max_samples=256000, ctx=ctx)
pcm = torch.from_numpy(synth_batch(0, B)).cuda()
ids = torch.empty((B, m.rows(256000)), dtype=torch.int32, device="cuda")
m.forward_pcm_dev(pcm.data_ptr(), B, 256000, ids.data_ptr())
torch.cuda.synchronize()
n0 = ctx.launch_count()
torch.cuda.nvtx.range_push("profiled")
m.forward_pcm_dev(pcm.data_ptr(), B, 256000, ids.data_ptr())
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("launches in profiled forward:", ctx.launch_count() - n0)
