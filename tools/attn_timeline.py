"""Diagnostic: prints the phase timeline (SM clock) of one attention CTA for the BASELINE geometry."""
import os, sys
os.environ["LELE_B200_ATTN_DBG"] = "1"; os.environ["LELE_B200_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lele_b200 import Context, SenseVoice
from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob, synth_batch
B = 64
cfg = SenseVoiceConfig(n_layers=2, n_stage1=1)
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = Context(0, st.cuda_stream)
m = SenseVoice(build_blob(cfg, seed=1), max_clips=B, max_samples=256000, ctx=ctx)
pcm = torch.from_numpy(synth_batch(0, B)).cuda()
ids = torch.empty((B, m.rows(256000)), dtype=torch.int32, device="cuda")
m.forward_pcm_dev(pcm.data_ptr(), B, 256000, ids.data_ptr())
torch.cuda.synchronize()
