"""lele_b200 -- Blackwell (sm_100a) back-end for lele's AOT operator execution path.

Package layout (only what the hot path needs):
  csrc/                 hand-written CUDA kernels + the C ABI (include/lele_b200.h)
  kernels.py            host-side mirror of `lele::kernels::*`
  features.py           host-side mirror of `lele::features::*`
  sensevoice.py         model object over the SenseVoice-shaped graph runner
  sensevoice_weights.py synthetic weights blob + synthetic PCM (numpy only, no CUDA)
  tokenizer.py          examples/sensevoice tokenizer + on-device greedy-decode filter
  model_rs.py           parses a lele_gen-generated model.rs and replays it over the C ABI
  e2e.py                the reference's e2e golden checks (examples/*/tests/e2e_test.rs) over the replay; SKIP while files are missing
  vad.py                streaming VAD caller (examples/silero): chunk loop with carried state, segment logic

Importing the package loads liblele_b200.so and raises if it is missing: there is no CPU
fallback.  (`lele_b200/build.py` is run as a script / loaded by path, so it works before the .so exists.)
"""
from ._lib import LeleB200Error, SO_PATH, lib  # noqa: F401  (fails loudly when the .so is absent)
from . import features, kernels  # noqa: F401
from .sensevoice import SenseVoice  # noqa: F401
from .kernels import Context, default_context  # noqa: F401
from .tokenizer import Tokenizer  # noqa: F401
from .vad import StreamingVad, VadConfig  # noqa: F401
