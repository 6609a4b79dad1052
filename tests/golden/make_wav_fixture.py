"""tests/golden/zh.wav is the reference's own audio fixture (fixtures/zh.wav: mono 16 kHz s16, 89 472 samples = 5.592 s), the clip
BASELINE.json configs[0] names ("Silero VAD on fixtures/zh.wav").  It is test DATA (not source): copied byte for byte so that the
config-1 plumbing tests can run on the GPU box, where /root/reference does not exist.

Run in the build container:  python tests/golden/make_wav_fixture.py
"""
import hashlib
import os
import shutil

SRC = "/root/reference/fixtures/zh.wav"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zh.wav")

if __name__ == "__main__":
    shutil.copyfile(SRC, DST)
    print(DST, os.path.getsize(DST), "bytes, sha256", hashlib.sha256(open(DST, "rb").read()).hexdigest()[:16])
