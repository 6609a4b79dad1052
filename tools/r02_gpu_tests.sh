#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r02_pytest.log
grep -E "passed|failed|PARITY" gpurun_out/r02_pytest.log | tail -5
