"""GPU replay of the generated-style statement forms in tests/model_forms.py over the C ABI (CudaOps), against the same replay on
the CPU oracle.  f32 bar: 1e-4 relative with a 1e-4 * max|ref| floor, as in tests/test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import model_forms as MF            # noqa: E402
from tests.test_gpu_parity import close        # noqa: E402


def test_math_statement_forms_replay():
    from lele_b200 import model_rs as MR
    prog, blob, x = MF.math_forms(MR)
    tg, tr = [], []
    got = MR.run_program(prog, blob, [x], MR.CudaOps(), trace=tg)
    ref = MR.run_program(prog, blob, [x], MF.R, trace=tr)
    for (n1, op1, a), (n2, op2, b) in zip(tg, tr):
        assert (n1, op1) == (n2, op2)
        close(a, b)
    for a, b in zip(got, ref):
        close(a, b)


def test_recurrent_statement_forms_replay():
    from lele_b200 import model_rs as MR
    prog, blob, x = MF.recurrent_forms(MR)
    got = MR.run_program(prog, blob, [x], MR.CudaOps())
    ref = MR.run_program(prog, blob, [x], MF.R)
    for a, b in zip(got, ref):
        close(a, b)
