import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _bootstrap_native():
    """The lele_b200 package refuses to import without liblele_b200.so (no CPU fallback), so a fresh
    checkout builds it (nvcc cross-compiles without a GPU) before any test module imports the package.
    build.py is loaded by file path for the same reason."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_lele_b200_build", os.path.join(ROOT, "lele_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


SO_PATH = _bootstrap_native()
