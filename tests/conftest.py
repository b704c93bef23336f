import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build checker/tool libraries if missing (the product .so is built by __graft_entry__.build())."""
    from pdmp3_b200 import build
    if not os.path.exists(os.path.join(ROOT, "tools", "libp3synth.so")):
        build.build_tools()
    if not os.path.exists(os.path.join(ROOT, "oracle", "libp3_oracle.so")):
        build.build_oracle()
    yield


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import pdmp3_b200
    ctx = pdmp3_b200.Context(0)
    yield ctx
    ctx.close()
