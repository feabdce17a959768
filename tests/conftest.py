import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run under gpurun)")
    # the suites drive built artefacts (libnezha_gpu.so, libnezha_host.so, nzproxy, the oracle):
    # build them once if a fresh checkout has none (nvcc cross-compiles without a GPU)
    need = [os.path.join(ROOT, "nezha_b200", "_build", n) for n in ("libnezha_gpu.so", "libnezha_host.so", "nzproxy", "nzwebhook", "nzload")]
    need += [os.path.join(ROOT, "oracle", "_build", n) for n in ("libnz_oracle.so", "nz_cpu_server")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
