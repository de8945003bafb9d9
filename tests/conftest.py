import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # A clean checkout has no binaries (they are git-ignored): build the CUDA library and the CPU oracle
    # in-tree once, exactly as __graft_entry__.build() does.  Both are no-ops when up to date.
    # (the build module is loaded by path: importing the package itself needs the library it builds)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_asora_build", os.path.join(ROOT, "pyc2ray_b200", "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_native()
    import oracle
    oracle.build()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
