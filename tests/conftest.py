"""pytest configuration: markers, repo paths, and lazy builders for the oracle / CUDA libraries."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a CPU box: they are skipped unless -m gpu selects them
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_package():
    """import go-jpeg2000_b200/ (hyphenated directory) under the module name go_jpeg2000_b200"""
    import importlib.util
    name = "go_jpeg2000_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "go-jpeg2000_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def j2k():
    return load_package()


@pytest.fixture(scope="session")
def gpu_ctx(j2k):
    ctx = j2k.Context(0)
    yield ctx
    ctx.close()
