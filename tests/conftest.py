import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libdcref.so (the compiled reference)")


def _gpu_count():
    try:
        from clustering_b200 import lib
        return lib.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a machine without a CUDA device skips the GPU tests instead of failing them
    # (`-m gpu` on the B200 box runs them; libdcb200.so itself has no CPU fallback and fails loudly)
    if any("gpu" in item.keywords for item in items) and _gpu_count() == 0:
        skip = pytest.mark.skip(reason="no CUDA device (libdcb200.so has no CPU fallback)")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from _oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from _oracle import Ref, build_ref
    if build_ref() is None:
        pytest.skip("compiled reference (oracle/_ref/libdcref.so) not available here")
    r = Ref()
    r.set_threads(1)
    return r


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
        return cache[name]
    return load
