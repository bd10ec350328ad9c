import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "ant-quantization_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np

    class G:
        def __init__(self):
            self._c = {}
            self.manifest = json.load(open(os.path.join(GOLDEN, "manifest.json")))

        def __getitem__(self, name):
            if name not in self._c:
                self._c[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
            return self._c[name]

        def json(self, name):
            return json.load(open(os.path.join(GOLDEN, name + ".json")))
    return G()
