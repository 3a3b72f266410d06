import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "medical-sam2_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _true_fp32_references():
    """the torch statements the kernels are checked against must run in real fp32 (cuDNN convs default to TF32)."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def pytest_configure(config):
    _true_fp32_references()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ------------------------------------------------------------------ measured parity errors (profiles/r2_parity_table.txt)
_PARITY = {}


def record_parity(tag, values):
    """tests call this with their measured max-abs errors; with MS2_PARITY_TABLE=<path> the session writes them there."""
    _PARITY[tag] = values


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("MS2_PARITY_TABLE")
    if not path or not _PARITY:
        return
    import json
    with open(path, "w") as fh:
        for tag in sorted(_PARITY):
            fh.write(json.dumps({"case": tag, **{k: (round(v, 7) if isinstance(v, float) else v)
                                                 for k, v in _PARITY[tag].items()}}) + "\n")
