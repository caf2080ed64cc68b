import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle_sht():
    from oracle import ref_sht
    ref_sht.build()
    return ref_sht


def _cuda_ready():
    """a CUDA device and the built product library: what every `gpu`-marked test needs"""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as ex:   # pragma: no cover
        return False, "torch unavailable: %r" % (ex,)
    so = os.path.join(ROOT, 'plancklens_b200', 'csrc', 'libplk_b200.so')
    if not os.path.exists(so):
        return False, "libplk_b200.so is not built"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only host: GPU tests are reported as skipped, not as 100+ PlkError failures.
    (With a device present nothing is skipped: a missing library then fails loudly, as the product does.)"""
    ok, why = _cuda_ready()
    if ok:
        return
    try:
        import torch
        has_dev = torch.cuda.is_available()
    except Exception:
        has_dev = False
    if has_dev:
        return
    skip = pytest.mark.skip(reason="needs a B200: " + why)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
