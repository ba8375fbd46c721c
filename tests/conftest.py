import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout at /root/reference")


# C3D_GPU_TESTS_ON_EMU=1: dry-run the -m gpu tests WITHOUT a GPU -- module-level DEV becomes "cpu" and every test body runs
# inside the CPU emulation of the kernels (tests/_emu.py).  It proves nothing about the hardware; it shakes out the tests'
# own Python (shapes, keyword names, tolerances against the oracle) before GPU minutes are spent on them:
#   C3D_GPU_TESTS_ON_EMU=1 python -m pytest tests -m gpu -q -k "optim or inference"
EMU_DRYRUN = os.environ.get("C3D_GPU_TESTS_ON_EMU", "0") == "1"


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            if EMU_DRYRUN:
                if hasattr(item.module, "DEV"):
                    item.module.DEV = "cpu"
            else:
                item.add_marker(skip_gpu)


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_protocol(item):
    if not (EMU_DRYRUN and "gpu" in item.keywords):
        yield
        return
    import torch
    import cips3d_b200
    from _emu import emulated
    mods = [m for m in (cips3d_b200.generator, cips3d_b200.discriminator, cips3d_b200.pigan) if hasattr(m, "_require_cuda")]
    saved = [m._require_cuda for m in mods]
    sync = torch.cuda.synchronize
    for m in mods:
        m._require_cuda = lambda *a, **k: None        # the modules refuse CPU tensors by design
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        with emulated(async_mode=0, sms=2):
            yield
    finally:
        torch.cuda.synchronize = sync
        for m, f in zip(mods, saved):
            m._require_cuda = f
