import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# GAS_EMU=1: run the `gpu` tests without a GPU, against tests/emu/_build/libgas_b200_emu.so — the product's .cu sources compiled by
# g++ and executed on the CPU (tests/emu/README.md).  Test infrastructure only: the binding is pointed at the emulation library
# here, in the test harness; the product (godot-audio-spatializer_b200/lib.py) knows nothing about it and has no fallback.
EMULATED = os.environ.get("GAS_EMU") == "1"

# tests that need real hardware even so: peers over NVLink / CUDA IPC, the "no device" error path
_NOT_EMULATED = ("test_comm_gpu.py", "test_shim_ab_gpu.py", "test_host_cpp.py")


def _install_emulation():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    import torch_shim
    lib_path = build_emu.build()
    import gaspkg
    pkg = gaspkg.load()
    from godot_audio_spatializer_b200 import lib as gas_lib
    gas_lib.LIB_PATH = lib_path
    gas_lib._lib = None
    torch_shim.install()
    return pkg


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    if EMULATED:
        _install_emulation()


def pytest_collection_modifyitems(config, items):
    if not EMULATED:
        return
    skip = pytest.mark.skip(reason="needs real hardware (not emulated)")
    for it in items:
        if os.path.basename(str(it.fspath)) in _NOT_EMULATED:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gas():
    import gaspkg
    return gaspkg.load()


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.load()
    return _orc
