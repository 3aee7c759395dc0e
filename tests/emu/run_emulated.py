"""TEST INFRASTRUCTURE: runs a Python script of this repository (bench.py, tools/...) against the CPU emulation of the library.

    python tests/emu/run_emulated.py bench.py --voices 512 --frames 128 --steps 16 --warmup 8 --no-configs

The binding is pointed at tests/emu/_build/libgas_b200_emu.so, CUDA devices of torch are redirected to the CPU (torch_shim) and
child processes the script starts with sys.executable go through this launcher too.  Nothing it prints is a measurement.
"""
import os
import runpy
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def install():
    # under torchrun: one emulated device per local rank, device allocations shareable between the rank processes
    lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    if lws > 1:
        os.environ.setdefault("GAS_EMU_DEVICES", str(lws))
        os.environ.setdefault("GAS_EMU_IPC", "1")
    for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import build_emu
    import torch_shim
    lib_path = build_emu.build()
    import gaspkg
    gaspkg.load()
    from godot_audio_spatializer_b200 import lib as gas_lib
    gas_lib.LIB_PATH = lib_path
    gas_lib._lib = None
    torch_shim.install()
    real_popen = subprocess.Popen

    class Popen(real_popen):  # children started with sys.executable run emulated too
        def __init__(self, cmd, *a, **k):
            if isinstance(cmd, (list, tuple)) and len(cmd) >= 2 and cmd[0] == sys.executable and str(cmd[1]).endswith(".py"):
                cmd = [sys.executable, os.path.abspath(__file__)] + list(cmd[1:])
            super().__init__(cmd, *a, **k)

    subprocess.Popen = Popen


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
