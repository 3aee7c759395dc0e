"""TEST INFRASTRUCTURE: lets code written for `torch.device("cuda", i)` run against the CPU emulation of the library.

The emulated "device memory" is host memory, so a CPU tensor's data_ptr() is a valid device pointer for libgas_b200_emu.so.
install() redirects CUDA devices to the CPU and turns the handful of torch.cuda calls the tests and bench.py make into no-ops.
"""
import time

import torch

_installed = False


class _FakeEvent:
    def __init__(self, enable_timing=False, **_):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _FakeStream:
    def __init__(self, *a, **k):
        self.cuda_stream = 0

    def synchronize(self):
        pass

    def wait_event(self, ev):
        pass

    def wait_stream(self, st):
        pass

    def record_event(self, ev=None):
        ev = ev or _FakeEvent()
        ev.record()
        return ev

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install():
    global _installed
    if _installed:
        return
    _installed = True
    real_device = torch.device

    class _Meta(type):
        def __instancecheck__(cls, obj):
            return isinstance(obj, real_device)

        def __call__(cls, *a, **k):
            d = real_device(*a, **k)
            return real_device("cpu") if d.type == "cuda" else d

    class device(metaclass=_Meta):
        pass

    torch.device = device
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        # `.to(cuda_device)` on the real thing copies; keep that (device buffers must not alias the numpy arrays they came from)
        out = real_to(self, *a, **k)
        return out.clone() if out is self or out.data_ptr() == self.data_ptr() else out

    torch.Tensor.to = to
    real_generator = torch.Generator

    def generator(device=None):
        return real_generator()

    torch.Generator = generator
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    # factory functions called with device="cuda" / "cuda:0"
    def _cpu_device_kw(fn):
        def wrapped(*a, **k):
            d = k.get("device")
            if isinstance(d, str) and d.startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    for name in ("tensor", "zeros", "ones", "empty", "full", "rand", "randn", "arange", "zeros_like", "empty_like"):
        setattr(torch, name, _cpu_device_kw(getattr(torch, name)))
    # several emulated ranks: the process group runs over gloo
    import torch.distributed as dist
    real_init = dist.init_process_group

    def init_process_group(backend=None, *a, **k):
        k.pop("device_id", None)
        return real_init("gloo", *a, **k)

    dist.init_process_group = init_process_group
    c = torch.cuda
    c.is_available = lambda: True
    import os
    c.device_count = lambda: int(os.environ.get("GAS_EMU_DEVICES", "1"))
    c.synchronize = lambda *a, **k: None
    c.set_device = lambda *a, **k: None
    c.current_device = lambda: 0
    c.Event = _FakeEvent
    c.Stream = _FakeStream
    c.ExternalStream = _FakeStream
    c.current_stream = lambda *a, **k: _FakeStream()
    c.stream = lambda s: s
    c.empty_cache = lambda: None
    c.get_device_name = lambda *a, **k: "emulated sm_100 (tests/emu)"
    c.mem_get_info = lambda *a, **k: (8 << 30, 8 << 30)
