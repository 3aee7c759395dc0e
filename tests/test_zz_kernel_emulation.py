"""The SOURCE of the two kernels that were finished after the round's GPU budget was spent (csrc/gas_resample.cu, csrc/gas_bus.cu),
compiled by g++ against tests/emu/cuda_emu.h and executed on the CPU — one std::thread per CUDA thread, one block at a time — and
compared bit for bit with the oracle.  Not a substitute for the GPU tests of tests/test_zz_resample.py / test_zz_busgraph.py: it
checks the kernels' indexing, control flow, barriers and arithmetic, not the launch plumbing around them."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import scenarios as S

abi = S.abi
HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu():
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    so = os.path.join(EMU, "libkernels_emu.so")
    srcs = [os.path.join(EMU, "kernels_emu.cpp"), os.path.join(EMU, "cuda_emu.h"),
            os.path.join(S.ROOT, "godot-audio-spatializer_b200", "csrc", "gas_resample.cu"),
            os.path.join(S.ROOT, "godot-audio-spatializer_b200", "csrc", "gas_bus.cu")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in srcs):
        subprocess.check_call([gxx, "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I" + CUDA_INC, "-o", so, srcs[0], "-lpthread"])
    return C.CDLL(so)


def _clip(n, seed):
    rs = np.random.RandomState(seed)
    t = np.arange(n, dtype=np.float64)
    return (0.4 * np.sin(2 * np.pi * (110.0 + 30.0 * seed) * t / 44100.0)[:, None] + 0.1 * rs.randn(n, 2)).astype(np.float32)


@pytest.mark.parametrize("loop", [False, True])
def test_resampler_kernel_source_matches_the_literal_loop(emu, orc, loop):
    F, V, blocks = 512, 12, 4
    rng = np.random.RandomState(7)
    for case in range(4):
        n = 700 + int(rng.randint(0, 2500))
        rate = [44100.0, 48000.0, 22050.0, 96000.0][case]
        pcm = _clip(n, case)
        start = rng.randint(0, n - 128, size=V).astype(np.int32)
        pos = np.zeros(V, dtype=np.uint64)
        want = [orc.Resampler(pcm, rate, loop=loop, start_frame=int(s_)) for s_ in start]
        for b in range(blocks):
            pitch = rng.uniform(0.5, 2.0, size=V).astype(np.float32)
            rows = np.full((V, F, 2), 9.0, dtype=np.float32)
            mixed = np.full(V, -7, dtype=np.int32)
            emu.emu_resample(pcm.ctypes.data_as(C.c_void_p), n, int(loop), C.c_float(rate), C.c_float(48000.0), V, start.ctypes.data_as(C.c_void_p),
                             pitch.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), F, rows.ctypes.data_as(C.c_void_p),
                             mixed.ctypes.data_as(C.c_void_p))
            for i in range(V):
                w, nw = want[i].mix(F, float(pitch[i]), 48000.0)
                assert mixed[i] == nw, f"case {case} block {b} voice {i}: {mixed[i]} valid frames, literal loop {nw}"
                np.testing.assert_array_equal(rows[i], w, err_msg=f"case {case} block {b} voice {i}")
        for r in want:
            r.close()


def test_bus_graph_kernel_source_matches_the_oracle(emu, orc):
    rng = np.random.default_rng(5)
    for B, Cc, F in ((3, 1, 128), (6, 4, 512), (2, 3, 64)):
        bus = rng.standard_normal((B, Cc, F, 2)).astype(np.float32)
        lay = [dict(volume_db=float(-1.5 * b), mute=(b == 1 and B > 2), send=max(0, b - 2)) for b in range(B)]
        want = orc.bus_graph(bus, lay)
        # what gas_bus_layout_set hands the kernel: linear volumes after mute / solo, resolved sends
        # (expf(volume_db * 0.115...f), the same libm call on both sides: gas_api.cu gas_bus_layout_set / orc_db_to_linear_f)
        vol = np.array([0.0 if l_.get("mute") else orc.load().orc_db_to_linear_f(C.c_float(l_["volume_db"])) for l_ in lay], dtype=np.float32)
        send = np.array([l_["send"] if 0 <= l_["send"] < b else 0 for b, l_ in enumerate(lay)], dtype=np.int32)
        got = bus.copy()
        emu.emu_bus_graph(B, Cc, F, vol.ctypes.data_as(C.c_void_p), send.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(got, want)
