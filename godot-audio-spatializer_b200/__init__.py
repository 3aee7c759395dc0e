"""godot-audio-spatializer_b200 — B200-native batched spatial-audio mixer.

The hot path of BuzzLord/godot-audio-spatializer (per-instance gain computation, volume-ramped /
filtered mixing of AudioFrame buffers into bus channels) as hand-written sm_100a CUDA behind a C ABI
(include/gas.h).  This package holds the CUDA sources (csrc/), the C++ host mirror of the reference
API (host/) and the ctypes binding used by tests and bench.py.

The directory name carries a hyphen, so import it through ``gaspkg.load()`` at the repo root, which
registers it as ``godot_audio_spatializer_b200``.
"""
from . import abi, shard, synth  # noqa: F401
from .lib import GasError, LIB_PATH, PROTOTYPES, load  # noqa: F401
from .mixer import Mixer  # noqa: F401
