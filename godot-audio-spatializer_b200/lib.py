"""ctypes binding of libgas_b200.so (the C ABI declared in include/gas.h).

The library is the product: if it is missing or was built from a different header this module raises —
there is no Python, PyTorch or CPU fallback for any entry point.
"""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgas_b200.so")

_vp, _i32, _f32, _sz, _u64 = C.c_void_p, C.c_int32, C.c_float, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); every symbol include/gas.h declares
PROTOTYPES = {
    "gas_abi_version": (C.c_int, []),
    "gas_abi_sizeof": (_sz, [_i32]),
    "gas_config_defaults": (None, [_vp]),
    "gas_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "gas_destroy": (None, [_vp]),
    "gas_last_error": (C.c_char_p, [_vp]),
    "gas_set_speaker_mode": (C.c_int, [_vp, _i32]),
    "gas_set_mix_rate": (C.c_int, [_vp, _f32]),
    "gas_set_global_panning_strength": (C.c_int, [_vp, _f32]),
    "gas_get_channel_count": (C.c_int, [_vp]),
    "gas_spatializer_defaults": (None, [_vp]),
    "gas_spatializer_set": (C.c_int, [_vp, _i32, _vp]),
    "gas_instance_init": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_instance_start": (C.c_int, [_vp, _i32, _vp]),
    "gas_instance_stop": (C.c_int, [_vp, _i32, _vp]),
    "gas_voice_init": (C.c_int, [_vp, _i32, _vp]),
    "gas_gain_compute": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp]),
    "gas_gain_compute_device": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp]),
    "gas_listeners_set": (C.c_int, [_vp, _i32, _vp]),
    "gas_areas_set": (C.c_int, [_vp, _i32, _vp]),
    "gas_capture_begin": (C.c_int, [_vp]),
    "gas_capture_end": (C.c_int, [_vp, C.POINTER(_i32)]),
    "gas_graph_launch": (C.c_int, [_vp, _i32]),
    "gas_graph_destroy": (C.c_int, [_vp, _i32]),
    "gas_profile_enable": (C.c_int, [_vp, _i32]),
    "gas_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "gas_params_set": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_params_get": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_effect_params_set": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_mix_block": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "gas_mix_block_device": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "gas_step_device": (C.c_int, [_vp, _vp, _i32, _vp]),
    "gas_step_join_device": (C.c_int, [_vp]),
    "gas_process_frames": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32]),
    "gas_mix_channel": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32]),
    "gas_mix_block_stream": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "gas_mix_block_stream_device": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "gas_bus_layout_set": (C.c_int, [_vp, _i32, _vp]),
    "gas_bus_graph_device": (C.c_int, [_vp, _vp, _i32]),
    "gas_bus_graph": (C.c_int, [_vp, _vp, _i32]),
    "gas_source_set": (C.c_int, [_vp, _i32, _vp, _i32, _f32, _i32]),
    "gas_voice_play": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "gas_resample_block_device": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i32, _vp]),
    "gas_mix_block_resident": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "gas_set_playback_disable_threshold_db": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_voice_life_export": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_voice_life_import": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_status_flags": (C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    "gas_sync": (C.c_int, [_vp]),
    "gas_mix_stream": (_vp, [_vp]),
    "gas_gain_stream": (_vp, [_vp]),
    "gas_kernel_launches": (_u64, [_vp]),
    "gas_voice_state_export": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_voice_state_import": (C.c_int, [_vp, _i32, _vp, _vp]),
    "gas_comm_export": (C.c_int, [_vp, _vp, _sz]),
    "gas_comm_open": (C.c_int, [_vp, _i32, _i32, _vp, _sz]),
    "gas_reduce_bus_device": (C.c_int, [_vp, _vp, _i32]),
    "gas_reduce_bus_begin_device": (C.c_int, [_vp, _vp, _i32]),
    "gas_reduce_bus_end_device": (C.c_int, [_vp, _vp, _i32]),
    "gas_reduce_bus_exchange_device": (C.c_int, [_vp, _vp, _vp, _i32]),
    "gas_comm_close": (C.c_int, [_vp]),
}

_lib = None


class GasError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"gas status {status}: {message}")
        self.status = status


def load():
    """Load libgas_b200.so, bind every exported symbol and verify the record layouts."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C godot-audio-spatializer_b200/csrc`). There is no fallback implementation.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gas_abi_version() != 3:
        raise ImportError(f"libgas_b200.so has ABI version {lib.gas_abi_version()}, this binding expects 3")
    abi.check_layout(lib.gas_abi_sizeof, "libgas_b200.so")
    _lib = lib
    return lib


def check(status, ctx=None):
    if status != 0:
        msg = load().gas_last_error(ctx)
        raise GasError(status, msg.decode() if msg else "")
