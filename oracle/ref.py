"""ctypes loader of oracle/_ref/libgas_ref.so: the reference module's OWN sources (/root/reference/*.cpp,
unmodified) compiled against the godot-lite stand-in headers (oracle/godot_lite/) and driven by
oracle/ref_harness.cpp.  Same Python surface as OracleMixer / the product's Mixer, so one scenario driver
plays all three.

TEST INFRASTRUCTURE ONLY: imported by tests/ (to pin the restated oracle against reference code) and by
tests/golden/make_golden.py; never by the product package.  The library can only be BUILT where
/root/reference exists (this container); the built .so is git-ignored and travels to the GPU box with gpurun.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gaspkg  # noqa: E402

abi = gaspkg.load().abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgas_ref.so")
LIB_PATH_GPU = os.path.join(_HERE, "_ref", "libgas_ref_gpu.so")  # + integration/godot_module (the GPU-backed spatializer classes)
REFERENCE_DIR = os.environ.get("GAS_REFERENCE_DIR", "/root/reference")
_lib = None
_lib_gpu = None
_vp, _i32, _f32 = C.c_void_p, C.c_int32, C.c_float


def sources_present():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "audio_spatializer_3d.cpp"))


def available():
    return os.path.exists(LIB_PATH) or sources_present()


def build(force=False):
    """(Re)build _ref when the reference tree is present; otherwise use the prebuilt library as is."""
    if not sources_present():
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise FileNotFoundError("oracle/_ref/libgas_ref.so is not built and the reference tree is not present")
    deps = [os.path.join(_HERE, "ref_harness.cpp"), os.path.join(os.path.dirname(_HERE), "include", "gas.h")]
    gl = os.path.join(_HERE, "godot_lite")
    deps += [os.path.join(gl, f) for f in os.listdir(gl) if f.endswith(".h")]
    deps += [os.path.join(REFERENCE_DIR, f) for f in os.listdir(REFERENCE_DIR) if f.endswith((".cpp", ".h"))]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in deps):
        return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE_DIR}"])
    return LIB_PATH


def build_gpu():
    """The same harness with the Godot-module shim linked in (needs the product library to link against)."""
    if not sources_present():
        if os.path.exists(LIB_PATH_GPU):
            return LIB_PATH_GPU
        raise FileNotFoundError("oracle/_ref/libgas_ref_gpu.so is not built and the reference tree is not present")
    subprocess.check_call(["make", "-C", _HERE, "-s", "refgpu", f"REF={REFERENCE_DIR}"])
    return LIB_PATH_GPU


def gpu_available():
    return os.path.exists(LIB_PATH_GPU) or sources_present()


def _bind(lib):
    sig = {
        "ref_create": (_vp, [_vp]),
        "ref_destroy": (None, [_vp]),
        "ref_error_count": (C.c_int, []),
        "ref_last_error": (C.c_char_p, []),
        "ref_registered_classes": (C.c_int, []),
        "ref_set_speaker_mode": (C.c_int, [_vp, C.c_int]),
        "ref_set_mix_rate": (C.c_int, [_vp, _f32]),
        "ref_set_global_panning_strength": (C.c_int, [_vp, _f32]),
        "ref_set_server_lookahead": (C.c_int, [_vp, C.c_int]),
        "ref_use_gpu_shim": (C.c_int, [_vp, C.c_int]),
        "ref_has_gpu_shim": (C.c_int, []),
        "ref_spatializer_set": (C.c_int, [_vp, C.c_int, _vp]),
        "ref_instance_init": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "ref_instance_start": (C.c_int, [_vp, C.c_int, _vp]),
        "ref_instance_stop": (C.c_int, [_vp, C.c_int, _vp]),
        "ref_voice_init": (C.c_int, [_vp, C.c_int, _vp]),
        "ref_gain_compute": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
        "ref_params_set": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "ref_params_get": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "ref_effect_params_set": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "ref_mix_block": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
        "ref_mix_block_stream": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
        "ref_voice_state_export": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "ref_get_attenuation_db": (_f32, [_vp, _f32, _f32, _f32]),
        "ref_calc_output_vol": (None, [C.c_int, _f32, _f32, _vp, _vp]),
        "ref_spcap_calculate": (None, [C.c_int, _vp, _f32, _vp, _vp]),
        "ref_get_bus_map": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
        "ref_process_frames_3d": (None, [_vp, _vp, _f32, _vp, _vp, C.c_int]),
        "ref_mix_channel_3d": (None, [_vp, _vp, _f32, C.c_int, _vp, _vp, C.c_int]),
        "ref_filter_prepare_coefficients": (None, [C.c_int, _f32, _f32, _f32, C.c_int, _f32, _vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    global _lib
    if _lib is None:
        build()
        _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def load_gpu():
    global _lib_gpu
    if _lib_gpu is None:
        build_gpu()
        _lib_gpu = _bind(C.CDLL(LIB_PATH_GPU))
    return _lib_gpu


def _arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class RefError(RuntimeError):
    pass


class RefMixer:
    """Same methods as OracleMixer, computed by the reference module's own code."""

    def __init__(self, gpu_shim=False, **config):
        """gpu_shim: AudioSpatializer3D resources become AudioSpatializer3DGPU (integration/godot_module): the reference's
        plumbing around the GPU-backed classes.  One such world at a time; needs a B200."""
        self._lib = load_gpu() if gpu_shim else load()
        self.config = abi.config_defaults(**config)
        self._w = self._lib.ref_create(_ptr(self.config.reshape(1)))
        if not self._w:
            raise RefError("ref_create: invalid configuration")
        if gpu_shim:
            self._ck(self._lib.ref_use_gpu_shim(self._w, 1))

    def close(self):
        if getattr(self, "_w", None):
            self._lib.ref_destroy(self._w)
            self._w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st != 0:
            raise RefError(f"reference harness status {st}: {self._lib.ref_last_error().decode()}")

    @property
    def channels(self):
        return int(self.config["speaker_mode"]) + 1

    @property
    def num_buses(self):
        return int(self.config["num_buses"])

    @property
    def error_count(self):
        """ERR_FAIL_* / ERR_PRINT hits inside the module and the stand-in since the library was loaded."""
        return int(self._lib.ref_error_count())

    def set_speaker_mode(self, mode):
        self._ck(self._lib.ref_set_speaker_mode(self._w, int(mode)))
        self.config["speaker_mode"] = mode

    def set_mix_rate(self, hz):
        self._ck(self._lib.ref_set_mix_rate(self._w, float(hz)))
        self.config["mix_rate"] = hz

    def set_global_panning_strength(self, s):
        self._ck(self._lib.ref_set_global_panning_strength(self._w, float(s)))
        self.config["global_panning_strength"] = s

    def set_server_lookahead(self, on):
        self._ck(self._lib.ref_set_server_lookahead(self._w, int(bool(on))))

    def spatializer_set(self, slot, spat):
        s = _arr(spat, abi.spatializer).reshape(1)
        self._ck(self._lib.ref_spatializer_set(self._w, int(slot), _ptr(s)))

    def instance_init(self, instances, spatializers):
        i = _arr(instances, np.int32)
        s = np.broadcast_to(_arr(spatializers, np.int32), i.shape).copy()
        self._ck(self._lib.ref_instance_init(self._w, i.size, _ptr(i), _ptr(s)))

    def instance_start(self, instances):
        i = _arr(instances, np.int32)
        self._ck(self._lib.ref_instance_start(self._w, i.size, _ptr(i)))

    def instance_stop(self, instances):
        i = _arr(instances, np.int32)
        self._ck(self._lib.ref_instance_stop(self._w, i.size, _ptr(i)))

    def voice_init(self, voices):
        v = _arr(voices, np.int32)
        self._ck(self._lib.ref_voice_init(self._w, v.size, _ptr(v)))

    def gain_compute(self, emitters, listeners, areas=None, want_params=True):
        e = _arr(emitters, abi.emitter).reshape(-1)
        l = _arr(listeners, abi.listener).reshape(-1)
        a = _arr(areas, abi.area).reshape(-1) if areas is not None else None
        out = np.zeros(e.size, dtype=abi.params) if want_params else None
        self._ck(self._lib.ref_gain_compute(self._w, e.size, _ptr(e), l.size, _ptr(l), 0 if a is None else a.size, _ptr(a), _ptr(out)))
        return out

    def params_set(self, instances, params):
        i = _arr(instances, np.int32)
        p = _arr(params, abi.params).reshape(-1)
        self._ck(self._lib.ref_params_set(self._w, i.size, _ptr(i), _ptr(p)))

    def params_get(self, instances):
        i = _arr(instances, np.int32)
        out = np.zeros(i.size, dtype=abi.params)
        self._ck(self._lib.ref_params_get(self._w, i.size, _ptr(i), _ptr(out)))
        return out

    def effect_params_set(self, instances, chains):
        i = _arr(instances, np.int32)
        c = _arr(chains, abi.effect_chain).reshape(-1)
        self._ck(self._lib.ref_effect_params_set(self._w, i.size, _ptr(i), _ptr(c)))

    @staticmethod
    def _src(src):
        s = np.asarray(src)
        if s.dtype == abi.frame:
            s = s.view(np.float32).reshape(s.shape + (2,))
        return np.ascontiguousarray(s, dtype=np.float32)

    def mix_block(self, voices, src, frames=None, want_peaks=True, **_):
        """Block mode.  The reference keeps a voice's peak in a local variable, so peaks come back as zeros."""
        v = _arr(voices, abi.voice).reshape(-1)
        s = self._src(src)
        rows = 0 if s.size == 0 else s.shape[0]
        if frames is None:
            frames = s.shape[1]
        bus = np.zeros((self.num_buses, self.channels, frames, 2), dtype=np.float32)
        self._ck(self._lib.ref_mix_block(self._w, v.size, _ptr(v), _ptr(s), rows, int(frames), _ptr(bus)))
        return bus, (np.zeros((v.size, 2), dtype=np.float32) if want_peaks else None)

    def mix_block_stream(self, voices, src, mixed_frames, frames=None):
        """Stream mode: returns (bus, state) with state bit 0 = still active, bit 1 = has_frames, 0 = deleted."""
        v = _arr(voices, abi.voice).reshape(-1)
        s = self._src(src)
        rows = 0 if s.size == 0 else s.shape[0]
        if frames is None:
            frames = s.shape[1]
        mf = _arr(mixed_frames, np.int32)
        assert mf.size == v.size
        bus = np.zeros((self.num_buses, self.channels, frames, 2), dtype=np.float32)
        act = np.zeros(max(v.size, 1), dtype=np.int32)
        self._ck(self._lib.ref_mix_block_stream(self._w, v.size, _ptr(v), _ptr(s), rows, int(frames), C.c_void_p(mf.ctypes.data), _ptr(bus), _ptr(act)))
        return bus, act[: v.size]

    def voice_state_export(self, voices):
        v = _arr(voices, np.int32)
        out = np.zeros(v.size, dtype=abi.voice_state)
        self._ck(self._lib.ref_voice_state_export(self._w, v.size, _ptr(v), _ptr(out)))
        return out
