"""ctypes loader of the CPU oracle (oracle/libgas_oracle.so) with the same Python surface as the
product's Mixer, so a parity test can drive both with one call sequence.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.  Pinned against the reference's own code by oracle/_ref
(oracle/ref.py, tests/test_oracle_vs_ref.py); see gas_oracle.h for what that covers.
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
LIB_PATH = os.path.join(_HERE, "libgas_oracle.so")
_lib = None
_vp, _i32, _f32 = C.c_void_p, C.c_int32, C.c_float


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("gas_oracle.c", "gas_oracle_mix.inc", "gas_oracle.h")]
    src.append(os.path.join(os.path.dirname(_HERE), "include", "gas.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src):
        return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    sig = {
        "orc_create": (_vp, [_vp]),
        "orc_destroy": (None, [_vp]),
        "orc_set_speaker_mode": (C.c_int, [_vp, C.c_int]),
        "orc_set_mix_rate": (C.c_int, [_vp, _f32]),
        "orc_set_global_panning_strength": (C.c_int, [_vp, _f32]),
        "orc_spatializer_set": (C.c_int, [_vp, C.c_int, _vp]),
        "orc_instance_init": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_instance_start": (C.c_int, [_vp, C.c_int, _vp]),
        "orc_instance_stop": (C.c_int, [_vp, C.c_int, _vp]),
        "orc_voice_init": (C.c_int, [_vp, C.c_int, _vp]),
        "orc_gain_compute": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
        "orc_params_set": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_params_get": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_effect_params_set": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_mix_block": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
        "orc_mix_block_stream": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int]),
        "orc_set_playback_disable_threshold_db": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_voice_state_export": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_voice_state_import": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "orc_last_mix_seconds": (C.c_double, [_vp]),
        "orc_last_gain_seconds": (C.c_double, [_vp]),
        "orc_max_threads": (C.c_int, []),
        "orc_sizeof": (C.c_size_t, [_i32]),
        "orc_db_to_linear_f": (_f32, [_f32]),
        "orc_linear_to_db_f": (_f32, [_f32]),
        "orc_get_attenuation_db": (_f32, [_vp, _f32, _f32, _f32]),
        "orc_calc_output_vol_stereo": (None, [_vp, _f32, _vp]),
        "orc_calc_output_vol_surround": (None, [C.c_int, _vp, _f32, _vp]),
        "orc_calc_output_vol": (None, [C.c_int, _f32, _f32, _vp, _vp]),
        "orc_spcap_effective_speakers": (None, [C.c_int, _vp]),
        "orc_spcap_calculate": (None, [C.c_int, _vp, _f32, _vp]),
        "orc_filter_prepare_coefficients": (None, [C.c_int, _f32, _f32, _f32, C.c_int, _f32, _vp]),
        "orc_get_bus_map": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
        "orc_process_frames_3d": (None, [_vp, _vp, _f32, _vp, _vp, C.c_int]),
        "orc_mix_channel_3d": (None, [_vp, _vp, _f32, C.c_int, _vp, _vp, C.c_int]),
        "orc_process_frames_effect": (None, [_vp, _vp, _f32, _vp, _vp, C.c_int]),
        "orc_bus_graph": (None, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
        "orc_resampler_begin": (_vp, [_vp, C.c_int, C.c_int, _f32, C.c_int]),
        "orc_resampler_mix": (C.c_int, [_vp, _vp, _f32, _f32, C.c_int]),
        "orc_resampler_free": (None, [_vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    abi.check_layout(lib.orc_sizeof, "libgas_oracle.so")
    _lib = lib
    return lib


def _arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class OracleError(RuntimeError):
    pass


class OracleMixer:
    """Same methods as godot_audio_spatializer_b200.Mixer, computed by the scalar CPU oracle."""

    def __init__(self, **config):
        self._lib = load()
        self.config = abi.config_defaults(**config)
        self._w = self._lib.orc_create(_ptr(self.config.reshape(1)))
        if not self._w:
            raise OracleError("orc_create: invalid configuration")
        self.last_bus64 = None

    def close(self):
        if getattr(self, "_w", None):
            self._lib.orc_destroy(self._w)
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
            raise OracleError(f"oracle status {st}")

    @property
    def channels(self):
        return int(self.config["speaker_mode"]) + 1

    @property
    def num_buses(self):
        return int(self.config["num_buses"])

    def set_speaker_mode(self, mode):
        self._ck(self._lib.orc_set_speaker_mode(self._w, int(mode)))
        self.config["speaker_mode"] = mode

    def set_mix_rate(self, hz):
        self._ck(self._lib.orc_set_mix_rate(self._w, float(hz)))
        self.config["mix_rate"] = hz

    def set_global_panning_strength(self, s):
        self._ck(self._lib.orc_set_global_panning_strength(self._w, float(s)))
        self.config["global_panning_strength"] = s

    def spatializer_set(self, slot, spat):
        s = _arr(spat, abi.spatializer).reshape(1)
        self._ck(self._lib.orc_spatializer_set(self._w, int(slot), _ptr(s)))

    def instance_init(self, instances, spatializers):
        i = _arr(instances, np.int32)
        s = np.broadcast_to(_arr(spatializers, np.int32), i.shape).copy()
        self._ck(self._lib.orc_instance_init(self._w, i.size, _ptr(i), _ptr(s)))

    def instance_start(self, instances):
        i = _arr(instances, np.int32)
        self._ck(self._lib.orc_instance_start(self._w, i.size, _ptr(i)))

    def instance_stop(self, instances):
        i = _arr(instances, np.int32)
        self._ck(self._lib.orc_instance_stop(self._w, i.size, _ptr(i)))

    def voice_init(self, voices):
        v = _arr(voices, np.int32)
        self._ck(self._lib.orc_voice_init(self._w, v.size, _ptr(v)))

    def gain_compute(self, emitters, listeners, areas=None, want_params=True):
        e = _arr(emitters, abi.emitter).reshape(-1)
        l = _arr(listeners, abi.listener).reshape(-1)
        a = _arr(areas, abi.area).reshape(-1) if areas is not None else None
        out = np.zeros(e.size, dtype=abi.params) if want_params else None
        self._ck(self._lib.orc_gain_compute(self._w, e.size, _ptr(e), l.size, _ptr(l), 0 if a is None else a.size, _ptr(a), _ptr(out)))
        return out

    def params_set(self, instances, params):
        i = _arr(instances, np.int32)
        p = _arr(params, abi.params).reshape(-1)
        self._ck(self._lib.orc_params_set(self._w, i.size, _ptr(i), _ptr(p)))

    def params_get(self, instances):
        i = _arr(instances, np.int32)
        out = np.zeros(i.size, dtype=abi.params)
        self._ck(self._lib.orc_params_get(self._w, i.size, _ptr(i), _ptr(out)))
        return out

    def effect_params_set(self, instances, chains):
        i = _arr(instances, np.int32)
        c = _arr(chains, abi.effect_chain).reshape(-1)
        self._ck(self._lib.orc_effect_params_set(self._w, i.size, _ptr(i), _ptr(c)))

    def mix_block(self, voices, src, frames=None, want_peaks=True, shadow=False, threads=1, want_bus=True):
        v = _arr(voices, abi.voice).reshape(-1)
        s = np.asarray(src)
        if s.dtype == abi.frame:
            s = s.view(np.float32).reshape(s.shape + (2,))
        s = np.ascontiguousarray(s, dtype=np.float32)
        rows = 0 if s.size == 0 else s.shape[0]
        if frames is None:
            frames = s.shape[1]
        bus = np.zeros((self.num_buses, self.channels, frames, 2), dtype=np.float32) if want_bus else None
        peaks = np.zeros((v.size, 2), dtype=np.float32) if want_peaks else None
        bus64 = np.zeros((self.num_buses, self.channels, frames, 2), dtype=np.float64) if shadow else None
        if shadow:
            threads = 1  # the float64 shadow is only computed by the single-threaded loop (orc_mix_block)
        self._ck(self._lib.orc_mix_block(self._w, v.size, _ptr(v), _ptr(s), rows, int(frames), _ptr(bus), _ptr(peaks), _ptr(bus64), int(threads)))
        self.last_bus64 = bus64
        return bus, peaks

    def mix_block_stream(self, voices, src, mixed_frames, frames=None, threads=1):
        """Stream form (voice lifecycle inside): returns (bus, status) with status bit 0 = active, bit 1 = has_frames."""
        v = _arr(voices, abi.voice).reshape(-1)
        s = np.asarray(src)
        if s.dtype == abi.frame:
            s = s.view(np.float32).reshape(s.shape + (2,))
        s = np.ascontiguousarray(s, dtype=np.float32)
        rows = 0 if s.size == 0 else s.shape[0]
        if frames is None:
            frames = s.shape[1]
        mf = _arr(mixed_frames, np.int32)
        assert mf.size == v.size
        bus = np.zeros((self.num_buses, self.channels, frames, 2), dtype=np.float32)
        peaks = np.zeros((max(v.size, 1), 2), dtype=np.float32)
        status = np.zeros(max(v.size, 1), dtype=np.int32)
        self._ck(self._lib.orc_mix_block_stream(self._w, v.size, _ptr(v), _ptr(s), rows, int(frames), C.c_void_p(mf.ctypes.data), _ptr(bus),
                                                _ptr(peaks), _ptr(status), int(threads)))
        self.last_peaks = peaks[: v.size]
        return bus, status[: v.size]

    # ---- device-resident sources: the oracle twin of gas_source_set / gas_voice_play / gas_mix_block_resident -----------------
    def source_set(self, slot, pcm, sample_rate, loop=False):
        if not hasattr(self, "_sources"):
            self._sources, self._players = {}, {}
        self._sources[int(slot)] = (np.ascontiguousarray(np.asarray(pcm, dtype=np.float32).reshape(-1, 2)), float(sample_rate), bool(loop))

    def voice_play(self, voices, sources, start_frames=None):
        v = _arr(voices, np.int32).reshape(-1)
        s = np.broadcast_to(_arr(sources, np.int32), v.shape)
        st = np.zeros(v.shape, dtype=np.int32) if start_frames is None else np.broadcast_to(_arr(start_frames, np.int32), v.shape)
        for vi, si, fi in zip(v, s, st):
            old = self._players.pop(int(vi), None)
            if old is not None:
                old.close()
            if si >= 0:
                pcm, rate, loop = self._sources[int(si)]
                self._players[int(vi)] = Resampler(pcm, rate, loop=loop, start_frame=int(fi))

    def mix_block_resident(self, voices, frames, threads=1):
        """upstream's resampler per voice at the instance's current pitch_scale (audio_spatializer.cpp:375-378), then the stream form."""
        v = _arr(voices, abi.voice).reshape(-1)
        rows = np.zeros((max(v.size, 1), frames, 2), dtype=np.float32)
        mixed = np.zeros(max(v.size, 1), dtype=np.int32)
        inst = np.unique(v["instance"]) if v.size else np.zeros(0, dtype=np.int32)
        pitch = dict(zip(inst.tolist(), self.params_get(inst)["pitch_scale"].tolist())) if inst.size else {}
        vv = v.copy()
        for i in range(v.size):
            pl = self._players.get(int(v["voice"][i])) if hasattr(self, "_players") else None
            if pl is not None and v["src_row"][i] >= 0:
                rows[i], mixed[i] = pl.mix(frames, pitch[int(v["instance"][i])], float(self.config["mix_rate"]))
            vv["src_row"][i] = i if v["src_row"][i] >= 0 else -1
        return self.mix_block_stream(vv, rows, mixed[: v.size], frames, threads=threads)

    def set_playback_disable_threshold_db(self, instances, db):
        i = _arr(instances, np.int32)
        d = np.broadcast_to(_arr(db, np.float32), i.shape).copy()
        self._ck(self._lib.orc_set_playback_disable_threshold_db(self._w, i.size, _ptr(i), _ptr(d)))

    def voice_state_export(self, voices):
        v = _arr(voices, np.int32)
        out = np.zeros(v.size, dtype=abi.voice_state)
        self._ck(self._lib.orc_voice_state_export(self._w, v.size, _ptr(v), _ptr(out)))
        return out

    def voice_state_import(self, voices, states):
        v = _arr(voices, np.int32)
        s = _arr(states, abi.voice_state).reshape(-1)
        self._ck(self._lib.orc_voice_state_import(self._w, v.size, _ptr(v), _ptr(s)))

    @property
    def last_mix_seconds(self):
        return float(self._lib.orc_last_mix_seconds(self._w))

    @property
    def last_gain_seconds(self):
        return float(self._lib.orc_last_gain_seconds(self._w))


def bus_graph(bus, buses):
    """upstream bus graph after the mix on bus [n_buses, channels, frames, 2] (a copy is returned); buses as for Mixer.bus_layout_set."""
    lib = load()
    out = np.ascontiguousarray(np.asarray(bus, dtype=np.float32)).copy()
    nb, ch, fr = out.shape[:3]
    vol = np.array([b.get("volume_db", 0.0) for b in buses], dtype=np.float32)
    mute = np.array([int(bool(b.get("mute", False))) for b in buses], dtype=np.int32)
    solo = np.array([int(bool(b.get("solo", False))) for b in buses], dtype=np.int32)
    send = np.array([int(b.get("send", 0)) for b in buses], dtype=np.int32)
    assert len(buses) == nb
    lib.orc_bus_graph(nb, ch, fr, _ptr(vol), _ptr(mute), _ptr(solo), _ptr(send), _ptr(out))
    return out


class Resampler:
    """upstream AudioStreamPlaybackResampled over PCM (oracle/gas_oracle.c, recalled): one playing voice."""

    def __init__(self, pcm, sample_rate, loop=False, start_frame=0):
        self._lib = load()
        self._pcm = np.ascontiguousarray(np.asarray(pcm, dtype=np.float32).reshape(-1, 2))  # kept alive: the C side holds the pointer
        self._r = self._lib.orc_resampler_begin(_ptr(self._pcm), self._pcm.shape[0], int(bool(loop)), float(sample_rate), int(start_frame))
        if not self._r:
            raise OracleError("orc_resampler_begin failed")

    def mix(self, frames, rate_scale, target_rate):
        out = np.zeros((frames, 2), dtype=np.float32)
        n = self._lib.orc_resampler_mix(self._r, _ptr(out), float(rate_scale), float(target_rate), int(frames))
        return out, n

    def close(self):
        if self._r:
            self._lib.orc_resampler_free(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def max_threads():
    return int(load().orc_max_threads())
