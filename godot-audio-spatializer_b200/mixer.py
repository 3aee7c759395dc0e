"""Mixer — thin numpy-facing wrapper of the C ABI (one gas_ctx).

Method names follow include/gas.h one to one; each docstring names the reference interface the call
replaces.  Array arguments are numpy structured arrays with the dtypes of ``abi`` (or anything
convertible); device-side variants take raw device pointers (e.g. ``tensor.data_ptr()``).
"""
import ctypes as C

import numpy as np

from . import abi
from .lib import check, load


def _arr(x, dtype):
    a = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    return a


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class StepNext(C.Structure):
    """gas_step_next (include/gas.h)"""
    _fields_ = [("n_emitters", C.c_int32), ("d_emitters", C.c_void_p), ("n_voices", C.c_int32), ("d_voices", C.c_void_p),
                ("src_rows", C.c_int32), ("frames", C.c_int32), ("d_bus_out", C.c_void_p), ("d_peaks", C.c_void_p)]


class Mixer:
    def __init__(self, **config):
        """gas_create.  Keyword arguments override abi.config_defaults()."""
        self._lib = load()
        self.config = abi.config_defaults(**config)
        self._ctx = C.c_void_p()
        st = self._lib.gas_create(_ptr(self.config.reshape(1)), C.byref(self._ctx))
        if st != 0:
            msg = self._lib.gas_last_error(None)
            from .lib import GasError
            raise GasError(st, msg.decode() if msg else "")

    # ---- lifetime -----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.gas_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def channels(self):
        return int(self._lib.gas_get_channel_count(self._ctx))

    @property
    def num_buses(self):
        return int(self.config["num_buses"])

    @property
    def kernel_launches(self):
        return int(self._lib.gas_kernel_launches(self._ctx))

    @property
    def mix_stream(self):
        return self._lib.gas_mix_stream(self._ctx)

    @property
    def gain_stream(self):
        return self._lib.gas_gain_stream(self._ctx)

    def _ck(self, st):
        check(st, self._ctx)

    # ---- globals (AudioServer::get_speaker_mode / get_mix_rate, 3d_panning_strength) -----------------
    def set_speaker_mode(self, mode):
        self._ck(self._lib.gas_set_speaker_mode(self._ctx, int(mode)))
        self.config["speaker_mode"] = mode

    def set_mix_rate(self, hz):
        self._ck(self._lib.gas_set_mix_rate(self._ctx, float(hz)))
        self.config["mix_rate"] = hz

    def set_global_panning_strength(self, s):
        self._ck(self._lib.gas_set_global_panning_strength(self._ctx, float(s)))
        self.config["global_panning_strength"] = s

    # ---- resources / slots ----------------------------------------------------------------------------
    def spatializer_set(self, slot, spat):
        """AudioSpatializer3D property setters with their validation (audio_spatializer_3d.cpp:654-765)."""
        s = _arr(spat, abi.spatializer).reshape(1)
        self._ck(self._lib.gas_spatializer_set(self._ctx, int(slot), _ptr(s)))

    def instance_init(self, instances, spatializers):
        """AudioSpatializer::instantiate (audio_spatializer_3d.cpp:645-652)."""
        i = _arr(instances, np.int32)
        s = np.broadcast_to(_arr(spatializers, np.int32), i.shape).copy()
        self._ck(self._lib.gas_instance_init(self._ctx, i.size, _ptr(i), _ptr(s)))

    def instance_start(self, instances):
        """Proxy playbacks registered with AudioServer (audio_spatializer.cpp:75-95)."""
        i = _arr(instances, np.int32)
        self._ck(self._lib.gas_instance_start(self._ctx, i.size, _ptr(i)))

    def instance_stop(self, instances):
        i = _arr(instances, np.int32)
        self._ck(self._lib.gas_instance_stop(self._ctx, i.size, _ptr(i)))

    def voice_init(self, voices):
        """instantiate_playback_data (audio_spatializer_3d.cpp:200-204)."""
        v = _arr(voices, np.int32)
        self._ck(self._lib.gas_voice_init(self._ctx, v.size, _ptr(v)))

    # ---- gain side --------------------------------------------------------------------------------------
    def gain_compute(self, emitters, listeners, areas=None, want_params=True):
        """Batched calculate_spatialization + update_spatializer_parameters
        (audio_spatializer_3d.cpp:277-489, audio_spatializer.cpp:258-272)."""
        e = _arr(emitters, abi.emitter).reshape(-1)
        l = _arr(listeners, abi.listener).reshape(-1)
        a = _arr(areas, abi.area).reshape(-1) if areas is not None else None
        out = np.zeros(e.size, dtype=abi.params) if want_params else None
        self._ck(self._lib.gas_gain_compute(self._ctx, e.size, _ptr(e), l.size, _ptr(l),
                                            0 if a is None else a.size, _ptr(a), _ptr(out)))
        return out

    def gain_compute_device(self, n, d_emitters, listeners=None, areas=None, d_out_params=0):
        """Device-resident emitters; listeners/areas None => the resident copies (listeners_set/areas_set)."""
        l = _arr(listeners, abi.listener).reshape(-1) if listeners is not None else None
        a = _arr(areas, abi.area).reshape(-1) if areas is not None else None
        self._ck(self._lib.gas_gain_compute_device(self._ctx, int(n), C.c_void_p(d_emitters), 0 if l is None else l.size, _ptr(l),
                                                   0 if a is None else a.size, _ptr(a), C.c_void_p(d_out_params)))

    def listeners_set(self, listeners):
        l = _arr(listeners, abi.listener).reshape(-1)
        self._ck(self._lib.gas_listeners_set(self._ctx, l.size, _ptr(l)))

    def areas_set(self, areas):
        a = _arr(areas, abi.area).reshape(-1)
        self._ck(self._lib.gas_areas_set(self._ctx, a.size, _ptr(a)))

    # ---- CUDA-graph capture / per-kernel timing ---------------------------------------------------------
    def capture_begin(self):
        self._ck(self._lib.gas_capture_begin(self._ctx))

    def capture_end(self):
        g = C.c_int32(-1)
        self._ck(self._lib.gas_capture_end(self._ctx, C.byref(g)))
        return int(g.value)

    def graph_launch(self, graph):
        self._ck(self._lib.gas_graph_launch(self._ctx, int(graph)))

    def graph_destroy(self, graph):
        self._ck(self._lib.gas_graph_destroy(self._ctx, int(graph)))

    def profile_enable(self, on=True):
        self._ck(self._lib.gas_profile_enable(self._ctx, 1 if on else 0))

    def profile_read(self):
        """{kind: (total_ms, launches)} for prologue / mix_stream (K2) / mix_voice (K3) / gain (K1) / none (an event pair
        around nothing, once per mix block: what the timer itself reads)."""
        ms = (C.c_double * 5)()
        n = (C.c_uint64 * 5)()
        self._ck(self._lib.gas_profile_read(self._ctx, ms, n))
        names = ("prologue", "mix_stream", "mix_voice", "gain", "none")
        return {names[k]: (float(ms[k]), int(n[k])) for k in range(5)}

    def params_set(self, instances, params):
        """set_spatializer_parameters + bus-map push (audio_spatializer.cpp:258-272, :558-564)."""
        i = _arr(instances, np.int32)
        p = _arr(params, abi.params).reshape(-1)
        if p.size != i.size:
            raise ValueError("one gas_params per instance")
        self._ck(self._lib.gas_params_set(self._ctx, i.size, _ptr(i), _ptr(p)))

    def params_get(self, instances):
        i = _arr(instances, np.int32)
        out = np.zeros(i.size, dtype=abi.params)
        self._ck(self._lib.gas_params_get(self._ctx, i.size, _ptr(i), _ptr(out)))
        return out

    def effect_params_set(self, instances, chains):
        """What a _process_effects override writes into its effects (audio_spatializer_effect.cpp:39)."""
        i = _arr(instances, np.int32)
        c = _arr(chains, abi.effect_chain).reshape(-1)
        if c.size != i.size:
            raise ValueError("one gas_effect_chain per instance")
        self._ck(self._lib.gas_effect_params_set(self._ctx, i.size, _ptr(i), _ptr(c)))

    # ---- mix side -----------------------------------------------------------------------------------------
    def mix_block(self, voices, src, frames=None, want_peaks=True):
        """Batched _mix_from_playback_list + AudioServer bus accumulate (audio_spatializer.cpp:326-471).

        src: float32 [rows, frames, 2] (or abi.frame [rows, frames]).  Returns (bus, peaks) with
        bus float32 [num_buses, channels, frames, 2] and peaks float32 [n_voices, 2] (or None)."""
        v = _arr(voices, abi.voice).reshape(-1)
        s = np.asarray(src)
        if s.dtype == abi.frame:
            s = s.view(np.float32).reshape(s.shape + (2,))
        s = np.ascontiguousarray(s, dtype=np.float32)
        if s.size == 0:
            rows = 0
            if frames is None:
                raise ValueError("frames is required without source rows")
        else:
            if s.ndim != 3 or s.shape[2] != 2:
                raise ValueError("src must be [rows, frames, 2]")
            rows = s.shape[0]
            if frames is None:
                frames = s.shape[1]
            elif frames != s.shape[1]:
                raise ValueError("src row length must equal frames")
        bus = np.empty((self.num_buses, self.channels, frames, 2), dtype=np.float32)
        peaks = np.zeros((v.size, 2), dtype=np.float32) if want_peaks else None
        self._ck(self._lib.gas_mix_block(self._ctx, v.size, _ptr(v), _ptr(s), rows, int(frames), _ptr(bus), _ptr(peaks)))
        return bus, peaks

    def mix_block_host_ptr(self, n_voices, voices_ptr, src_ptr, src_rows, frames, bus_ptr, peaks_ptr=0):
        """gas_mix_block with caller-owned (e.g. pinned) host buffers given as raw addresses."""
        self._ck(self._lib.gas_mix_block(self._ctx, int(n_voices), C.c_void_p(voices_ptr), C.c_void_p(src_ptr), int(src_rows),
                                         int(frames), C.c_void_p(bus_ptr), C.c_void_p(peaks_ptr)))

    def mix_block_device(self, n_voices, d_voices, d_src, src_rows, src_row_stride, frames, d_bus_out, d_peaks=0):
        """gas_mix_block_device: asynchronous on the mix stream, everything device-resident."""
        self._ck(self._lib.gas_mix_block_device(self._ctx, int(n_voices), C.c_void_p(d_voices), C.c_void_p(d_src), int(src_rows),
                                                int(src_row_stride), int(frames), C.c_void_p(d_bus_out), C.c_void_p(d_peaks)))

    def bus_layout_set(self, buses):
        """gas_bus_layout_set: buses = list of dict(volume_db=0.0, mute=False, solo=False, send=0), one per bus of the context."""
        d = np.zeros(len(buses), dtype=abi.bus_desc)
        for i, b in enumerate(buses):
            d[i]["volume_db"] = b.get("volume_db", 0.0)
            d[i]["mute"] = int(bool(b.get("mute", False)))
            d[i]["solo"] = int(bool(b.get("solo", False)))
            d[i]["send"] = int(b.get("send", 0))
        self._ck(self._lib.gas_bus_layout_set(self._ctx, d.size, _ptr(d)))

    def bus_graph_device(self, d_bus, frames):
        """gas_bus_graph_device: the bus graph in place on device-resident bus buffers, asynchronous on the mix stream."""
        self._ck(self._lib.gas_bus_graph_device(self._ctx, C.c_void_p(d_bus), int(frames)))

    def bus_graph(self, bus):
        """gas_bus_graph: the bus graph over host bus buffers [num_buses, channels, frames, 2]; returns the processed copy."""
        b = np.ascontiguousarray(np.asarray(bus, dtype=np.float32)).copy()
        self._ck(self._lib.gas_bus_graph(self._ctx, _ptr(b), int(b.shape[2])))
        return b

    def source_set(self, slot, pcm, sample_rate, loop=False):
        """gas_source_set: a PCM clip (float32 [n, 2]) becomes device-resident source `slot`."""
        p = np.ascontiguousarray(np.asarray(pcm, dtype=np.float32).reshape(-1, 2))
        self._ck(self._lib.gas_source_set(self._ctx, int(slot), _ptr(p), p.shape[0], float(sample_rate), int(bool(loop))))

    def voice_play(self, voices, sources, start_frames=None):
        """gas_voice_play: begin_resample of voices[i] on sources[i] from start_frames[i]."""
        v = _arr(voices, np.int32).reshape(-1)
        s = np.broadcast_to(_arr(sources, np.int32), v.shape).copy()
        st = None if start_frames is None else np.broadcast_to(_arr(start_frames, np.int32), v.shape).copy()
        self._ck(self._lib.gas_voice_play(self._ctx, v.size, _ptr(v), _ptr(s), _ptr(st) if st is not None else None))

    def resample_block_device(self, n_voices, d_voices, frames, d_rows, row_stride, src_rows, d_mixed_frames):
        """gas_resample_block_device: asynchronous on the mix stream, device pointers."""
        self._ck(self._lib.gas_resample_block_device(self._ctx, int(n_voices), C.c_void_p(d_voices), int(frames), C.c_void_p(d_rows),
                                                     int(row_stride), int(src_rows), C.c_void_p(d_mixed_frames)))

    def mix_block_resident(self, voices, frames):
        """gas_mix_block_resident: resample the voices' resident sources and mix them through the stream form.  Returns (bus, status)."""
        v = _arr(voices, abi.voice).reshape(-1)
        bus = np.zeros((self.num_buses, self.channels, frames, 2), dtype=np.float32)
        status = np.zeros(max(v.size, 1), dtype=np.int32)
        self._ck(self._lib.gas_mix_block_resident(self._ctx, v.size, _ptr(v), int(frames), _ptr(bus), _ptr(status)))
        return bus, status[: v.size]

    def mix_block_resident_host_ptr(self, n_voices, voices_ptr, frames, bus_ptr, status_ptr=0):
        """gas_mix_block_resident with caller-owned (pinned) host buffers: bench.py's e2e leg."""
        self._ck(self._lib.gas_mix_block_resident(self._ctx, int(n_voices), C.c_void_p(voices_ptr), int(frames), C.c_void_p(bus_ptr),
                                                  C.c_void_p(status_ptr) if status_ptr else None))

    def step_device(self, d_src=0, src_row_stride=0, next=None):
        """gas_step_device: streams the block planned by the previous call from d_src and prepares `next` in the same launch.
        next: dict(n_voices, d_voices, src_rows, frames, d_bus_out[, d_peaks, n_emitters, d_emitters]) of device pointers, or None
        to end the run."""
        if next is None:
            self._ck(self._lib.gas_step_device(self._ctx, C.c_void_p(d_src), int(src_row_stride), None))
            return
        nx = StepNext(int(next.get("n_emitters", 0)), C.c_void_p(next.get("d_emitters", 0) or None), int(next["n_voices"]),
                      C.c_void_p(next["d_voices"]), int(next["src_rows"]), int(next["frames"]), C.c_void_p(next["d_bus_out"]),
                      C.c_void_p(next.get("d_peaks", 0) or None))
        self._ck(self._lib.gas_step_device(self._ctx, C.c_void_p(d_src or None), int(src_row_stride), C.byref(nx)))

    def step_join_device(self):
        """gas_step_join_device: the mix stream waits for the outstanding voice-parallel kernels of pipelined steps."""
        self._ck(self._lib.gas_step_join_device(self._ctx))

    def process_frames(self, instance, voice, src):
        """gas_process_frames: the reference's process_frames virtual on one voice (audio_spatializer_3d.cpp:491-552,
        audio_spatializer_effect.cpp:33-77).  src float32 [frames, 2]; returns out [frames, 2]."""
        s = np.ascontiguousarray(np.asarray(src, dtype=np.float32))
        out = np.empty_like(s)
        self._ck(self._lib.gas_process_frames(self._ctx, int(instance), int(voice), _ptr(out), _ptr(s), s.shape[0]))
        return out

    def mix_channel(self, instance, voice, channel, src):
        """gas_mix_channel: the reference's mix_channel virtual on one voice and one channel pair (audio_spatializer_3d.cpp:554-609)."""
        s = np.ascontiguousarray(np.asarray(src, dtype=np.float32))
        out = np.empty_like(s)
        self._ck(self._lib.gas_mix_channel(self._ctx, int(instance), int(voice), int(channel), _ptr(out), _ptr(s), s.shape[0]))
        return out

    def mix_block_stream(self, voices, src, mixed_frames, frames=None):
        """gas_mix_block_stream: the voice lifecycle of _mix_from_playback_list on the device (audio_spatializer.cpp:353-408,
        :464-469).  src row r = the mixed_frames[i] frames AudioStreamPlayback::mix returned for voice i this block.
        Returns (bus, status) with status = VOICE_ACTIVE | VOICE_HAS_FRAMES per voice after the block."""
        v = _arr(voices, abi.voice).reshape(-1)
        s = np.asarray(src)
        if s.dtype == abi.frame:
            s = s.view(np.float32).reshape(s.shape + (2,))
        s = np.ascontiguousarray(s, dtype=np.float32)
        rows = 0 if s.size == 0 else s.shape[0]
        if frames is None:
            frames = s.shape[1]
        mf = _arr(mixed_frames, np.int32)
        if mf.size != v.size:
            raise ValueError("one frame count per voice")
        bus = np.empty((self.num_buses, self.channels, frames, 2), dtype=np.float32)
        status = np.zeros(max(v.size, 1), dtype=np.int32)
        self._ck(self._lib.gas_mix_block_stream(self._ctx, v.size, _ptr(v), _ptr(s), rows, int(frames), C.c_void_p(mf.ctypes.data), _ptr(bus),
                                                C.c_void_p(status.ctypes.data)))
        return bus, status[: v.size]

    def mix_block_stream_device(self, n_voices, d_voices, d_src, src_rows, src_row_stride, frames, d_mixed_frames, d_bus_out, d_status=0):
        self._ck(self._lib.gas_mix_block_stream_device(self._ctx, int(n_voices), C.c_void_p(d_voices), C.c_void_p(d_src), int(src_rows),
                                                       int(src_row_stride), int(frames), C.c_void_p(d_mixed_frames), C.c_void_p(d_bus_out),
                                                       C.c_void_p(d_status)))

    def set_playback_disable_threshold_db(self, instances, db):
        """AudioSpatializerInstance::set_playback_disable_threshold_db (audio_spatializer.cpp:576-582)."""
        i = _arr(instances, np.int32)
        d = np.broadcast_to(_arr(db, np.float32), i.shape).copy()
        self._ck(self._lib.gas_set_playback_disable_threshold_db(self._ctx, i.size, _ptr(i), _ptr(d)))

    def voice_life_export(self, voices):
        v = _arr(voices, np.int32)
        out = np.zeros(v.size, dtype=abi.voice_life)
        self._ck(self._lib.gas_voice_life_export(self._ctx, v.size, _ptr(v), _ptr(out)))
        return out

    def voice_life_import(self, voices, life):
        v = _arr(voices, np.int32)
        s = _arr(life, abi.voice_life).reshape(-1)
        self._ck(self._lib.gas_voice_life_import(self._ctx, v.size, _ptr(v), _ptr(s)))

    def status_flags(self):
        f = C.c_uint32(0)
        self._ck(self._lib.gas_status_flags(self._ctx, C.byref(f)))
        return int(f.value)

    def sync(self):
        self._ck(self._lib.gas_sync(self._ctx))

    # ---- persistent state ------------------------------------------------------------------------------------
    def voice_state_export(self, voices):
        v = _arr(voices, np.int32)
        out = np.zeros(v.size, dtype=abi.voice_state)
        self._ck(self._lib.gas_voice_state_export(self._ctx, v.size, _ptr(v), _ptr(out)))
        return out

    def voice_state_import(self, voices, states):
        v = _arr(voices, np.int32)
        s = _arr(states, abi.voice_state).reshape(-1)
        if s.size != v.size:
            raise ValueError("one gas_voice_state per voice")
        self._ck(self._lib.gas_voice_state_import(self._ctx, v.size, _ptr(v), _ptr(s)))

    # ---- multi-GPU exchange -----------------------------------------------------------------------------------
    def comm_export(self):
        buf = (C.c_ubyte * 64)()
        self._ck(self._lib.gas_comm_export(self._ctx, buf, 64))
        return bytes(buf)

    def comm_open(self, rank, handles):
        blob = b"".join(handles)
        self._ck(self._lib.gas_comm_open(self._ctx, int(rank), len(handles), blob, 64))

    def reduce_bus_device(self, d_bus, frames):
        """gas_reduce_bus_device: d_bus (partial sums of this rank) becomes the sum over all ranks; asynchronous."""
        self._ck(self._lib.gas_reduce_bus_device(self._ctx, C.c_void_p(d_bus), int(frames)))

    def reduce_bus_begin_device(self, d_bus, frames):
        self._ck(self._lib.gas_reduce_bus_begin_device(self._ctx, C.c_void_p(d_bus), int(frames)))

    def reduce_bus_end_device(self, d_bus, frames):
        self._ck(self._lib.gas_reduce_bus_end_device(self._ctx, C.c_void_p(d_bus), int(frames)))

    def reduce_bus_exchange_device(self, d_partial, d_prev_sum, frames):
        """finish(previous block) -> d_prev_sum, push(d_partial), on the exchange stream (gas_reduce_bus_exchange_device)."""
        self._ck(self._lib.gas_reduce_bus_exchange_device(self._ctx, C.c_void_p(d_partial), C.c_void_p(d_prev_sum), int(frames)))

    def comm_close(self):
        self._ck(self._lib.gas_comm_close(self._ctx))
