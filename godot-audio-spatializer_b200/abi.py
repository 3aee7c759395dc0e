"""numpy mirrors of the POD records in include/gas.h (C layout, natural alignment).

Every dtype here is checked against ``gas_abi_sizeof`` when the library is loaded, so a layout drift
between the header and this file fails loudly instead of corrupting a call.
"""
import numpy as np

MAX_CHANNELS_PER_BUS = 4        # reference audio_spatializer.h:48
LOOKAHEAD_BUFFER_SIZE = 64      # :49
MAX_BUSES_PER_PLAYBACK = 6      # :50
MAX_LISTENERS = 8
MAX_EFFECTS = 4
MAX_FILTER_STAGES = 4
MAX_BUSES = 16

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_NCCL, ERR_NO_DEVICE = range(7)

SPEAKER_MODE_STEREO, SPEAKER_SURROUND_31, SPEAKER_SURROUND_51, SPEAKER_SURROUND_71 = range(4)
ATTENUATION_INVERSE_DISTANCE, ATTENUATION_INVERSE_SQUARE_DISTANCE, ATTENUATION_LOGARITHMIC, ATTENUATION_DISABLED = range(4)
DOPPLER_TRACKING_DISABLED, DOPPLER_TRACKING_IDLE_STEP, DOPPLER_TRACKING_PHYSICS_STEP = range(3)
SPATIALIZER_3D, SPATIALIZER_EFFECT = range(2)
(FILTER_BANDPASS, FILTER_HIGHPASS, FILTER_LOWPASS, FILTER_NOTCH, FILTER_PEAK, FILTER_BANDLIMIT,
 FILTER_LOWSHELF, FILTER_HIGHSHELF) = range(8)
VOICE_WANT_PEAK = 1

f4, i4, u4 = np.float32, np.int32, np.uint32

frame = np.dtype([("l", f4), ("r", f4)], align=True)
effect = np.dtype([("mode", i4), ("cutoff_hz", f4), ("resonance", f4), ("gain", f4), ("stages", i4)], align=True)
effect_chain = np.dtype([("n_effects", i4), ("effects", effect, (MAX_EFFECTS,))], align=True)
spatializer = np.dtype([
    ("kind", i4), ("attenuation_model", i4), ("unit_size", f4), ("max_distance", f4), ("panning_strength", f4),
    ("area_mask", u4), ("emission_angle_enabled", i4), ("emission_angle", f4),
    ("emission_angle_filter_attenuation_db", f4), ("attenuation_filter_cutoff_hz", f4),
    ("attenuation_filter_db", f4), ("doppler_tracking", i4), ("doppler_speed_of_sound", f4),
    ("mix_channel_mode", i4), ("effect_gain_binding", i4), ("chain", effect_chain)], align=True)
listener = np.dtype([("basis", f4, (9,)), ("origin", f4, (3,)), ("velocity", f4, (3,))], align=True)
area = np.dtype([("override_bus", i4), ("bus", i4), ("use_reverb", i4), ("reverb_bus", i4), ("reverb_amount", f4),
                 ("reverb_uniformity", f4), ("closest_point", f4, (MAX_LISTENERS, 3))], align=True)
emitter = np.dtype([("instance", i4), ("spatializer", i4), ("area", i4), ("bus", i4), ("origin", f4, (3,)),
                    ("basis_z", f4, (3,)), ("velocity", f4, (3,)), ("volume_db", f4), ("max_db", f4),
                    ("pitch_scale", f4)], align=True)
params = np.dtype([("mix_volumes", f4, (MAX_CHANNELS_PER_BUS, 2)), ("pitch_scale", f4), ("linear_attenuation", f4),
                   ("attenuation_filter_cutoff_hz", f4), ("update_parameters", i4), ("n_bus", i4),
                   ("bus", i4, (MAX_BUSES_PER_PLAYBACK,)),
                   ("bus_volumes", f4, (MAX_BUSES_PER_PLAYBACK, MAX_CHANNELS_PER_BUS, 2))], align=True)
voice = np.dtype([("voice", i4), ("instance", i4), ("src_row", i4), ("flags", u4)], align=True)
processor_state = np.dtype([("b0", f4), ("b1", f4), ("b2", f4), ("a1", f4), ("a2", f4),
                            ("ha1", f4), ("ha2", f4), ("hb1", f4), ("hb2", f4)], align=True)
voice_state = np.dtype([("prev_mix_volumes", f4, (MAX_CHANNELS_PER_BUS, 2)),
                        ("filter_processors", processor_state, (2 * MAX_CHANNELS_PER_BUS,)),
                        ("effect_history", f4, (MAX_EFFECTS, 2, MAX_FILTER_STAGES, 4))], align=True)
voice_life = np.dtype([("lookahead", frame, (LOOKAHEAD_BUFFER_SIZE,)), ("flags", u4)], align=True)
bus_desc = np.dtype([("volume_db", f4), ("mute", i4), ("solo", i4), ("send", i4)], align=True)  # gas_bus_desc
step_next = np.dtype([("n_emitters", i4), ("d_emitters", np.uint64), ("n_voices", i4), ("d_voices", np.uint64), ("src_rows", i4), ("frames", i4),
                      ("d_bus_out", np.uint64), ("d_peaks", np.uint64)], align=True)  # gas_step_next (pointers as 64-bit integers)
VOICE_ACTIVE, VOICE_HAS_FRAMES = 1, 2
STATUS_CLASS_OVERFLOW = 1
config = np.dtype([("device", i4), ("max_instances", i4), ("max_voices", i4), ("max_frames", i4),
                   ("max_spatializers", i4), ("num_buses", i4), ("speaker_mode", i4), ("mix_rate", f4),
                   ("global_panning_strength", f4)], align=True)

# gas_struct_id order (include/gas.h)
STRUCT_IDS = [frame, effect, effect_chain, spatializer, listener, area, emitter, params, voice,
              processor_state, voice_state, config, voice_life, bus_desc, step_next]


def check_layout(sizeof_fn, who):
    """sizeof_fn(struct_id) -> C sizeof; raises if any numpy mirror disagrees."""
    for sid, dt in enumerate(STRUCT_IDS):
        c = int(sizeof_fn(sid))
        if c != dt.itemsize:
            raise RuntimeError(f"{who}: struct id {sid} is {c} bytes in C but {dt.itemsize} in abi.py")


def spatializer_defaults(**kw):
    """AudioSpatializer3D property defaults (reference audio_spatializer_3d.h:171-188)."""
    s = np.zeros((), dtype=spatializer)
    s["kind"] = SPATIALIZER_3D
    s["attenuation_model"] = ATTENUATION_INVERSE_DISTANCE
    s["unit_size"] = 10.0
    s["max_distance"] = 0.0
    s["panning_strength"] = 1.0
    s["area_mask"] = 1
    s["emission_angle_enabled"] = 0
    s["emission_angle"] = 45.0
    s["emission_angle_filter_attenuation_db"] = -12.0
    s["attenuation_filter_cutoff_hz"] = 5000.0
    s["attenuation_filter_db"] = -24.0
    s["doppler_tracking"] = DOPPLER_TRACKING_DISABLED
    s["doppler_speed_of_sound"] = 343.0
    s["mix_channel_mode"] = 0
    s["effect_gain_binding"] = -1
    for k, v in kw.items():
        s[k] = v
    return s


def config_defaults(**kw):
    c = np.zeros((), dtype=config)
    c["device"] = 0
    c["max_instances"] = 1024
    c["max_voices"] = 1024
    c["max_frames"] = 512          # upstream AudioServer buffer_size
    c["max_spatializers"] = 16
    c["num_buses"] = 2
    c["speaker_mode"] = SPEAKER_MODE_STEREO
    c["mix_rate"] = 44100.0        # upstream AudioServer default mix rate
    c["global_panning_strength"] = 0.5  # audio/general/3d_panning_strength default
    for k, v in kw.items():
        c[k] = v
    return c


def identity_listener(origin=(0, 0, 0), velocity=(0, 0, 0)):
    l = np.zeros((), dtype=listener)
    l["basis"] = np.eye(3, dtype=f4).ravel()
    l["origin"] = origin
    l["velocity"] = velocity
    return l
