"""Seeded synthetic workload: sources, emitter geometry, voice lists.

Everything is a pure function of (seed, index) through splitmix64, so the same inputs can be rebuilt
on any machine (tests, bench, the CPU baseline) without shipping data.
"""
import numpy as np

from . import abi

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_BASE = 0xA5D10


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays."""
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def u01(seed, idx):
    """uniform [0,1) doubles from (seed, idx) pairs (broadcast)."""
    with np.errstate(over="ignore"):
        h = splitmix64(np.asarray(seed, dtype=np.uint64) * np.uint64(0x100000001B3) + np.asarray(idx, dtype=np.uint64))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def make_sources(n_voices, frames, block=0, mix_rate=48000.0, voice0=0, noise=0.1):
    """float32 [n_voices, frames, 2]: per voice three sines (80 Hz..12 kHz, random phase) plus uniform
    noise, peak amplitude uniform in [0.05, 0.5]; continuous across blocks."""
    v = (np.arange(n_voices, dtype=np.uint64) + np.uint64(voice0 + SEED_BASE))[:, None]
    amp = 0.05 + 0.45 * u01(v, 1)
    i = (np.arange(frames, dtype=np.float64) + float(block) * frames)[None, :]
    out = np.zeros((n_voices, frames, 2), dtype=np.float64)
    for k in range(3):
        f = 80.0 * (12000.0 / 80.0) ** u01(v, 10 + k)
        ph_l = 2 * np.pi * u01(v, 20 + k)
        ph_r = 2 * np.pi * u01(v, 30 + k)
        w = 2 * np.pi * f / mix_rate
        out[:, :, 0] += np.sin(w * i + ph_l)
        out[:, :, 1] += np.sin(w * i + ph_r)
    out *= (1.0 - noise) / 3.0
    gi = (np.arange(frames, dtype=np.uint64) + np.uint64(block * frames))[None, :]
    out[:, :, 0] += noise * (2.0 * u01(v * np.uint64(2), gi) - 1.0)
    out[:, :, 1] += noise * (2.0 * u01(v * np.uint64(2) + np.uint64(1), gi) - 1.0)
    out *= amp[:, :, None]
    return out.astype(np.float32)


def make_emitters(n, block=0, dt=512.0 / 48000.0, instance0=0, spatializer=0, bus=0, area_fraction=0.0, seed0=0,
                  r_min=0.5, r_max=120.0, speed_max=10.0):
    """abi.emitter[n]: positions uniform in a spherical shell r in [r_min, r_max], moving with a
    per-emitter velocity (<= speed_max m/s) so that volumes change every block; volume_db U[-12, 0],
    max_db 3, pitch 1.  Emitters whose hash falls under area_fraction reference area 0."""
    s = (np.arange(n, dtype=np.uint64) + np.uint64(seed0 + instance0 + 0x51A7))
    e = np.zeros(n, dtype=abi.emitter)
    r = r_min + (r_max - r_min) * u01(s, 1) ** (1.0 / 3.0)
    cz = 2.0 * u01(s, 2) - 1.0
    az = 2.0 * np.pi * u01(s, 3)
    sx = np.sqrt(np.maximum(0.0, 1.0 - cz * cz))
    pos = np.stack([r * sx * np.cos(az), r * cz, r * sx * np.sin(az)], axis=1)
    vel = np.stack([2.0 * u01(s, 4 + k) - 1.0 for k in range(3)], axis=1) * (speed_max / np.sqrt(3.0))
    pos = pos + vel * (dt * block)
    e["instance"] = np.arange(n, dtype=np.int32) + instance0
    e["spatializer"] = spatializer
    e["area"] = np.where(u01(s, 8) < area_fraction, 0, -1).astype(np.int32)
    e["bus"] = bus
    e["origin"] = pos.astype(np.float32)
    bz = np.stack([2.0 * u01(s, 11 + k) - 1.0 for k in range(3)], axis=1)
    bz /= np.maximum(1e-9, np.linalg.norm(bz, axis=1, keepdims=True))
    e["basis_z"] = bz.astype(np.float32)
    e["velocity"] = vel.astype(np.float32)
    e["volume_db"] = (-12.0 * u01(s, 15)).astype(np.float32)
    e["max_db"] = 3.0
    e["pitch_scale"] = 1.0
    return e


def make_voices(n, voice0=0, instance0=0, voices_per_instance=1, flags=0):
    """abi.voice[n]: voice slot voice0+j, instance instance0 + j // voices_per_instance, source row j."""
    v = np.zeros(n, dtype=abi.voice)
    j = np.arange(n, dtype=np.int32)
    v["voice"] = voice0 + j
    v["instance"] = instance0 + j // voices_per_instance
    v["src_row"] = j
    v["flags"] = flags
    return v


def rotated_listener(yaw=0.6, origin=(3.0, 1.5, -2.0), velocity=(0.0, 0.0, 0.0)):
    """A yaw-rotated, translated listener (the second listener geometry of SURVEY §8d)."""
    l = np.zeros((), dtype=abi.listener)
    c, s = np.cos(yaw), np.sin(yaw)
    l["basis"] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32).ravel()
    l["origin"] = origin
    l["velocity"] = velocity
    return l


def reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0, override_bus=False, bus=0, n_listeners=1):
    a = np.zeros((), dtype=abi.area)
    a["override_bus"] = int(override_bus)
    a["bus"] = bus
    a["use_reverb"] = 1
    a["reverb_bus"] = reverb_bus
    a["reverb_amount"] = amount
    a["reverb_uniformity"] = uniformity
    cp = np.zeros((abi.MAX_LISTENERS, 3), dtype=np.float32)
    for k in range(n_listeners):
        cp[k] = (4.0 + k, 0.5, -6.0)
    a["closest_point"] = cp
    return a
