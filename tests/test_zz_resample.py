"""The resampler in front of the path: upstream AudioStreamPlaybackResampled::mix over device-resident PCM.

The oracle restates upstream's loop literally (internal buffer, refills, end-of-stream rule) from Godot 4.x as recalled — the engine
is not part of the reference tree, so this row is pinned by known answers, not by reference code.  The CUDA kernel evaluates the
closed form of the same loop and must reproduce the oracle bit for bit: rows, the count of valid frames, across blocks, for pitch
scales 0.5 .. 2, clips that end inside a block (at every alignment against the 128-frame internal buffer) and looping clips."""
import numpy as np
import pytest

import scenarios as S

abi, synth = S.abi, S.synth


def _clip(n, seed):
    rs = np.random.RandomState(seed)
    t = np.arange(n, dtype=np.float64)
    x = 0.4 * np.sin(2 * np.pi * (110.0 + 30.0 * seed) * t / 44100.0)[:, None] + 0.1 * rs.randn(n, 2)
    return x.astype(np.float32)


def test_oracle_resampler_known_answers(orc):
    """Unit rate: the resampler is a two-frame delay (mu = 0 selects y1 = S[g - 2]); half rate: every other output sits exactly on a
    source frame; the count of valid frames ends 4 source frames early, upstream's rule."""
    pcm = _clip(1000, 1)
    r = orc.Resampler(pcm, 48000.0)
    out, n = r.mix(512, 1.0, 48000.0)
    assert n == 512
    np.testing.assert_array_equal(out[2:], pcm[:510])
    np.testing.assert_array_equal(out[:2], 0.0)
    out2, n2 = r.mix(512, 1.0, 48000.0)
    np.testing.assert_array_equal(out2[:486], pcm[510:996])
    # 1000 = 7 * 128 + 104: the refill of buffer 7 returns 104 frames; the first output with 4 + (g % 128) >= 104 is g = 996
    assert n2 == 996 - 512
    r.close()
    r = orc.Resampler(pcm, 24000.0)
    out, n = r.mix(256, 1.0, 48000.0)  # increment 0.5
    assert n == 256
    np.testing.assert_array_equal(out[4::2], pcm[:126])
    r.close()


def closed_form_block(pcm, n_frames, loop, sample_rate, start, pos, pitch, mix_rate, frames):
    """The kernel's arithmetic (csrc/gas_resample.cu, k_resample) in numpy, operation for operation: returns (rows, valid, new pos)."""
    f32 = np.float32
    rate = f32(f32(sample_rate) * f32(pitch)) * f32(1.0)
    incd = (float(rate) / float(f32(mix_rate))) * 65536.0
    inc = int(incd) if incd > 0.0 else 0
    i = np.arange(frames, dtype=np.uint64)
    p = np.uint64(pos) + i * np.uint64(inc)
    g = (p >> np.uint64(16)).astype(np.int64)
    mu = (p & np.uint64(0xffff)).astype(np.float32) / f32(65536.0)

    def at(x):
        a = start + x
        ok = x >= 0
        if loop:
            a = np.where(a >= n_frames, a % n_frames, a)
        else:
            ok &= a < n_frames
        out = np.zeros((frames, 2), dtype=np.float32)
        out[ok] = pcm[a[ok]]
        return out

    y0, y1, y2, y3 = at(g - 3), at(g - 2), at(g - 1), at(g)
    n_rel = n_frames - start
    k_end, end_val = n_rel >> 7, n_rel - ((n_rel >> 7) << 7)
    bad = (~np.bool_(loop)) & ((g >> 7) >= k_end) & ((4 + (g & 127)) >= end_val)
    valid = int(np.argmax(bad)) if bad.any() else frames
    mu = mu[:, None]
    mu2 = mu * mu
    h11 = mu2 * (mu - f32(1.0))
    z = mu2 - h11
    h01 = z - h11
    h10 = mu - z
    out = (y1 + (y2 - y1) * h01) + ((y2 - y0) * h10 + (y3 - y1) * h11) * f32(0.5)
    return out.astype(np.float32), valid, int(pos) + frames * inc


@pytest.mark.parametrize("loop", [False, True])
def test_closed_form_of_the_kernel_equals_the_literal_loop(orc, loop):
    """The CUDA kernel does not walk upstream's internal buffer: it evaluates the closed form.  Same arithmetic in numpy against the
    oracle's literal loop, bit for bit, over clip lengths that end at every alignment against the 128-frame buffer."""
    F = 512
    rng = np.random.RandomState(5)
    for case in range(40):
        n = 128 + int(rng.randint(0, 3000))
        start = int(rng.randint(0, max(1, n - 128)))
        if not loop and n - start < 128:
            start = 0
        rate = [44100.0, 48000.0, 22050.0, 96000.0][case % 4]
        pcm = _clip(n, case)
        r = orc.Resampler(pcm, rate, loop=loop, start_frame=start)
        pos = 0
        for b in range(5):
            pitch = np.float32(rng.uniform(0.5, 2.0))
            want, n_want = r.mix(F, float(pitch), 48000.0)
            got, n_got, pos = closed_form_block(pcm, n, loop, rate, start, pos, pitch, 48000.0, F)
            assert n_got == n_want, f"case {case} block {b}: {n_got} valid frames, literal loop {n_want}"
            np.testing.assert_array_equal(got, want, err_msg=f"case {case} block {b}")
        r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("loop", [False, True])
def test_cuda_resampler_matches_oracle_bit_for_bit(gas, orc, loop):
    import torch
    V, F, blocks = 48, 512, 5
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=abi.SPEAKER_MODE_STEREO, mix_rate=48000.0)
    rates = [44100.0, 48000.0, 22050.0, 32000.0]
    lens = [1200 + 37 * k for k in range(8)]  # ends at every alignment against the 128-frame buffer, inside blocks 1-3
    clips = [_clip(lens[k], k) for k in range(8)]
    pitches = np.linspace(0.5, 2.0, V).astype(np.float32)
    starts = np.array([(7 * i) % 64 for i in range(V)], dtype=np.int32)
    dev = torch.device("cuda", 0)
    voices = synth.make_voices(V)
    with gas.Mixer(**cfg) as m:
        inst = np.arange(V, dtype=np.int32)
        m.spatializer_set(0, abi.spatializer_defaults())
        m.instance_init(inst, 0)
        p = np.zeros(V, dtype=abi.params)
        p["pitch_scale"] = pitches
        p["mix_volumes"] = 1.0
        m.params_set(inst, p)
        m.voice_init(inst)
        for k in range(8):
            m.source_set(k, clips[k], rates[k % 4], loop=loop)
        m.voice_play(inst, inst % 8, starts)
        want = [orc.Resampler(clips[i % 8], rates[(i % 8) % 4], loop=loop, start_frame=int(starts[i])) for i in range(V)]
        d_voices = torch.from_numpy(voices.view(np.uint8).copy()).to(dev)
        d_rows = torch.full((V, F, 2), 9.0, device=dev)
        d_mixed = torch.full((V,), -7, device=dev, dtype=torch.int32)
        for b in range(blocks):
            if b == 2:  # the pitch scale is read per block from the instance's current parameters
                pitches = pitches[::-1].copy()
                p["pitch_scale"] = pitches
                m.params_set(inst, p)
            m.resample_block_device(V, d_voices.data_ptr(), F, d_rows.data_ptr(), F, V, d_mixed.data_ptr())
            m.sync()
            rows, mixed = d_rows.cpu().numpy(), d_mixed.cpu().numpy()
            for i in range(V):
                w, n = want[i].mix(F, float(pitches[i]), 48000.0)
                assert mixed[i] == n, f"block {b} voice {i}: {mixed[i]} valid frames, oracle {n}"
                np.testing.assert_array_equal(rows[i], w, err_msg=f"block {b} voice {i}")
        for r in want:
            r.close()
        if not loop:
            assert (mixed < F).any()  # the fast voices have run off the end of their clips by now


def _resident_scene(V, F, blocks):
    clips = [_clip(1500 + 211 * k, 10 + k) for k in range(6)]
    ems = [synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.5) for b in range(blocks)]
    for e in ems:
        e["pitch_scale"] = np.linspace(0.5, 2.0, V).astype(np.float32)  # AudioStreamPlayerSpatial::pitch_scale, passed through
    return clips, ems


def _play_resident(mm, V, F, blocks, clips, ems, rate=44100.0):
    """The same calls on the CUDA Mixer and on the oracle: returns per-block (bus, status) and the voices alive before each block."""
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5)], dtype=abi.area)
    inst = np.arange(V, dtype=np.int32)
    voices = synth.make_voices(V)
    mm.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1, attenuation_filter_db=-18.0))
    mm.instance_init(inst, 0)
    mm.gain_compute(ems[0], listeners, areas, want_params=False)
    mm.instance_start(inst)
    mm.voice_init(inst)
    for k, c in enumerate(clips):
        mm.source_set(k, c, rate)
    mm.voice_play(inst, inst % len(clips))
    active = np.ones(V, dtype=bool)
    out = []
    for b in range(blocks):
        mm.gain_compute(ems[b], listeners, areas, want_params=False)
        live = voices[active].copy()
        live["src_row"] = np.arange(live.size)
        bus, status = mm.mix_block_resident(live, F)
        out.append((bus, status.copy(), active.copy()))
        alive = (status & abi.VOICE_ACTIVE) != 0
        idx = np.nonzero(active)[0]
        active[idx[~alive]] = False
    return out


def test_oracle_resident_path_lifecycle(orc):
    """The oracle twin of gas_mix_block_resident (upstream resampler per voice + stream form): clips end inside the run, voices keep
    their filter tails for a while and are then deactivated; a 48 kHz clip at pitch 1 is the stream form fed with the clip itself,
    two frames late."""
    V, F, blocks = 24, 512, 8
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=abi.SPEAKER_SURROUND_51, mix_rate=48000.0)
    clips, ems = _resident_scene(V, F, blocks)
    with orc.OracleMixer(**cfg) as o:
        res = _play_resident(o, V, F, blocks, clips, ems)
    n_active = [int(r[2].sum()) for r in res]
    assert n_active[0] == V and n_active[-1] < V and all(a >= b for a, b in zip(n_active, n_active[1:]))
    assert all(np.isfinite(r[0]).all() for r in res) and np.abs(res[0][0]).max() > 0
    # unit rate: rows = the clip delayed by two frames
    for e in ems:
        e["pitch_scale"] = 1.0
    clip = _clip(4096, 3)
    with orc.OracleMixer(**cfg) as a, orc.OracleMixer(**cfg) as b:
        got = _play_resident(a, V, F, 2, [clip], ems, rate=48000.0)
        listeners = np.array([abi.identity_listener()], dtype=abi.listener)
        areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5)], dtype=abi.area)
        inst = np.arange(V, dtype=np.int32)
        b.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1, attenuation_filter_db=-18.0))
        b.instance_init(inst, 0)
        b.gain_compute(ems[0], listeners, areas, want_params=False)
        b.instance_start(inst)
        b.voice_init(inst)
        delayed = np.concatenate([np.zeros((2, 2), np.float32), clip])
        for blk in range(2):
            b.gain_compute(ems[blk], listeners, areas, want_params=False)
            rows = np.broadcast_to(delayed[blk * F:(blk + 1) * F], (V, F, 2)).copy()
            want_bus, want_status = b.mix_block_stream(synth.make_voices(V), rows, np.full(V, F, dtype=np.int32), F)
            np.testing.assert_array_equal(got[blk][0], want_bus)
            np.testing.assert_array_equal(got[blk][1], want_status)


@pytest.mark.gpu
def test_resident_mix_matches_the_oracle_twin(gas, orc):
    """gas_mix_block_resident (resample + lifecycle + mix on the device) against the oracle twin, call for call: bus buffers within
    tolerance, routing and lifecycle status identical, until clips have ended and their tails have died."""
    V, F, blocks = 96, 512, 8
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=abi.SPEAKER_SURROUND_51, mix_rate=48000.0)
    clips, ems = _resident_scene(V, F, blocks)
    with gas.Mixer(**cfg) as m, orc.OracleMixer(**cfg) as o:
        got = _play_resident(m, V, F, blocks, clips, ems)
        want = _play_resident(o, V, F, blocks, clips, ems)
    for b, ((gb, gs, ga), (wb, ws, wa)) in enumerate(zip(got, want)):
        np.testing.assert_array_equal(ga, wa, err_msg=f"block {b}: live voices")
        np.testing.assert_array_equal(gs, ws, err_msg=f"block {b}: lifecycle status")
        assert np.array_equal(S.routing(gb), S.routing(wb)), f"block {b}: routing"
        ok, worst, nbad = S.sample_close(gb, wb)
        assert ok, f"block {b}: {nbad} samples out of tolerance (worst {worst:.3e})"
    assert not want[-1][2].all()  # clips ended and tails died inside the run
