"""N > 1 host logic on CPU: two gloo ranks each mix their shard of the voices (with the CPU oracle standing in
for a GPU) and all-reduce the partial bus buffers; the sum must match the unsharded mix within the north-star
tolerance (float summation order differs across shards)."""
import os
import socket
import sys

import numpy as np
import pytest

import scenarios as S

abi = S.abi
shard = S.gas.shard


def test_instance_ranges_partition_everything():
    for n in (0, 1, 7, 64, 1000, 16384):
        for w in (1, 2, 3, 4, 8):
            spans = [shard.instance_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            if n:
                inst = np.arange(n)
                own = shard.owner_of_instance(inst, n, w)
                for r, (lo, hi) in enumerate(spans):
                    assert np.all(own[lo:hi] == r)


def test_shard_voices_keeps_instances_together():
    v = S.synth.make_voices(45, voices_per_instance=3)
    v["src_row"][::7] = -1
    seen = []
    for r in range(2):
        loc, idx = shard.shard_voices(v, 15, 2, r)
        lo, hi = shard.instance_range(15, 2, r)
        assert np.all((v["instance"][idx] >= lo) & (v["instance"][idx] < hi))
        assert np.array_equal(loc["instance"], v["instance"][idx] - lo)
        assert np.array_equal(loc["src_row"] < 0, v["src_row"][idx] < 0)
        seen.extend(idx.tolist())
    assert sorted(seen) == list(range(45))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, F, blocks = 48, 256, 3
        sc = S.default_scenario(voices=V, frames=F, speaker_mode=abi.SPEAKER_SURROUND_51, spat=dict(mix_channel_mode=1),
                                area=dict(reverb_bus=1, amount=0.5), area_fraction=0.5, blocks=blocks)
        lo, hi = shard.instance_range(V, world, rank)
        n_loc = hi - lo
        voices, idx = shard.shard_voices(S.synth.make_voices(V), V, world, rank)
        cfg = S.config_of(sc)
        cfg.update(max_instances=max(n_loc, 1), max_voices=max(n_loc, 1))
        listeners = np.array([abi.identity_listener()], dtype=abi.listener)
        areas = np.array([S.synth.reverb_area(**sc["area"])], dtype=abi.area)
        out = []
        with orc.OracleMixer(**cfg) as o:
            inst = np.arange(n_loc, dtype=np.int32)
            o.spatializer_set(0, S.make_spatializer(sc))
            o.instance_init(inst, 0)
            for b in range(blocks):
                em = S.synth.make_emitters(V, block=b, dt=F / sc["mix_rate"], area_fraction=sc["area_fraction"])[lo:hi].copy()
                em["instance"] -= lo
                o.gain_compute(em, listeners, areas, want_params=False)
                if b == 0:
                    o.instance_start(inst)
                    o.voice_init(inst)
                src = S.synth.make_sources(V, F, block=b, mix_rate=sc["mix_rate"])[idx]
                bus, _ = o.mix_block(voices, src, F, want_peaks=False)
                t = torch.from_numpy(bus.copy())
                shard.reduce_bus(t, dist)
                out.append(t.numpy())
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


def test_two_ranks_sum_to_the_unsharded_mix(orc):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    V, F, blocks = 48, 256, 3
    sc = S.default_scenario(voices=V, frames=F, speaker_mode=abi.SPEAKER_SURROUND_51, spat=dict(mix_channel_mode=1),
                            area=dict(reverb_bus=1, amount=0.5), area_fraction=0.5, blocks=blocks)
    with orc.OracleMixer(**S.config_of(sc)) as o:
        want = S.run(o, sc, collect_state=False)["bus"]
    for b in range(blocks):
        assert np.array_equal(S.routing(got[b]), S.routing(want[b]))
        ok, worst, nbad = S.sample_close(got[b], want[b])
        assert ok, f"block {b}: {nbad} samples out of tolerance (worst {worst:.3e})"
