#!/usr/bin/env python
"""N-GPU correctness check of the peer-memory bus reduce (run under torchrun on one box):
every rank mixes its shard of the voices on its own B200, gas_reduce_bus_device sums the partial bus buffers
over NVLink, and every rank compares the result with the CPU oracle's unsharded mix.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as S  # noqa: E402
from oracle import orc  # noqa: E402

gas, abi, synth = S.gas, S.abi, S.synth


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, F, blocks, mode = 1024, 256, 4, abi.SPEAKER_SURROUND_51
    sc = S.default_scenario(voices=V, frames=F, speaker_mode=mode, spat=dict(mix_channel_mode=1), area=dict(reverb_bus=1, amount=0.5),
                            area_fraction=0.5, blocks=blocks, force_filter_off=False)
    lo, hi = gas.shard.instance_range(V, world, rank)
    n_loc = hi - lo
    voices, idx = gas.shard.shard_voices(synth.make_voices(V), V, world, rank)
    cfg = S.config_of(sc)
    cfg.update(max_instances=n_loc, max_voices=n_loc, device=local)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(**sc["area"])], dtype=abi.area)
    dev = torch.device("cuda", local)
    got = []
    with gas.Mixer(**cfg) as m:
        handles = [None] * world
        dist.all_gather_object(handles, m.comm_export())
        m.comm_open(rank, handles)
        dist.barrier()
        inst = np.arange(n_loc, dtype=np.int32)
        m.spatializer_set(0, S.make_spatializer(sc))
        m.instance_init(inst, 0)
        d_voices = torch.from_numpy(voices.view(np.uint8).copy()).to(dev)
        d_bus = torch.zeros((cfg["num_buses"], mode + 1, F, 2), device=dev, dtype=torch.float32)
        for b in range(blocks):
            em = synth.make_emitters(V, block=b, dt=F / sc["mix_rate"], area_fraction=sc["area_fraction"])[lo:hi].copy()
            em["instance"] -= lo
            m.gain_compute(em, listeners, areas, want_params=False)
            if b == 0:
                m.instance_start(inst)
                m.voice_init(inst)
            src = torch.from_numpy(synth.make_sources(V, F, block=b, mix_rate=sc["mix_rate"])[idx].copy()).to(dev)
            m.mix_block_device(n_loc, d_voices.data_ptr(), src.data_ptr(), n_loc, F, F, d_bus.data_ptr())
            m.reduce_bus_device(d_bus.data_ptr(), F)
            m.sync()
            got.append(d_bus.cpu().numpy().copy())
        m.sync()
        dist.barrier()
        # the pipelined variant: one block in flight on the exchange stream, sums arrive one call later
        got2 = []
        m.instance_init(inst, 0)
        d_part = [torch.zeros_like(d_bus) for _ in range(2)]
        d_sum = [torch.zeros_like(d_bus) for _ in range(2)]
        for b in range(blocks):
            em = synth.make_emitters(V, block=b, dt=F / sc["mix_rate"], area_fraction=sc["area_fraction"])[lo:hi].copy()
            em["instance"] -= lo
            m.gain_compute(em, listeners, areas, want_params=False)
            if b == 0:
                m.instance_start(inst)
                m.voice_init(inst)
            src = torch.from_numpy(synth.make_sources(V, F, block=b, mix_rate=sc["mix_rate"])[idx].copy()).to(dev)
            m.mix_block_device(n_loc, d_voices.data_ptr(), src.data_ptr(), n_loc, F, F, d_part[b % 2].data_ptr())
            m.reduce_bus_exchange_device(d_part[b % 2].data_ptr(), d_sum[(b + 1) % 2].data_ptr(), F)  # pushes b, finishes b - 1
            m.sync()
            if b >= 1:
                got2.append(d_sum[(b + 1) % 2].cpu().numpy().copy())
        m.reduce_bus_end_device(d_sum[blocks % 2].data_ptr(), F)
        m.sync()
        got2.append(d_sum[blocks % 2].cpu().numpy().copy())
        dist.barrier()
    with orc.OracleMixer(**S.config_of(sc)) as o:
        want = S.run(o, sc, collect_state=False)["bus"]
    ok_all = True
    for b in range(blocks):
        ok, worst, nbad = S.sample_close(got[b], want[b])
        routing = np.array_equal(S.routing(got[b]), S.routing(want[b]))
        ok_all &= ok and routing
        print(f"rank {rank} block {b}: ok={ok} routing={routing} worst_abs_err={worst:.3e} bad={nbad}", flush=True)
    for b in range(blocks):
        ok, worst, nbad = S.sample_close(got2[b], want[b])
        ok_all &= ok
        print(f"rank {rank} block {b} (pipelined exchange): ok={ok} worst_abs_err={worst:.3e} bad={nbad}", flush=True)
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0 and int(flag.item()) == 1:
        print(f"multi-GPU check ok: {world} ranks, in-order and pipelined exchange match the unsharded oracle", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
