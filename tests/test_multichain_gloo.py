"""N>1 host logic on CPU: the chain sharding / gather of parallelHMCSampler (parallelHMC.jl:10-49) with a
world_size-2 gloo group.  The per-chain sampler is stubbed (no GPU here); what is exercised is the rank
assignment, the absence of any data-path collective, and the final gather + output."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from hmcmt2d_b200 import api


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def fake_sampler(mesh, data, inv, prior, nsamples=None, seed=0, device=0, **kw):
        calls.append((seed, device))
        model = np.full((3, nsamples), float(seed))
        st = api.HMCStatus(nsamples, 0, np.ones(nsamples, bool), np.zeros((4, nsamples + 1)))
        return model, st, np.zeros((2, nsamples + 1), complex)

    real = api.runHMCSampler
    api.runHMCSampler = fake_sampler
    try:
        inv = api.InvDataModel(None, None, np.zeros(3), np.zeros(3), None, None, None)
        models, stats, datas = api.parallelHMCSampler(None, None, inv, api.HMCPrior(), pids=[0, 1, 2, 3, 4], nsamples=2,
                                                      outdir=outdir if rank == 0 else None)
    finally:
        api.runHMCSampler = real
    assert [s for s, _ in calls] == [k + 1 for k in range(5) if k % world == rank]       # chains rank, rank+world, ...
    assert all(d == rank for _, d in calls)                                                # on this rank's device
    assert len(models) == 5 and [m[0, 0] for m in models] == [1.0, 2.0, 3.0, 4.0, 5.0]      # every rank sees all chains
    dist.barrier()
    dist.destroy_process_group()


def test_parallel_sampler_shards_chains_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == sorted([f"hmcsamples_id{k}.{e}" for k in range(1, 6) for e in ("model", "data")]
                                                  + [f"hmcstatistics_id{k}.log" for k in range(1, 6)])
