"""Frequency sharding on CPU: world_size-2 gloo group, the per-rank plan is stubbed (no GPU here).
Exercised: the round-robin frequency partition, the renumbering of freqID / observation rows, and the single
sum-all-reduce that rebuilds [gradient | misfit | predicted data] of the full problem on every rank."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from hmcmt2d_b200 import api, synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakePlan:
    """Stands in for api.Plan: 'physics' that is additive over frequencies, like the real data gradient."""

    def __init__(self, mesh, data, inv, prior, nChains=1, device=0):
        self.data, self.inv, self.nAC = data, inv, len(inv.strModel)

    def forward_gradient(self, m):
        d = self.data
        f = d.freqs[d.freqID - 1]
        pred = (f * d.rxID + 1j * d.dtID) * np.sum(m)
        g = sum(np.cos(fr * np.arange(self.nAC)) for fr in d.freqs) * m
        phi = np.array([np.sum(np.abs(pred - self.inv.obsData) ** 2 * self.inv.dataW ** 2)])
        return pred[None, :], phi, g[None, :]

    def close(self):
        pass


def _problem():
    mesh, data, inv, prior = synthetic.make_problem(24, 20, 5, nRx=3)
    return mesh, data, inv, prior


def _worker(rank, world, port):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    real = api.Plan
    api.Plan = _FakePlan
    try:
        mesh, data, inv, prior = _problem()
        m = np.linspace(-5.0, -4.0, len(inv.strModel))
        sp = api.FreqShardedPlan(mesh, data, inv, prior, rank, world, balance="frequencies")
        assert list(sp.plan.data.freqs) == list(data.freqs[rank::world])
        assert sp.plan.data.freqID.min() == 1 and sp.plan.data.freqID.max() == len(data.freqs[rank::world])
        pred, phi, g = sp.forward_gradient(m)
        fpred, fphi, fg = _FakePlan(mesh, data, inv, prior).forward_gradient(m)
        assert np.allclose(pred, fpred[0], rtol=1e-14, atol=0)
        assert abs(phi - fphi[0]) <= 1e-12 * abs(fphi[0])
        assert np.allclose(g, fg[0], rtol=1e-12, atol=1e-12)
    finally:
        api.Plan = real
    dist.barrier()
    dist.destroy_process_group()


def test_frequency_shards_world2():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def _worker_systems(rank, world, port):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    real = api.Plan
    api.Plan = _FakePlan
    try:
        mesh, data, inv, prior = _problem()
        m = np.linspace(-5.0, -4.0, len(inv.strModel))
        sp = api.FreqShardedPlan(mesh, data, inv, prior, rank, world)            # default: balanced (frequency, mode) systems
        assert not sp.in_library_nccl                                             # gloo group: the exchange stays with torch.distributed
        assert sp.plan.data.dataComp == [data.dataComp[rank]] and list(sp.plan.data.freqs) == list(data.freqs)
        assert (sp.plan.data.dtID == 1).all() and len(sp.rows) == len(inv.obsData) // 2
        pred, phi, g = sp.forward_gradient(m)
        assert np.count_nonzero(pred) == len(inv.obsData)                         # every observation row filled by exactly one rank
    finally:
        api.Plan = real
    dist.barrier()
    dist.destroy_process_group()


def test_system_shards_world2():
    mp.spawn(_worker_systems, args=(2, _free_port()), nprocs=2, join=True)


def test_system_partition_is_balanced_and_complete():
    mesh, data, inv, prior = synthetic.make_problem(24, 20, 6, nRx=3)
    for world in (2, 4, 6):
        seen = np.zeros(len(inv.obsData), int)
        counts = []
        for rank in range(world):
            sub, sinv, rows = api.shard_systems(data, inv, rank, world)
            seen[rows] += 1
            counts.append(len(sub.freqs) * len(sub.dataComp))
            assert np.array_equal(sinv.obsData, inv.obsData[rows]) and len(sub.dataID) == len(sub.freqs) * 3
            assert np.array_equal(sub.freqs[sub.freqID - 1], data.freqs[data.freqID[rows] - 1])
        assert (seen == 1).all() and max(counts) - min(counts) <= (0 if 12 % world == 0 else 1)
    sub, sinv, rows = api.shard_systems(data, inv, 1, 3)                          # odd world: frequency round-robin
    assert list(sub.freqs) == list(data.freqs[1::3]) and len(sub.dataComp) == 2


def test_partition_covers_every_observation_once():
    mesh, data, inv, prior = _problem()
    for world in (1, 2, 3, 5):
        seen = np.zeros(len(inv.obsData), int)
        for rank in range(world):
            sub, sinv, rows = api.shard_frequencies(data, inv, rank, world)
            seen[rows] += 1
            assert np.array_equal(sub.freqs[sub.freqID - 1], data.freqs[data.freqID[rows] - 1])
            assert np.array_equal(sinv.obsData, inv.obsData[rows]) and len(sub.dataID) == len(sub.freqs) * 3 * 2
        assert (seen == 1).all()
    try:
        api.shard_frequencies(data, inv, 5, 6)
        assert False
    except ValueError:
        pass
