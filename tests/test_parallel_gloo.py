"""CPU, world_size = 2, gloo: the N>1 host logic (graph sharding + the single flat-gradient
all-reduce) without any GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gotennet_b200.parallel import FlatGradBuffer, shard_bounds, take_shard
from gotennet_b200.synthetic import synth_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.SiLU(), torch.nn.Linear(16, 4))
    # shard a synthetic batch by molecule; each rank back-propagates its own shard
    z, pos, batch = synth_batch("qm9", 10, seed=3)
    n_atoms = torch.bincount(batch)
    lo, hi = shard_bounds((n_atoms.double() ** 2).tolist(), world)[rank]
    zs, ps, bs = take_shard(z, pos, batch, lo, hi)
    feats = torch.cat([ps, zs.float().unsqueeze(1), torch.ones(ps.shape[0], 4)], dim=1)
    loss = net(feats).pow(2).sum()
    loss.backward()
    fb = FlatGradBuffer(net.parameters())
    fb.all_reduce()
    fb.unpack()
    total = torch.tensor([float(loss)])
    dist.all_reduce(total)
    q.put((rank, fb.flat.clone(), float(total), (lo, hi), int(zs.numel())))
    dist.destroy_process_group()


def test_sharded_backward_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the whole batch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.SiLU(), torch.nn.Linear(16, 4))
    z, pos, batch = synth_batch("qm9", 10, seed=3)
    feats = torch.cat([pos, z.float().unsqueeze(1), torch.ones(pos.shape[0], 4)], dim=1)
    loss = net(feats).pow(2).sum()
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(res[0][1], res[1][1])                      # every rank holds the same reduced buffer
    assert torch.allclose(res[0][1], ref, rtol=1e-5, atol=1e-5)        # = gradient of the un-sharded batch
    assert abs(res[0][2] - float(loss)) < 1e-3 * abs(float(loss))
    assert res[0][3][1] == res[1][3][0] and res[0][3][0] == 0 and res[1][3][1] == 10
    assert res[0][4] + res[1][4] == z.numel()


def test_shard_bounds_properties():
    w = [float(i % 7 + 1) for i in range(100)]
    for world in (1, 2, 3, 8, 150):
        b = shard_bounds(w, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == 100
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        if world <= 8:
            loads = [sum(w[lo:hi]) for lo, hi in b]
            assert max(loads) - min(loads) <= 2 * max(w)
    assert shard_bounds([], 4) == [(0, 0)] * 4
