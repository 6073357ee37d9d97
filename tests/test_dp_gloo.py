"""Data-parallel host logic on CPU with the gloo backend, world_size 2: the flat bucket all-reduce sums the
per-rank gradients, and shard-averaged gradients equal the oracle's full-batch gradient when BN is not involved
(per-rank BN statistics are DDP semantics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from feature_level_style_transfer_for_tsc_b200.train_step import FlatParameters
    torch.manual_seed(0)
    a, b = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    flat = FlatParameters([(list(a.parameters()), 0.001), (list(b.parameters()), 0.003)])
    assert flat.group_end[-1] == flat.flat_p.numel() and flat.group_end[0] % 4 == 0
    assert all(p.data.data_ptr() >= flat.flat_p.data_ptr() for p in flat.params)
    x = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10.0
    shard = x[rank * 4:(rank + 1) * 4]                       # rank r takes samples [r*B/N, (r+1)*B/N)
    flat.zero_grad()
    b(torch.relu(a(shard))).pow(2).mean().backward()          # local mean over the shard
    local = flat.flat_g.clone()
    world_n = flat.all_reduce_sum()
    assert world_n == world
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat.flat_g, sum(gathered))
    # average of the shard gradients == full-batch gradient (equal shard sizes)
    a2, b2 = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    a2.load_state_dict(a.state_dict()); b2.load_state_dict(b.state_dict())
    b2(torch.relu(a2(x))).pow(2).mean().backward()
    full = torch.cat([torch.nn.functional.pad(p.grad.flatten(), (0, (-p.numel()) % 4)) for p in list(a2.parameters()) + list(b2.parameters())])
    assert torch.allclose(flat.flat_g / world, full, atol=1e-6)
    ret[rank] = True
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
