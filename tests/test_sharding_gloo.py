"""The multi-GPU path is a pure partition of independent pouches (no data-path collective).  This
runs the host-side sharding logic with world_size 2 over gloo on CPU: each rank takes its LPT
shard, "computes" it with the CPU oracle standing in for the kernel, and the per-pouch digests
gathered on rank 0 must equal the unsharded run's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _digests_for(indices, T, steps):
    from oracle import pouch_oracle as O
    from calcium_b200.parallel import frames_digest
    from calcium_b200.ensemble import make_specs
    specs = make_specs(6, size="xsmall")
    out = {}
    for i in indices:
        s = specs[i]
        z = np.zeros(s.geometry.n_cells)
        f = O.simulate_c(s.geometry.csr, O.pack_params(s.params), z, z, s.r0, s.vplc, T, steps)
        out[i] = frames_digest(f)
    return out


def _worker(rank, world, port, T, steps, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from calcium_b200.parallel import shard_indices
    mine = shard_indices([200] * 6, T, world)[rank]
    d = _digests_for(mine, T, steps)
    payload = torch.zeros(6, 2, dtype=torch.int64)
    for i, v in d.items():
        payload[i] = torch.from_numpy(v.view(np.int64))
    dist.barrier()
    dist.reduce(payload, dst=0, op=dist.ReduceOp.SUM)      # disjoint shards: sum == union
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)               # the bench's max-over-ranks pattern
    if rank == 0:
        ret["payload"] = payload.numpy().copy()
        ret["max_count"] = float(t.item())
    dist.destroy_process_group()


def test_two_rank_partition_matches_single_rank():
    from oracle import pouch_oracle as O
    O.build()
    T, steps = 400, [0, 200, 399]
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, T, steps, ret), nprocs=2, join=True)
        payload, max_count = ret["payload"], ret["max_count"]
    want = _digests_for(range(6), T, steps)
    for i in range(6):
        assert np.array_equal(payload[i].view(np.uint64), want[i]), i
    assert max_count == 3.0
