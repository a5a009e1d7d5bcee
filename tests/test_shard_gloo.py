"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): partition / scatter / gather of the sharding layer.
The per-rank solve is stood in by the oracle (the CUDA kernel needs a GPU); what is under test is that any
partition reproduces the single-process result bitwise (T5) including ragged and empty shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qppvm_b200 import gen, shard
from qppvm_b200.layout import CONFIGS, layout


def test_partition_covers_batch_exactly():
    for batch in (0, 1, 7, 64, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.partition(batch, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == batch
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert shard.partition(1 << 20, 8, 3) == (3 * 131072, 131072)      # 131 072 per GPU at G = 8


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    desc = CONFIGS[1]["desc"]
    L = layout(desc)
    recs = torch.from_numpy(gen.generate(desc, batch, 123)) if rank == 0 else None

    def solve(r):
        if r.shape[0] == 0:
            return torch.empty((0, L.out_doubles), dtype=torch.float64)
        return torch.from_numpy(oracle.solve_batch(desc, r.numpy(), threads=1)[0])
    full = shard.solve_sharded(solve, recs, batch, L.rec_doubles, torch.device("cpu"))
    if rank == 0:
        ref = oracle.solve_batch(desc, recs.numpy(), threads=1)[0]
        q.put(bool(np.array_equal(full.numpy(), ref)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 37), (3, 2), (2, 64)])
def test_scatter_solve_gather_matches_single_process(oracle_mod, world, batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
