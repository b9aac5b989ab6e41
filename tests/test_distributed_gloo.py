"""world_size-2 CPU test (gloo) of the only collective on the path: the final token gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, N, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from d3pm_b200.distributed import gather_tokens, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(B * N, dtype=torch.int64).reshape(B, N) * 7 % 4097   # the "single-GPU" result
        b, e = shard_range(B, world, rank)
        got = gather_tokens(full[b:e].clone(), B)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), got.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5])  # equal and ragged shards
def test_gather_tokens_world2(tmp_path, B):
    N, world = 16, 2
    mp.spawn(_worker, args=(world, _free_port(), B, N, str(tmp_path)), nprocs=world, join=True)
    want = (torch.arange(B * N, dtype=torch.int64).reshape(B, N) * 7 % 4097).numpy()
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)
