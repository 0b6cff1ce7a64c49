"""world_size-2 CPU test (gloo) of the multi-GPU plumbing: shard ranges, weight broadcast, example gather, max-over-ranks."""
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


def _worker(rank, world, port, K, Q, H, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rebel_b200 import dist as D
    lo, hi = D.shard_range(rank, world, K)
    w = D.broadcast_weights(np.arange(1000, dtype=np.float32) if rank == 0 else None, 1000, "cpu")
    assert np.array_equal(w, np.arange(1000, dtype=np.float32))
    ids = np.arange(lo, hi, dtype=np.float32)
    q = np.repeat(ids[:, None], Q, 1)
    v = np.repeat(-ids[:, None], H, 1)
    res = D.gather_examples(q, v, "cpu")
    # recursive evaluation: contiguous strategy-id blocks per rank, float32 partial sums reduced to rank 0
    from rebel_b200.recursive_eval import reduce_sums, strategy_ids
    blocks = [strategy_ids(r, world, 4097) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 4097 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    ss, sr = reduce_sums(torch.full((5, 3, 4), float(rank + 1)), torch.full((5, 3, 1), 0.5), "cpu")
    if rank == 0:
        assert torch.all(ss == sum(range(1, world + 1))) and torch.all(sr == 0.5 * world)
    m = D.max_over_ranks(10.0 + rank, "cpu")
    assert m == 10.0 + world - 1
    if rank == 0:
        aq, av = res
        assert aq.shape == (world * K, Q) and av.shape == (world * K, H)
        assert np.array_equal(aq[:, 0], np.arange(world * K)) and np.array_equal(av[:, 0], -np.arange(world * K))
        open(os.path.join(out_dir, "ok"), "w").write("1")
    else:
        assert res is None
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world, K, Q, H = 2, 37, 27, 6
    mp.spawn(_worker, args=(world, _free_port(), K, Q, H, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def test_single_process_passthrough():
    from rebel_b200 import dist as D
    assert D.shard_range(3, 8, 2048) == (6144, 8192)
    q, v = D.gather_examples(np.ones((4, 2, 5), np.float32), np.ones((4, 2, 3), np.float32), "cpu")
    assert q.shape == (8, 5) and v.shape == (8, 3)
    assert D.max_over_ranks(3.5, "cpu") == 3.5
