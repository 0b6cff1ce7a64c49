"""Multi-GPU plumbing of the data-generation path: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  Self-play games are independent (the reference runs them as unrelated threads,
rela/context.h:48-55), so the data path has NO collective: rank r owns games / subgames [r*K, (r+1)*K).  The two exchanges
the reference performs through host memory inside one process become collectives:

* fresh value-net weights, trainer -> every generator (ModelLocker::updateModel, rela/model_locker.h:69-79):
  `broadcast_weights` — one flat fp32 buffer (~300 KB) from rank 0;
* training examples, generators -> the trainer's replay (PrioritizedReplay::add, rela/prioritized_replay.h:247-261):
  `gather_examples` — packed [n, Q+H] fp32 blocks to rank 0.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(rank, world, per_rank):
    """Global subgame / game ids owned by `rank` (weak scaling: every rank owns `per_rank` of them)."""
    assert 0 <= rank < world
    return rank * per_rank, (rank + 1) * per_rank


def _is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_weights(flat, numel, device, src=0):
    """Rank `src` passes the flat weight buffer (numpy fp32), the others pass None; everyone gets a numpy copy."""
    if not _is_dist():
        return np.ascontiguousarray(flat, np.float32)
    if dist.get_rank() == src:
        t = torch.from_numpy(np.ascontiguousarray(flat, np.float32)).to(device)
        assert t.numel() == numel
    else:
        t = torch.empty(numel, dtype=torch.float32, device=device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def gather_examples(queries, values, device, dst=0):
    """queries [n,Q], values [n,H] (numpy fp32) of this rank -> on rank `dst` the concatenation over ranks (rank order),
    elsewhere None.  Without a process group returns the inputs."""
    q = np.ascontiguousarray(queries, np.float32).reshape(len(queries), -1) if queries.ndim == 2 else queries.reshape(-1, queries.shape[-1])
    v = values.reshape(-1, values.shape[-1])
    if not _is_dist():
        return q, v
    blk = torch.from_numpy(np.concatenate([q, v], 1)).to(device)
    out = [torch.empty_like(blk) for _ in range(dist.get_world_size())] if dist.get_rank() == dst else None
    dist.gather(blk, out, dst=dst)
    if out is None:
        return None
    allb = torch.cat(out, 0).cpu().numpy()
    return allb[:, :q.shape[1]], allb[:, q.shape[1]:]


def max_over_ranks(value, device):
    if not _is_dist():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    if not _is_dist():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
