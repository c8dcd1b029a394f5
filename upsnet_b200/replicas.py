"""Multi-GPU inference = independent replicas: one process per GPU, full model
replica, image i -> rank i mod world, NO data-path collective.  torch.distributed is used only for
the barrier around a timed region and for the max-over-ranks reduction of its duration.
(The reference is single-process / thread-per-GPU and re-broadcasts all parameters every forward:
lib/utils/data_parallel.py:103-116.)"""
import torch
import torch.distributed as dist


def image_indices_for_rank(num_images, rank, world):
    """Round-robin partition used by upsnet_end2end_test.py:224-239 (one sample per GPU per step)."""
    return list(range(rank, num_images, world))


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (device timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks(value, device=None):
    """The python float of every rank, in rank order (per-rank timings: names the straggler of a scaling run)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.zeros(dist.get_world_size(), dtype=torch.float64, device=device)
    t[dist.get_rank()] = float(value)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def gather_counts(local_count, device=None):
    """Per-rank processed-unit counts (whole-job throughput = sum / max time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(local_count)]
    t = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device)
    t[dist.get_rank()] = int(local_count)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]
