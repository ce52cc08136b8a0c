"""Host-side sharding of the projection path across ranks.

The path has no exchange step: a projection (sample b, view, pose candidate)
reads one cloud and one pose and writes its own grid / mask / depth.  Ranks
therefore own disjoint SAMPLES -- all ``replicas = step_size x candidates``
projections of a sample stay on one GPU, so the reduction of point gradients
over the replicas of a cloud (autograd of ``tf_repeat_0``,
model_pc_to.py:47-56, 302-306) never crosses a GPU -- and no data-path
collective is issued.  ``torch.distributed`` is used only to fence the timed
region and to take the max of the per-rank device times.
"""
import torch


def sample_range(n_samples, rank, world):
    """Contiguous block of samples owned by ``rank`` (sizes differ by at most one,
    the first ``n_samples % world`` ranks take the extra sample)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    if n_samples < 0:
        raise ValueError("n_samples must be >= 0")
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def projection_range(n_samples, replicas, rank, world):
    """[lo, hi) over the flattened projection axis P = n_samples * replicas
    (sample-major, the layout ``tf_repeat_0`` produces)."""
    lo, hi = sample_range(n_samples, rank, world)
    return lo * replicas, hi * replicas


def shard(tensors, n_samples, replicas, rank, world):
    """Slice every [P, ...] tensor of a dict (None passes through) to this rank's
    projections.  Tensors are views: nothing is copied."""
    lo, hi = projection_range(n_samples, replicas, rank, world)
    out = {}
    for k, t in tensors.items():
        if t is None:
            out[k] = None
            continue
        if t.shape[0] != n_samples * replicas:
            raise ValueError("%s: leading dimension %d is not n_samples * replicas = %d"
                             % (k, t.shape[0], n_samples * replicas))
        out[k] = t[lo:hi]
    return out


def max_over_ranks(ms, device=None):
    """Device time of the slowest rank (identity without a process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_rate(units_per_rank_step, steps, ms_max, world):
    """Whole-job throughput: the units ALL ranks processed / the slowest rank's time."""
    return world * units_per_rank_step * steps / (ms_max * 1e-3)
