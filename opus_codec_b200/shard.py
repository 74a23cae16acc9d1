"""Multi-GPU sharding of a stream batch: streams are independent units, so rank g of G owns the contiguous
index range [g*S/G, (g+1)*S/G) for its whole life and there is NO data-path collective.  torch.distributed is used only
for the barrier around timed regions and the max-over-ranks of device timings."""


def stream_range(rank, world, n_streams):
    """Contiguous, balanced partition: returns (first, count) of rank's streams."""
    base, rem = divmod(n_streams, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def max_over_ranks(value_ms, device=None):
    """MAX-reduce a timing over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def aggregate_throughput(units_per_rank, world, ms_max):
    """Whole-job throughput for weak scaling: every rank processed units_per_rank units in at most ms_max milliseconds."""
    return units_per_rank * world / (ms_max / 1000.0)
