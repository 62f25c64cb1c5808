"""Work partitioning for multi-GPU runs (host logic only; no device code).

Two modes:
* view sharding -- one depth map, source views split over ranks, one all-reduce per cascade stage;
* replicas      -- independent reference images per rank, no collective (what the reference does
  with SLURM array jobs, scripts/submit_depthmap.py:35-75).
"""


def view_range(n_views: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of source views owned by ``rank`` (first ranks get the extra)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_views, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def replica_range(n_items: int, rank: int, world: int):
    """Reference images owned by ``rank`` in replica mode."""
    return view_range(n_items, rank, world)
