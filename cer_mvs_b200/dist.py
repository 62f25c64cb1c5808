"""Work partitioning for multi-GPU runs (host logic only; no device code).

Two modes:
* sharded build -- one depth map, the (source view, hypothesis) units of the cost-volume build split over ranks, one
  all-reduce per cascade stage (works for more ranks than views: BASELINE configs[4], 7 views on 8 GPUs);
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


def unit_range(n_units: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of (view, hypothesis) build units owned by ``rank``; view-major order
    (unit = view * D + hypothesis), so a rank touches as few source views as possible."""
    return view_range(n_units, rank, world)


def views_of_units(unit_begin: int, unit_end: int, D: int):
    """[view_begin, view_end) touched by the unit run (empty run -> (0, 0))."""
    if unit_end <= unit_begin:
        return 0, 0
    return unit_begin // D, (unit_end - 1) // D + 1
