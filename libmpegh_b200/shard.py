"""Stream-index partitioning across GPUs.

Streams are independent (all state is per decoder handle), so the path shards with no data-path collective:
rank r of `world` owns one contiguous block of stream indices and keeps those streams' device state.  The only
cross-rank communication is bookkeeping (barriers / max-over-ranks timing in bench.py)."""


def stream_range(rank: int, world: int, n_streams: int):
    """Contiguous [begin, end) block of rank `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world <= 0 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad rank/world/n_streams")
    base, extra = divmod(n_streams, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def owner_of(stream: int, world: int, n_streams: int) -> int:
    """Rank that owns stream index `stream` under stream_range()."""
    if not (0 <= stream < n_streams):
        raise ValueError("stream index out of range")
    base, extra = divmod(n_streams, world)
    split = extra * (base + 1)
    if stream < split:
        return stream // (base + 1)
    return extra + (stream - split) // base
