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


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_cpus(pci_bus_id: str):
    """CPUs of the NUMA node a GPU hangs on (from sysfs), or None when the platform does not say.
    pci_bus_id as nvidia-smi / CUDA print it, e.g. '00000000:1B:00.0'."""
    import os
    dom, rest = pci_bus_id.lower().split(":", 1)
    dev = f"{dom[-4:]}:{rest}"
    try:
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except (OSError, ValueError):
        return None
    allowed = os.sched_getaffinity(0)
    cpus = [c for c in cpus if c in allowed]
    return cpus or None


def bind_to_gpu_numa(device_index: int):
    """Pin the calling process to the CPUs next to GPU `device_index` (the host side of a
    rank -- parse threads, pinned staging buffers, which are placed by first touch -- belongs on the GPU's NUMA
    node, otherwise every rank's PCIe traffic crosses the socket interconnect).  Returns the CPU list or None."""
    import os
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
    except (OSError, subprocess.SubprocessError):
        return None
    if not out:
        return None
    cpus = gpu_numa_cpus(out.splitlines()[0].strip())
    if cpus:
        os.sched_setaffinity(0, cpus)
    return cpus
