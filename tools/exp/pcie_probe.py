"""PCIe probe: 25 MB H2D alone, D2H alone, both at once on two streams (pinned host memory)."""
import time
import torch

n = 6144 * 1024
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, device="cuda")
d_b = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def chunked(k=8):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1):
            d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)


mb = n * 4 / 1e6
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both), ("both, 8 chunks", chunked)):
    ms = run(fn)
    print(f"{name:16s} {ms:.3f} ms  ({mb / ms:.1f} GB/s per direction)")


def dep_chunks(k):
    """H2D chunks on s1; D2H chunk i on s2 after H2D chunk i (event) -- the fd_execute pattern without kernels"""
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1):
                d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s1)
            s2.wait_event(ev)
            with torch.cuda.stream(s2):
                h_out[i * c:(i + 1) * c].copy_(d_a[i * c:(i + 1) * c], non_blocking=True)
    return f


def lanes(k, n_lanes):
    """chunk i entirely on stream i % n_lanes: H2D then D2H in stream order, no events"""
    c = n // k
    ss = [torch.cuda.Stream() for _ in range(n_lanes)]
    def f():
        for i in range(k):
            with torch.cuda.stream(ss[i % n_lanes]):
                d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
                h_out[i * c:(i + 1) * c].copy_(d_a[i * c:(i + 1) * c], non_blocking=True)
    return f


for k in (2, 4, 8, 16):
    print(f"dep chunks k={k:2d}     {run(dep_chunks(k)):.3f} ms")
for k, l in ((8, 2), (8, 3), (8, 4), (16, 4), (4, 2)):
    print(f"lanes k={k:2d} l={l}      {run(lanes(k, l)):.3f} ms")
