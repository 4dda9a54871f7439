#!/usr/bin/env python3
"""Where the mixed-sequence variant of the FD step loses time: the same 6144 units with only one kind of non-ONLY_LONG
frame switched in at a time.  Diagnostic, prints ms per launch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fd_cases  # noqa: E402
from libmpegh_b200.lib import Lib  # noqa: E402

lib = Lib().create(0)
U, G = 6144, 8
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)
coef = [torch.randn(U, 1024, device="cuda") * 500 for _ in range(G)]
ov = [torch.zeros(U, 1024, device="cuda") for _ in range(G)]
out = [torch.empty(U, 1024, device="cuda") for _ in range(G)]


def side(kinds):
    seq = np.zeros(U, np.int32)
    r = rng.random(U)
    for k, (lo, hi) in kinds.items():
        seq[(r > lo) & (r <= hi)] = k
    z = np.zeros(U, np.int32)
    s = np.stack([seq, (rng.random(U) < 0.5).astype(np.int32), (rng.random(U) < 0.5).astype(np.int32), z, z], -1)
    return torch.from_numpy(fd_cases.pack_side(s).astype(np.int32)).cuda()


def run(name, kinds, steps=2000):
    sd = [side(kinds) for _ in range(G)]
    for i in range(100):
        lib.fd_synth_dev(coef[i % G].data_ptr(), sd[i % G].data_ptr(), ov[i % G].data_ptr(), out[i % G].data_ptr(), U, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        lib.fd_synth_dev(coef[i % G].data_ptr(), sd[i % G].data_ptr(), ov[i % G].data_ptr(), out[i % G].data_ptr(), U, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / steps * 1000:7.2f} us", flush=True)


run("ONLY_LONG", {})
run("3% LONG_START", {1: (0.92, 0.95)})
run("3% EIGHT_SHORT", {2: (0.95, 0.98)})
run("2% LONG_STOP", {3: (0.98, 1.0)})
run("1% EIGHT_SHORT", {2: (0.97, 0.98)})
run("10% EIGHT_SHORT", {2: (0.88, 0.98)})
run("100% EIGHT_SHORT", {2: (-1.0, 1.0)})
run("mixed (bench)", {1: (0.92, 0.95), 2: (0.95, 0.98), 3: (0.98, 1.0)})
lib.close()
