"""Kernel time vs number of units (latency vs throughput regime). Diagnostic only."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libmpegh_b200.lib import Lib
lib = Lib().create(0)
G = 6
U = 6144
coef = [torch.randn(U, 1024, device="cuda") * 500 for _ in range(G)]
ov = [torch.zeros(U, 1024, device="cuda") for _ in range(G)]
out = [torch.empty(U, 1024, device="cuda") for _ in range(G)]
side = torch.zeros(U, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for n in [int(x) for x in sys.argv[1:]] or [3, 148 * 3, 148 * 6, 148 * 12, 148 * 24, 148 * 42, 6144]:
    for i in range(5):
        lib.fd_synth_dev(coef[i % G].data_ptr(), side.data_ptr(), ov[i % G].data_ptr(), out[i % G].data_ptr(), n, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 60
    e0.record()
    for i in range(K):
        lib.fd_synth_dev(coef[i % G].data_ptr(), side.data_ptr(), ov[i % G].data_ptr(), out[i % G].data_ptr(), n, st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / K
    print(f"n_units {n:6d}: {us:8.2f} us/launch  {n * 16384 / us / 1e3:8.1f} GB/s")
