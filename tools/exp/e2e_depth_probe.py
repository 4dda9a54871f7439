"""e2e of configs[1] through mpeghb200_fd_submit / _wait with 1, 2 (and, when the library is built with kDepth 3, 3)
frames in flight, pinned host buffers."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libmpegh_b200.lib import Lib
import fd_cases
S, C = 256, 24
U = S * C
lib = Lib().create(0)
sess = lib.fd_open(S, C)
rng = np.random.default_rng(0)
n_host = 6
h_coef = [torch.empty(S, C, 1024).pin_memory() for _ in range(n_host)]
for h in h_coef: h.normal_(0, 500)
h_out = [torch.empty(S, C, 1024).pin_memory() for _ in range(n_host)]
side = np.zeros((U, 5), np.int32)
h_side = np.ascontiguousarray(fd_cases.pack_side(side).reshape(S, C))
def sub(i):
    k = i % n_host
    return lib.c.mpeghb200_fd_submit(sess.h, h_coef[k].data_ptr(), h_side.ctypes.data, h_out[k].data_ptr())
def wait():
    lib.check(lib.c.mpeghb200_fd_wait(sess.h))
for depth in (1, 2, 3, 4):
    # can the library hold `depth` frames?
    ok = True
    for i in range(depth):
        if sub(i) != 0: ok = False; break
    for i in range(depth): wait()
    if not ok:
        print("depth", depth, "not supported by this build"); continue
    n = 300
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(depth - 1): lib.check(sub(i))
        for i in range(depth - 1, n):
            lib.check(sub(i)); wait()
        for i in range(depth - 1): wait()
        dt = (time.perf_counter() - t0) / n
    print(f"depth {depth}: {dt*1e3:.3f} ms/step  {U*1024/dt/1e9:.2f} G ch-samples/s  {U*4096/dt/1e9:.1f} GB/s each way")
sess.close(); lib.close()
