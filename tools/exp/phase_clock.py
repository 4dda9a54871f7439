import os, sys, ctypes
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libmpegh_b200.lib import Lib
lib = Lib().create(0)
U = 6144
coef = torch.randn(U, 1024, device="cuda") * 500
ov = torch.zeros(U, 1024, device="cuda")
out = torch.empty(U, 1024, device="cuda")
side = torch.zeros(U, dtype=torch.int32, device="cuda")
big = torch.empty(64 << 20, device="cuda")
st = torch.cuda.current_stream().cuda_stream
buf = (ctypes.c_longlong * 32)()
for n in [3, 444, 6144]:
    for rep in range(3):
        if rep == 2:
            big.zero_()  # flush L2
        torch.cuda.synchronize()
        lib.fd_synth_dev(coef.data_ptr(), side.data_ptr(), ov.data_ptr(), out.data_ptr(), n, st)
        torch.cuda.synchronize()
        lib.c.mpeghb200_debug_read(buf)
        t = list(buf)
        rel = [x - t[0] for x in t]
        print(f"n={n} rep={rep} A:", rel[1:4], "B:", rel[8:11], "C:", rel[16:19], flush=True)
