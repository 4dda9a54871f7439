import os, sys
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
st = torch.cuda.current_stream().cuda_stream
print("ptrs", hex(coef.data_ptr()), hex(ov.data_ptr()), hex(out.data_ptr()), hex(side.data_ptr()), st, flush=True)
for n in [6144, 3072, 444, 7, 6, 5, 4, 3, 2, 1]:
    try:
        lib.fd_synth_dev(coef.data_ptr(), side.data_ptr(), ov.data_ptr(), out.data_ptr(), n, st)
        torch.cuda.synchronize()
        print("n", n, "ok", flush=True)
    except Exception as e:
        print("n", n, "FAILED", str(e)[:100], flush=True)
        break
