"""Times mpeghb200_fd_synth_frames_dev for F = 1, 2, 4, 8 on configs[1]'s shape (256 streams x 24 ch), inputs rotated
over groups so that the working set exceeds the L2.  Prints channel-samples/s and GB/s of the F-dependent
algorithmic traffic (8192 + 8192 / F bytes per channel-frame)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from libmpegh_b200.lib import Lib
import fd_cases

U = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
lib = Lib().create(0)
st = torch.cuda.current_stream()
rng = np.random.default_rng(1)
for mixed in (False, True):
    for F in (1, 2, 4, 8, 16):
        G = max(2, 32 // F)
        coef = [torch.randn(F, U, 1024, device="cuda") * 500 for _ in range(G)]
        out = [torch.empty(F, U, 1024, device="cuda") for _ in range(G)]
        ovl = [torch.zeros(U, 1024, device="cuda") for _ in range(G)]
        sides = []
        for g in range(G):
            seq = np.zeros((F, U), np.int32)
            if mixed:
                r = rng.random((F, U))
                seq[r > 0.92] = 1; seq[r > 0.95] = 2; seq[r > 0.98] = 3
            sh = (rng.random((F, U)) < 0.5).astype(np.int32)
            z = np.zeros((F, U), np.int32)
            sides.append(torch.from_numpy(fd_cases.pack_side(np.stack([seq, sh, sh, z, z], -1)).astype(np.int32)).cuda())
        def step(i):
            g = i % G
            lib.fd_synth_frames_dev(coef[g].data_ptr(), sides[g].data_ptr(), ovl[g].data_ptr(), out[g].data_ptr(), U, F,
                                    st.cuda_stream)
        n = max(200, 4000 // F)
        for i in range(50): step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(n): step(i)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        cs = U * F * 1024 / (ms * 1e-3)
        print(f"mixed={int(mixed)} F={F:2d} U={U}: {ms*1e3:8.1f} us/launch  {cs/1e9:7.1f} G ch-samples/s  "
              f"{U*F*(8192+8192/F)/(ms*1e-3)/1e9:7.0f} GB/s algorithmic(F)  {U*F*16384/(ms*1e-3)/1e9:7.0f} GB/s at 16 KB/unit",
              flush=True)
        del coef, out, ovl, sides
        torch.cuda.empty_cache()
lib.close()
