#!/usr/bin/env python3
"""Per-stage throughput of the kernels beyond the headline FD synthesis, at the BASELINE.json config shapes
(one GPU).  Prints one JSON line per stage: ms per step, algorithmic GB/s (SURVEY.md 8d byte counts) and the
fraction of the measured HBM peak.  Diagnostic companion of bench.py; numbers go to profiles/."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libmpegh_b200.lib import Lib  # noqa: E402
from oracle import loader  # noqa: E402  (layout / matrix fixtures only)

lib = Lib().create(0)
st = torch.cuda.current_stream().cuda_stream
peak = 6549.4
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, steps=30, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(steps):
        flush.zero_()  # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def report(name, ms, bytes_per_step, units, unit_name, extra=None):
    gbs = bytes_per_step / (ms * 1e-3) / 1e9
    line = {"stage": name, "ms_per_step": ms, "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak,
            unit_name + "_per_s": units / (ms * 1e-3)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


# ---- config 5: 32 objects -> 7.1.4 (512 streams per GPU) + limiter + 24-bit PCM --------------------------
S, O = 512, 32
lay = lib.obj_layout_create(loader.obj_layout_golden(19))
n_spk = 12
rng = np.random.default_rng(4567)
import obj_cases  # noqa: E402
params = obj_cases.random_params(rng, 2, S * O, spread_prob=0.0).reshape(2, S, O, 7)
sides = [torch.from_numpy(lib.obj_side_from_polar(params[k]).view(np.float32).reshape(S, O, 16).copy()).cuda() for k in range(2)]
x = torch.randn(S, O, 1024, device="cuda") * 3000
out = torch.empty(S, n_spk, 1024, device="cuda")
state = torch.empty(S * lib.obj_state_floats(lay, O), device="cuda")
scratch = torch.empty(S * lib.obj_scratch_floats(lay, O), device="cuda")
lib.obj_state_init_dev(lay, O, state.data_ptr(), S)
ms = timeit(lambda i: lib.obj_render_dev(lay, O, sides[i % 2].data_ptr(), x.data_ptr(), out.data_ptr(), state.data_ptr(),
                                         scratch.data_ptr(), S, st))
report("obj_render 32 obj -> 7.1.4, 512 streams", ms, S * (O * 4096 + n_spk * 4096), S * O * 1024, "object_samples",
       {"fp32_unfused_Gops": S * O * 11 * 1024 * 3 / (ms * 1e-3) / 1e9})

p = lib.limiter_params(n_spk)
lst = torch.empty(S * lib.limiter_state_floats(p), device="cuda")
lib.limiter_state_init_dev(p, lst.data_ptr(), S)
lim_out = torch.empty_like(out)
out.normal_(0, 9000.0)
ms = timeit(lambda i: lib.limiter_dev(p, out.data_ptr(), lim_out.data_ptr(), lst.data_ptr(), S, None, st))
report("limiter 12 ch, 512 streams (engaged)", ms, S * (2 * n_spk * 4096 + 2 * 240 * (n_spk + 1) * 4), S * n_spk * 1024,
       "channel_samples")
out.normal_(0, 300.0)
lib.limiter_state_init_dev(p, lst.data_ptr(), S)
ms = timeit(lambda i: lib.limiter_dev(p, out.data_ptr(), lim_out.data_ptr(), lst.data_ptr(), S, None, st))
report("limiter 12 ch, 512 streams (quiet: smoother skipped)", ms, S * (2 * n_spk * 4096 + 2 * 240 * (n_spk + 1) * 4),
       S * n_spk * 1024, "channel_samples")
pcm = torch.empty(S, 1024 * n_spk * 3, dtype=torch.uint8, device="cuda")
ms = timeit(lambda i: lib.pcm_pack_dev(lim_out.data_ptr(), n_spk, 24, pcm.data_ptr(), S, None, None, None, st))
report("pcm_pack 24-bit 12 ch, 512 streams", ms, S * (n_spk * 4096 + n_spk * 1024 * 3), S * n_spk * 1024, "channel_samples")

# ---- config 3: HOA order 4 (25 coeffs) + NFC -> 22.2, 1024 streams --------------------------------------
import hoa_cases  # noqa: E402
g = np.load(os.path.join(ROOT, "tests", "golden", "hoa_ren_golden.npz"))
k = hoa_cases.key(hoa_cases.CONFIGS[0])
for use_nfc in (1, 0):
    S = 1024
    h = lib.hoa_renderer_create(g[k + "_M"], g[k + "_nfc"] if use_nfc else None)
    hin = torch.randn(S, 25, 1024, device="cuda") * 2000
    hout = torch.empty(S, 24, 1024, device="cuda")
    hst = torch.zeros(S * 25 * 14, device="cuda")
    ms = timeit(lambda i: lib.hoa_render_dev(h, hin.data_ptr(), hout.data_ptr(), hst.data_ptr(), S, st))
    report(f"hoa_render order 4 -> 22.2, 1024 streams, nfc={use_nfc}", ms, S * (25 + 24) * 4096, S * 24 * 1024,
           "output_channel_samples", {"fp32_unfused_Gops": S * 24 * 25 * 1024 * 2 / (ms * 1e-3) / 1e9})
    lib.hoa_renderer_destroy(h)
lib.close()
