"""Generates tests/golden/tail_golden.npz from the UNMODIFIED reference build (oracle/_ref): limiter frames
(impeghd_peak_limiter_process) and packed PCM (ia_core_coder_samples_sat).  Run in the dev container."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tail_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref, ns = loader.load_ref(), loader.load_ref_ns()
rng = np.random.default_rng(20261019)
n_ch = 5
x = tail_cases.limiter_signal(rng, 10, n_ch)
h = ref.ref_limiter_create(n_ch, 48000, 1)
y = np.zeros_like(x)
for f in range(x.shape[0]):
    t = x[f].copy()
    ref.ref_limiter_process(h, t, 1024)
    y[f] = t
ref.ref_limiter_destroy(h)
out = {"lim_in": x, "lim_out": y}
p = tail_cases.pcm_signal(rng, 3)
cmap = np.array([2, 0, 1], np.int32)
out["pcm_in"], out["pcm_map"], out["pcm_delay"] = p, cmap, np.int32(7)
for bits in (16, 24, 32):
    buf = np.zeros(3 * 1024 * 4, np.uint8)
    d = np.array([7], np.int32)
    nb = ns.ref_samples_sat(p, 1024, 3, bits, cmap.copy(), d, buf)
    out[f"pcm_out{bits}"] = buf[:nb].copy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tail_golden.npz"), **out)
print("wrote tail_golden.npz", {k: v.shape for k, v in out.items()})
