#!/usr/bin/env python3
"""Golden vectors for the sub-band DRC gain application: outputs of the UNMODIFIED reference build (oracle/_ref,
needs /root/reference) on the seeded cases of tests/drc_cases.py.
Run from the repo root: python tests/golden/make_drc_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import drc_cases  # noqa: E402
from oracle import loader  # noqa: E402

SEED, N = 4242, 8
ref = loader.load_ref()
assert ref is not None, "reference build missing"
out = {"seed": SEED, "n_cases": N}
rng = np.random.default_rng(SEED)
for i in range(N):
    fr = drc_cases.random_frame(rng)
    out[f"re_{i}"], out[f"im_{i}"] = drc_cases.run_oracle(ref.ref_drc_gains_subband, fr, (0, 0))
# the fused chain (domain switcher t2f -> gains -> f2t), three frames through the reference
out["chain_seed"], out["chain_frames"] = 99, 3
win = np.ascontiguousarray(loader.golden_rom()["sine_256"], np.float32)
out["chain_out"] = drc_cases.stft_chain(None, win, np.random.default_rng(99), 3, ref=ref)
# time-domain multi-band DRC through the reference (static functions: the -Dstatic= build)
ref_ns = loader.load_ref_ns()
out["td_seed"], out["td_cases"] = 3100, 4
for i in range(4):
    rng2 = np.random.default_rng(3100 + i)
    setup = drc_cases.random_td_frame_setup(rng2)
    frames = drc_cases.td_frames(rng2, setup, 2)
    out[f"td_out_{i}"] = drc_cases.td_reference(ref_ns, setup, frames)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "drc_golden.npz"), **out)
print("wrote drc_golden.npz")
