#!/usr/bin/env python3
"""Golden vectors for the multichannel coding tool: outputs of the UNMODIFIED reference build (oracle/_ref, needs
/root/reference) on the seeded cases of tests/mct_cases.py, plus the reference's angle tables.
Run from the repo root: python tests/golden/make_mct_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mct_cases  # noqa: E402
from oracle import loader  # noqa: E402

SEED, N = 20261, 6
ref = loader.load_ref()
assert ref is not None, "reference build missing"
out = {"seed": SEED, "n_cases": N}
s, c = np.zeros(65, np.float32), np.zeros(65, np.float32)
ref.ref_mct_rom(s, c)
out["sin_alpha"], out["cos_alpha"] = s, c
rng = np.random.default_rng(SEED)
for i in range(N):
    boxes = mct_cases.random_group(rng)
    x = (rng.standard_normal((24, 1024)) * 500).astype(np.float32)
    out[f"in_{i}"] = x
    out[f"out_{i}"] = mct_cases.apply_group(ref.ref_mct_process, x.copy(), boxes)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mct_golden.npz"), **out)
print("wrote mct_golden.npz")
