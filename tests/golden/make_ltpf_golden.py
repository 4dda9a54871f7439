"""LTPF vectors from the unmodified reference (ia_core_coder_usac_ltpf_process via oracle/ref_harness_cpe.c) into
tests/golden/ltpf_golden.npz: per channel a 12-frame side-info sequence, input and output."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ltpf_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()


def run_reference(side, x, sample_rate=48000):
    h = ref.ref_ltpf_create(sample_rate)
    y = x.copy()
    for f in range(len(side)):
        ref.ref_ltpf_process(h, int(side[f, 0]), int(side[f, 1]), int(side[f, 2]), y[f])
    ref.ref_ltpf_destroy(h)
    return y


if __name__ == "__main__":
    out = {}
    h = ref.ref_ltpf_create(48000)
    cfg = np.zeros(5, np.int32)
    ref.ref_ltpf_cfg(h, cfg)
    ref.ref_ltpf_destroy(h)
    out["cfg_48000"] = cfg
    C, F = 8, 12
    sides, xs, ys = [], [], []
    for c in range(C):
        rng = np.random.default_rng(6600 + c)
        side, x = ltpf_cases.random_sequence(rng, F), ltpf_cases.signal(rng, F)
        sides.append(side), xs.append(x), ys.append(run_reference(side, x))
    out["side"], out["in"], out["out"] = np.stack(sides), np.stack(xs), np.stack(ys)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ltpf_golden.npz"), **out)
    print("wrote ltpf_golden.npz", cfg)
