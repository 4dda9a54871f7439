"""Dumps the reference-designed format-converter downmix tensors (impeghd_format_conv_init_data, or the VBAP
fallback of impeghd_format_conv_init for layouts the rule table does not cover) for the test layout pairs, plus
converted frames from impeghd_format_converter_process, into tests/golden/fc_golden.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fc_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
out = {}
for i, pair in enumerate(fc_cases.PAIRS):
    h, D = loader.ref_fc_new(ref, *pair)
    k = fc_cases.key(pair)
    out[k + "_D"] = D
    if i < 3:
        rng = np.random.default_rng(4567 + i)
        x = fc_cases.signal(rng, 3, D.shape[1])
        y = np.zeros((3, D.shape[0], 1024), np.float32)
        for f in range(3):
            assert ref.ref_fc_process(h, x[f].reshape(-1), y[f].reshape(-1)) == 0
        out[k + "_in"], out[k + "_out"] = x, y
    ref.ref_fc_destroy(h)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fc_golden.npz"), **out)
print("wrote fc_golden.npz", len(out))
