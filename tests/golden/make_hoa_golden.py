"""Dumps the reference-designed HOA render matrices and NFC coefficients (impeghd_hoa_ren_renderer_init,
impeghd_hoa_ren_nfc_init) for the test configurations, plus rendered frames from
impeghd_hoa_ren_block_process, into tests/golden/hoa_ren_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hoa_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
out = {}
for cfg in hoa_cases.CONFIGS:
    cicp, order, lfe, radius = cfg
    err = ctypes.c_int()
    h = ref.ref_hoa_ren_create(cicp, order, lfe, radius, 48000, ctypes.byref(err))
    assert err.value == 0, cfg
    M, nfc, n_out = loader.ref_hoa_ren_tables(ref, h, order)
    k = hoa_cases.key(cfg)
    out[k + "_M"], out[k + "_nfc"] = M, nfc
    if cfg in hoa_cases.CONFIGS[:3]:
        rng = np.random.default_rng(2345 + cicp + order)
        x = hoa_cases.hoa_signal(rng, 3, M.shape[1])
        y = np.zeros((3, n_out, 1024), np.float32)
        for f in range(3):
            assert ref.ref_hoa_ren_process(h, x[f].reshape(-1), M.shape[1], 1024, 1, y[f].reshape(-1)) == 0
        out[k + "_in"], out[k + "_out"] = x, y
    ref.ref_hoa_ren_destroy(h)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hoa_ren_golden.npz"), **out)
print("wrote hoa_ren_golden.npz", len(out))
