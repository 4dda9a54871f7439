"""Dumps the tables the reference's impeghd_hoa_spatial_init designs (mode matrices, fade windows, gain-correction
window powers) for the test configurations, plus decoded frames of seeded side-info sequences from the
reference's stage functions, into tests/golden/hoasp_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hoasp_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
out = {}
shared_done = False
for ci, cfg in enumerate(hoasp_cases.CONFIGS):
    err = ctypes.c_int()
    h = ref.ref_hoasp_create(hoasp_cases.cfg_array(cfg), ctypes.byref(err))
    assert err.value == 0
    t = loader.ref_hoasp_tables(ref, h)
    k = hoasp_cases.key(cfg)
    nc, low = int(t["info"][2]), int(t["info"][0])
    out[k + "_info"] = t["info"]
    out[k + "_order_matrix"] = t["order_matrix"][:low * low].copy()
    out[k + "_fine"] = t["fine"].reshape(900, 49)[:, :nc].copy()       # columns beyond n_coef are never read
    out[k + "_coarse"] = t["coarse"][:nc * nc].copy()
    out[k + "_wins"] = t["wins"]
    if not shared_done:
        out["dyn_pow"], out["dyn_win"], out["pow2"] = t["dyn_pow"], t["dyn_win"], t["pow2"]
        shared_done = True
    if ci < 3:
        rng = np.random.default_rng(2345 + ci)
        gen = hoasp_cases.SideGen(cfg, rng)
        F = 6
        x = hoasp_cases.signals(rng, F, cfg[1])
        sides = np.zeros(F, hoasp_cases.SIDE_DTYPE)
        y = np.zeros((F, nc, 1024), np.float32)
        for f in range(F):
            sides[f] = gen.next()
            assert ref.ref_hoasp_process(h, sides[f:f + 1].ctypes.data, x[f].reshape(-1), y[f].reshape(-1)) == 0
        out[k + "_sides"] = sides.view(np.uint8).reshape(F, -1)
        out[k + "_in"], out[k + "_out"] = x, y
    ref.ref_hoasp_destroy(h)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hoasp_golden.npz"), **out)
print("wrote hoasp_golden.npz", len(out))
