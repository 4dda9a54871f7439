"""Dumps the reference object renderer's layout tables (after impeghd_obj_renderer_dec_init) for the CICP
layouts the tests use, plus a few rendered frames, into tests/golden/obj_layouts.npz / obj_golden.npz.
These tables are what a maintainer hands to mpeghb200_obj_layout_create; the product never
recomputes the triangulation."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import obj_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
lay, gold = {}, {}
for cicp in obj_cases.CICPS:
    err = ctypes.c_int()
    h = ref.ref_obj_create(cicp, ctypes.byref(err))
    assert err.value == 0, cicp
    for k, v in loader.ref_obj_layout_arrays(ref, h).items():
        lay[f"cicp{cicp}_{k}"] = v
    if cicp in (6, 19):
        rng = np.random.default_rng(4567 + cicp)
        F, O = 4, 7
        params = obj_cases.random_params(rng, F, O)
        x = obj_cases.object_audio(rng, F, O)
        n_spk = int(lay[f"cicp{cicp}_counts"][0])
        out = np.zeros((F, n_spk, 1024), np.float32)
        gains = np.zeros((F, O, 44), np.float32)
        for f in range(F):
            assert ref.ref_obj_render(h, params[f].reshape(-1), O, x[f].reshape(-1), out[f].reshape(-1),
                                      gains[f].reshape(-1)) == 0
        gold[f"cicp{cicp}_params"], gold[f"cicp{cicp}_in"] = params, x
        gold[f"cicp{cicp}_out"], gold[f"cicp{cicp}_gains"] = out, gains
    ref.ref_obj_destroy(h)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "obj_layouts.npz"), **lay)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "obj_golden.npz"), **gold)
print("wrote obj_layouts.npz, obj_golden.npz")
