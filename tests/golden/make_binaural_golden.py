"""Runs the unmodified reference binaural renderer (oracle/_ref: impeghd_binaural_renderer_init +
impeghd_binaural_struct_process_ld_ola fed 1024 samples per call) on two seeded cases and stores filters,
inputs and outputs in tests/golden/binaural_golden.npz.  Case a goes through the reference's own init."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bin_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
out = {}
for tag, (n_in, length, nb, via_init, T) in {"a": (5, 1024, 1, 1, 4), "b": (24, 4096, 1, 0, 2)}.items():
    rng = np.random.default_rng(3456 + n_in)
    f = bin_cases.with_cutoffs(rng, bin_cases.brir_set(rng, n_in, length, nb))
    err = ctypes.c_int()
    h = ref.ref_bin_create(n_in, length, nb, f["taps_direct"].reshape(-1), f["taps_diffuse"].reshape(-1),
                           f["direct_fc"].reshape(-1).copy(), np.ascontiguousarray(f["diffuse_fc"]).reshape(-1),
                           f["weight"], via_init, ctypes.byref(err))
    assert err.value == 0
    info = np.zeros(8, np.int32)
    ref.ref_bin_info(h, info)
    B = int(info[2])
    x = bin_cases.signal(rng, T, n_in, B)
    stream = np.concatenate(list(x), axis=1)
    y, buf = [], np.zeros((2, 16384), np.float32)
    for c in range(stream.shape[1] // 1024):
        n = ref.ref_bin_process(h, np.ascontiguousarray(stream[:, c * 1024:(c + 1) * 1024]).reshape(-1), 1024,
                                buf.reshape(-1), 16384)
        if n:
            y.append(buf[:, :n].copy())
    ref.ref_bin_destroy(h)
    for k, v in f.items():
        out[f"{tag}_{k}"] = v
    out[f"{tag}_in"], out[f"{tag}_out"] = x, np.stack(y)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "binaural_golden.npz"), **out)
print("wrote binaural_golden.npz", {k: v.shape for k, v in out.items()})
