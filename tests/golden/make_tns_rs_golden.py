"""TNS and resampler vectors from the unmodified reference (ia_core_coder_tns_apply, impeghd_resample) plus the
reference's resampler filter ROM, into tests/golden/tns_rs_golden.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tns_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
out = {}
f2, f3 = np.zeros(60, np.float32), np.zeros(60, np.float32)
ref.ref_rs_filters(f2, f3)
out["rs_filter_2"], out["rs_filter_3"] = f2, f3
for up, down in ((2, 1), (3, 1), (3, 2)):
    rng = np.random.default_rng(up * 10 + down)
    h = ref.ref_rs_create(up, down)
    x = rng.normal(0, 5000, (3, 1024)).astype(np.float32)
    y = np.zeros((3, 1024 * up // down), np.float32)
    for f in range(3):
        ref.ref_rs_process(h, x[f], y[f])
    ref.ref_rs_destroy(h)
    out[f"rs_{up}_{down}_in"], out[f"rs_{up}_{down}_out"] = x, y
for islong in (1, 0):
    rng = np.random.default_rng(500 + islong)
    sfb = tns_cases.SFB_LONG if islong else tns_cases.SFB_SHORT
    tmax = tns_cases.TNS_MAX_LONG if islong else tns_cases.TNS_MAX_SHORT
    N = 12
    xs, ys, nf, fl = [], [], [], []
    for _ in range(N):
        n_filt, filt = tns_cases.random_tns(rng, islong)
        x = tns_cases.spectrum(rng)
        y = x.copy()
        assert ref.ref_tns_apply(y, islong, n_filt, filt.reshape(-1), sfb, len(sfb), tmax, len(sfb)) == 0
        xs.append(x), ys.append(y), nf.append(n_filt), fl.append(filt)
    tag = "long" if islong else "short"
    out[f"tns_{tag}_in"], out[f"tns_{tag}_out"] = np.stack(xs), np.stack(ys)
    out[f"tns_{tag}_nfilt"], out[f"tns_{tag}_filt"] = np.stack(nf), np.stack(fl)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tns_rs_golden.npz"), **out)
print("wrote tns_rs_golden.npz", len(out))
