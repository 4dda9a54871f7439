"""Joint-stereo vectors from the unmodified reference: the MDST filter ROM of the complex-prediction tool and the
phase-2 output (time signal + saved spectra) of channel pairs over consecutive frames, through
ia_core_coder_cplx_prev_mdct_dmx + ia_core_coder_data_process (oracle/ref_harness_cpe.c), into
tests/golden/stereo_golden.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stereo_cases  # noqa: E402
import tns_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()
out = {}
curr, prev = np.zeros(4 * 2 * 2 * 7, np.float32), np.zeros(2 * 2 * 7, np.float32)
ref.ref_mdst_rom(curr, prev)
out["mdst_curr"], out["mdst_prev"] = curr.reshape(4, 2, 2, 7), prev.reshape(2, 2, 7)
for islong, want in ((1, stereo_cases.SFB_LONG), (0, stereo_cases.SFB_SHORT)):
    t = np.zeros(64, np.int16)
    n = ref.ref_cpe_sfb_table(stereo_cases.SAMPLING_RATE_IDX, islong, t, 64)
    assert np.array_equal(t[:n], want), "48 kHz sfb table differs from tns_cases"

N_PAIRS, N_FRAMES = 6, 10


def run_reference(frames, coef):
    h = ref.ref_cpe_create(stereo_cases.SAMPLING_RATE_IDX, tns_cases.TNS_MAX_LONG, tns_cases.TNS_MAX_SHORT)
    pcm, spec = np.zeros((len(frames), 2, 1024), np.float32), np.zeros((len(frames), 2, 1024), np.float32)
    for f, fr in enumerate(frames):
        err = ref.ref_cpe_process(h, fr["sd"].tobytes(), fr["grp"], fr["tns_on_lr"],
                                  np.ascontiguousarray(fr["n_filt"].reshape(-1)),
                                  np.ascontiguousarray(fr["filt"].reshape(-1)), fr["max_sfb"], coef[f, 0], coef[f, 1],
                                  pcm[f, 0], pcm[f, 1], spec[f, 0], spec[f, 1])
        assert err == 0, err
    ref.ref_cpe_destroy(h)
    return pcm, spec


if __name__ == "__main__":
    for p in range(N_PAIRS):
        rng = np.random.default_rng(7100 + p)
        frames = stereo_cases.random_pair_sequence(rng, N_FRAMES)
        coef = rng.normal(0, 500, (N_FRAMES, 2, 1024)).astype(np.float32)
        pcm, spec = run_reference(frames, coef)
        out[f"p{p}_coef"], out[f"p{p}_pcm"], out[f"p{p}_spec"] = coef, pcm, spec
        out[f"p{p}_sd"] = np.array([fr["sd"] for fr in frames], stereo_cases.STEREO_SIDE).view(np.uint8).reshape(
            N_FRAMES, -1)
        for k in ("grp", "n_filt", "filt", "max_sfb"):
            out[f"p{p}_{k}"] = np.stack([fr[k] for fr in frames])
        out[f"p{p}_tns_on_lr"] = np.array([fr["tns_on_lr"] for fr in frames], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stereo_golden.npz"), **out)
    print("wrote stereo_golden.npz", len(out))
