"""Sub-band DRC gains: oracle restatement of impd_drc_apply_gains_subband (decoder/impd_drc_process.c:243) against
the reference build and the committed golden vectors."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import drc_cases  # noqa: E402
from oracle import loader  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "drc_golden.npz")


@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_reference(seed):
    r = loader.load_ref()
    if r is None:
        pytest.skip("reference build not available")
    o = loader.load_oracle()
    rng = np.random.default_rng(700 + seed)
    for _ in range(8):
        fr = drc_cases.random_frame(rng)
        a = drc_cases.run_oracle(o.oracle_drc_gains_subband, fr)
        # where the gains sit inside lpcm_gains must not matter: both delay modes, some gain delay
        b = drc_cases.run_oracle(r.ref_drc_gains_subband, fr, (int(rng.integers(0, 2)), int(rng.integers(0, 400))))
        assert loader.bits_equal(a[0], b[0]) and loader.bits_equal(a[1], b[1])


def test_oracle_matches_golden():
    o = loader.load_oracle()
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["seed"]))
    for i in range(int(g["n_cases"])):
        fr = drc_cases.random_frame(rng)
        re, im = drc_cases.run_oracle(o.oracle_drc_gains_subband, fr)
        assert loader.bits_equal(re, g[f"re_{i}"]) and loader.bits_equal(im, g[f"im_{i}"])
