"""HOA renderer restatement (oracle/oracle_hoa.c) against the unmodified reference and its golden vectors."""
import ctypes
import os

import numpy as np
import pytest

import hoa_cases
from oracle import loader
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "hoa_ren_golden.npz")


@pytest.mark.parametrize("cfg", hoa_cases.CONFIGS, ids=hoa_cases.key)
def test_render_vs_reference(ref, cfg):
    cicp, order, lfe, radius = cfg
    err = ctypes.c_int()
    h = ref.ref_hoa_ren_create(cicp, order, lfe, radius, 48000, ctypes.byref(err))
    assert err.value == 0
    M, nfc, n_out = loader.ref_hoa_ren_tables(ref, h, order)
    g = np.load(GOLD)
    assert bits_equal(M, g[hoa_cases.key(cfg) + "_M"]) and bits_equal(nfc, g[hoa_cases.key(cfg) + "_nfc"])
    assert M.shape == (n_out, (order + 1) ** 2)
    rng = np.random.default_rng(order * 7 + cicp)
    x = hoa_cases.hoa_signal(rng, 5, M.shape[1])
    state = np.zeros((M.shape[1], 14), np.float32)
    clipped = False
    for f in range(5):
        want = np.zeros((n_out, 1024), np.float32)
        assert ref.ref_hoa_ren_process(h, x[f].reshape(-1), M.shape[1], 1024, 1, want.reshape(-1)) == 0
        got = loader.oracle_hoa_render(M, x[f], nfc if radius > 0 else None, state)
        assert bits_equal(got, want), (cfg, f)
        clipped |= bool((want == 32767.0).any() or (want == -32768.0).any())
    assert clipped
    ref.ref_hoa_ren_destroy(h)


def test_hoa_golden(oracle):
    g = np.load(GOLD)
    for cfg in hoa_cases.CONFIGS[:3]:
        k = hoa_cases.key(cfg)
        M, nfc, x, y = g[k + "_M"], g[k + "_nfc"], g[k + "_in"], g[k + "_out"]
        state = np.zeros((M.shape[1], 14), np.float32)
        for f in range(x.shape[0]):
            assert bits_equal(loader.oracle_hoa_render(M, x[f], nfc if cfg[3] > 0 else None, state), y[f]), (k, f)
