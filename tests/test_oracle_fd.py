"""The plain-C restatement (oracle/oracle_fd.c) against (a) committed golden vectors produced by the
unmodified reference and (b) the reference build itself, bit for bit."""
import os

import numpy as np
import pytest

import fd_cases
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fd_golden.npz")


def test_oracle_vs_golden_combinations(oracle):
    g = np.load(GOLD)
    for i, s in enumerate(g["combos_side"]):
        coef = g["combos_coef"][i].copy()
        ov = g["combos_ov_in"][i].copy()
        out = np.zeros(1024, np.float32)
        assert oracle.oracle_fd_frm_dec(coef, ov, out, *[int(v) for v in s]) == 0
        assert bits_equal(out, g["combos_out"][i]), ("out", s)
        assert bits_equal(ov, g["combos_ov_out"][i]), ("overlap", s)


def test_oracle_vs_golden_chain(oracle):
    g = np.load(GOLD)
    side, coef = g["chain_side"], g["chain_coef"]
    F, C = side.shape[:2]
    ov = np.zeros((C, 1024), np.float32)
    out = np.zeros((F, C, 1024), np.float32)
    assert oracle.oracle_fd_frm_dec_batch(coef, ov, out, np.ascontiguousarray(side), F, C) == 0
    assert bits_equal(out, g["chain_out"])
    assert bits_equal(ov, g["chain_ov_final"])


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024])
def test_oracle_ifft_vs_reference(oracle, ref, n):
    rng = np.random.default_rng(n)
    for scale in (1.0, 1e4, 1e-3):
        re = (rng.normal(0, scale, n)).astype(np.float32)
        im = (rng.normal(0, scale, n)).astype(np.float32)
        a, b, c, d = re.copy(), im.copy(), re.copy(), im.copy()
        oracle.oracle_cplx_ifft_pow2(a, b, n)
        ref.ref_rad2_cplx_ifft(c, d, n)
        assert bits_equal(a, c) and bits_equal(b, d)


def test_oracle_fd_vs_reference_random_chain(oracle, ref):
    """40 frames x 4 channels, random (also illegal) sequence transitions, 30 % kernel switching."""
    rng = np.random.default_rng(77)
    F, C = 40, 4
    for legal in (True, False):
        side = fd_cases.random_side(rng, F, C, kernel_switch_prob=0.3, legal=legal)
        coef = fd_cases.random_coef(rng, (F, C, 1024))
        ov_a = np.zeros((C, 1024), np.float32)
        ov_b = np.zeros((C, 1024), np.float32)
        out_a = np.zeros((F, C, 1024), np.float32)
        out_b = np.zeros((F, C, 1024), np.float32)
        h = ref.ref_fd_create()
        assert ref.ref_fd_frm_dec_batch(h, coef, ov_b, out_b, side, F, C) == 0
        ref.ref_fd_destroy(h)
        assert oracle.oracle_fd_frm_dec_batch(coef, ov_a, out_a, side, F, C) == 0
        assert bits_equal(out_a, out_b) and bits_equal(ov_a, ov_b)


def test_oracle_fd_edge_inputs(oracle, ref):
    """Zero spectra (signed zeros), denormals and full-scale values keep bit parity."""
    h = ref.ref_fd_create()
    cases = [np.zeros(1024, np.float32), -np.zeros(1024, np.float32), np.full(1024, 1e-41, np.float32),
             np.full(1024, 3.0e7, np.float32), np.where(np.arange(1024) % 2, -32768.0, 32767.0).astype(np.float32)]
    for coef in cases:
        for s in fd_cases.all_combinations()[::7]:
            ov0 = np.linspace(-1, 1, 1024).astype(np.float32)
            ova, ovb = ov0.copy(), ov0.copy()
            oa, ob = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
            oracle.oracle_fd_frm_dec(coef.copy(), ova, oa, *[int(v) for v in s])
            ref.ref_fd_frm_dec(h, coef, ovb, ob, *[int(v) for v in s])
            assert bits_equal(oa, ob) and bits_equal(ova, ovb), s
    ref.ref_fd_destroy(h)
