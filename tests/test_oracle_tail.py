"""The plain-C restatement of limiter / PCM packer (oracle/oracle_tail.c) against the unmodified reference
functions (oracle/_ref, skipped where that build is absent) and against committed vectors made from it."""
import os

import numpy as np
import pytest

import tail_cases
from oracle import loader
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tail_golden.npz")


@pytest.mark.parametrize("n_ch,limiter_on", [(1, 1), (2, 1), (12, 1), (24, 1), (6, 0)])
def test_limiter_vs_reference(ref, n_ch, limiter_on):
    rng = np.random.default_rng(100 + n_ch)
    x = tail_cases.limiter_signal(rng, 24, n_ch)
    h = ref.ref_limiter_create(n_ch, 48000, limiter_on)
    st = loader.oracle_limiter_new(n_ch, 48000, limiter_on)
    c = np.zeros(3, np.float32)
    ref.ref_limiter_constants(h, c)
    assert (c[0], c[1], c[2]) == (st.attack_const, st.release_const, st.threshold)
    engaged = False
    for f in range(x.shape[0]):
        want = x[f].copy()
        ref.ref_limiter_process(h, want, 1024)
        got = loader.oracle_limiter_run(st, x[f])
        assert bits_equal(got, want), (n_ch, f)
        engaged |= st.pre_smoothed_gain < 1.0
    ref.ref_limiter_destroy(h)
    assert engaged or not limiter_on


def test_limiter_quiet_signal_is_pure_delay(oracle):
    rng = np.random.default_rng(3)
    x = (rng.normal(0, 100.0, size=(3, 4, 1024))).astype(np.float32)
    st = loader.oracle_limiter_new(4)
    y = np.concatenate([loader.oracle_limiter_run(st, x[f]) for f in range(3)], axis=1)
    xin = np.concatenate(list(x), axis=1)
    assert bits_equal(y[:, 240:], xin[:, :-240]) and not y[:, :240].any()


@pytest.mark.parametrize("pcmsize", [16, 24, 32])
def test_samples_sat_vs_reference(pcmsize):
    ns = loader.load_ref_ns()
    if ns is None:
        pytest.skip("oracle/_ref/libmpeghref_ns.so not built")
    rng = np.random.default_rng(pcmsize)
    for n_ch, delay, cmap in [(1, 0, None), (2, 240, [1, 0]), (12, 1500, None), (24, 1024, list(range(23, -1, -1))),
                              (6, 1, [-1, 2, 1, -1, -1, -1])]:
        x = tail_cases.pcm_signal(rng, n_ch)
        m = np.full(n_ch, -1, np.int32) if cmap is None else np.array(cmap, np.int32)
        d = np.array([delay], np.int32)
        want = np.zeros(n_ch * 1024 * 4, np.uint8)
        nb = ns.ref_samples_sat(x, 1024, n_ch, pcmsize, m, d, want)
        got, d2 = loader.oracle_samples_sat(x, pcmsize, cmap, delay)
        assert nb == got.size and d2 == int(d[0])
        assert np.array_equal(got, want[:nb]), (pcmsize, n_ch)


def test_tail_golden(oracle):
    """Vectors produced by tests/golden/make_tail_golden.py from the reference build."""
    g = np.load(GOLD)
    st = loader.oracle_limiter_new(int(g["lim_in"].shape[1]))
    for f in range(g["lim_in"].shape[0]):
        assert bits_equal(loader.oracle_limiter_run(st, g["lim_in"][f]), g["lim_out"][f]), f
    for bits in (16, 24, 32):
        got, _ = loader.oracle_samples_sat(g["pcm_in"], bits, g["pcm_map"], int(g["pcm_delay"]))
        assert np.array_equal(got, g[f"pcm_out{bits}"])
