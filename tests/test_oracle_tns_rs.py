"""TNS and resampler restatements (oracle/oracle_tns_rs.c) against the unmodified reference."""
import numpy as np
import pytest

import tns_cases
from oracle import loader
from oracle.loader import bits_equal


@pytest.mark.parametrize("islong", [1, 0])
def test_tns_vs_reference(ref, oracle, islong):
    rng = np.random.default_rng(40 + islong)
    sfb = tns_cases.SFB_LONG if islong else tns_cases.SFB_SHORT
    tmax = tns_cases.TNS_MAX_LONG if islong else tns_cases.TNS_MAX_SHORT
    n_active = 0
    for trial in range(60):
        n_filt, filt = tns_cases.random_tns(rng, islong)
        x = tns_cases.spectrum(rng)
        nbands = int(rng.integers(len(sfb) // 2, len(sfb) + 1))
        want = x.copy()
        assert ref.ref_tns_apply(want, islong, n_filt, filt.reshape(-1), sfb, len(sfb), tmax, nbands) == 0
        got, resolved = loader.oracle_tns_apply(x, islong, n_filt, filt, sfb, tmax, nbands)
        assert bits_equal(got, want), trial
        n_active += len(resolved)
    assert n_active > 30


@pytest.mark.parametrize("up,down", [(2, 1), (3, 1), (3, 2)])
def test_resampler_vs_reference(ref, oracle, up, down):
    f2, f3 = np.zeros(60, np.float32), np.zeros(60, np.float32)
    ref.ref_rs_filters(f2, f3)
    h = ref.ref_rs_create(up, down)
    o = loader.OracleResampler(up, down, f2, f3)
    rng = np.random.default_rng(up * 10 + down)
    for f in range(5):
        x = rng.normal(0, 5000, 1024).astype(np.float32)
        want = np.zeros(1024 * up // down, np.float32)
        ref.ref_rs_process(h, x, want)
        assert bits_equal(o.process(x), want), (up, down, f)
    ref.ref_rs_destroy(h)


def test_tns_lpc_host_helper_matches_oracle(so_lib, oracle):
    """mpeghb200_tns_lpc_from_coef (host code of the product, no GPU) against the restatement, all index values."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        res = int(rng.integers(3, 5))
        order = int(rng.integers(1, 16))
        lim = 1 << (res - 1)
        coef = rng.integers(-lim, lim, 15).astype(np.int32)
        want = np.zeros(16, np.float32)
        oracle.oracle_tns_lpc(order, res, coef, want)
        got = so_lib.tns_lpc_from_coef(order, res, coef)
        assert bits_equal(got[:order], want[:order])


def test_tns_rs_golden(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tns_rs_golden.npz"))
    for tag, islong in (("long", 1), ("short", 0)):
        sfb = tns_cases.SFB_LONG if islong else tns_cases.SFB_SHORT
        tmax = tns_cases.TNS_MAX_LONG if islong else tns_cases.TNS_MAX_SHORT
        for u in range(g[f"tns_{tag}_in"].shape[0]):
            got, _ = loader.oracle_tns_apply(g[f"tns_{tag}_in"][u], islong, g[f"tns_{tag}_nfilt"][u],
                                             g[f"tns_{tag}_filt"][u], sfb, tmax, len(sfb))
            assert bits_equal(got, g[f"tns_{tag}_out"][u]), (tag, u)
    for up, down in ((2, 1), (3, 1), (3, 2)):
        o = loader.OracleResampler(up, down, g["rs_filter_2"], g["rs_filter_3"])
        for f in range(3):
            assert bits_equal(o.process(g[f"rs_{up}_{down}_in"][f]), g[f"rs_{up}_{down}_out"][f])
