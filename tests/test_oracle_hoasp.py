"""HOA spatial decoder restatement (oracle/oracle_hoasp.c) against the unmodified reference stage functions."""
import ctypes

import numpy as np
import pytest

import hoasp_cases
from oracle import loader
from oracle.loader import bits_equal


def _ref_new(ref, cfg):
    err = ctypes.c_int()
    h = ref.ref_hoasp_create(hoasp_cases.cfg_array(cfg), ctypes.byref(err))
    assert err.value == 0, cfg
    return h


@pytest.mark.parametrize("cfg", hoasp_cases.CONFIGS, ids=hoasp_cases.key)
def test_spatial_vs_reference(ref, oracle, cfg):
    h = _ref_new(ref, cfg)
    tables = loader.ref_hoasp_tables(ref, h)
    # the reference tabulates 2^-k as 17-digit decimals: exact only down to 2^-17, so the table is an input
    assert np.array_equal(tables["pow2"][64:82], np.exp2(-np.arange(18.0)).astype(np.float32))
    o = loader.OracleHoasp(cfg, tables)
    rng = np.random.default_rng(sum(cfg) + 11)
    gen = hoasp_cases.SideGen(cfg, rng, wild_exp=True)
    F = 14
    x = hoasp_cases.signals(rng, F, cfg[1])
    seen = set()
    for f in range(F):
        sd = gen.next()
        want = np.zeros((o.n_coef, 1024), np.float32)
        assert ref.ref_hoasp_process(h, sd.ctypes.data, x[f].reshape(-1), want.reshape(-1)) == 0
        got = o.process(sd, x[f])
        assert bits_equal(got, want), (cfg, f, float(np.abs(got - want).max()))
        seen |= {k for k in ("n_vec", "n_dir", "n_enable", "n_disable") if sd[k] > 0}
        if sd["pred_type"].any():
            seen.add("pred")
    if cfg[2] >= 4:
        assert {"n_vec", "n_dir", "n_enable", "n_disable", "pred"} <= seen, seen
    ref.ref_hoasp_destroy(h)


def test_hoasp_golden(oracle):
    """Committed vectors from the reference build (tests/golden/make_hoasp_golden.py); no reference needed."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hoasp_golden.npz"))
    for cfg in hoasp_cases.CONFIGS[:3]:
        k = hoasp_cases.key(cfg)
        o = loader.OracleHoasp(cfg, hoasp_cases.golden_tables(g, cfg))
        sides = g[k + "_sides"].view(hoasp_cases.SIDE_DTYPE).reshape(-1)
        for f in range(len(sides)):
            assert bits_equal(o.process(sides[f:f + 1], g[k + "_in"][f]), g[k + "_out"][f]), (k, f)
