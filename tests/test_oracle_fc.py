"""Format converter restatement (oracle/oracle_fc.c) against the unmodified reference."""
import os

import numpy as np
import pytest

import fc_cases
from oracle import loader
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fc_golden.npz")


@pytest.mark.parametrize("pair", fc_cases.PAIRS, ids=fc_cases.key)
def test_fc_vs_reference(ref, oracle, pair):
    h, D = loader.ref_fc_new(ref, *pair)
    g = np.load(GOLD)
    assert bits_equal(D, g[fc_cases.key(pair) + "_D"])
    o = loader.OracleFc(D)
    rng = np.random.default_rng(pair[0] * 31 + pair[1])
    x = fc_cases.signal(rng, 6, o.n_in)
    for f in range(6):
        want = np.zeros((o.n_out, 1024), np.float32)
        assert ref.ref_fc_process(h, x[f].reshape(-1), want.reshape(-1)) == 0
        got = o.process(x[f])
        assert bits_equal(got, want), (pair, f, float(np.abs(got - want).max()))
    ref.ref_fc_destroy(h)


@pytest.mark.parametrize("n_in,n_out", [(24, 12), (3, 5), (1, 1), (24, 24)])
def test_fc_dense_matrix_vs_reference(ref, oracle, n_in, n_out):
    h, _ = loader.ref_fc_new(ref, 6, 2)
    rng = np.random.default_rng(n_in * 100 + n_out)
    D = fc_cases.dense_matrix(rng, n_in, n_out)
    ref.ref_fc_set_matrix(h, n_in, n_out, D.reshape(-1))
    o = loader.OracleFc(D)
    x = fc_cases.signal(rng, 4, n_in)
    for f in range(4):
        want = np.zeros((n_out, 1024), np.float32)
        assert ref.ref_fc_process(h, x[f].reshape(-1), want.reshape(-1)) == 0
        assert bits_equal(o.process(x[f]), want), (f,)
    ref.ref_fc_destroy(h)


def test_fc_golden(oracle):
    g = np.load(GOLD)
    for pair in fc_cases.PAIRS[:3]:
        k = fc_cases.key(pair)
        o = loader.OracleFc(g[k + "_D"])
        for f in range(g[k + "_in"].shape[0]):
            assert bits_equal(o.process(g[k + "_in"][f]), g[k + "_out"][f]), (k, f)
