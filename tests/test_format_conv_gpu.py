"""CUDA format converter (through the C ABI) against the oracle restatement (pinned to the reference by
tests/test_oracle_fc.py) and the committed reference vectors; downmix tensors are the reference-designed ones."""
import os

import numpy as np
import pytest

import fc_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "fc_golden.npz")


class DevFc:
    def __init__(self, lib, D, n_streams):
        import torch
        self.t, self.lib, self.S = torch, lib, n_streams
        self.n_out, self.n_in = D.shape[:2]
        self.h = lib.format_conv_create(D)
        self.state = torch.zeros(n_streams * lib.format_conv_state_floats(self.h), device="cuda")

    def run(self, x):
        t = self.t
        d_in = t.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        d_out = t.full((self.S, self.n_out, 1024), float("nan"), device="cuda")
        self.lib.format_conv_dev(self.h, d_in.data_ptr(), d_out.data_ptr(), self.state.data_ptr(), self.S,
                                 t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return d_out.cpu().numpy()

    def close(self):
        self.lib.format_conv_destroy(self.h)


def test_golden(gpu_lib):
    g = np.load(GOLD)
    for pair in fc_cases.PAIRS[:3]:
        k = fc_cases.key(pair)
        r = DevFc(gpu_lib, g[k + "_D"], 1)
        for f in range(g[k + "_in"].shape[0]):
            got = r.run(g[k + "_in"][f][None])[0]
            assert bits_equal(got, g[k + "_out"][f]), (k, f, float(np.abs(got - g[k + "_out"][f]).max()))
        r.close()


@pytest.mark.parametrize("pair", fc_cases.PAIRS, ids=fc_cases.key)
def test_streams_vs_oracle(gpu_lib, oracle, pair):
    g = np.load(GOLD)
    D = g[fc_cases.key(pair) + "_D"]
    S, F = 4, 5
    rng = np.random.default_rng(pair[0] + 7 * pair[1])
    r = DevFc(gpu_lib, D, S)
    orc = [loader.OracleFc(D) for _ in range(S)]
    for f in range(F):
        x = np.stack([fc_cases.signal(rng, 1, D.shape[1])[0] for _ in range(S)])
        if f == 2:
            x[1] = 0   # a silent frame: energies decay, EQ hits its floor
        got = r.run(x)
        for s in range(S):
            want = orc[s].process(x[s])
            assert bits_equal(got[s], want), (pair, f, s, float(np.abs(got[s] - want).max()))
    r.close()


@pytest.mark.parametrize("n_in,n_out", [(24, 24), (3, 5), (1, 1), (24, 12)])
def test_dense_matrices(gpu_lib, oracle, n_in, n_out):
    rng = np.random.default_rng(n_in * 100 + n_out)
    D = fc_cases.dense_matrix(rng, n_in, n_out)
    r = DevFc(gpu_lib, D, 2)
    orc = [loader.OracleFc(D) for _ in range(2)]
    for f in range(3):
        x = fc_cases.signal(rng, 2, n_in)
        got = r.run(x)
        for s in range(2):
            assert bits_equal(got[s], orc[s].process(x[s])), (f, s)
    r.close()


def test_full_size_properties(gpu_lib, oracle):
    """BASELINE config 5's bed conversion: 22.2 -> 7.1.4, 512 streams; permutation invariance + spot checks."""
    g = np.load(GOLD)
    D = g[fc_cases.key((13, 19)) + "_D"]
    S = 512
    rng = np.random.default_rng(4567)
    x = rng.normal(0, 3000.0, size=(2, S, 24, 1024)).astype(np.float32)
    r = DevFc(gpu_lib, D, S)
    out = [r.run(x[0]), r.run(x[1])]
    r.close()
    perm = rng.permutation(S)
    r2 = DevFc(gpu_lib, D, S)
    out_p = [r2.run(x[0][perm]), r2.run(x[1][perm])]
    r2.close()
    assert bits_equal(out_p[0], out[0][perm]) and bits_equal(out_p[1], out[1][perm])
    for s in rng.choice(S, 4, replace=False):
        o = loader.OracleFc(D)
        assert bits_equal(out[0][s], o.process(x[0, s])) and bits_equal(out[1][s], o.process(x[1, s])), s
