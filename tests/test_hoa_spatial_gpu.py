"""CUDA HOA spatial decoder (through the C ABI) against the oracle restatement (pinned to the reference by
tests/test_oracle_hoasp.py) and the committed reference vectors; tables are the reference-designed ones."""
import os

import numpy as np
import pytest

import hoasp_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "hoasp_golden.npz")


class DevHoasp:
    def __init__(self, lib, cfg, tables, n_streams):
        import torch
        self.t, self.lib, self.S, self.cfg = torch, lib, n_streams, cfg
        self.h = lib.hoa_spatial_create(cfg, tables)
        self.nc = lib.hoa_spatial_num_coeffs(self.h)
        self.state = torch.empty(n_streams * lib.hoa_spatial_state_bytes(self.h), dtype=torch.uint8, device="cuda")
        lib.hoa_spatial_state_init_dev(self.h, self.state.data_ptr(), n_streams, torch.cuda.current_stream().cuda_stream)

    def run(self, sides, x):
        t = self.t
        d_side = t.from_numpy(np.ascontiguousarray(sides).view(np.uint8).reshape(self.S, -1).copy()).cuda()
        d_in = t.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        d_out = t.full((self.S, self.nc, 1024), float("nan"), device="cuda")
        self.lib.hoa_spatial_dev(self.h, d_side.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), self.state.data_ptr(),
                                 self.S, t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return d_out.cpu().numpy()

    def close(self):
        self.lib.hoa_spatial_destroy(self.h)


def test_golden(gpu_lib):
    g = np.load(GOLD)
    for cfg in hoasp_cases.CONFIGS[:3]:
        k = hoasp_cases.key(cfg)
        r = DevHoasp(gpu_lib, cfg, hoasp_cases.golden_tables(g, cfg), 1)
        sides = g[k + "_sides"].view(hoasp_cases.SIDE_DTYPE).reshape(-1)
        for f in range(len(sides)):
            got = r.run(sides[f:f + 1], g[k + "_in"][f][None])[0]
            assert bits_equal(got, g[k + "_out"][f]), (k, f, float(np.abs(got - g[k + "_out"][f]).max()))
        r.close()


@pytest.mark.parametrize("cfg", hoasp_cases.CONFIGS, ids=hoasp_cases.key)
def test_streams_vs_oracle(gpu_lib, oracle, cfg):
    """5 independent streams x 12 frames with different side-info walks, state carried on the device."""
    g = np.load(GOLD)
    tables = hoasp_cases.golden_tables(g, cfg)
    S, F = 5, 12
    rng = np.random.default_rng(sum(cfg) + 3)
    gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(100 * s + cfg[0])) for s in range(S)]
    orc = [loader.OracleHoasp(cfg, tables) for _ in range(S)]
    r = DevHoasp(gpu_lib, cfg, tables, S)
    for f in range(F):
        sides = np.array([gen.next() for gen in gens], hoasp_cases.SIDE_DTYPE)
        x = hoasp_cases.signals(rng, S, cfg[1])
        got = r.run(sides, x)
        for s in range(S):
            want = orc[s].process(sides[s:s + 1], x[s])
            assert bits_equal(got[s], want), (cfg, f, s, float(np.abs(got[s] - want).max()))
    r.close()


def test_large_exponents_within_one_ulp(gpu_lib, oracle):
    """|exponent| > 6 takes the reference's per-sample double pow(): device pow is not bit-reproducible."""
    g = np.load(GOLD)
    cfg = hoasp_cases.CONFIGS[0]
    tables = hoasp_cases.golden_tables(g, cfg)
    rng = np.random.default_rng(77)
    gen = hoasp_cases.SideGen(cfg, rng, wild_exp=True)
    o = loader.OracleHoasp(cfg, tables)
    r = DevHoasp(gpu_lib, cfg, tables, 1)
    for f in range(10):
        sd = np.array([gen.next()], hoasp_cases.SIDE_DTYPE)
        x = hoasp_cases.signals(rng, 1, cfg[1], sigma=100.0)
        got, want = r.run(sd, x)[0], o.process(sd, x[0])
        tol = 4 * np.spacing(np.maximum(np.abs(want), np.float32(1e-3)).astype(np.float32)) * 16
        assert np.all(np.abs(got - want) <= tol), f
    r.close()


def test_full_size_properties(gpu_lib, oracle):
    """BASELINE config 3 shape: 1024 streams, order 4, 16 transport channels; permutation invariance + spot checks."""
    g = np.load(GOLD)
    cfg = hoasp_cases.CONFIGS[0]
    tables = hoasp_cases.golden_tables(g, cfg)
    S = 1024
    rng = np.random.default_rng(2345)
    gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(s)) for s in range(S)]
    frames = []
    for f in range(2):
        frames.append((np.array([gen.next() for gen in gens], hoasp_cases.SIDE_DTYPE),
                       rng.normal(0, 2000.0, size=(S, 16, 1024)).astype(np.float32)))
    r = DevHoasp(gpu_lib, cfg, tables, S)
    out = [r.run(sd, x) for sd, x in frames]
    r.close()
    perm = rng.permutation(S)
    r2 = DevHoasp(gpu_lib, cfg, tables, S)
    out_p = [r2.run(sd[perm], x[perm]) for sd, x in frames]
    r2.close()
    assert bits_equal(out_p[0], out[0][perm]) and bits_equal(out_p[1], out[1][perm])
    for s in rng.choice(S, 6, replace=False):
        o = loader.OracleHoasp(cfg, tables)
        for f in range(2):
            assert bits_equal(out[f][s], o.process(frames[f][0][s:s + 1], frames[f][1][s])), (s, f)
