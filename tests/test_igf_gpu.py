"""CUDA IGF (mono path, through the C ABI) against the committed reference vectors and the oracle."""
import os

import numpy as np
import pytest

import igf_cases
from oracle.loader import bits_equal
from test_oracle_igf import run_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "igf_golden.npz")


class GpuIgf:
    """n_units channels advanced one frame per call, TNF buffers carried on the device."""

    def __init__(self, lib, n_units):
        import torch
        self.torch, self.lib, self.n = torch, lib, n_units
        self.bands = lib.bands_create(igf_cases.SFB_LONG, igf_cases.SFB_SHORT)
        self.state = torch.zeros(n_units, 2048, device="cuda")

    def close(self):
        self.lib.bands_destroy(self.bands)

    def frame(self, sds, coef, tiles):
        torch = self.torch
        sds = np.array(sds, igf_cases.IGF_SIDE)
        sds["unit"] = np.arange(self.n)
        d_side = torch.from_numpy(sds.view(np.uint8).reshape(self.n, -1).copy()).cuda()
        d_coef = torch.from_numpy(np.ascontiguousarray(coef, np.float32)).cuda()
        d_tiles = torch.from_numpy(np.ascontiguousarray(tiles, np.float32)).cuda()
        d_seed = torch.zeros(self.n, dtype=torch.int32, device="cuda")
        self.lib.igf_dev(self.bands, d_coef.data_ptr(), d_tiles.data_ptr(), d_side.data_ptr(), self.state.data_ptr(),
                         d_seed.data_ptr(), self.n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_tiles.cpu().numpy(), tiles), "source tiles must not be written"
        return d_coef.cpu().numpy(), d_seed.cpu().numpy().view(np.uint32)


def test_whitening_rom(gpu_lib):
    g = np.load(GOLD)
    t = gpu_lib.igf_whitening_rom()
    assert bits_equal(t[0][:63], g["whiten_pos"][:63]) and bits_equal(t[1], g["whiten_neg"])


def test_igf_golden(gpu_lib):
    """The six reference configurations run side by side, one channel each, eight consecutive frames."""
    g = np.load(GOLD)
    C = len(igf_cases.CONFIGS)
    sds = [g[f"c{ci}_sd"].view(igf_cases.IGF_SIDE).reshape(-1) for ci in range(C)]
    gi = GpuIgf(gpu_lib, C)
    try:
        for k in range(len(sds[0])):
            y, seed = gi.frame([sds[ci][k] for ci in range(C)], np.stack([g[f"c{ci}_in"][k] for ci in range(C)]),
                               np.stack([g[f"c{ci}_tiles"][k] for ci in range(C)]))
            for ci in range(C):
                assert bits_equal(y[ci], g[f"c{ci}_out"][k]), (ci, k)
                assert int(seed[ci]) == int(g[f"c{ci}_seed"][k]), (ci, k)
    finally:
        gi.close()


def test_igf_batch_vs_oracle(gpu_lib, oracle):
    """600 channels x 5 frames, every configuration, long and short blocks, carried TNF buffers."""
    g = np.load(GOLD)
    U, F = 600, 5
    rngs = [np.random.default_rng(77000 + u) for u in range(U)]
    cfgs = [u % len(igf_cases.CONFIGS) for u in range(U)]
    states = [np.zeros(2048, np.float32) for _ in range(U)]
    gi = GpuIgf(gpu_lib, U)
    try:
        for f in range(F):
            sds, coefs, tls = [], [], []
            for u in range(U):
                sd, coef, tiles = igf_cases.random_frame(rngs[u], igf_cases.CONFIGS[cfgs[u]], g[f"c{cfgs[u]}_grid"],
                                                         win_type=int(rngs[u].random() < 0.3), all_zero_prob=0.1)
                sds.append(sd), coefs.append(coef), tls.append(tiles)
            y, seed = gi.frame(sds, np.stack(coefs), np.stack(tls))
            for u in range(U):
                want, want_seed = run_oracle(oracle, sds[u], coefs[u], tls[u], states[u])
                assert bits_equal(y[u], want), (u, f, int(sds[u]["win_type"]), int(sds[u]["all_zero"]))
                assert int(seed[u]) == want_seed, (u, f)
    finally:
        gi.close()
