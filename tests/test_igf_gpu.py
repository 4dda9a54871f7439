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


class GpuIgfPairs:
    """n channel pairs through the joint-stereo IGF chain: fill, optional M/S over the IGF range (spectra and TNF
    buffers), TNF -- units 2p (left), 2p+1 (right)."""

    def __init__(self, lib, n_pairs):
        import torch
        self.torch, self.lib, self.n = torch, lib, n_pairs
        self.bands = lib.bands_create(igf_cases.SFB_LONG, igf_cases.SFB_SHORT)
        self.tnf = torch.zeros(2 * n_pairs, 2048, device="cuda")
        self.dummy = torch.zeros(n_pairs, 2048, device="cuda")

    def close(self):
        self.lib.bands_destroy(self.bands)

    def frame(self, sdl, sdr, masks, coef, tiles, ms_after):
        import stereo_cases
        torch, lib, n = self.torch, self.lib, self.n
        st = torch.cuda.current_stream().cuda_stream
        sides = np.zeros(2 * n, igf_cases.IGF_SIDE)
        sides[0::2], sides[1::2] = sdl, sdr
        sides["unit"] = np.arange(2 * n)
        d_side = torch.from_numpy(sides.view(np.uint8).reshape(2 * n, -1).copy()).cuda()
        d_coef = torch.from_numpy(np.ascontiguousarray(coef, np.float32).reshape(2 * n, 1024)).cuda()
        d_tiles = torch.from_numpy(np.ascontiguousarray(tiles, np.float32).reshape(2 * n, 4, 1024)).cuda()
        d_mask = torch.from_numpy(np.ascontiguousarray(masks, np.uint8)).cuda()
        d_seed = torch.zeros(2 * n, dtype=torch.int32, device="cuda")
        lib.igf_stereo_dev(self.bands, d_coef.data_ptr(), d_tiles.data_ptr(), d_side.data_ptr(), d_mask.data_ptr(),
                           self.tnf.data_ptr(), d_seed.data_ptr(), n, st)
        # stereo tool over the IGF band range on the spectra, and on the TNF spectra of long blocks with TNF enabled
        ms = np.zeros(n, stereo_cases.STEREO_SIDE)
        ms_t = np.zeros(n, stereo_cases.STEREO_SIDE)
        for p in range(n):
            ms[p]["left"], ms[p]["right"], ms[p]["slot"] = 2 * p, 2 * p + 1, p
            ms[p]["tool"] = 1 if ms_after[p] else 0
            ms[p]["is_short"] = sdl[p]["win_type"]
            ms[p]["first_band"], ms[p]["last_band"] = sdl[p]["sfb_start"], sdl[p]["sfb_stop"]
            ms[p]["used"] = igf_cases.mask_per_window(masks[p], sdl[p])
            ms_t[p] = ms[p]
            ms_t[p]["left"], ms_t[p]["right"] = 4 * p, 4 * p + 2  # TNF spectra inside [unit][2048] viewed as [.][1024]
            if not (sdl[p]["use_tnf"] and not sdl[p]["win_type"]):
                ms_t[p]["tool"] = 0
        d_ms = torch.from_numpy(ms.view(np.uint8).reshape(n, -1).copy()).cuda()
        d_ms_t = torch.from_numpy(ms_t.view(np.uint8).reshape(n, -1).copy()).cuda()
        lib.stereo_dev(self.bands, d_coef.data_ptr(), d_ms.data_ptr(), self.dummy.data_ptr(), n, st)
        lib.stereo_dev(self.bands, self.tnf.data_ptr(), d_ms_t.data_ptr(), self.dummy.data_ptr(), n, st)
        lib.igf_tnf_dev(self.bands, d_coef.data_ptr(), d_side.data_ptr(), self.tnf.data_ptr(), 2 * n, st)
        torch.cuda.synchronize()
        return d_coef.cpu().numpy().reshape(n, 2, 1024), d_seed.cpu().numpy().view(np.uint32).reshape(n, 2)


def test_igf_stereo_golden(gpu_lib):
    g = np.load(GOLD)
    C = len(igf_cases.CONFIGS)
    sdl = [g[f"s{ci}_sdl"].view(igf_cases.IGF_SIDE).reshape(-1) for ci in range(C)]
    sdr = [g[f"s{ci}_sdr"].view(igf_cases.IGF_SIDE).reshape(-1) for ci in range(C)]
    gp = GpuIgfPairs(gpu_lib, C)
    try:
        for k in range(len(sdl[0])):
            y, seeds = gp.frame([sdl[ci][k] for ci in range(C)], [sdr[ci][k] for ci in range(C)],
                                np.stack([g[f"s{ci}_mask"][k] for ci in range(C)]),
                                np.stack([g[f"s{ci}_in"][k] for ci in range(C)]),
                                np.stack([g[f"s{ci}_tiles"][k] for ci in range(C)]),
                                [int(g[f"s{ci}_ms"][k]) for ci in range(C)])
            for ci in range(C):
                assert bits_equal(y[ci], g[f"s{ci}_out"][k]), (ci, k)
                assert np.array_equal(seeds[ci], g[f"s{ci}_seeds"][k]), (ci, k)
    finally:
        gp.close()


def test_igf_stereo_batch_vs_oracle(gpu_lib, oracle):
    """300 pairs x 4 frames against the oracle chain (fill, M/S on spectra and TNF buffers, TNF)."""
    g = np.load(GOLD)
    P, F = 300, 4
    rngs = [np.random.default_rng(88000 + p) for p in range(P)]
    cfgs = [p % len(igf_cases.CONFIGS) for p in range(P)]
    pairs = [igf_cases.OracleIgfPair() for _ in range(P)]
    gp = GpuIgfPairs(gpu_lib, P)
    try:
        for f in range(F):
            fr = [igf_cases.random_pair_frame(rngs[p], igf_cases.CONFIGS[cfgs[p]], g[f"c{cfgs[p]}_grid"],
                                              win_type=int(rngs[p].random() < 0.3)) for p in range(P)]
            ms_after = [int(rngs[p].integers(2)) for p in range(P)]
            y, seeds = gp.frame([x[0] for x in fr], [x[1] for x in fr], np.stack([x[2] for x in fr]),
                                np.stack([x[3] for x in fr]), np.stack([x[4] for x in fr]), ms_after)
            for p in range(P):
                want, want_seeds = pairs[p].frame(fr[p][0], fr[p][1], fr[p][2], fr[p][3], fr[p][4], ms_after[p])
                assert bits_equal(y[p], want), (p, f, int(fr[p][0]["win_type"]), ms_after[p])
                assert np.array_equal(seeds[p], want_seeds), (p, f)
    finally:
        gp.close()
