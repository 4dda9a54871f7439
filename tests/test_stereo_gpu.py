"""CUDA joint-stereo tools (through the C ABI) composed with the TNS and FD-synthesis kernels into phase 2 of a
channel-pair element, against the committed reference vectors and the oracle."""
import os

import numpy as np
import pytest

import stereo_cases
import tns_cases
from oracle.loader import bits_equal
from test_oracle_stereo import golden_frames

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "stereo_golden.npz")


class GpuPairs:
    """n_pairs channel pairs advanced one frame per call: units 2p (left), 2p+1 (right)."""

    def __init__(self, lib, n_pairs):
        import torch
        self.torch, self.lib, self.n = torch, lib, n_pairs
        self.h = lib.bands_create(stereo_cases.SFB_LONG, stereo_cases.SFB_SHORT)
        self.state = torch.zeros(n_pairs, lib.stereo_state_floats(), device="cuda")
        self.ovl = torch.zeros(2 * n_pairs, 1024, device="cuda")
        self.stream = torch.cuda.current_stream().cuda_stream

    def close(self):
        self.lib.bands_destroy(self.h)

    def _tns(self, d_coef, frames, which):
        from libmpegh_b200.lib import TNS_FILTER_DTYPE
        torch = self.torch
        chains = []
        for p, fr in enumerate(frames):
            if fr["tns_on_lr"] != which:
                continue
            islong = int(fr["sd"]["is_short"] == 0)
            for ch in range(2):
                chains += tns_cases.resolve_filters(self.lib, 2 * p + ch, islong, fr["n_filt"][ch], fr["filt"][ch],
                                                    int(fr["max_sfb"][ch]), self.lib.tns_lpc_from_coef)
        if not chains:
            return
        recs = np.array([r for ch in chains for r in ch], TNS_FILTER_DTYPE)
        first = np.cumsum([0] + [len(ch) for ch in chains]).astype(np.int32)
        d_f = torch.from_numpy(recs.view(np.uint8).reshape(len(recs), -1).copy()).cuda()
        d_first = torch.from_numpy(first).cuda()
        self.lib.tns_dev(d_coef.data_ptr(), d_f.data_ptr(), d_first.data_ptr(), len(chains), self.stream)

    def frame(self, frames, coef):
        """frames: one dict per pair; coef [n_pairs, 2, 1024] -> (pcm [n_pairs, 2, 1024], spectra)"""
        torch, lib = self.torch, self.lib
        from libmpegh_b200.lib import fd_side
        sds = np.array([fr["sd"] for fr in frames], stereo_cases.STEREO_SIDE)
        sds["left"] = 2 * np.arange(self.n)
        sds["right"] = 2 * np.arange(self.n) + 1
        sds["slot"] = np.arange(self.n)
        d_pairs = torch.from_numpy(sds.view(np.uint8).reshape(self.n, -1).copy()).cuda()
        d_coef = torch.from_numpy(np.ascontiguousarray(coef, np.float32).reshape(2 * self.n, 1024)).cuda()
        # ia_core_coder_data_process order: the pairs whose TNS runs before the stereo tools, the tools, the rest
        self._tns(d_coef, frames, 0)
        lib.stereo_dev(self.h, d_coef.data_ptr(), d_pairs.data_ptr(), self.state.data_ptr(), self.n, self.stream)
        self._tns(d_coef, frames, 1)
        lib.stereo_save_dev(self.h, d_coef.data_ptr(), d_pairs.data_ptr(), self.state.data_ptr(), self.n, self.stream)
        spec = d_coef.cpu().numpy().reshape(self.n, 2, 1024)
        side = np.repeat(fd_side(sds["window_sequence"].astype(np.int32), sds["window_shape"].astype(np.int32),
                                 sds["window_shape_prev"].astype(np.int32), 0, 0), 2).astype(np.int32)
        d_side = torch.from_numpy(side).cuda()
        d_out = torch.empty(2 * self.n, 1024, device="cuda")
        lib.fd_synth_dev(d_coef.data_ptr(), d_side.data_ptr(), self.ovl.data_ptr(), d_out.data_ptr(), 2 * self.n,
                         self.stream)
        torch.cuda.synchronize()
        return d_out.cpu().numpy().reshape(self.n, 2, 1024), spec


def test_cpe_golden(gpu_lib):
    """The six reference pair sequences run side by side as one batch of six pairs."""
    g = np.load(GOLD)
    seqs = [golden_frames(g, p) for p in range(6)]
    gp = GpuPairs(gpu_lib, 6)
    try:
        for f in range(len(seqs[0])):
            coef = np.stack([g[f"p{p}_coef"][f] for p in range(6)])
            pcm, spec = gp.frame([s[f] for s in seqs], coef)
            for p in range(6):
                sd = seqs[p][f]["sd"]
                bins = 128 if sd["is_short"] else 1024
                assert bits_equal(spec[p][:, 1024 - bins:], g[f"p{p}_spec"][f][:, 1024 - bins:]), (p, f, "spectrum")
                assert bits_equal(pcm[p], g[f"p{p}_pcm"][f]), (p, f, int(sd["tool"]), int(sd["is_short"]))
    finally:
        gp.close()


def test_cpe_batch_vs_oracle(gpu_lib, oracle):
    """300 pairs x 6 frames of random side info against the oracle chain, spectra and PCM bit for bit."""
    P, F = 300, 6
    seqs, coefs, pairs = [], [], []
    for p in range(P):
        rng = np.random.default_rng(31000 + p)
        seqs.append(stereo_cases.random_pair_sequence(rng, F))
        coefs.append(rng.normal(0, 600, (F, 2, 1024)).astype(np.float32))
        pairs.append(stereo_cases.OraclePair())
    gp = GpuPairs(gpu_lib, P)
    try:
        for f in range(F):
            pcm, spec = gp.frame([s[f] for s in seqs], np.stack([c[f] for c in coefs]))
            for p in range(P):
                want_pcm, want_spec = pairs[p].frame(seqs[p][f], coefs[p][f, 0], coefs[p][f, 1])
                assert bits_equal(spec[p], np.stack(want_spec)), (p, f, "spectrum")
                assert bits_equal(pcm[p], want_pcm), (p, f)
    finally:
        gp.close()


def test_stereo_edge_cases(gpu_lib, oracle):
    """tool 0 leaves the spectra alone; save modes 0/1/2; signed zeros survive (the '+ 0' of the MDST estimate)."""
    import torch
    lib = gpu_lib
    h = lib.bands_create(stereo_cases.SFB_LONG, stereo_cases.SFB_SHORT)
    rng = np.random.default_rng(5)
    n = 8
    sds = np.zeros(n, stereo_cases.STEREO_SIDE)
    coef = rng.normal(0, 300, (2 * n, 1024)).astype(np.float32)
    coef[:, ::7] = -0.0
    coef[3] = 0.0
    for p in range(n):
        sd, _ = stereo_cases.random_frame(rng, [0, 2][p % 2], p & 1, (p >> 1) & 1, tool=[0, 1, 3, 3][p % 4])
        sd["left"], sd["right"], sd["slot"] = 2 * p, 2 * p + 1, p
        sd["save"] = p % 3
        if sd["tool"] == 3:
            sd["complex_coef"], sd["use_prev_frame"], sd["prev_dmx"] = 1, 1, p % 2
        sds[p] = sd
    state0 = rng.normal(0, 100, (n, 2048)).astype(np.float32)
    d_coef, d_state = torch.from_numpy(coef.copy()).cuda(), torch.from_numpy(state0.copy()).cuda()
    d_pairs = torch.from_numpy(sds.view(np.uint8).reshape(n, -1).copy()).cuda()
    st = torch.cuda.current_stream().cuda_stream
    lib.stereo_dev(h, d_coef.data_ptr(), d_pairs.data_ptr(), d_state.data_ptr(), n, st)
    lib.stereo_save_dev(h, d_coef.data_ptr(), d_pairs.data_ptr(), d_state.data_ptr(), n, st)
    torch.cuda.synchronize()
    got, got_state = d_coef.cpu().numpy(), d_state.cpu().numpy()
    for p in range(n):
        l, r, stt = coef[2 * p].copy(), coef[2 * p + 1].copy(), state0[p].copy()
        sfb = stereo_cases.SFB_SHORT if sds[p]["is_short"] else stereo_cases.SFB_LONG
        oracle.oracle_stereo_apply(sds[p].tobytes(), np.ascontiguousarray(sfb, np.int16), len(sfb), stt, l, r)
        oracle.oracle_stereo_save(sds[p].tobytes(), stt, l, r)
        assert bits_equal(got[2 * p], l) and bits_equal(got[2 * p + 1], r), p
        assert bits_equal(got_state[p], stt), p
    lib.stereo_dev(h, 0, 0, 0, 0, st)  # zero pairs: accepted, nothing launched
    lib.bands_destroy(h)
