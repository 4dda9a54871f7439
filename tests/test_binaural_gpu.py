"""CUDA binaural renderer (through the C ABI) against the oracle restatement (which tests/test_oracle_binaural.py
pins to the unmodified reference) and the committed reference vectors."""
import os

import numpy as np
import pytest

import bin_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "binaural_golden.npz")


class DevBin:
    def __init__(self, lib, filters, n_streams, sets=None):
        """filters: one dict (shared set) or a list of dicts (one set each); sets: per-stream set index"""
        import torch
        self.t, self.lib, self.S = torch, lib, n_streams
        fl = filters if isinstance(filters, list) else [filters]

        def st(k):
            return np.stack([f[k] for f in fl])
        self.h = lib.binaural_create(st("taps_direct"), st("taps_diffuse"), st("direct_fc"), st("diffuse_fc"),
                                     st("weight"))
        self.B = lib.binaural_block_len(self.h)
        self.n_in = fl[0]["taps_direct"].shape[1]
        self.nb = fl[0]["taps_diffuse"].shape[0]
        self.n_sets = len(fl)
        self.state = torch.zeros(n_streams * lib.binaural_state_floats(self.h), device="cuda")
        self.sets = None if sets is None else torch.from_numpy(np.asarray(sets, np.int32)).cuda()

    def run(self, x, layout=0, in_scale=1.0, out_scale=1.0):
        t = self.t
        d_in = t.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        d_out = t.full((self.S, 2, self.B), float("nan"), device="cuda")
        self.lib.binaural_block_dev(self.h, d_in.data_ptr(), d_out.data_ptr(), self.state.data_ptr(), self.S, layout,
                                    in_scale, out_scale, None if self.sets is None else self.sets.data_ptr(),
                                    t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return d_out.cpu().numpy()

    def close(self):
        self.lib.binaural_destroy(self.h)


def _oracle(f):
    return loader.OracleBinaural(f["taps_direct"], f["taps_diffuse"], f["direct_fc"], f["diffuse_fc"], f["weight"])


@pytest.mark.parametrize("n_in,length,nb,cut", [(24, 4096, 1, False), (5, 1024, 1, True), (21, 4096, 0, False),
                                                (2, 1024, 2, True), (11, 3000, 1, False), (32, 700, 0, True),
                                                # transform lengths of the generic path: 1024, 512, 128, 64, 16; 16384
                                                (4, 300, 1, True), (6, 512, 1, False), (3, 200, 0, True), (2, 64, 1, False),
                                                (5, 20, 1, False), (2, 8, 0, False), (3, 5000, 1, True),
                                                (2, 8192, 0, False)])
def test_streams_vs_oracle(gpu_lib, oracle, n_in, length, nb, cut):
    """3 independent streams x 4 blocks, state carried on the device; filter spectra compared first.  The oracle is
    pinned to the unmodified reference for every one of these transform lengths (tests/test_oracle_binaural.py),
    including 16384, where the reference's e^{+j} transform has no column stage."""
    rng = np.random.default_rng(n_in * 7 + length + nb)
    f = bin_cases.brir_set(rng, n_in, length, nb)
    if cut:
        f = bin_cases.with_cutoffs(rng, f)
    S, T = 3, 4
    r = DevBin(gpu_lib, f, S)
    o = [_oracle(f) for _ in range(S)]
    d, df = gpu_lib.binaural_spectra(r.h, 1, n_in, nb)
    assert bits_equal(d[0], o[0].h_direct)
    if nb:
        assert bits_equal(df[0], o[0].h_diffuse)
    x = np.stack([bin_cases.signal(rng, T, n_in, r.B) for _ in range(S)], axis=1)  # [T, S, n_in, B]
    for t in range(T):
        got = r.run(x[t])
        for s in range(S):
            want = o[s].block(x[t, s])
            assert bits_equal(got[s], want), (t, s, float(np.abs(got[s] - want).max()))
    r.close()


def test_frame_layout_and_pcm_scale(gpu_lib, oracle):
    """Decoder-style feed: 4 stacked 1024-sample frames at 16-bit scale, output scaled back (decode_main.c:2825-2853)."""
    rng = np.random.default_rng(99)
    n_in, S = 6, 4
    f = bin_cases.brir_set(rng, n_in, 4096, 1)
    r = DevBin(gpu_lib, f, S)
    o = [_oracle(f) for _ in range(S)]
    for blk in range(2):
        frames = rng.normal(0, 3000.0, size=(4, S, n_in, 1024)).astype(np.float32)
        got = r.run(frames, layout=1, in_scale=1.0 / 32768.0, out_scale=32768.0)
        for s in range(S):
            x = (np.concatenate(list(frames[:, s]), axis=1) / np.float32(32768.0)).astype(np.float32)
            want = o[s].block(x) * np.float32(32768.0)
            assert bits_equal(got[s], want), (blk, s)
    r.close()


def test_per_stream_filter_sets(gpu_lib, oracle):
    rng = np.random.default_rng(5)
    n_in, S = 4, 5
    fl = [bin_cases.brir_set(rng, n_in, 1024, 1) for _ in range(3)]
    sets = [2, 0, 1, 1, 2]
    r = DevBin(gpu_lib, fl, S, sets)
    o = [_oracle(fl[k]) for k in sets]
    for t in range(3):
        x = bin_cases.signal(rng, 1, S * n_in, r.B)[0].reshape(S, n_in, r.B)
        got = r.run(x)
        for s in range(S):
            assert bits_equal(got[s], o[s].block(x[s])), (t, s)
    r.close()


def test_golden(gpu_lib):
    """Vectors produced by the unmodified reference (tests/golden/make_binaural_golden.py)."""
    g = np.load(GOLD)
    for tag in ("a", "b"):
        f = {k: g[f"{tag}_{k}"] for k in ("taps_direct", "taps_diffuse", "direct_fc", "diffuse_fc", "weight")}
        r = DevBin(gpu_lib, f, 1)
        x, y = g[f"{tag}_in"], g[f"{tag}_out"]
        for t in range(x.shape[0]):
            assert bits_equal(r.run(x[t][None])[0], y[t]), (tag, t)
        r.close()


def test_unsupported_lengths(gpu_lib):
    """1025 .. 2048 taps -> the 4096-point transform, the one length the reference itself has no defined result for."""
    from libmpegh_b200.lib import MpeghError
    rng = np.random.default_rng(1)
    for length in (2048, 1025):
        f = bin_cases.brir_set(rng, 2, length, 0)
        with pytest.raises(MpeghError) as e:
            DevBin(gpu_lib, f, 1)
        assert (e.value.code & 0xFFFFFFFF) == 0xFFFF8005 and "4096" in str(e.value)


def test_full_size_properties(gpu_lib, oracle):
    """BASELINE config 4 shape (reduced stream count to keep the filter upload small): 24 inputs, 4096 taps,
    per-stream filter sets; permutation invariance + oracle spot checks."""
    rng = np.random.default_rng(3456)
    n_in, S, n_sets = 24, 64, 8
    fl = [bin_cases.brir_set(rng, n_in, 4096, 1) for _ in range(n_sets)]
    sets = rng.integers(0, n_sets, S)
    x = rng.normal(0, 0.1, size=(2, S, n_in, 4096)).astype(np.float32)
    r = DevBin(gpu_lib, fl, S, sets)
    out = [r.run(x[0]), r.run(x[1])]
    r.close()
    perm = rng.permutation(S)
    r2 = DevBin(gpu_lib, fl, S, sets[perm])
    out_p = [r2.run(x[0][perm]), r2.run(x[1][perm])]
    r2.close()
    assert bits_equal(out_p[0], out[0][perm]) and bits_equal(out_p[1], out[1][perm])
    for s in rng.choice(S, 3, replace=False):
        o = _oracle(fl[sets[s]])
        assert bits_equal(out[0][s], o.block(x[0, s])) and bits_equal(out[1][s], o.block(x[1, s])), s

