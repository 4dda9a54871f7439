"""CUDA limiter + PCM packer (through the C ABI) against the oracle and the committed reference vectors.
Bit-exact: float32 compared as uint32, PCM bytes compared byte for byte."""
import os

import numpy as np
import pytest

import tail_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tail_golden.npz")


class DevLimiter:
    def __init__(self, lib, n_streams, n_ch, limiter_on=1):
        import torch
        self.torch, self.lib, self.S, self.C = torch, lib, n_streams, n_ch
        self.p = lib.limiter_params(n_ch, 48000, limiter_on)
        self.state = torch.empty(n_streams * lib.limiter_state_floats(self.p), device="cuda")
        lib.limiter_state_init_dev(self.p, self.state.data_ptr(), n_streams)

    def run(self, x, gain=None):
        """x [S, C, 1024] -> limited [S, C, 1024]"""
        t = self.torch
        d_in = t.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        d_out = t.empty_like(d_in)
        d_g = None if gain is None else t.from_numpy(np.ascontiguousarray(gain, dtype=np.float32)).cuda()
        self.lib.limiter_dev(self.p, d_in.data_ptr(), d_out.data_ptr(), self.state.data_ptr(), self.S,
                             None if d_g is None else d_g.data_ptr(), t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return d_out.cpu().numpy()


def test_limiter_golden(gpu_lib):
    g = np.load(GOLD)
    F, C = g["lim_in"].shape[:2]
    lim = DevLimiter(gpu_lib, 1, C)
    for f in range(F):
        assert bits_equal(lim.run(g["lim_in"][f][None])[0], g["lim_out"][f]), f


@pytest.mark.parametrize("n_ch,limiter_on", [(1, 1), (2, 1), (12, 1), (24, 1), (5, 0)])
def test_limiter_streams_vs_oracle(gpu_lib, oracle, n_ch, limiter_on):
    """19 independent streams, 12 frames, state carried on the device; quiet frames take the skip path."""
    rng = np.random.default_rng(7 + n_ch)
    S, F = 19, 12
    x = np.stack([tail_cases.limiter_signal(rng, F, n_ch) for _ in range(S)], axis=1)  # [F, S, C, 1024]
    x[:, 3] *= np.float32(0.01)  # one stream never reaches the threshold
    sts = [loader.oracle_limiter_new(n_ch, 48000, limiter_on) for _ in range(S)]
    lim = DevLimiter(gpu_lib, S, n_ch, limiter_on)
    for f in range(F):
        got = lim.run(x[f])
        for s in range(S):
            assert bits_equal(got[s], loader.oracle_limiter_run(sts[s], x[f, s])), (f, s)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 7])
def test_limiter_launch_variants(oracle, variant, monkeypatch):
    """Every launch form of the limiter (MPEGHB200_LIM_VARIANT: 128 / 256 threads per stream, the release fast path of the
    smoother, batched channel loads for small launches -- the default, 7, takes the last one at these sizes and the plain
    256-thread kernel beyond four streams per SM) against the oracle, with a loudness gain, over engaged, releasing and
    quiet frames."""
    from libmpegh_b200.lib import Lib
    monkeypatch.setenv("MPEGHB200_LIM_VARIANT", str(variant))
    lib = Lib().create(0)
    try:
        rng = np.random.default_rng(100 + variant)
        S, C, F = 7, 12, 10
        x = np.stack([tail_cases.limiter_signal(rng, F, C) for _ in range(S)], axis=1)
        x[:, 2] *= np.float32(0.01)
        gains = np.array([oracle.oracle_loudness_gain(db) for db in (-6.0, 0.5, 3.0, 12.0, 0.0, -1.0, 2.0)], np.float32)
        sts = [loader.oracle_limiter_new(C) for _ in range(S)]
        lim = DevLimiter(lib, S, C)
        for f in range(F):
            got = lim.run(x[f], gains)
            for s in range(S):
                assert bits_equal(got[s], loader.oracle_limiter_run(sts[s], x[f, s] * gains[s])), (variant, f, s)
    finally:
        lib.close()


def test_limiter_with_loudness_gain(gpu_lib, oracle):
    rng = np.random.default_rng(11)
    S, C, F = 4, 6, 5
    x = np.stack([tail_cases.limiter_signal(rng, F, C) for _ in range(S)], axis=1)
    gains = np.array([oracle.oracle_loudness_gain(db) for db in (-6.0, 0.5, 3.0, 12.0)], np.float32)
    sts = [loader.oracle_limiter_new(C) for _ in range(S)]
    lim = DevLimiter(gpu_lib, S, C)
    for f in range(F):
        got = lim.run(x[f], gains)
        for s in range(S):
            want = loader.oracle_limiter_run(sts[s], x[f, s] * gains[s])  # float32 * float32: one rounding
            assert bits_equal(got[s], want), (f, s)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_pcm_pack_vs_oracle(gpu_lib, oracle, bits):
    import torch
    rng = np.random.default_rng(bits)
    for n_ch, cmap, delays in [(1, None, [0, 0, 0]), (2, [1, 0], [0, 240, 1024]), (12, None, [7, 1500, 256]),
                               (24, list(range(23, -1, -1)), [0, 1, 1023])]:
        S = len(delays)
        x = np.stack([tail_cases.pcm_signal(rng, n_ch) for _ in range(S)])
        d_in = torch.from_numpy(x).cuda()
        bps = bits // 8
        d_out = torch.zeros(S, 1024 * n_ch * bps, dtype=torch.uint8, device="cuda")
        d_delay = torch.tensor(delays, dtype=torch.int32, device="cuda")
        d_nb = torch.zeros(S, dtype=torch.int32, device="cuda")
        gpu_lib.pcm_pack_dev(d_in.data_ptr(), n_ch, bits, d_out.data_ptr(), S, cmap, d_delay.data_ptr(), d_nb.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out, nb, nd = d_out.cpu().numpy(), d_nb.cpu().numpy(), d_delay.cpu().numpy()
        for s in range(S):
            want, d2 = loader.oracle_samples_sat(x[s], bits, cmap, delays[s])
            assert nb[s] == want.size and nd[s] == d2, (n_ch, s)
            assert np.array_equal(out[s, :want.size], want), (bits, n_ch, s)


def test_pcm_pack_golden(gpu_lib):
    import torch
    g = np.load(GOLD)
    d_in = torch.from_numpy(g["pcm_in"][None]).cuda()
    for bits in (16, 24, 32):
        d_out = torch.zeros(1, 1024 * 3 * (bits // 8), dtype=torch.uint8, device="cuda")
        d_delay = torch.tensor([int(g["pcm_delay"])], dtype=torch.int32, device="cuda")
        gpu_lib.pcm_pack_dev(d_in.data_ptr(), 3, bits, d_out.data_ptr(), 1, g["pcm_map"], d_delay.data_ptr(), None,
                             torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = g[f"pcm_out{bits}"]
        assert np.array_equal(d_out.cpu().numpy()[0, :want.size], want), bits


def test_fd_limiter_pack_chain_24bit(gpu_lib, oracle):
    """FD synthesis -> limiter -> 24-bit PCM on the device vs the same chain in the oracle: final PCM identical
    (north_star: +-1 LSB at 24 bit; asserted here with tolerance 0)."""
    import torch
    import fd_cases
    rng = np.random.default_rng(2026)
    S, C, F = 3, 4, 6
    U = S * C
    side = fd_cases.random_side(rng, F, U)
    coef = fd_cases.random_coef(rng, (F, U, 1024), sigma=9000.0)  # loud enough to engage the limiter
    ov = np.zeros((U, 1024), np.float32)
    td = np.zeros((F, U, 1024), np.float32)
    assert oracle.oracle_fd_frm_dec_batch(coef, ov, td, side, F, U) == 0
    sts = [loader.oracle_limiter_new(C) for _ in range(S)]
    lim = DevLimiter(gpu_lib, S, C)
    d_ov = torch.zeros(U, 1024, device="cuda")
    d_delay = torch.full((S,), 240, dtype=torch.int32, device="cuda")
    delays = [240] * S
    st = torch.cuda.current_stream().cuda_stream
    for f in range(F):
        d_coef = torch.from_numpy(coef[f]).cuda()
        d_side = torch.from_numpy(fd_cases.pack_side(side[f]).astype(np.int32)).cuda()
        d_td = torch.empty(U, 1024, device="cuda")
        d_lim = torch.empty(U, 1024, device="cuda")
        d_pcm = torch.zeros(S, 1024 * C * 3, dtype=torch.uint8, device="cuda")
        d_nb = torch.zeros(S, dtype=torch.int32, device="cuda")
        gpu_lib.fd_synth_dev(d_coef.data_ptr(), d_side.data_ptr(), d_ov.data_ptr(), d_td.data_ptr(), U, st)
        gpu_lib.limiter_dev(lim.p, d_td.data_ptr(), d_lim.data_ptr(), lim.state.data_ptr(), S, None, st)
        gpu_lib.pcm_pack_dev(d_lim.data_ptr(), C, 24, d_pcm.data_ptr(), S, None, d_delay.data_ptr(), d_nb.data_ptr(), st)
        torch.cuda.synchronize()
        pcm, nb = d_pcm.cpu().numpy(), d_nb.cpu().numpy()
        for s in range(S):
            y = loader.oracle_limiter_run(sts[s], td[f, s * C:(s + 1) * C])
            want, delays[s] = loader.oracle_samples_sat(y, 24, None, delays[s])
            assert nb[s] == want.size and np.array_equal(pcm[s, :want.size], want), (f, s)


@pytest.mark.parametrize("n_ch,bits,limiter_on,cmap", [(2, 16, 1, [1, 0]), (12, 24, 1, None), (24, 24, 1, list(range(23, -1, -1))),
                                                       (5, 32, 0, None), (1, 24, 1, None)])
def test_fused_limiter_pack_equals_the_two_kernels(gpu_lib, n_ch, bits, limiter_on, cmap):
    """mpeghb200_limiter_pack_dev (the chain tail) = mpeghb200_limiter_dev + mpeghb200_pcm_pack_dev: PCM bytes, byte
    counts, delay counters and the limiter state identical over 6 frames of 9 streams (start-up delays 0 .. 1500,
    a loudness gain per stream)."""
    import torch
    rng = np.random.default_rng(100 * n_ch + bits)
    S, F = 9, 6
    x = np.stack([tail_cases.limiter_signal(rng, F, n_ch) for _ in range(S)], axis=1)  # [F, S, C, 1024]
    gains = torch.from_numpy(rng.uniform(0.5, 2.0, S).astype(np.float32)).cuda()
    delays = [0, 240, 1, 1023, 1024, 1500, 7, 256, 512]
    a, b = DevLimiter(gpu_lib, S, n_ch, limiter_on), DevLimiter(gpu_lib, S, n_ch, limiter_on)
    da = torch.tensor(delays, dtype=torch.int32, device="cuda")
    db = da.clone()
    st = torch.cuda.current_stream().cuda_stream
    nbytes = 1024 * n_ch * (bits // 8)
    scratch = torch.empty(S * gpu_lib.limiter_pack_scratch_floats(b.p), device="cuda")
    for f in range(F):
        d_in = torch.from_numpy(x[f]).cuda()
        d_lim = torch.empty_like(d_in)
        pa = torch.zeros(S, nbytes, dtype=torch.uint8, device="cuda")
        pb = torch.zeros(S, nbytes, dtype=torch.uint8, device="cuda")
        na = torch.zeros(S, dtype=torch.int32, device="cuda")
        nb = torch.zeros(S, dtype=torch.int32, device="cuda")
        gpu_lib.limiter_dev(a.p, d_in.data_ptr(), d_lim.data_ptr(), a.state.data_ptr(), S, gains.data_ptr(), st)
        gpu_lib.pcm_pack_dev(d_lim.data_ptr(), n_ch, bits, pa.data_ptr(), S, cmap, da.data_ptr(), na.data_ptr(), st)
        gpu_lib.limiter_pack_dev(b.p, d_in.data_ptr(), b.state.data_ptr(), scratch.data_ptr(), bits, pb.data_ptr(), S,
                                 gains.data_ptr(), cmap, db.data_ptr(), nb.data_ptr(), st)
        torch.cuda.synchronize()
        na_, nb_ = na.cpu().numpy(), nb.cpu().numpy()
        assert np.array_equal(na_, nb_) and np.array_equal(da.cpu().numpy(), db.cpu().numpy()), f
        pa_, pb_ = pa.cpu().numpy(), pb.cpu().numpy()
        for s in range(S):
            assert np.array_equal(pa_[s, :na_[s]], pb_[s, :nb_[s]]), (f, s)
        assert np.array_equal(a.state.cpu().numpy().view(np.uint32), b.state.cpu().numpy().view(np.uint32)), f
