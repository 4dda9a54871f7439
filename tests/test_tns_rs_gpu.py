"""CUDA TNS and resampler (through the C ABI) against the oracle and the committed reference vectors."""
import os

import numpy as np
import pytest

import tns_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tns_rs_golden.npz")


def _run_tns(lib, coef, chains):
    import torch
    from libmpegh_b200.lib import TNS_FILTER_DTYPE
    recs = [r for ch in chains for r in ch]
    first = np.cumsum([0] + [len(ch) for ch in chains]).astype(np.int32)
    d_coef = torch.from_numpy(np.ascontiguousarray(coef, np.float32)).cuda()
    if recs:
        arr = np.array(recs, TNS_FILTER_DTYPE)
        d_f = torch.from_numpy(arr.view(np.uint8).reshape(len(recs), -1).copy()).cuda()
        d_first = torch.from_numpy(first).cuda()
        lib.tns_dev(d_coef.data_ptr(), d_f.data_ptr(), d_first.data_ptr(), len(chains),
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_coef.cpu().numpy()


def test_tns_golden(gpu_lib):
    g = np.load(GOLD)
    for tag, islong in (("long", 1), ("short", 0)):
        x, y, nf, fl = (g[f"tns_{tag}_{k}"] for k in ("in", "out", "nfilt", "filt"))
        nb = len(tns_cases.SFB_LONG if islong else tns_cases.SFB_SHORT)
        chains = []
        for u in range(x.shape[0]):
            chains += tns_cases.resolve_filters(gpu_lib, u, islong, nf[u], fl[u], nb, gpu_lib.tns_lpc_from_coef)
        assert len(chains) > 3
        got = _run_tns(gpu_lib, x, chains)
        assert bits_equal(got, y), tag


def test_tns_batch_vs_oracle(gpu_lib, oracle):
    """2000 channel-frames, long and short mixed, random band limits."""
    rng = np.random.default_rng(9)
    U = 2000
    x = rng.normal(0, 500, (U, 1024)).astype(np.float32)
    want = x.copy()
    chains = []
    for u in range(U):
        islong = int(rng.random() < 0.7)
        if rng.random() < 0.4:
            continue  # no TNS on this channel
        sfb = tns_cases.SFB_LONG if islong else tns_cases.SFB_SHORT
        tmax = tns_cases.TNS_MAX_LONG if islong else tns_cases.TNS_MAX_SHORT
        n_filt, filt = tns_cases.random_tns(rng, islong)
        nbands = int(rng.integers(len(sfb) // 2, len(sfb) + 1))
        want[u], _ = loader.oracle_tns_apply(x[u], islong, n_filt, filt, sfb, tmax, nbands)
        chains += tns_cases.resolve_filters(gpu_lib, u, islong, n_filt, filt, nbands, gpu_lib.tns_lpc_from_coef)
    got = _run_tns(gpu_lib, x, chains)
    assert bits_equal(got, want)
    assert bits_equal(_run_tns(gpu_lib, x, []), x)


class DevRs:
    def __init__(self, lib, up, down, f2, f3, n_units):
        import torch
        self.t, self.lib, self.U = torch, lib, n_units
        self.h = lib.resampler_create(up, down, f3 if up == 3 else f2, f2 if down == 2 else None)
        self.n_out = lib.resampler_out_len(self.h)
        self.state = torch.zeros(n_units * lib.resampler_state_floats(self.h), device="cuda")

    def run(self, x):
        t = self.t
        d_in = t.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
        d_out = t.full((self.U, self.n_out), float("nan"), device="cuda")
        self.lib.resample_dev(self.h, d_in.data_ptr(), d_out.data_ptr(), self.state.data_ptr(), self.U,
                              t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return d_out.cpu().numpy()

    def close(self):
        self.lib.resampler_destroy(self.h)


@pytest.mark.parametrize("up,down", [(2, 1), (3, 1), (3, 2)])
def test_resampler(gpu_lib, oracle, up, down):
    g = np.load(GOLD)
    f2, f3 = g["rs_filter_2"], g["rs_filter_3"]
    r = DevRs(gpu_lib, up, down, f2, f3, 1)
    x, y = g[f"rs_{up}_{down}_in"], g[f"rs_{up}_{down}_out"]
    for f in range(x.shape[0]):
        assert bits_equal(r.run(x[f][None])[0], y[f]), (up, down, f)
    r.close()
    U, F = 37, 4
    rng = np.random.default_rng(up + down)
    r = DevRs(gpu_lib, up, down, f2, f3, U)
    orc = [loader.OracleResampler(up, down, f2, f3) for _ in range(U)]
    for f in range(F):
        xs = rng.normal(0, 5000, (U, 1024)).astype(np.float32)
        got = r.run(xs)
        for u in range(U):
            assert bits_equal(got[u], orc[u].process(xs[u])), (f, u)
    r.close()


def test_resampler_rejects_other_ratios(gpu_lib):
    from libmpegh_b200.lib import MpeghError
    with pytest.raises(MpeghError):
        gpu_lib.resampler_create(1, 2, np.zeros(60, np.float32), np.zeros(60, np.float32))
