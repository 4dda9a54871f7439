"""CUDA LTPF (through the C ABI) against the committed reference vectors and the oracle."""
import os

import numpy as np
import pytest

import ltpf_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ltpf_golden.npz")


def run_gpu(lib, sides, x):
    """sides [C, F, 3], x [C, F, 1024] -> y [C, F, 1024]; channels in parallel, frames in order, state on the device"""
    import torch
    from libmpegh_b200.lib import LTPF_SIDE_DTYPE
    C, F = sides.shape[:2]
    state = torch.zeros(C, lib.ltpf_state_floats(48000), device="cuda")
    y = np.zeros_like(x)
    st = torch.cuda.current_stream().cuda_stream
    for f in range(F):
        sd = np.zeros(C, LTPF_SIDE_DTYPE)
        sd["unit"] = np.arange(C)
        sd["data_present"], sd["pitch_lag_idx"], sd["gain_idx"] = sides[:, f, 0], sides[:, f, 1], sides[:, f, 2]
        d_side = torch.from_numpy(sd.view(np.uint8).reshape(C, -1).copy()).cuda()
        d = torch.from_numpy(np.ascontiguousarray(x[:, f])).cuda()
        lib.ltpf_dev(48000, d.data_ptr(), d_side.data_ptr(), state.data_ptr(), C, st)
        torch.cuda.synchronize()
        y[:, f] = d.cpu().numpy()
    return y


def test_ltpf_golden(gpu_lib):
    g = np.load(GOLD)
    y = run_gpu(gpu_lib, g["side"], g["in"])
    for c in range(y.shape[0]):
        for f in range(y.shape[1]):
            assert bits_equal(y[c, f], g["out"][c, f]), (c, f, g["side"][c, f].tolist())


def test_ltpf_batch_vs_oracle(gpu_lib, oracle):
    """1500 channels x 8 frames, every transition mode many times over."""
    C, F = 1500, 8
    sides = np.zeros((C, F, 3), np.int32)
    x = np.zeros((C, F, 1024), np.float32)
    for c in range(C):
        rng = np.random.default_rng(51000 + c)
        sides[c], x[c] = ltpf_cases.random_sequence(rng, F), ltpf_cases.signal(rng, F)
    y = run_gpu(gpu_lib, sides, x)
    for c in range(C):
        o = loader.OracleLtpf()
        for f in range(F):
            want = o.process(sides[c, f, 0], sides[c, f, 1], sides[c, f, 2], x[c, f])
            assert bits_equal(y[c, f], want), (c, f, sides[c, f].tolist())
