"""Multichannel coding tool boxes on the GPU (mpeghb200_mct_dev, csrc/mct.cu) against the oracle restatement of
impeghd_mc_process (decoder/impeghd_multichannel.c:899) and the reference-generated golden vectors; float32 results
compared as uint32."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import mct_cases  # noqa: E402
from oracle.loader import bits_equal  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "mct_golden.npz")


def run_gpu(gpu_lib, spec, groups):
    """spec [S][24][1024]; groups: list of box lists, group s works on the rows of stream s"""
    import torch
    from libmpegh_b200.lib import MCT_BOX_DTYPE
    recs, first = [], [0]
    for s, boxes in enumerate(groups):
        recs += mct_cases.resolve_group(gpu_lib, boxes, s * spec.shape[1])
        first.append(len(recs))
    rec = np.array(recs, MCT_BOX_DTYPE) if recs else np.zeros(1, MCT_BOX_DTYPE)
    d_spec = torch.from_numpy(spec.reshape(-1, 1024).copy()).cuda()
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
    d_first = torch.from_numpy(np.array(first, np.int32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.mct_dev(d_spec.data_ptr(), d_rec.data_ptr(), d_first.data_ptr(), len(groups), st)
    torch.cuda.synchronize()
    return d_spec.cpu().numpy().reshape(spec.shape)


def test_golden(gpu_lib):
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["seed"]))
    xs, groups = [], []
    for i in range(int(g["n_cases"])):
        groups.append(mct_cases.random_group(rng))
        xs.append((rng.standard_normal((24, 1024)) * 500).astype(np.float32))
    got = run_gpu(gpu_lib, np.stack(xs), groups)
    for i in range(len(xs)):
        assert bits_equal(got[i], g[f"out_{i}"]), i


@pytest.mark.parametrize("streams,seed", [(1, 0), (7, 1), (64, 2), (256, 3)])
def test_matches_oracle(gpu_lib, oracle, streams, seed):
    rng = np.random.default_rng(900 + seed)
    spec = (rng.standard_normal((streams, 24, 1024)) * 500).astype(np.float32)
    spec[:, :, ::29] = -0.0  # signed zeros survive the way the reference leaves them
    groups = [mct_cases.random_group(rng, max_boxes=16) for _ in range(streams)]
    want = np.stack([mct_cases.apply_group(oracle.oracle_mct_process, spec[s].copy(), groups[s])
                     for s in range(streams)])
    got = run_gpu(gpu_lib, spec, groups)
    assert bits_equal(got, want)


def test_empty_and_untouched(gpu_lib):
    import torch
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.mct_dev(0, 0, 0, 0, st)  # no groups: nothing launched, no error
    rng = np.random.default_rng(5)
    spec = (rng.standard_normal((2, 24, 1024)) * 500).astype(np.float32)
    # a group without boxes and a box whose mask selects nothing leave the spectra bit-identical
    bx = mct_cases.random_group(rng)[0]
    bx["fullband"], bx["mask"] = 0, np.zeros_like(bx["mask"])
    got = run_gpu(gpu_lib, spec, [[], [bx]])
    assert bits_equal(got, spec)
