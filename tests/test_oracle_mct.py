"""Multichannel coding tool: oracle restatement vs the reference build, the committed golden vectors, and the host
half (mpeghb200_mct_resolve) of the CUDA library -- no GPU needed."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import mct_cases  # noqa: E402
from oracle import loader  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mct_golden.npz")


def _spec(rng, n_ch=24):
    return (rng.standard_normal((n_ch, 1024)) * 500).astype(np.float32)


def test_rom_matches_reference():
    o = loader.load_oracle()
    a, b = np.zeros(65, np.float32), np.zeros(65, np.float32)
    o.oracle_mct_rom(a, b)
    g = np.load(GOLD)
    assert loader.bits_equal(a, g["sin_alpha"]) and loader.bits_equal(b, g["cos_alpha"])
    r = loader.load_ref()
    if r is not None:
        c, d = np.zeros(65, np.float32), np.zeros(65, np.float32)
        r.ref_mct_rom(c, d)
        assert loader.bits_equal(a, c) and loader.bits_equal(b, d)


@pytest.mark.parametrize("seed", range(8))
def test_oracle_matches_reference(seed):
    r = loader.load_ref()
    if r is None:
        pytest.skip("reference build not available")
    o = loader.load_oracle()
    rng = np.random.default_rng(300 + seed)
    for _ in range(6):
        boxes = mct_cases.random_group(rng)
        x = _spec(rng)
        a = mct_cases.apply_group(o.oracle_mct_process, x.copy(), boxes)
        b = mct_cases.apply_group(r.ref_mct_process, x.copy(), boxes)
        assert loader.bits_equal(a, b)
        assert not loader.bits_equal(a, x)


def test_oracle_matches_golden():
    o = loader.load_oracle()
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["seed"]))
    for i in range(int(g["n_cases"])):
        boxes = mct_cases.random_group(rng)
        x = _spec(rng)
        assert loader.bits_equal(x, g[f"in_{i}"])
        assert loader.bits_equal(mct_cases.apply_group(o.oracle_mct_process, x.copy(), boxes), g[f"out_{i}"])


@pytest.mark.parametrize("seed", range(4))
def test_resolve_host_half(so_lib, seed):
    """mpeghb200_mct_resolve + a numpy replay of the kernel's arithmetic == oracle (C ABI, host only)."""
    lib = so_lib
    o = loader.load_oracle()
    sin_t, cos_t = lib.mct_rom()
    a, b = np.zeros(65, np.float32), np.zeros(65, np.float32)
    o.oracle_mct_rom(a, b)
    assert loader.bits_equal(sin_t, a) and loader.bits_equal(cos_t, b)
    rng = np.random.default_rng(500 + seed)
    for _ in range(5):
        boxes = mct_cases.random_group(rng)
        x = _spec(rng)
        want = mct_cases.apply_group(o.oracle_mct_process, x.copy(), boxes)
        got = mct_cases.apply_records_numpy(x.copy(), mct_cases.resolve_group(lib, boxes, 0), sin_t, cos_t)
        assert loader.bits_equal(got, want)


def test_resolve_rejects_bad_side_info(so_lib):
    from libmpegh_b200 import lib as L
    lib = so_lib
    sfb = mct_cases.SFB_LONG
    with pytest.raises(L.MpeghError):  # rotation angle index beyond the table
        lib.mct_resolve(1, 0, 1, 1, 1024, 1, sfb, np.array([65], np.int32), np.array([1], np.int32), 0, 1)
    with pytest.raises(L.MpeghError):  # a pair must be two different channels
        lib.mct_resolve(1, 0, 1, 1, 1024, 1, sfb, np.array([3], np.int32), np.array([1], np.int32), 2, 2)
