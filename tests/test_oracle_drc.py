"""Sub-band DRC gains: oracle restatement of impd_drc_apply_gains_subband (decoder/impd_drc_process.c:243) against
the reference build and the committed golden vectors."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import drc_cases  # noqa: E402
from oracle import loader  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "drc_golden.npz")


@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_reference(seed):
    r = loader.load_ref()
    if r is None:
        pytest.skip("reference build not available")
    o = loader.load_oracle()
    rng = np.random.default_rng(700 + seed)
    for _ in range(8):
        fr = drc_cases.random_frame(rng)
        a = drc_cases.run_oracle(o.oracle_drc_gains_subband, fr)
        # where the gains sit inside lpcm_gains must not matter: both delay modes, some gain delay
        b = drc_cases.run_oracle(r.ref_drc_gains_subband, fr, (int(rng.integers(0, 2)), int(rng.integers(0, 400))))
        assert loader.bits_equal(a[0], b[0]) and loader.bits_equal(a[1], b[1])


def test_oracle_matches_golden():
    o = loader.load_oracle()
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["seed"]))
    for i in range(int(g["n_cases"])):
        fr = drc_cases.random_frame(rng)
        re, im = drc_cases.run_oracle(o.oracle_drc_gains_subband, fr)
        assert loader.bits_equal(re, g[f"re_{i}"]) and loader.bits_equal(im, g[f"im_{i}"])


def _win():
    return np.ascontiguousarray(loader.golden_rom()["sine_256"], np.float32)


@pytest.mark.parametrize("n_ch", [1, 5, 24])
def test_domain_switcher_matches_reference(n_ch):
    """oracle_ds_t2f / oracle_ds_f2t vs impeghd_domain_switcher_process (decoder/ia_core_coder_process.c:505) over
    consecutive frames (carried input history and output overlap)."""
    r = loader.load_ref()
    if r is None:
        pytest.skip("reference build not available")
    o = loader.load_oracle()
    win = _win()
    rng = np.random.default_rng(40 + n_ch)
    h = r.ref_ds_create(n_ch)
    ih, oh = np.zeros((n_ch, 256), np.float32), np.zeros((n_ch, 256), np.float32)
    for _ in range(4):
        x = (rng.standard_normal((n_ch, 1024)) * 3000).astype(np.float32)
        re, im, re2, im2 = (np.zeros((n_ch, 1024), np.float32) for _ in range(4))
        assert r.ref_ds_t2f(h, x.reshape(-1), re.reshape(-1), im.reshape(-1)) == 0
        o.oracle_ds_t2f(win, ih.reshape(-1), x.reshape(-1), re2.reshape(-1), im2.reshape(-1), n_ch)
        assert loader.bits_equal(re, re2) and loader.bits_equal(im, im2)
        g = (rng.random(re.shape) * 2).astype(np.float32)
        re *= g
        im *= g
        y, y2 = np.zeros_like(re), np.zeros_like(re)
        assert r.ref_ds_f2t(h, re.reshape(-1), im.reshape(-1), y.reshape(-1)) == 0
        o.oracle_ds_f2t(win, oh.reshape(-1), re.reshape(-1), im.reshape(-1), y2.reshape(-1), n_ch)
        assert loader.bits_equal(y, y2)
    r.ref_ds_destroy(h)


def test_stft_chain_matches_golden():
    """t2f -> sub-band gains -> f2t, three frames with carried state, against vectors made by the reference."""
    o = loader.load_oracle()
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["chain_seed"]))
    out = drc_cases.stft_chain(o, _win(), rng, int(g["chain_frames"]))
    assert loader.bits_equal(out, g["chain_out"])


@pytest.mark.parametrize("seed", range(8))
def test_td_multiband_matches_reference(seed):
    """oracle_drc_td_channel vs impd_drc_filter_banks_process + impd_drc_apply_gains_and_add (static, reached through
    the -Dstatic= build), three frames with carried filter memories."""
    r = loader.load_ref_ns()
    if r is None:
        pytest.skip("reference build not available")
    o = loader.load_oracle()
    rng = np.random.default_rng(60 + seed)
    setup = drc_cases.random_td_frame_setup(rng)
    frames = drc_cases.td_frames(rng, setup, 3)
    assert loader.bits_equal(drc_cases.td_oracle(o, setup, frames), drc_cases.td_reference(r, setup, frames))


@pytest.mark.parametrize("seed", range(4))
def test_td_whole_process_matches_oracle(seed):
    """the same through impd_drc_td_process of the STOCK build (oracle/ref_harness_drc.c: impd_drc_get_gain's buffer
    bookkeeping in front, no spline nodes): what the drop-in's whole-step interposer is tested against."""
    r = loader.load_ref()
    if r is None:
        pytest.skip("reference build not available")
    o = loader.load_oracle()
    rng = np.random.default_rng(160 + seed)
    setup = drc_cases.random_td_frame_setup(rng)
    frames = drc_cases.td_frames(rng, setup, 3)
    assert loader.bits_equal(drc_cases.td_oracle(o, setup, frames), drc_cases.td_reference_whole(r, setup, frames))


def test_td_multiband_matches_golden():
    o = loader.load_oracle()
    g = np.load(GOLD)
    for i in range(int(g["td_cases"])):
        rng = np.random.default_rng(int(g["td_seed"]) + i)
        setup = drc_cases.random_td_frame_setup(rng)
        frames = drc_cases.td_frames(rng, setup, 2)
        assert loader.bits_equal(drc_cases.td_oracle(o, setup, frames), g[f"td_out_{i}"])
