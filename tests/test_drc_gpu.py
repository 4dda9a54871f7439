"""Sub-band DRC gains on the GPU (mpeghb200_gain_subband_dev, csrc/mix.cu) against the oracle restatement of
impd_drc_apply_gains_subband (decoder/impd_drc_process.c:243) and the reference-generated golden vectors."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import drc_cases  # noqa: E402
from oracle.loader import bits_equal  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "drc_golden.npz")


def run_gpu(gpu_lib, frames):
    """all frames in one launch: channels, gain sequences and weight rows concatenated"""
    import torch
    recs, gains, weights, re, im = [], [], [], [], []
    seq0 = row0 = 0
    for fr in frames:
        recs.append(drc_cases.channel_records(fr, seq0, row0))
        gains.append(fr["gains"]), weights.append(fr["weights"].reshape(-1, 256))
        re.append(fr["re"]), im.append(fr["im"])
        seq0 += fr["gains"].shape[0]
        row0 += 4 * fr["n_groups"]
    rec = np.concatenate(recs)
    d_re = torch.from_numpy(np.concatenate(re)).cuda()
    d_im = torch.from_numpy(np.concatenate(im)).cuda()
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
    d_g = torch.from_numpy(np.concatenate(gains)).cuda()
    d_w = torch.from_numpy(np.concatenate(weights)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.gain_subband_dev(d_re.data_ptr(), d_im.data_ptr(), d_rec.data_ptr(), d_g.data_ptr(), d_w.data_ptr(),
                             len(rec), st)
    torch.cuda.synchronize()
    return d_re.cpu().numpy(), d_im.cpu().numpy()


def test_golden(gpu_lib):
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["seed"]))
    frames = [drc_cases.random_frame(rng) for _ in range(int(g["n_cases"]))]
    re, im = run_gpu(gpu_lib, frames)
    c = 0
    for i, fr in enumerate(frames):
        assert bits_equal(re[c:c + fr["n_ch"]], g[f"re_{i}"]) and bits_equal(im[c:c + fr["n_ch"]], g[f"im_{i}"]), i
        c += fr["n_ch"]


@pytest.mark.parametrize("n_frames,n_ch,seed", [(1, 1, 0), (9, None, 1), (256, 24, 2)])
def test_matches_oracle(gpu_lib, oracle, n_frames, n_ch, seed):
    rng = np.random.default_rng(1200 + seed)
    frames = [drc_cases.random_frame(rng, n_ch) for _ in range(n_frames)]
    re, im = run_gpu(gpu_lib, frames)
    c = 0
    for fr in frames:
        a, b = drc_cases.run_oracle(oracle.oracle_drc_gains_subband, fr)
        assert bits_equal(re[c:c + fr["n_ch"]], a) and bits_equal(im[c:c + fr["n_ch"]], b)
        c += fr["n_ch"]


def test_empty(gpu_lib):
    import torch
    gpu_lib.gain_subband_dev(0, 0, 0, 0, 0, 0, torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("n_ch,n_frames,seed", [(12, 3, 99), (1, 2, 5), (24 * 64, 2, 6)])
def test_fused_stft_chain(gpu_lib, oracle, n_ch, n_frames, seed):
    """mpeghb200_drc_stft_dev (domain switcher t2f -> sub-band gains -> f2t in one launch, carried state) against the
    oracle chain; the (12, 3, 99) case is also the reference-made golden."""
    import torch
    from oracle import loader
    win = np.ascontiguousarray(loader.golden_rom()["sine_256"], np.float32)
    frames = []
    want = drc_cases.stft_chain(oracle, win, np.random.default_rng(seed), n_frames, n_ch=n_ch, frames_out=frames)
    if (n_ch, n_frames, seed) == (12, 3, 99):
        assert bits_equal(want, np.load(GOLD)["chain_out"])
    st = torch.cuda.current_stream().cuda_stream
    d_state = torch.zeros(n_ch * 512, device="cuda")
    for f, (x, sets) in enumerate(frames):
        recs, gains, weights = [], [], []
        seq0 = row0 = 0
        for fr in sets:
            recs.append(drc_cases.channel_records(fr, seq0, row0))
            gains.append(fr["gains"]), weights.append(fr["weights"].reshape(-1, 256))
            seq0 += fr["gains"].shape[0]
            row0 += 4 * fr["n_groups"]
        d_rec = torch.from_numpy(np.concatenate(recs).view(np.uint8).reshape(-1)).cuda()
        d_g = torch.from_numpy(np.concatenate(gains)).cuda()
        d_w = torch.from_numpy(np.concatenate(weights)).cuda()
        d_x = torch.from_numpy(x.copy()).cuda()
        gpu_lib.drc_stft_dev(d_x.data_ptr(), d_state.data_ptr(), d_rec.data_ptr(), len(sets), d_g.data_ptr(),
                             d_w.data_ptr(), n_ch, st)
        torch.cuda.synchronize()
        assert bits_equal(d_x.cpu().numpy(), want[f]), f


@pytest.mark.parametrize("n_ch,seed", [(1, 3), (7, 4), (24 * 16, 5)])
def test_stft_halves(gpu_lib, oracle, n_ch, seed):
    """mpeghb200_stft_dev: the two halves of the domain switcher as separate launches (what the drop-in library uses
    around the interposed gain application) against the oracle's t2f / f2t (pinned to impeghd_domain_switcher_process by
    tests/test_oracle_drc.py), three frames with the carried input slot and output overlap."""
    import torch
    from oracle import loader
    win = np.ascontiguousarray(loader.golden_rom()["sine_256"], np.float32)
    rng = np.random.default_rng(seed)
    ih, oh = np.zeros((n_ch, 256), np.float32), np.zeros((n_ch, 256), np.float32)
    st = torch.cuda.current_stream().cuda_stream
    d_state = torch.zeros(n_ch * 512, device="cuda")
    for f in range(3):
        x = (rng.standard_normal((n_ch, 1024)) * 3000).astype(np.float32)
        re, im = np.zeros((n_ch, 1024), np.float32), np.zeros((n_ch, 1024), np.float32)
        oracle.oracle_ds_t2f(win, ih.reshape(-1), x.reshape(-1), re.reshape(-1), im.reshape(-1), n_ch)
        d_x = torch.from_numpy(x.copy()).cuda()
        d_spec = torch.full((n_ch, 2048), float("nan"), device="cuda")
        gpu_lib.stft_dev(d_x.data_ptr(), d_spec.data_ptr(), d_state.data_ptr(), True, n_ch, st)
        torch.cuda.synchronize()
        spec = d_spec.cpu().numpy()
        assert bits_equal(spec[:, :1024], re) and bits_equal(spec[:, 1024:], im), f
        # a gain per bin in between, like the DRC step, so that the way back does not just undo the way there
        g = (rng.random((n_ch, 2048)) + 0.5).astype(np.float32)
        spec2 = (spec * g).astype(np.float32)
        y = np.zeros((n_ch, 1024), np.float32)
        oracle.oracle_ds_f2t(win, oh.reshape(-1), np.ascontiguousarray(spec2[:, :1024]).reshape(-1),
                             np.ascontiguousarray(spec2[:, 1024:]).reshape(-1), y.reshape(-1), n_ch)
        d_spec2 = torch.from_numpy(spec2).cuda()
        d_y = torch.full((n_ch, 1024), float("nan"), device="cuda")
        gpu_lib.stft_dev(d_y.data_ptr(), d_spec2.data_ptr(), d_state.data_ptr(), False, n_ch, st)
        torch.cuda.synchronize()
        assert bits_equal(d_y.cpu().numpy(), y), f


def run_td_gpu(gpu_lib, setups_frames):
    """several independent streams (setup, frames) in one launch per frame index; -> list of [n_frames][n_ch][1024]"""
    import torch
    from libmpegh_b200.lib import DRC_TD_BANK_DTYPE, DRC_TD_CHANNEL_DTYPE
    chans, banks = [], []
    seq0 = bank0 = 0
    for setup, _ in setups_frames:
        first = np.concatenate([[0], np.cumsum(setup["banks"]["n_bands"][:setup["n_groups"]])])
        ch = np.zeros(setup["n_ch"], DRC_TD_CHANNEL_DTYPE)
        for c, g in enumerate(setup["group_of_ch"]):
            ch[c]["bank"] = bank0 + (int(g) if g >= 0 else setup["n_groups"])
            ch[c]["first_seq"] = seq0 + int(first[g]) if g >= 0 else -1
        chans.append(ch)
        banks.append(setup["banks"].astype(DRC_TD_BANK_DTYPE))
        seq0 += max(setup["n_seq"], 1)
        bank0 += len(setup["banks"])
    chan = np.concatenate(chans)
    n_ch = len(chan)
    d_chan = torch.from_numpy(chan.view(np.uint8).reshape(-1)).cuda()
    d_banks = torch.from_numpy(np.concatenate(banks).view(np.uint8).reshape(-1)).cuda()
    d_state = torch.zeros(n_ch * gpu_lib.drc_td_state_floats(), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    n_frames = len(setups_frames[0][1])
    outs = []
    for f in range(n_frames):
        d_x = torch.from_numpy(np.concatenate([fr[f][0] for _, fr in setups_frames])).cuda()
        d_g = torch.from_numpy(np.concatenate([fr[f][1] for _, fr in setups_frames])).cuda()
        gpu_lib.drc_td_dev(d_x.data_ptr(), d_state.data_ptr(), d_chan.data_ptr(), d_banks.data_ptr(), d_g.data_ptr(),
                           n_ch, st)
        torch.cuda.synchronize()
        outs.append(d_x.cpu().numpy())
    res, c = [], 0
    for setup, _ in setups_frames:
        res.append(np.stack([o[c:c + setup["n_ch"]] for o in outs]))
        c += setup["n_ch"]
    return res


def test_td_multiband_golden(gpu_lib):
    g = np.load(GOLD)
    sf = []
    for i in range(int(g["td_cases"])):
        rng = np.random.default_rng(int(g["td_seed"]) + i)
        setup = drc_cases.random_td_frame_setup(rng)
        sf.append((setup, drc_cases.td_frames(rng, setup, 2)))
    got = run_td_gpu(gpu_lib, sf)
    for i in range(len(sf)):
        assert bits_equal(got[i], g[f"td_out_{i}"]), i


@pytest.mark.parametrize("n_streams,seed", [(1, 0), (40, 1), (256, 2)])
def test_td_multiband_matches_oracle(gpu_lib, oracle, n_streams, seed):
    rng = np.random.default_rng(2200 + seed)
    sf = []
    for _ in range(n_streams):
        setup = drc_cases.random_td_frame_setup(rng)
        sf.append((setup, drc_cases.td_frames(rng, setup, 3 if n_streams < 100 else 2)))
    got = run_td_gpu(gpu_lib, sf)
    for (setup, frames), y in zip(sf, got):
        assert bits_equal(y, drc_cases.td_oracle(oracle, setup, frames))
