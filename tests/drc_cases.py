"""Seeded cases for the sub-band DRC gain application (decoder/impd_drc_process.c:243)."""
import numpy as np


def random_frame(rng, n_ch=None):
    """One stream-frame in the STFT256 domain: channel groups with 1..4 bands, some channels without DRC."""
    n_ch = int(rng.integers(1, 17)) if n_ch is None else n_ch
    n_groups = int(rng.integers(1, 5))
    band_count = rng.integers(1, 5, n_groups).astype(np.int32)
    group_of_ch = rng.integers(-1, n_groups, n_ch).astype(np.int32)
    n_seq = int(band_count.sum())
    gains = (rng.random((n_seq, 4)) * 2).astype(np.float32)
    weights = rng.random((n_groups, 4, 256)).astype(np.float32)
    re = (rng.standard_normal((n_ch, 1024)) * 1000).astype(np.float32)
    im = (rng.standard_normal((n_ch, 1024)) * 1000).astype(np.float32)
    re[:, ::31] = -0.0
    return dict(n_ch=n_ch, n_groups=n_groups, band_count=band_count, group_of_ch=group_of_ch, gains=gains,
                weights=weights, re=re, im=im)


def run_oracle(fn, fr, ref_args=None):
    re, im = fr["re"].copy(), fr["im"].copy()
    args = [re.reshape(-1), im.reshape(-1), fr["n_ch"], fr["group_of_ch"], fr["band_count"], fr["n_groups"],
            np.ascontiguousarray(fr["gains"]).reshape(-1), np.ascontiguousarray(fr["weights"]).reshape(-1), 4, 256]
    if ref_args is None:
        fn(*args, 1)
    else:
        assert fn(*args, 3, *ref_args) == 0  # SUBBAND_DOMAIN_MODE_STFT256, delay_mode, gain_delay_samples
    return re, im


def channel_records(fr, seq0=0, row0=0):
    """-> DRC_SB_CHANNEL_DTYPE records of the frame's channels; its gain sequences start at seq0 of the gains array
    and its weight rows (4 per group here) at row0 of the weights array"""
    from libmpegh_b200.lib import DRC_SB_CHANNEL_DTYPE
    first = np.concatenate([[0], np.cumsum(fr["band_count"])[:-1]])
    rec = np.zeros(fr["n_ch"], DRC_SB_CHANNEL_DTYPE)
    for c, g in enumerate(fr["group_of_ch"]):
        if g < 0:
            rec[c]["first_seq"] = -1
        else:
            rec[c]["first_seq"] = seq0 + int(first[g])
            rec[c]["band_count"] = int(fr["band_count"][g])
            rec[c]["weight_row"] = row0 + 4 * int(g)
    return rec


def stft_chain(oracle, win, rng, n_frames, n_ch=12, ref=None, frames_out=None):
    """Frames of the fused step (domain switcher t2f -> gains of 1..2 DRC sets -> f2t) with carried state, through the
    oracle (or, ref given, through the reference: ref_ds_* and ref_drc_gains_subband).  -> [n_frames][n_ch][1024];
    frames_out collects the per-frame side info for the GPU test."""
    ih, oh = np.zeros((n_ch, 256), np.float32), np.zeros((n_ch, 256), np.float32)
    h = ref.ref_ds_create(n_ch) if ref is not None else None
    outs = []
    for _ in range(n_frames):
        x = (rng.standard_normal((n_ch, 1024)) * 3000).astype(np.float32)
        sets = [random_frame(rng, n_ch) for _ in range(int(rng.integers(1, 3)))]
        if frames_out is not None:
            frames_out.append((x, sets))
        re, im = np.zeros((n_ch, 1024), np.float32), np.zeros((n_ch, 1024), np.float32)
        if ref is not None:
            assert ref.ref_ds_t2f(h, x.reshape(-1), re.reshape(-1), im.reshape(-1)) == 0
        else:
            oracle.oracle_ds_t2f(win, ih.reshape(-1), x.reshape(-1), re.reshape(-1), im.reshape(-1), n_ch)
        for fr in sets:
            fr = dict(fr, re=re, im=im)
            if ref is not None:
                re, im = run_oracle(ref.ref_drc_gains_subband, fr, (0, 128))
            else:
                re, im = run_oracle(oracle.oracle_drc_gains_subband, fr)
        y = np.zeros((n_ch, 1024), np.float32)
        if ref is not None:
            assert ref.ref_ds_f2t(h, re.reshape(-1), im.reshape(-1), y.reshape(-1)) == 0
        else:
            oracle.oracle_ds_f2t(win, oh.reshape(-1), re.reshape(-1), im.reshape(-1), y.reshape(-1), n_ch)
        outs.append(y)
    if h is not None:
        ref.ref_ds_destroy(h)
    return np.stack(outs)


# ---- time-domain multi-band DRC (decoder/impd_drc_process.c:399, :68; decoder/impd_drc_filter_bank.c) ---------
def random_biquad(rng):
    """a stable second-order section (a1, a2, b0, b1, b2)"""
    r, th = rng.uniform(0.3, 0.92), rng.uniform(0.05, 3.0)
    return np.array([-2 * r * np.cos(th), r * r, rng.uniform(0.2, 1.0), rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)],
                    np.float32)


def random_td_frame_setup(rng, n_ch=None):
    """channel groups with 1..4 bands, a bank per group plus one for ungrouped channels (phase alignment cascades)"""
    from oracle.loader import DRC_BANK_DTYPE
    n_ch = int(rng.integers(1, 13)) if n_ch is None else n_ch
    n_groups = int(rng.integers(1, 4))
    banks = np.zeros(n_groups + 1, DRC_BANK_DTYPE)
    for g in range(n_groups + 1):
        banks[g]["n_bands"] = int(rng.integers(1, 5)) if g < n_groups else 1
        banks[g]["n_cascade"] = int(rng.choice([0, 0, 2, 3, 5]))
        banks[g]["f"] = np.stack([random_biquad(rng) for _ in range(8)])
        banks[g]["ap"] = np.stack([random_biquad(rng) for _ in range(9)])
    group_of_ch = rng.integers(-1, n_groups, n_ch).astype(np.int32)
    return dict(n_ch=n_ch, n_groups=n_groups, banks=banks, group_of_ch=group_of_ch,
                n_seq=int(banks["n_bands"][:n_groups].sum()))


def td_frames(rng, setup, n_frames):
    return [((rng.standard_normal((setup["n_ch"], 1024)) * 3000).astype(np.float32),
             (rng.random((max(setup["n_seq"], 1), 1024)) + 0.5).astype(np.float32)) for _ in range(n_frames)]


def td_oracle(oracle, setup, frames):
    """-> [n_frames][n_ch][1024] through oracle_drc_td_channel, state carried per channel"""
    first = np.concatenate([[0], np.cumsum(setup["banks"]["n_bands"][:setup["n_groups"]])])
    state = np.zeros((setup["n_ch"], 17 * 4), np.float32)
    outs = []
    for x, gains in frames:
        y = x.copy()
        for c in range(setup["n_ch"]):
            g = int(setup["group_of_ch"][c])
            bank = setup["banks"][g if g >= 0 else setup["n_groups"]:][:1]
            row = np.ascontiguousarray(y[c])
            gp = None
            if g >= 0:
                gg = np.ascontiguousarray(gains[int(first[g]):int(first[g]) + int(bank["n_bands"][0])])
                gp = gg.ctypes.data
            oracle.oracle_drc_td_channel(bank.ctypes.data, state[c], row, gp)
            y[c] = row
        outs.append(y)
    return np.stack(outs)


def td_reference(ref_ns, setup, frames):
    h = ref_ns.ref_drctd_create(setup["n_ch"], setup["group_of_ch"], setup["n_groups"], setup["banks"].ctypes.data)
    outs = []
    for x, gains in frames:
        y = x.copy()
        assert ref_ns.ref_drctd_process(h, y.reshape(-1), np.ascontiguousarray(gains).reshape(-1)) == 0
        outs.append(y)
    ref_ns.ref_drctd_destroy(h)
    return np.stack(outs)


def td_reference_whole(ref, setup, frames):
    """the same through impd_drc_td_process (gain-buffer bookkeeping of impd_drc_get_gain included); stock library"""
    h = ref.ref_drctd_whole_create(setup["n_ch"], setup["group_of_ch"], setup["n_groups"], setup["banks"].ctypes.data)
    outs = []
    for x, gains in frames:
        y = x.copy()
        assert ref.ref_drctd_whole_process(h, y.reshape(-1), np.ascontiguousarray(gains).reshape(-1)) == 0
        outs.append(y)
    ref.ref_drctd_whole_destroy(h)
    return np.stack(outs)
