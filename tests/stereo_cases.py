"""Shared joint-stereo (channel-pair element) test cases: side-info generator and the oracle-level phase-2 chain
stereo tools -> TNS (before or after, tns_on_lr) -> spectrum save -> FD synthesis."""
import numpy as np

import fd_cases
import tns_cases

# byte layout of mpeghb200_stereo_pair (include/mpeghb200.h) == oracle_stereo_side (oracle/oracle.h)
STEREO_SIDE = np.dtype([
    ("left", "<i4"), ("right", "<i4"), ("slot", "<i4"),
    ("tool", "u1"), ("is_short", "u1"), ("window_sequence", "u1"), ("window_shape", "u1"),
    ("window_shape_prev", "u1"), ("pred_dir", "u1"), ("complex_coef", "u1"), ("use_prev_frame", "u1"),
    ("first_band", "u1"), ("last_band", "u1"), ("prev_dmx", "u1"), ("save", "u1"),
    ("used", "u1", 128), ("alpha_re", "<i2", 128), ("alpha_im", "<i2", 128)])
assert STEREO_SIDE.itemsize == 664

SFB_LONG, SFB_SHORT = tns_cases.SFB_LONG, tns_cases.SFB_SHORT
SAMPLING_RATE_IDX = 3  # 48 kHz


def random_groups(rng):
    """window -> group index for an EIGHT_SHORT frame (contiguous groups)."""
    cuts = rng.random(7) < 0.35
    g = np.zeros(8, np.int32)
    for w in range(1, 8):
        g[w] = g[w - 1] + int(cuts[w - 1])
    return g


def random_frame(rng, seq, shape, shape_prev, tool=None):
    """One frame of side info for a pair with a common window: -> (STEREO_SIDE record, win_group[8])."""
    sd = np.zeros((), STEREO_SIDE)
    is_short = int(seq == 2)
    n_sfb = len(SFB_SHORT) if is_short else len(SFB_LONG)
    grp = random_groups(rng) if is_short else np.zeros(8, np.int32)
    sd["left"], sd["right"], sd["slot"] = 0, 1, 0
    sd["tool"] = int(rng.choice([0, 1, 3], p=[0.15, 0.3, 0.55])) if tool is None else tool
    sd["is_short"], sd["window_sequence"], sd["window_shape"], sd["window_shape_prev"] = is_short, seq, shape, shape_prev
    sd["pred_dir"] = int(rng.integers(2))
    sd["complex_coef"] = int(rng.random() < 0.6)
    sd["use_prev_frame"] = int(rng.random() < 0.6) if sd["complex_coef"] else 0
    sd["prev_dmx"] = int(rng.random() < 0.85)
    sd["save"] = 1
    last = int(rng.integers(n_sfb // 2, n_sfb + 1))
    sd["first_band"], sd["last_band"] = 0, last
    n_win = 8 if is_short else 1
    stride = 16 if is_short else 0
    n_groups = int(grp.max()) + 1 if is_short else 1
    g_used = (rng.random((n_groups, n_sfb)) < 0.7).astype(np.uint8)
    g_used[:, last:] = 0
    g_re = rng.integers(-25, 26, (n_groups, n_sfb)).astype(np.int16)
    g_im = rng.integers(-25, 26, (n_groups, n_sfb)).astype(np.int16) * int(sd["complex_coef"])
    if sd["tool"] == 3:
        g_re[g_used == 0] = 0
        g_im[g_used == 0] = 0
    else:
        g_re[:], g_im[:] = 0, 0
    used, are, aim = np.zeros(128, np.uint8), np.zeros(128, np.int16), np.zeros(128, np.int16)
    for w in range(n_win):
        used[w * stride:w * stride + n_sfb] = g_used[grp[w]]
        are[w * stride:w * stride + n_sfb] = g_re[grp[w]]
        aim[w * stride:w * stride + n_sfb] = g_im[grp[w]]
    sd["used"], sd["alpha_re"], sd["alpha_im"] = used, are, aim
    return sd, grp


def random_pair_sequence(rng, n_frames, tns_prob=0.6):
    """A pair's frames: list of dicts {sd, grp, tns_on_lr, n_filt[2,8], filt[2,8,3,20], max_sfb[2]}."""
    side = fd_cases.random_side(rng, n_frames, 1)[:, 0]
    frames = []
    for f in range(n_frames):
        seq, shape, shape_prev = (int(v) for v in side[f, :3])
        sd, grp = random_frame(rng, seq, shape, shape_prev)
        islong = int(seq != 2)
        n_sfb = len(SFB_LONG) if islong else len(SFB_SHORT)
        n_filt, filt = np.zeros((2, 8), np.int32), np.zeros((2, 8, 3, 20), np.int32)
        for ch in range(2):
            if rng.random() < tns_prob:
                n_filt[ch], filt[ch] = tns_cases.random_tns(rng, islong)
        frames.append(dict(sd=sd, grp=grp, tns_on_lr=int(rng.integers(2)), n_filt=n_filt, filt=filt,
                           max_sfb=rng.integers(n_sfb // 2, n_sfb + 1, 2).astype(np.int32)))
    return frames


class OraclePair:
    """Oracle-level phase 2 of one pair with carried state (saved spectra, FD overlap)."""

    def __init__(self):
        from oracle import loader
        self.loader = loader
        self.lib = loader.load_oracle()
        self.state = np.zeros(2048, np.float32)  # oracle_stereo_state
        self.ovl = np.zeros((2, 1024), np.float32)

    def tools(self, fr, left, right):
        """stereo tools + TNS in the frame's order; returns the two final spectra (copies)."""
        sd = fr["sd"]
        islong = int(sd["is_short"] == 0)
        sfb = SFB_LONG if islong else SFB_SHORT
        tmax = tns_cases.TNS_MAX_LONG if islong else tns_cases.TNS_MAX_SHORT
        spec = [np.ascontiguousarray(left, np.float32).copy(), np.ascontiguousarray(right, np.float32).copy()]

        def tns():
            for ch in range(2):
                spec[ch], _ = self.loader.oracle_tns_apply(spec[ch], islong, fr["n_filt"][ch], fr["filt"][ch], sfb, tmax,
                                                           int(fr["max_sfb"][ch]))

        def stereo():
            self.lib.oracle_stereo_apply(sd.tobytes(), np.ascontiguousarray(sfb, np.int16), len(sfb), self.state,
                                         spec[0], spec[1])

        if fr["tns_on_lr"]:
            stereo(), tns()
        else:
            tns(), stereo()
        self.lib.oracle_stereo_save(sd.tobytes(), self.state, spec[0], spec[1])
        return spec

    def frame(self, fr, left, right):
        sd = fr["sd"]
        spec = self.tools(fr, left, right)
        out = np.zeros((2, 1024), np.float32)
        for ch in range(2):
            assert self.lib.oracle_fd_frm_dec(spec[ch].copy(), self.ovl[ch], out[ch], int(sd["window_sequence"]),
                                              int(sd["window_shape"]), int(sd["window_shape_prev"]), 0, 0) == 0
        return out, spec
