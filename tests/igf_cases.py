"""Shared IGF (intelligent gap filling, mono path) test cases."""
import numpy as np

import stereo_cases

# byte layout of mpeghb200_igf_side (include/mpeghb200.h) == oracle_igf_side (oracle/oracle.h)
IGF_SIDE = np.dtype([
    ("unit", "<i4"), ("seed", "<u4"),
    ("win_type", "u1"), ("sfb_start", "u1"), ("sfb_stop", "u1"), ("igf_min", "u1"),
    ("use_high_res", "u1"), ("use_whitening", "u1"), ("use_tnf", "u1"), ("all_zero", "u1"), ("is_tnf", "u1"),
    ("num_groups", "u1"), ("pad", "u1", 2),
    ("group_len", "u1", 8), ("tile_num", "u1", 4), ("whitening_level", "u1", 4),
    ("levels", "<f4", 128)])
assert IGF_SIDE.itemsize == 548

SFB_LONG, SFB_SHORT = stereo_cases.SFB_LONG, stereo_cases.SFB_SHORT

# (igf_start_index, igf_stop_index, use_high_res, use_whitening, use_enf) -- the configurations the vectors cover
CONFIGS = [(0, 15, 0, 1, 1), (4, 15, 1, 1, 1), (8, 10, 0, 1, 0), (2, 13, 1, 0, 1), (12, 15, 0, 1, 1), (6, 5, 1, 1, 1)]


def random_frame(rng, cfg, grid, win_type, all_zero_prob=0.05):
    """cfg = CONFIGS entry, grid = [2][4] {sfb_start, sfb_stop, igf_min, num_tiles} -> (side, coef, tiles)"""
    sd = np.zeros((), IGF_SIDE)
    g = grid[win_type]
    sfb = SFB_SHORT if win_type else SFB_LONG
    bins = 128 if win_type else 1024
    n_win = 8 if win_type else 1
    sd["seed"] = int(rng.integers(0, 2**32))
    sd["win_type"], sd["sfb_start"], sd["sfb_stop"], sd["igf_min"] = win_type, g[0], g[1], g[2]
    sd["use_high_res"], sd["use_whitening"], sd["use_tnf"] = cfg[2], cfg[3], cfg[4]
    sd["all_zero"] = int(rng.random() < all_zero_prob)
    sd["is_tnf"] = int(rng.random() < 0.6)
    if win_type:
        grp = stereo_cases.random_groups(rng)
        n_groups = int(grp.max()) + 1
        glen = np.bincount(grp, minlength=8).astype(np.uint8)
    else:
        n_groups, glen = 1, np.array([1, 0, 0, 0, 0, 0, 0, 0], np.uint8)
    sd["num_groups"], sd["group_len"] = n_groups, glen
    sd["tile_num"] = rng.integers(0, 4, 4)
    sd["whitening_level"] = rng.integers(0, 3, 4)
    lev = np.zeros(128, np.float32)
    stride = 16 if win_type else 0
    for gi in range(n_groups):
        lev[stride * gi:stride * gi + len(sfb)] = np.exp2(0.25 * rng.integers(0, 44, len(sfb))).astype(np.float32)
    sd["levels"] = lev
    bgn = int(sfb[g[0] - 1])
    coef = np.zeros((n_win, bins), np.float32)
    for w in range(n_win):
        c = rng.normal(0, 500, bins).astype(np.float32)
        c[rng.random(bins) < 0.15] = 0.0
        hi = np.round(rng.normal(0, 40, bins - bgn)).astype(np.float32)
        hi[rng.random(bins - bgn) < 0.75] = 0.0
        c[bgn:] = hi
        coef[w] = c
    tiles = rng.normal(0, 500, (4, 1024)).astype(np.float32)
    tiles[rng.random((4, 1024)) < 0.1] = 0.0
    if rng.random() < 0.1:
        tiles[int(rng.integers(4))] = 0.0  # a silent source tile: zero source energy, gain 0
    return sd, coef.reshape(1024), tiles


def random_pair_frame(rng, cfg, grid, win_type):
    """Joint-stereo IGF frame of a channel pair: -> (sd_l, sd_r, mask[128], coef[2,1024], tiles[2,4,1024]);
    both sides share grid, window type and grouping and are synchronised the way ia_core_coder_sync_data does."""
    sd_l, c_l, t_l = random_frame(rng, cfg, grid, win_type)
    sd_r, c_r, t_r = random_frame(rng, cfg, grid, win_type)
    sd_r["num_groups"], sd_r["group_len"] = sd_l["num_groups"], sd_l["group_len"]
    if sd_l["all_zero"] and not sd_r["all_zero"]:
        sd_l["tile_num"], sd_l["whitening_level"], sd_l["is_tnf"] = sd_r["tile_num"], sd_r["whitening_level"], sd_r["is_tnf"]
    elif sd_r["all_zero"] and not sd_l["all_zero"]:
        sd_r["tile_num"], sd_r["whitening_level"], sd_r["is_tnf"] = sd_l["tile_num"], sd_l["whitening_level"], sd_l["is_tnf"]
    n_sfb = len(SFB_SHORT) if win_type else len(SFB_LONG)
    stride = 16 if win_type else 0
    mask = np.zeros(128, np.uint8)
    for g in range(int(sd_l["num_groups"])):
        mask[stride * g:stride * g + n_sfb] = rng.random(n_sfb) < 0.5
    return sd_l, sd_r, mask, np.stack([c_l, c_r]), np.stack([t_l, t_r])


def mask_reference_layout(mask, sd):
    """oracle layout (16 per group) -> the reference's ms_used layout (sfb_per_sbk per group)"""
    win_type = int(sd["win_type"])
    n_sfb = len(SFB_SHORT) if win_type else len(SFB_LONG)
    stride = 16 if win_type else 0
    out = np.zeros(8 * 64, np.uint8)
    for g in range(int(sd["num_groups"])):
        out[g * n_sfb:(g + 1) * n_sfb] = mask[stride * g:stride * g + n_sfb]
    return out


def mask_per_window(mask, sd):
    """oracle per-group mask -> per-window `used` array of a STEREO_SIDE record"""
    win_type = int(sd["win_type"])
    if not win_type:
        return mask.copy()
    out = np.zeros(128, np.uint8)
    w = 0
    for g in range(int(sd["num_groups"])):
        for _ in range(int(sd["group_len"][g])):
            out[16 * w:16 * w + 16] = mask[16 * g:16 * g + 16]
            w += 1
    return out


class OracleIgfPair:
    """Oracle-level joint-stereo IGF of one pair with carried TNF buffers."""

    def __init__(self):
        from oracle import loader
        self.lib = loader.load_oracle()
        self.tnf = np.zeros((2, 2048), np.float32)

    def frame(self, sd_l, sd_r, mask, coef, tiles, ms_after):
        sfb = np.ascontiguousarray(SFB_SHORT if sd_l["win_type"] else SFB_LONG, np.int16)
        c = [coef[0].copy(), coef[1].copy()]
        t = [np.ascontiguousarray(tiles[0]).reshape(-1).copy(), np.ascontiguousarray(tiles[1]).reshape(-1).copy()]
        seeds = np.zeros(2, np.uint32)
        self.lib.oracle_igf_stereo(sd_l.tobytes(), sd_r.tobytes(), np.ascontiguousarray(mask), sfb, len(sfb), c[0], c[1],
                                   t[0], t[1], self.tnf[0], self.tnf[1], seeds)
        if ms_after:
            ms = np.zeros((), stereo_cases.STEREO_SIDE)
            ms["tool"], ms["is_short"] = 1, sd_l["win_type"]
            ms["first_band"], ms["last_band"] = sd_l["sfb_start"], sd_l["sfb_stop"]
            ms["used"] = mask_per_window(mask, sd_l)
            dummy = np.zeros(2048, np.float32)
            self.lib.oracle_stereo_apply(ms.tobytes(), sfb, len(sfb), dummy, c[0], c[1])
            if sd_l["use_tnf"] and not sd_l["win_type"]:
                a, b = self.tnf[0][:1024].copy(), self.tnf[1][:1024].copy()
                self.lib.oracle_stereo_apply(ms.tobytes(), sfb, len(sfb), dummy, a, b)
                self.tnf[0][:1024], self.tnf[1][:1024] = a, b
        self.lib.oracle_igf_tnf(sd_l.tobytes(), sfb, len(sfb), c[0], self.tnf[0])
        self.lib.oracle_igf_tnf(sd_r.tobytes(), sfb, len(sfb), c[1], self.tnf[1])
        return np.stack(c), seeds
