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
