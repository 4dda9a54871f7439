"""IGF restatement (oracle/oracle_igf.c) pinned against the unmodified reference: committed vectors
(tests/golden/igf_golden.npz) everywhere, live ia_core_coder_igf_mono/_tnf where oracle/_ref is built."""
import os

import numpy as np

import igf_cases
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "igf_golden.npz")


def run_oracle(oracle, sd, coef, tiles, state):
    """state: the channel's carried TNF buffers, np.zeros(2048, float32) at stream start"""
    sfb = igf_cases.SFB_SHORT if sd["win_type"] else igf_cases.SFB_LONG
    y, t = coef.copy(), np.ascontiguousarray(tiles, np.float32).copy()
    seed = oracle.oracle_igf_mono(sd.tobytes(), np.ascontiguousarray(sfb, np.int16), len(sfb), y, t.reshape(-1), state)
    return y, seed


def test_whitening_tables(oracle):
    g = np.load(GOLD)
    t = np.zeros(128, np.float32)
    oracle.oracle_igf_whitening_table(t)
    assert bits_equal(t[:63], g["whiten_pos"][:63])  # the ROM lists 63 of its 64 entries
    assert bits_equal(t[64:], g["whiten_neg"])


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    seen = set()
    for ci in range(len(igf_cases.CONFIGS)):
        sds = g[f"c{ci}_sd"].view(igf_cases.IGF_SIDE).reshape(-1)
        state = np.zeros(2048, np.float32)
        for k, sd in enumerate(sds):
            y, seed = run_oracle(oracle, sd, g[f"c{ci}_in"][k], g[f"c{ci}_tiles"][k], state)
            seen.add((int(sd["win_type"]), int(sd["is_tnf"] and sd["use_tnf"])))
            assert bits_equal(y, g[f"c{ci}_out"][k]), (ci, k)
            assert seed == int(g[f"c{ci}_seed"][k]), (ci, k)
            if not sd["all_zero"]:
                assert not bits_equal(y, g[f"c{ci}_in"][k])  # the gap was actually filled
    assert seen == {(0, 0), (0, 1), (1, 0)} or seen == {(0, 0), (0, 1), (1, 0), (1, 1)}


def test_oracle_matches_reference_live(oracle, ref):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_igf_golden as mg
    for ci, cfg in enumerate(igf_cases.CONFIGS):
        h, grid = mg.reference_grid(cfg)
        rng = np.random.default_rng(12000 + ci)
        state = np.zeros(2048, np.float32)
        for k in range(40):
            sd, coef, tiles = igf_cases.random_frame(rng, cfg, grid, win_type=int(rng.random() < 0.3))
            want = coef.copy()
            want_seed = ref.ref_igf_process(h, sd.tobytes(), want, tiles)
            y, seed = run_oracle(oracle, sd, coef, tiles, state)
            assert bits_equal(y, want), (ci, k, int(sd["win_type"]))
            assert seed == want_seed
        ref.ref_igf_destroy(h)


def test_stereo_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    for ci in range(len(igf_cases.CONFIGS)):
        sdl = g[f"s{ci}_sdl"].view(igf_cases.IGF_SIDE).reshape(-1)
        sdr = g[f"s{ci}_sdr"].view(igf_cases.IGF_SIDE).reshape(-1)
        pair = igf_cases.OracleIgfPair()
        for k in range(len(sdl)):
            y, seeds = pair.frame(sdl[k], sdr[k], g[f"s{ci}_mask"][k], g[f"s{ci}_in"][k], g[f"s{ci}_tiles"][k],
                                  int(g[f"s{ci}_ms"][k]))
            assert bits_equal(y, g[f"s{ci}_out"][k]), (ci, k, int(sdl[k]["win_type"]))
            assert np.array_equal(seeds, g[f"s{ci}_seeds"][k]), (ci, k)


def test_stereo_oracle_matches_reference_live(oracle, ref):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_igf_golden as mg
    for ci, cfg in enumerate(igf_cases.CONFIGS):
        h, grid = mg.reference_grid(cfg)
        rng = np.random.default_rng(14000 + ci)
        pair = igf_cases.OracleIgfPair()
        for k in range(30):
            sd_l, sd_r, mask, coef, tiles = igf_cases.random_pair_frame(rng, cfg, grid, win_type=int(rng.random() < 0.3))
            ms_after = int(rng.integers(2))
            want = coef.copy()
            want_seeds = np.zeros(2, np.uint32)
            ref.ref_igf_stereo_process(h, sd_l.tobytes(), sd_r.tobytes(), igf_cases.mask_reference_layout(mask, sd_l),
                                       ms_after, want[0], want[1], np.ascontiguousarray(tiles[0]).reshape(-1),
                                       np.ascontiguousarray(tiles[1]).reshape(-1), want_seeds)
            y, seeds = pair.frame(sd_l, sd_r, mask, coef, tiles, ms_after)
            assert bits_equal(y, want), (ci, k, int(sd_l["win_type"]), ms_after)
            assert np.array_equal(seeds, want_seeds), (ci, k)
        ref.ref_igf_destroy(h)
