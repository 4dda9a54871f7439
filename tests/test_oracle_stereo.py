"""Joint-stereo restatement (oracle/oracle_stereo.c) pinned against the unmodified reference: committed vectors
(tests/golden/stereo_golden.npz) everywhere, live ia_core_coder_data_process where oracle/_ref is built."""
import os

import numpy as np
import pytest

import stereo_cases
import tns_cases
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "stereo_golden.npz")


def golden_frames(g, p):
    sd = g[f"p{p}_sd"].view(stereo_cases.STEREO_SIDE).reshape(-1)
    return [dict(sd=sd[f], grp=np.ascontiguousarray(g[f"p{p}_grp"][f]), tns_on_lr=int(g[f"p{p}_tns_on_lr"][f]),
                 n_filt=g[f"p{p}_n_filt"][f], filt=g[f"p{p}_filt"][f], max_sfb=g[f"p{p}_max_sfb"][f])
            for f in range(len(sd))]


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    tools_seen = set()
    for p in range(6):
        frames = golden_frames(g, p)
        pair = stereo_cases.OraclePair()
        for f, fr in enumerate(frames):
            out, spec = pair.frame(fr, g[f"p{p}_coef"][f, 0], g[f"p{p}_coef"][f, 1])
            sd = fr["sd"]
            tools_seen.add((int(sd["tool"]), int(sd["complex_coef"]), int(sd["use_prev_frame"]), int(sd["is_short"])))
            assert bits_equal(out, g[f"p{p}_pcm"][f]), (p, f, sd["tool"], sd["is_short"])
            bins = 128 if sd["is_short"] else 1024
            want = g[f"p{p}_spec"][f]
            for ch in range(2):
                assert bits_equal(spec[ch][1024 - bins:], want[ch][1024 - bins:]), (p, f, ch)
    # the vectors exercise M/S, real and complex prediction with and without the previous frame, long and short
    assert {t[0] for t in tools_seen} == {0, 1, 3}
    assert (3, 1, 1, 0) in tools_seen and (3, 1, 1, 1) in tools_seen and (3, 0, 0, 0) in tools_seen


def test_oracle_matches_reference_live(oracle, ref):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_stereo_golden as msg
    for seed in range(12):
        rng = np.random.default_rng(9000 + seed)
        frames = stereo_cases.random_pair_sequence(rng, 8)
        coef = rng.normal(0, 800, (8, 2, 1024)).astype(np.float32)
        pcm, _ = msg.run_reference(frames, coef)
        pair = stereo_cases.OraclePair()
        for f, fr in enumerate(frames):
            out, _ = pair.frame(fr, coef[f, 0], coef[f, 1])
            assert bits_equal(out, pcm[f]), (seed, f)


def test_save_zeros_and_keep(oracle):
    pair = stereo_cases.OraclePair()
    rng = np.random.default_rng(1)
    l, r = (rng.normal(0, 100, 1024).astype(np.float32) for _ in range(2))
    sd = np.zeros((), stereo_cases.STEREO_SIDE)
    sd["save"] = 1
    oracle.oracle_stereo_save(sd.tobytes(), pair.state, l, r)
    assert np.array_equal(pair.state[:1024], l) and np.array_equal(pair.state[1024:], r)
    sd["is_short"], sd["save"] = 1, 2
    oracle.oracle_stereo_save(sd.tobytes(), pair.state, l, r)
    assert np.all(pair.state[896:1024] == 0) and np.array_equal(pair.state[:896], l[:896])
    sd["save"] = 0
    before = pair.state.copy()
    oracle.oracle_stereo_save(sd.tobytes(), pair.state, r, l)
    assert np.array_equal(before, pair.state)
