"""LTPF restatement (oracle/oracle_ltpf.c) pinned against the unmodified reference."""
import os

import numpy as np

import ltpf_cases
from oracle import loader
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ltpf_golden.npz")


def modes_of(side):
    """transition mode per frame: 0 same, 2 fade-in, 3 fade-out, 1 changed (ZIR), -1 both off"""
    out, prev = [], (0, 0, 0)
    for p, lag, g in side:
        cur = (int(p), int(lag), int(g)) if p else (0, 0, 0)
        if not cur[0] and not prev[0]:
            out.append(-1)
        elif cur == prev:
            out.append(0)
        elif not prev[0]:
            out.append(2)
        elif not cur[0]:
            out.append(3)
        else:
            out.append(1)
        prev = cur
    return out


def test_cfg(oracle):
    g = np.load(GOLD)
    o = loader.OracleLtpf(48000)
    assert list(o.cfg) == list(g["cfg_48000"])


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    seen = set()
    for c in range(g["side"].shape[0]):
        o = loader.OracleLtpf()
        seen |= set(modes_of(g["side"][c]))
        for f in range(g["side"].shape[1]):
            s = g["side"][c, f]
            y = o.process(s[0], s[1], s[2], g["in"][c, f])
            assert bits_equal(y, g["out"][c, f]), (c, f)
    assert {0, 1, 2, 3} <= seen


def test_oracle_matches_reference_live(oracle, ref):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_ltpf_golden as mg
    for seed in range(10):
        rng = np.random.default_rng(7700 + seed)
        side, x = ltpf_cases.random_sequence(rng, 16), ltpf_cases.signal(rng, 16)
        want = mg.run_reference(side, x)
        o = loader.OracleLtpf()
        for f in range(16):
            assert bits_equal(o.process(side[f, 0], side[f, 1], side[f, 2], x[f]), want[f]), (seed, f)
