"""Shared LTPF test cases: per-frame side info (data_present, pitch_lag_idx, gain_idx) sequences that visit every
transition mode (off->on fade-in, on->off fade-out, same parameters, changed parameters with ZIR)."""
import numpy as np


def random_sequence(rng, n_frames):
    side = np.zeros((n_frames, 3), np.int32)
    present, lag, gidx = 0, 0, 0
    for f in range(n_frames):
        r = rng.random()
        if r < 0.25:
            present = 0
        elif r < 0.55 and present:
            pass  # same parameters as the previous frame
        else:
            present, lag, gidx = 1, int(rng.integers(0, 512)), int(rng.integers(0, 4))
        side[f] = (present, lag if present else int(rng.integers(0, 512)), gidx if present else int(rng.integers(0, 4)))
    return side


def signal(rng, n_frames):
    """tonal + noise, 16-bit scale: the filter has something periodic to act on"""
    t = np.arange(n_frames * 1024)
    f0 = rng.uniform(80, 400)
    x = 4000 * np.sin(2 * np.pi * f0 / 48000 * t) + 1500 * np.sin(2 * np.pi * 2.01 * f0 / 48000 * t)
    x += rng.normal(0, 300, t.size)
    return x.astype(np.float32).reshape(n_frames, 1024)
