"""Shared format-converter test cases."""
import numpy as np

# (cicp_in, cicp_out): BASELINE config 5's bed conversion first (22.2 -> 7.1.4)
PAIRS = [(13, 19), (13, 6), (19, 6), (6, 2), (14, 6), (13, 2), (12, 6)]


def key(p):
    return f"c{p[0]}_c{p[1]}"


def dense_matrix(rng, n_in, n_out):
    """Arbitrary frequency-dependent tensor with some structurally absent (band 0 == 0) pairs."""
    D = rng.uniform(-1.0, 1.0, size=(n_out, n_in, 257)).astype(np.float32)
    D[rng.random((n_out, n_in)) < 0.3, 0] = 0
    return D


def signal(rng, n_frames, n_ch, sigma=3000.0):
    x = rng.normal(0, sigma, size=(n_frames, n_ch, 1024)).astype(np.float32)
    t = np.arange(n_frames * 1024).reshape(n_frames, 1, 1024)
    x += (6000.0 * np.sin(2 * np.pi * 0.01 * t)).astype(np.float32)   # correlated part: exercises the EQ
    return x
