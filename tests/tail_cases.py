"""Shared test-case generation for the tail of the chain (limiter, PCM packer)."""
import numpy as np


def limiter_signal(rng, n_frames, n_ch, n=1024, sigma=6000.0, burst_prob=0.3):
    """Planar frames [n_frames, n_ch, n]: noise around -15 dBFS with loud bursts so the limiter engages,
    releases and re-engages; some frames are quiet (gain stays exactly 1)."""
    x = rng.normal(0.0, sigma, size=(n_frames, n_ch, n)).astype(np.float32)
    for f in range(n_frames):
        r = rng.random()
        if r < burst_prob:
            a = int(rng.integers(0, n - 64))
            w = int(rng.integers(8, 400))
            x[f, :, a:a + w] *= np.float32(rng.uniform(3.0, 12.0))
        elif r < burst_prob + 0.25:
            x[f] *= np.float32(0.05)
    return x


def pcm_signal(rng, n_ch, n=1024):
    """Values spanning far beyond full scale, exact clamp edges, tiny negatives (truncation toward zero)."""
    x = rng.normal(0.0, 20000.0, size=(n_ch, n)).astype(np.float32)
    edge = np.array([32767.0, 32767.996, 32768.0, -32768.0, -32768.004, -32769.0, 0.0, -0.0, -0.3, 0.999,
                     -0.999, 1e9, -1e9, 32767.5, -32767.99, 127.99609375], np.float32)
    x[:, :edge.size] = edge
    return x
