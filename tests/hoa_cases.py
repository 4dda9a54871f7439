"""Shared HOA-renderer test-case generation."""
import numpy as np

# (cicp, order, lfe_enabled, nfc_radius): config 3 of BASELINE.json first
CONFIGS = [(13, 4, 1, 1.5), (13, 4, 1, 0.0), (6, 3, 1, 0.0), (19, 4, 0, 0.8), (2, 1, 0, 0.0), (13, 6, 1, 2.0),
           (14, 5, 1, 1.0)]


def key(cfg):
    c, o, l, r = cfg
    return f"c{c}_o{o}_l{l}_r{int(r * 10)}"


def hoa_signal(rng, n_frames, n_coef, n=1024, sigma=2000.0):
    x = rng.normal(0.0, sigma, size=(n_frames, n_coef, n)).astype(np.float32)
    x[:, 0] *= np.float32(3.0)   # the omni coefficient dominates in real material
    if n_frames > 1:
        x[1] *= np.float32(60.0)  # drive the clip
    return x
