"""Shared object-renderer test-case generation."""
import numpy as np

CICPS = [2, 6, 13, 14, 19]  # stereo, 5.1, 22.2, 5.1.2, 7.1.4


def random_params(rng, n_frames, n_obj, spread_prob=0.3, moving=True):
    """[n_frames, n_obj, 7] = az, el, radius, gain, spread width/height/depth in the OAM decoder's descaled units."""
    p = np.zeros((n_frames, n_obj, 7), np.float32)
    az = rng.uniform(-180, 180, n_obj)
    el = rng.uniform(-45, 90, n_obj)
    for f in range(n_frames):
        if moving and f:
            az = (az + rng.normal(0, 8, n_obj) + 180) % 360 - 180
            el = np.clip(el + rng.normal(0, 4, n_obj), -90, 90)
        p[f, :, 0] = np.round(az * 2) / 2      # 1.5 / 3 degree grids descale to multiples like these
        p[f, :, 1] = np.round(el)
        p[f, :, 2] = 1.0
        p[f, :, 3] = rng.choice([1.0, 0.7079458, 0.5, 1.4125376], n_obj)
        sp = rng.random(n_obj) < spread_prob
        p[f, sp, 4] = rng.uniform(0.5, 170, sp.sum()).astype(np.float32)
        p[f, sp, 5] = rng.uniform(0.0, 80, sp.sum()).astype(np.float32)
        p[f, sp, 6] = rng.uniform(0.0, 10, sp.sum()).astype(np.float32)
    # exact poles / seams
    p[0, 0, :2] = (0.0, 90.0)
    if n_obj > 1:
        p[0, 1, :2] = (-180.0, 0.0)
    if n_obj > 2:
        p[0, 2, :2] = (30.0, 0.0)   # exactly on a loudspeaker
    return p


def object_audio(rng, n_frames, n_obj, sigma=3000.0):
    return rng.normal(0.0, sigma, size=(n_frames, n_obj, 1024)).astype(np.float32)
