"""Shared binaural-renderer test-case generation (synthetic BRIRs: exponentially decaying noise)."""
import numpy as np


def brir_set(rng, n_in, length, n_blocks, decay=6.0):
    """-> dict(taps_direct [2,n_in,len], taps_diffuse [n_blocks,2,len], direct_fc [2,n_in], diffuse_fc [2,n_blocks],
    weight [n_in])"""
    env = np.exp(-decay * np.arange(length) / length).astype(np.float32)
    td = (rng.normal(0, 0.2, size=(2, n_in, length)).astype(np.float32) * env)
    tdf = (rng.normal(0, 0.02, size=(n_blocks, 2, length)).astype(np.float32) * env)
    dfc = np.ones((2, n_in), np.float32)
    ffc = np.ones((2, max(n_blocks, 1)), np.float32)[:, :n_blocks]
    w = rng.uniform(0.5, 1.5, size=n_in).astype(np.float32)
    return dict(taps_direct=td, taps_diffuse=tdf, direct_fc=dfc, diffuse_fc=ffc, weight=w)


def with_cutoffs(rng, f):
    """Variant with per-filter cut-off factors below and above 1 (the latter are clamped by the init)."""
    g = dict(f)
    g["direct_fc"] = rng.uniform(0.3, 1.2, size=f["direct_fc"].shape).astype(np.float32)
    g["diffuse_fc"] = rng.uniform(0.3, 1.2, size=f["diffuse_fc"].shape).astype(np.float32)
    return g


def signal(rng, n_blocks_time, n_in, B, sigma=0.1):
    return rng.normal(0, sigma, size=(n_blocks_time, n_in, B)).astype(np.float32)
