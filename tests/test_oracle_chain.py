"""Whole-configuration chains: the oracle composition the GPU chains are checked against (tests/chain_cases.py) vs
the UNMODIFIED reference's stage functions chained in C (oracle/ref_chain.c), final PCM bytes compared.  Needs
oracle/_ref (skips without it)."""
import os

import numpy as np
import pytest

import bin_cases
import chain_cases
import fd_cases
import hoa_cases
import hoasp_cases
import obj_cases
from oracle import loader

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ns():
    lib = loader.load_ref_ns()
    if lib is None:
        pytest.skip("oracle/_ref/libmpeghref_ns.so not built (needs /root/reference)")
    return lib


def same(pcm, nb, want, tag):
    for s, w in enumerate(want):
        assert int(nb[s]) == w.size and np.array_equal(pcm[s, :w.size], w), (tag, s)


def test_hoa_chain_vs_reference(ns, oracle):
    g, gr = np.load(os.path.join(G, "hoasp_golden.npz")), np.load(os.path.join(G, "hoa_ren_golden.npz"))
    cfg, rc = hoasp_cases.CONFIGS[0], hoa_cases.CONFIGS[0]
    k = hoa_cases.key(rc)
    S, F = 3, 4
    orc = chain_cases.OracleChain(S, hoa=(cfg, hoasp_cases.golden_tables(g, cfg), gr[k + "_M"], gr[k + "_nfc"]),
                                  start_delay=240)
    rch = loader.RefChain("hoa", S, n_threads=2, cfg=cfg, ren=rc, side_bytes=hoasp_cases.SIDE_DTYPE.itemsize,
                          start_delay=240)
    assert rch.n_out == 24
    rng = np.random.default_rng(1)
    gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(s)) for s in range(S)]
    sides = np.array([[gen.next() for gen in gens] for _ in range(F)], hoasp_cases.SIDE_DTYPE)
    x = rng.normal(0, 4000.0, size=(F, S, cfg[1], 1024)).astype(np.float32)
    _, pcm, nb = rch.run(x, sides)
    for f in range(F):
        same(pcm[f], nb[f], orc.frame(hoa=x[f], hoa_sides=sides[f]), ("hoa", f))
    rch.close()


def test_object_chain_vs_reference(ns, oracle):
    S, F, O = 3, 4, 32
    orc = chain_cases.OracleChain(S, n_obj=O, obj_cicp=19, start_delay=240)
    rch = loader.RefChain("obj", S, n_threads=3, n_obj=O, cicp=19, n_spk=12, start_delay=240)
    rng = np.random.default_rng(2)
    params = np.stack([obj_cases.random_params(rng, F, O) for _ in range(S)], axis=1)
    x = np.stack([obj_cases.object_audio(rng, F, O) for _ in range(S)], axis=1)
    # the reference marks only its first 24 objects as "first frame" (MAX_NUM_OAM_OBJS); the oracle state does the same
    _, pcm, nb = rch.run(x, params)
    for f in range(F):
        same(pcm[f], nb[f], orc.frame(obj=x[f], obj_params=params[f]), ("obj", f))
    rch.close()


@pytest.mark.parametrize("length,nblk", [(1024, 1), (4096, 0)])
def test_binaural_chain_vs_reference(ns, oracle, length, nblk):
    S, n_in = 2, 24
    rng = np.random.default_rng(3 + length)
    filt = [bin_cases.brir_set(rng, n_in, length, nblk) for _ in range(2)]
    sets = [1, 0]
    orc = chain_cases.OracleChain(S, n_bed=n_in, binaural=[filt[i] for i in sets], start_delay=100)
    rch = loader.RefChain("bin", S, n_threads=2, filters=filt, sets=sets, start_delay=100)
    fpr = rch.fpr
    assert fpr == max(1, length // 1024)
    F = 3 * fpr
    x = rng.normal(0, 3000.0, size=(F, S, n_in, 1024)).astype(np.float32)
    _, pcm, nb = rch.run(x)
    for f in range(F):
        want = orc.frame(bed=x[f])
        if f % fpr == fpr - 1:
            same(pcm[f // fpr], nb[f // fpr], want, ("bin", f))
        else:
            assert all(w.size == 0 for w in want)
    rch.close()


def test_fd_chain_vs_reference(ns, oracle):
    S, C, F = 2, 24, 4
    orc = chain_cases.OracleChain(S, n_bed=C, spectra=True, start_delay=240)
    rch = loader.RefChain("fd", S, n_threads=2, n_ch=C, start_delay=240)
    rng = np.random.default_rng(4)
    side = fd_cases.random_side(rng, F, S * C, kernel_switch_prob=0.1).reshape(F, S, C, 5)
    coef = fd_cases.random_coef(rng, (F, S, C, 1024), 6000.0)
    _, pcm, nb = rch.run(coef, side)
    for f in range(F):
        same(pcm[f], nb[f], orc.frame(bed=coef[f], fd_side5={"bed": side[f]}), ("fd", f))
    rch.close()
