"""Long transforms and binaural renderer restatement (oracle/oracle_fft.c, oracle_binaural.c) against the
unmodified reference (oracle/_ref) on seeded inputs."""
import ctypes

import numpy as np
import pytest

import bin_cases
from oracle import loader
from oracle.loader import bits_equal


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024])
def test_fft_pow2_vs_reference(ref, oracle, n):
    rng = np.random.default_rng(n)
    for _ in range(3):
        re, im = rng.normal(0, 100, n).astype(np.float32), rng.normal(0, 100, n).astype(np.float32)
        a, b = re.copy(), im.copy()
        ref.ref_rad2_cplx_fft(a, b, n)
        c, d = re.copy(), im.copy()
        oracle.oracle_cplx_fft_pow2(c, d, n)
        assert bits_equal(a, c) and bits_equal(b, d)


@pytest.mark.parametrize("n", [2048, 8192, 16384])
def test_ifft_big_vs_reference(ref, oracle, n):
    rng = np.random.default_rng(n + 1)
    for real_only in (False, True):
        re = rng.normal(0, 1, n).astype(np.float32)
        im = np.zeros(n, np.float32) if real_only else rng.normal(0, 1, n).astype(np.float32)
        if real_only:
            re[n // 2:] = 0
        a, b = re.copy(), im.copy()
        ref.ref_cplx_ifft_8k(a, b, n)
        c, d = re.copy(), im.copy()
        assert oracle.oracle_cplx_ifft_big(c, d, n) == 0
        assert bits_equal(a, c) and bits_equal(b, d)


@pytest.mark.parametrize("n", [2048, 4096, 8192, 16384])
def test_fft_big_vs_reference(ref, oracle, n):
    rng = np.random.default_rng(n + 2)
    re, im = rng.normal(0, 1, n).astype(np.float32), rng.normal(0, 1, n).astype(np.float32)
    a, b = re.copy(), im.copy()
    ref.ref_cplx_fft_16k(a, b, n)
    c, d = re.copy(), im.copy()
    assert oracle.oracle_cplx_fft_big(c, d, n) == 0
    assert bits_equal(a, c) and bits_equal(b, d)


def test_unsupported_sizes(oracle):
    z = np.zeros(16384, np.float32)
    assert oracle.oracle_cplx_ifft_big(z, z.copy(), 4096) == -1


def test_ifft_4096_depends_on_stale_scratch(ref):
    """Why N = 4096 is refused everywhere (mpeghb200_binaural_create, the oracle): impeghd_cplx_ifft_4
    (decoder/impeghd_ifft.c:1212) reads x_i[4] and x_i[6] of 4-element rows; for the last row (i = 1023,
    ptr_data_i = &interim[(2 * 1023 + 1) * 4], :1336-1340) that is interim[8192] and interim[8194], two floats behind
    the 8192-float interim buffer, in scratch memory this transform never writes.  The same input therefore gives
    different outputs depending on what an earlier user left in the scratch; the other sizes do not care."""
    rng = np.random.default_rng(4096)
    for n in (2048, 4096, 8192, 16384):
        re, im = rng.normal(0, 1, n).astype(np.float32), rng.normal(0, 1, n).astype(np.float32)
        a, b = re.copy(), im.copy()
        ref.ref_cplx_ifft_8k_fill(a, b, n, 0.0)
        c, d = re.copy(), im.copy()
        ref.ref_cplx_ifft_8k_fill(c, d, n, 1000.0)
        same = bits_equal(a, c) and bits_equal(b, d)
        assert same == (n != 4096), n
        if n == 4096:  # the damage is confined to the bins the last row feeds: k = 1023 + 1024 j
            bad = np.nonzero((a.view(np.uint32) != c.view(np.uint32)) | (b.view(np.uint32) != d.view(np.uint32)))[0]
            assert set(bad) <= {1023, 2047, 3071, 4095} and len(bad) >= 1


def test_ifft_16384_is_not_a_transform(ref):
    """N = 16384 (direct filters of 4097 .. 8192 taps): impeghd_cplx_ifft_8k runs the sixteen 1024-point row transforms
    and the inter-stage twiddles but no 16-point column transform (decoder/impeghd_ifft.c:1318-1359 tests 8192, 4096 and
    2048 only).  Deterministic, restated by the oracle bit for bit (test_ifft_big_vs_reference), but not an e^{+j} DFT:
    a unit impulse at n = 1 does not come back as a complex exponential."""
    n = 16384
    re, im = np.zeros(n, np.float32), np.zeros(n, np.float32)
    re[1] = 1.0
    ref.ref_cplx_ifft_8k(re, im, n)
    k = np.arange(n)
    want = np.exp(2j * np.pi * k / n)
    assert np.abs((re + 1j * im) - want).max() > 0.5


def _ref_bin(ref, f, via_init):
    err = ctypes.c_int()
    td, tdf = f["taps_direct"], f["taps_diffuse"]
    nb = tdf.shape[0]
    h = ref.ref_bin_create(td.shape[1], td.shape[2], nb, td.reshape(-1), np.ascontiguousarray(tdf).reshape(-1),
                           f["direct_fc"].reshape(-1).copy(), np.ascontiguousarray(f["diffuse_fc"]).reshape(-1),
                           f["weight"], via_init, ctypes.byref(err))
    assert err.value == 0
    return h


def _ref_spectra(ref, h, n_in, nb):
    info = np.zeros(8, np.int32)
    ref.ref_bin_info(h, info)
    fft = int(info[0])
    d, df = np.zeros((2, n_in, fft), np.float32), np.zeros((max(nb, 1), 2, fft), np.float32)
    dl, dfl = np.zeros((2, n_in), np.int32), np.zeros((max(nb, 1), 2), np.int32)
    ref.ref_bin_spectra(h, d.reshape(-1), df.reshape(-1), dl.reshape(-1), dfl.reshape(-1))
    return info, d, df, dl, dfl


@pytest.mark.parametrize("length,nb", [(1024, 1), (4096, 1)])
def test_hand_fill_equals_reference_init(ref, length, nb):
    """The harness' fill of the persistent struct (needed for > 21 inputs) equals impeghd_binaural_renderer_init."""
    rng = np.random.default_rng(length + nb)
    f = bin_cases.with_cutoffs(rng, bin_cases.brir_set(rng, 21, length, nb))
    a, b = _ref_bin(ref, f, 1), _ref_bin(ref, f, 0)
    ia, ib = _ref_spectra(ref, a, 21, nb), _ref_spectra(ref, b, 21, nb)
    assert np.array_equal(ia[0], ib[0])
    for x, y in zip(ia[1:], ib[1:]):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    ref.ref_bin_destroy(a)
    ref.ref_bin_destroy(b)


@pytest.mark.parametrize("n_in,length,nb,cut", [(24, 4096, 1, False), (5, 1024, 1, True), (21, 4096, 0, False),
                                                (2, 1024, 1, True), (11, 3000, 1, False), (3, 5000, 1, True), (2, 8192, 0, False), (4, 300, 1, True), (2, 64, 0, False)])
def test_binaural_vs_reference(ref, oracle, n_in, length, nb, cut):
    rng = np.random.default_rng(n_in * 7 + length + nb)
    f = bin_cases.brir_set(rng, n_in, length, nb)
    if cut:
        f = bin_cases.with_cutoffs(rng, f)
    h = _ref_bin(ref, f, 0)
    info, d, df, dl, dfl = _ref_spectra(ref, h, n_in, nb)
    o = loader.OracleBinaural(f["taps_direct"], f["taps_diffuse"], f["direct_fc"], f["diffuse_fc"], f["weight"])
    assert o.fft == info[0] and o.B == info[2]
    assert bits_equal(o.h_direct, d) and np.array_equal(o.direct_len, dl)
    if nb:
        assert bits_equal(o.h_diffuse, df) and np.array_equal(o.diffuse_len, dfl)
    T = 5
    x = bin_cases.signal(rng, T, n_in, o.B)
    # feed the reference 1024 samples per call like the decoder does (ia_core_coder_decode_main.c:2837)
    stream = np.concatenate(list(x), axis=1)  # [n_in, T*B]
    got_ref = []
    buf = np.zeros((2, 16384), np.float32)
    for pos in range(0, stream.shape[1], 1024):
        feed = min(1024, stream.shape[1] - pos)
        n = ref.ref_bin_process(h, np.ascontiguousarray(stream[:, pos:pos + feed]).reshape(-1), feed,
                                buf.reshape(-1), 16384)
        assert n % o.B == 0  # short filters: several blocks of B < 1024 samples complete in one call
        for k in range(n // o.B):
            got_ref.append(buf[:, k * o.B:(k + 1) * o.B].copy())
    assert len(got_ref) == T
    for t in range(T):
        assert bits_equal(o.block(x[t]), got_ref[t]), (t,)
    ref.ref_bin_destroy(h)


def test_binaural_golden(oracle):
    """Committed vectors from the reference build (tests/golden/make_binaural_golden.py); no reference needed."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "binaural_golden.npz"))
    for tag in ("a", "b"):
        o = loader.OracleBinaural(g[f"{tag}_taps_direct"], g[f"{tag}_taps_diffuse"], g[f"{tag}_direct_fc"],
                                  g[f"{tag}_diffuse_fc"], g[f"{tag}_weight"])
        for t in range(g[f"{tag}_in"].shape[0]):
            assert bits_equal(o.block(g[f"{tag}_in"][t]), g[f"{tag}_out"][t]), (tag, t)
