"""mpeghb200_chain_*: whole BASELINE configurations behind one host-buffer call, compared BYTE FOR BYTE on the final
PCM with the same stage chain run by the oracle (tests/chain_cases.py; every oracle stage is pinned to the unmodified
reference by tests/test_oracle_*.py).  north_star's gate is +-1 LSB at 24 bit; the tolerance asserted here is 0."""
import os

import numpy as np
import pytest

import bin_cases
import chain_cases
import fc_cases
import fd_cases
import hoa_cases
import hoasp_cases
import obj_cases
from oracle import loader

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def check_records(got_pcm, got_bytes, want, tag):
    """got_pcm [S, record_bytes], got_bytes [S]; want: list over streams of byte arrays"""
    for s, w in enumerate(want):
        assert int(got_bytes[s]) == w.size, (tag, s, int(got_bytes[s]), w.size)
        assert np.array_equal(got_pcm[s, :w.size], w), (tag, s)


def hoa_setup(lib, cfg_idx=0, ren_idx=0):
    g, gr = np.load(os.path.join(G, "hoasp_golden.npz")), np.load(os.path.join(G, "hoa_ren_golden.npz"))
    cfg = hoasp_cases.CONFIGS[cfg_idx]
    tables = hoasp_cases.golden_tables(g, cfg)
    rc = hoa_cases.CONFIGS[ren_idx]
    k = hoa_cases.key(rc)
    M, nfc = gr[k + "_M"], (gr[k + "_nfc"] if rc[3] > 0 else None)
    return cfg, tables, M, nfc, lib.hoa_spatial_create(cfg, tables), lib.hoa_renderer_create(M, nfc)


def test_chain_hoa_config3(gpu_lib, oracle):
    """configs[2]: order 4, 16 transport channels -> spatial decode -> NFC + render to 22.2 -> limiter -> 24-bit PCM.
    6 streams x 6 frames, two frames per call, transport signals loud enough to engage the limiter."""
    cfg, tables, M, nfc, hsp, hren = hoa_setup(gpu_lib)
    S, F, T = 6, 6, cfg[1]
    ch = gpu_lib.chain_open(S, n_hoa_transport=T, hoa_spatial=hsp, hoa_renderer=hren, pcm_bits=24, limiter=1,
                            start_delay=240, max_frames_per_call=2)
    assert ch.n_out == 24 and ch.record_bytes == 1024 * 24 * 3 and ch.frames_per_record == 1
    orc = chain_cases.OracleChain(S, hoa=(cfg, tables, M, nfc), pcm_bits=24, start_delay=240)
    rng = np.random.default_rng(2345)
    gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(50 + s)) for s in range(S)]
    for f0 in range(0, F, 2):
        sides = np.array([[gen.next() for gen in gens] for _ in range(2)], hoasp_cases.SIDE_DTYPE)  # [2, S]
        x = rng.normal(0, 4000.0, size=(2, S, T, 1024)).astype(np.float32)
        pcm, nb = ch.execute(2, core=x, hoa_side=sides)
        for k in range(2):
            want = orc.frame(hoa=x[k], hoa_sides=sides[k])
            check_records(pcm[k], nb[k], want, ("hoa", f0 + k))
    ch.close()
    gpu_lib.hoa_renderer_destroy(hren)
    gpu_lib.hoa_spatial_destroy(hsp)


def test_chain_objects_config5(gpu_lib, oracle):
    """configs[4]: 32 objects -> VBAP gains + ramped mix to 7.1.4 -> limiter -> 24-bit PCM; 5 streams x 5 frames."""
    S, F, O = 5, 5, 32
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(19))
    ch = gpu_lib.chain_open(S, n_obj=O, obj_layout=lay, pcm_bits=24, limiter=1, start_delay=240)
    assert ch.n_out == 12
    orc = chain_cases.OracleChain(S, n_obj=O, obj_cicp=19, pcm_bits=24, start_delay=240)
    rng = np.random.default_rng(4567)
    params = np.stack([obj_cases.random_params(rng, F, O) for _ in range(S)], axis=1)  # [F, S, O, 7]
    x = np.stack([obj_cases.object_audio(rng, F, O) for _ in range(S)], axis=1)
    for f in range(F):
        side = gpu_lib.obj_side_from_polar(params[f])
        pcm, nb = ch.execute(1, core=x[f], obj_side=side)
        check_records(pcm[0], nb[0], orc.frame(obj=x[f], obj_params=params[f]), ("obj", f))
    ch.close()
    gpu_lib.obj_layout_destroy(lay)


def test_chain_full_size_properties(gpu_lib, oracle):
    """BASELINE.json's stream counts (configs[2]: 1024 HOA streams, configs[4]: 512 object streams per GPU) through a
    size-independent property: the batch is D distinct streams repeated S / D times in scrambled order, every copy must
    produce the bytes of its original, and the D originals are checked against the oracle chain.  Three frames, state
    carried on the device; covers the launch forms that depend on the stream count (limiter beyond four streams per SM)."""
    D = 4
    rng = np.random.default_rng(97)
    # ---- configs[2] ----
    cfg, tables, M, nfc, hsp, hren = hoa_setup(gpu_lib)
    S, T = 1024, cfg[1]
    perm = rng.permutation(S)
    src = perm % D  # stream s is a copy of distinct stream src[s]
    ch = gpu_lib.chain_open(S, n_hoa_transport=T, hoa_spatial=hsp, hoa_renderer=hren, pcm_bits=24, limiter=1, start_delay=240)
    orc = chain_cases.OracleChain(D, hoa=(cfg, tables, M, nfc), pcm_bits=24, start_delay=240)
    gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(70 + d)) for d in range(D)]
    for f in range(3):
        sides = np.array([gen.next() for gen in gens], hoasp_cases.SIDE_DTYPE)
        x = rng.normal(0, 4000.0, size=(D, T, 1024)).astype(np.float32)
        pcm, nb = ch.execute(1, core=x[src], hoa_side=sides[src])
        want = orc.frame(hoa=x, hoa_sides=sides)
        for d in range(D):
            first = int(np.flatnonzero(src == d)[0])
            check_records(pcm[0][first:first + 1], nb[0][first:first + 1], want[d:d + 1], ("hoa full", f, d))
            same = src == d
            assert np.all(nb[0][same] == nb[0][first]) and np.all(pcm[0][same] == pcm[0][first]), ("hoa copies", f, d)
    ch.close()
    gpu_lib.hoa_renderer_destroy(hren)
    gpu_lib.hoa_spatial_destroy(hsp)
    # ---- configs[4] ----
    S, O = 512, 32
    perm = rng.permutation(S)
    src = perm % D
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(19))
    ch = gpu_lib.chain_open(S, n_obj=O, obj_layout=lay, pcm_bits=24, limiter=1, start_delay=240)
    orc = chain_cases.OracleChain(D, n_obj=O, obj_cicp=19, pcm_bits=24, start_delay=240)
    params = np.stack([obj_cases.random_params(rng, 3, O) for _ in range(D)], axis=1)  # [F, D, O, 7]
    x = np.stack([obj_cases.object_audio(rng, 3, O) for _ in range(D)], axis=1)
    for f in range(3):
        side = gpu_lib.obj_side_from_polar(params[f][src])
        pcm, nb = ch.execute(1, core=x[f][src], obj_side=side)
        want = orc.frame(obj=x[f], obj_params=params[f])
        for d in range(D):
            first = int(np.flatnonzero(src == d)[0])
            check_records(pcm[0][first:first + 1], nb[0][first:first + 1], want[d:d + 1], ("obj full", f, d))
            same = src == d
            assert np.all(nb[0][same] == nb[0][first]) and np.all(pcm[0][same] == pcm[0][first]), ("obj copies", f, d)
    ch.close()
    gpu_lib.obj_layout_destroy(lay)


def test_chain_bed_objects_format_conv_16bit(gpu_lib, oracle):
    """22.2 bed + 8 objects rendered into 22.2, summed, format-converted to 7.1.4, limited, 16-bit PCM with a
    channel map; spectra in front (FD synthesis of bed and object signals on the device)."""
    S, F, O = 3, 4, 8
    D = np.load(os.path.join(G, "fc_golden.npz"))["c13_c19_D"]
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(13))
    fc = gpu_lib.format_conv_create(D)
    cmap = list(range(11, -1, -1))
    ch = gpu_lib.chain_open(S, core_is_spectra=1, n_bed_ch=24, n_obj=O, obj_layout=lay, format_conv=fc, pcm_bits=16,
                            limiter=1, out_ch_map=cmap)
    assert ch.n_out == 12
    orc = chain_cases.OracleChain(S, n_bed=24, n_obj=O, obj_cicp=13, fc_D=D, pcm_bits=16, spectra=True, ch_map=cmap)
    rng = np.random.default_rng(99)
    params = np.stack([obj_cases.random_params(rng, F, O) for _ in range(S)], axis=1)
    sb = fd_cases.random_side(rng, F, S * 24).reshape(F, S, 24, 5)
    so = fd_cases.random_side(rng, F, S * O).reshape(F, S, O, 5)
    for f in range(F):
        bed = fd_cases.random_coef(rng, (S, 24, 1024), 1500.0)
        obj = fd_cases.random_coef(rng, (S, O, 1024), 2500.0)
        sides = {"bed": sb[f], "obj": so[f]}
        pcm, nb = ch.execute(1, core=chain_cases.core_buffer(bed, obj), fd_side=chain_cases.side_buffer(sides),
                             obj_side=gpu_lib.obj_side_from_polar(params[f]))
        want = orc.frame(bed=bed, obj=obj, fd_side5=sides, obj_params=params[f])
        check_records(pcm[0], nb[0], want, ("bed+obj+fc", f))
    ch.close()
    gpu_lib.format_conv_destroy(fc)
    gpu_lib.obj_layout_destroy(lay)


@pytest.mark.parametrize("length,nb,spectra", [(1024, 1, False), (4096, 0, True)])
def test_chain_binaural(gpu_lib, oracle, length, nb, spectra):
    """configs[3]: 22.2 bed -> binaural renderer -> 24-bit PCM.  4096 taps: a block every 4 frames (one PCM record
    of 4096 samples per block); per-stream BRIR sets."""
    S, n_in = 3, 24
    rng = np.random.default_rng(3456 + length)
    filt = [bin_cases.brir_set(rng, n_in, length, nb) for _ in range(2)]
    sets = [0, 1, 0]

    def st(k):
        return np.stack([f[k] for f in filt])
    hb = gpu_lib.binaural_create(st("taps_direct"), st("taps_diffuse"), st("direct_fc"), st("diffuse_fc"), st("weight"))
    ch = gpu_lib.chain_open(S, core_is_spectra=int(spectra), n_bed_ch=n_in, binaural=hb, pcm_bits=24, limiter=1,
                            binaural_set_of_stream=sets, start_delay=100)
    fpr = ch.frames_per_record
    assert ch.n_out == 2 and fpr == (4 if length == 4096 else 1) and ch.record_bytes == fpr * 1024 * 2 * 3
    orc = chain_cases.OracleChain(S, n_bed=n_in, binaural=[filt[i] for i in sets], pcm_bits=24, spectra=spectra,
                                  start_delay=100)
    n_blocks = 3
    for b in range(n_blocks):
        sd = fd_cases.random_side(rng, fpr, S * n_in).reshape(fpr, S, n_in, 5)
        x = (fd_cases.random_coef(rng, (fpr, S, n_in, 1024), 800.0) if spectra
             else rng.normal(0, 3000.0, size=(fpr, S, n_in, 1024)).astype(np.float32))
        kw = {"fd_side": np.stack([chain_cases.side_buffer({"bed": sd[k]}) for k in range(fpr)])} if spectra else {}
        pcm, nbytes = ch.execute(fpr, core=x, **kw)
        want = None
        for k in range(fpr):
            want = orc.frame(bed=x[k], fd_side5={"bed": sd[k]})
            if k < fpr - 1:
                assert all(w.size == 0 for w in want)
        check_records(pcm[0], nbytes[0], want, ("binaural", length, b))
    ch.close()
    gpu_lib.binaural_destroy(hb)


def test_chain_fd_front_config2(gpu_lib, oracle):
    """configs[1] as a chain: spectra of a 22.2 bed -> FD synthesis -> loudness gain + limiter -> 24-bit PCM; three
    frames per call, two calls in flight (submit / submit / wait / wait) from page-locked buffers."""
    S, C, FPC, CALLS = 4, 24, 3, 4
    gain = float(oracle.oracle_loudness_gain(-3.0))
    ch = gpu_lib.chain_open(S, core_is_spectra=1, n_bed_ch=C, pcm_bits=24, limiter=1, start_delay=240,
                            loudness_gain=gain, max_frames_per_call=FPC)
    orc = chain_cases.OracleChain(S, n_bed=C, pcm_bits=24, start_delay=240, spectra=True, loudness_gain=gain)
    rng = np.random.default_rng(1234)
    side = fd_cases.random_side(rng, FPC * CALLS, S * C, kernel_switch_prob=0.1).reshape(CALLS, FPC, S, C, 5)
    coef = fd_cases.random_coef(rng, (CALLS, FPC, S, C, 1024), 6000.0)
    h_core = [gpu_lib.host_array((FPC, S, C, 1024)) for _ in range(2)]
    h_side = [gpu_lib.host_array((FPC, S, C), np.uint32) for _ in range(2)]
    h_pcm = [gpu_lib.host_array((FPC, S, ch.record_bytes), np.uint8) for _ in range(2)]
    h_nb = [gpu_lib.host_array((FPC, S), np.int32) for _ in range(2)]

    def submit(c):
        k = c % 2
        h_core[k][:] = coef[c]
        h_side[k][:] = fd_cases.pack_side(side[c])
        ch.submit(ch.frame(core=h_core[k], fd_side=h_side[k], pcm=h_pcm[k], pcm_bytes=h_nb[k]), FPC)

    def check(c):
        ch.wait()
        k = c % 2
        for f in range(FPC):
            want = orc.frame(bed=coef[c, f], fd_side5={"bed": side[c, f]})
            check_records(h_pcm[k][f], h_nb[k][f], want, ("fd", c, f))

    submit(0)
    for c in range(1, CALLS):
        submit(c)
        check(c - 1)
    check(CALLS - 1)
    ch.close()
    for a in h_core + h_side + h_pcm + h_nb:
        gpu_lib.host_free(a)


def test_chain_objects_plus_hoa(gpu_lib, oracle):
    """Objects + HOA scene (no bed): both rendered to 22.2 and summed (hoa + obj), limiter off as a pure delay,
    32-bit PCM."""
    cfg, tables, M, nfc, hsp, hren = hoa_setup(gpu_lib, 1, 1)  # T = 8, no NFC
    S, F, O, T = 2, 3, 5, cfg[1]
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(13))
    ch = gpu_lib.chain_open(S, n_obj=O, n_hoa_transport=T, obj_layout=lay, hoa_spatial=hsp, hoa_renderer=hren,
                            pcm_bits=32, limiter=2)
    orc = chain_cases.OracleChain(S, n_obj=O, obj_cicp=13, hoa=(cfg, tables, M, nfc), pcm_bits=32, limiter=2)
    rng = np.random.default_rng(5)
    gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(7 + s)) for s in range(S)]
    params = np.stack([obj_cases.random_params(rng, F, O) for _ in range(S)], axis=1)
    for f in range(F):
        sides = np.array([gen.next() for gen in gens], hoasp_cases.SIDE_DTYPE)
        xo = obj_cases.object_audio(rng, 1, S * O)[0].reshape(S, O, 1024)
        xh = rng.normal(0, 1500.0, size=(S, T, 1024)).astype(np.float32)
        pcm, nb = ch.execute(1, core=chain_cases.core_buffer(None, xo, xh), hoa_side=sides,
                             obj_side=gpu_lib.obj_side_from_polar(params[f]))
        want = orc.frame(obj=xo, hoa=xh, obj_params=params[f], hoa_sides=sides)
        check_records(pcm[0], nb[0], want, ("obj+hoa", f))
    ch.close()
    gpu_lib.obj_layout_destroy(lay)
    gpu_lib.hoa_renderer_destroy(hren)
    gpu_lib.hoa_spatial_destroy(hsp)


def test_chain_device_entry_matches_host_entry(gpu_lib):
    """mpeghb200_chain_run_dev (what bench.py times with resident inputs) and the host-buffer call produce the same
    bytes: two chains with identical state, one fed through each entry."""
    import torch
    S, O = 7, 32
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(19))
    a = gpu_lib.chain_open(S, n_obj=O, obj_layout=lay, pcm_bits=24)
    b = gpu_lib.chain_open(S, n_obj=O, obj_layout=lay, pcm_bits=24)
    rng = np.random.default_rng(8)
    for f in range(3):
        params = obj_cases.random_params(rng, 1, S * O)[0].reshape(S, O, 7)
        x = obj_cases.object_audio(rng, 1, S * O)[0].reshape(S, O, 1024)
        side = gpu_lib.obj_side_from_polar(params)
        pcm, nb = a.execute(1, core=x, obj_side=side)
        d_x = torch.from_numpy(x).cuda()
        d_side = torch.from_numpy(side.view(np.uint8).reshape(-1).copy()).cuda()
        d_pcm = torch.zeros(S, b.record_bytes, dtype=torch.uint8, device="cuda")
        d_nb = torch.zeros(S, dtype=torch.int32, device="cuda")
        b.run_dev(b.frame(core=d_x.data_ptr(), obj_side=d_side.data_ptr(), pcm=d_pcm.data_ptr(),
                          pcm_bytes=d_nb.data_ptr()), 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_pcm.cpu().numpy(), pcm[0]) and np.array_equal(d_nb.cpu().numpy(), nb[0]), f
    a.close()
    b.close()
    gpu_lib.obj_layout_destroy(lay)


def test_chain_argument_errors(gpu_lib):
    from libmpegh_b200.lib import MpeghError
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(19))
    with pytest.raises(MpeghError):
        gpu_lib.chain_open(4, n_obj=8)                       # objects without a layout
    with pytest.raises(MpeghError):
        gpu_lib.chain_open(4, n_bed_ch=24, n_obj=8, obj_layout=lay)  # 22.2 bed vs 7.1.4 object layout
    with pytest.raises(MpeghError):
        gpu_lib.chain_open(0, n_bed_ch=2)
    ch = gpu_lib.chain_open(2, n_bed_ch=2, pcm_bits=16)
    x = np.zeros((2, 2, 1024), np.float32)
    with pytest.raises(MpeghError):
        ch.execute(2, core=np.zeros((2, 2, 2, 1024), np.float32))   # more than max_frames_per_call
    pcm, nb = ch.execute(1, core=x)
    assert nb.tolist() == [[1024 * 2 * 2] * 2] and not pcm.any()
    ch.close()
    gpu_lib.obj_layout_destroy(lay)
