#!/usr/bin/env python3
"""Per-stage throughput of the kernels beyond the headline FD synthesis, at the BASELINE.json config shapes
(one GPU).  Prints one JSON line per stage: ms per step, algorithmic GB/s (SURVEY.md 8d byte counts) and the
fraction of the measured HBM peak.  Diagnostic companion of bench.py; numbers go to profiles/."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libmpegh_b200.lib import Lib  # noqa: E402
from oracle import loader  # noqa: E402  (layout / matrix fixtures only)

lib = Lib().create(0)
st = torch.cuda.current_stream().cuda_stream
peak = 6549.4
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


flush_sink = torch.zeros((), dtype=torch.int64, device="cuda")


def timeit(fn, steps=30, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(steps):
        # evict L2 between timed iterations by READING 256 MB: the L2 is left full of clean lines.  (Writing the
        # flush buffer instead leaves 126 MB of dirty lines whose write-back is then charged to the timed kernel,
        # ~20 us at HBM speed -- that inflated every sub-100-us stage of the first round-1 tables.)
        flush_sink.copy_(flush.view(torch.int64).sum())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def report(name, ms, bytes_per_step, units, unit_name, extra=None):
    gbs = bytes_per_step / (ms * 1e-3) / 1e9
    line = {"stage": name, "ms_per_step": ms, "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak,
            unit_name + "_per_s": units / (ms * 1e-3)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""


def binaural_stage():
    """config 4: 22.2 -> 2 ears, 4096-tap BRIRs, 1024 streams; per-stream filter sets and one shared set."""
    import bin_cases
    S, n_in, L = 1024, 24, 4096
    rng = np.random.default_rng(3456)
    for label, n_sets, nb in (("per-stream filters", S, 0), ("per-stream filters, 1 diffuse block", S, 1),
                              ("shared filters", 1, 0)):
        env = np.exp(-6.0 * np.arange(L) / L).astype(np.float32)
        td = rng.standard_normal((n_sets, 2, n_in, L), dtype=np.float32) * np.float32(0.2) * env
        tf = rng.standard_normal((n_sets, nb, 2, L), dtype=np.float32) * np.float32(0.02) * env
        h = lib.binaural_create(td, tf, np.ones((n_sets, 2, n_in), np.float32), np.ones((n_sets, 2, nb), np.float32),
                                np.ones((n_sets, n_in), np.float32))
        del td, tf
        B = lib.binaural_block_len(h)
        x = torch.randn(S, n_in, B, device="cuda") * 0.1
        y = torch.empty(S, 2, B, device="cuda")
        state = torch.zeros(S * lib.binaural_state_floats(h), device="cuda")
        sets = torch.arange(S, dtype=torch.int32, device="cuda") % n_sets
        ms = timeit(lambda i: lib.binaural_block_dev(h, x.data_ptr(), y.data_ptr(), state.data_ptr(), S, 0, 1.0, 1.0,
                                                     sets.data_ptr(), st), steps=10, warmup=2)
        N = 2 * B
        per_unit = n_in * B * 4 + 2 * B * 4 + 2 * 2 * B * 4 + (2 * n_in * N * 4 if n_sets > 1 else 0) + \
            nb * (2 * N * 4 * (n_sets > 1) + 2 * N * 4)
        report(f"binaural 22.2 -> 2, 4096 taps, 1024 streams, {label}", ms, S * per_unit, S * n_in * B,
               "input_channel_samples", {"stream_blocks_per_s": S / (ms * 1e-3),
                                         "audio_s_per_wall_s": S * B / 48000.0 / (ms * 1e-3),
                                         "algorithmic_bytes_per_unit": per_unit})
        lib.binaural_destroy(h)
        del x, y, state


def hoa_chain_stage():
    """config 3: 1024 streams, order 4, 16 (and 8) transport channels -> spatial decode -> NFC + render to 22.2."""
    import hoasp_cases
    import hoa_cases as hc
    g = np.load(os.path.join(ROOT, "tests", "golden", "hoasp_golden.npz"))
    gr = np.load(os.path.join(ROOT, "tests", "golden", "hoa_ren_golden.npz"))
    kr = hc.key(hc.CONFIGS[0])
    S = 1024
    for cfg in (hoasp_cases.CONFIGS[0], hoasp_cases.CONFIGS[1]):
        T = cfg[1]
        tables = hoasp_cases.golden_tables(g, cfg)
        h = lib.hoa_spatial_create(cfg, tables)
        state = torch.empty(S * lib.hoa_spatial_state_bytes(h), dtype=torch.uint8, device="cuda")
        lib.hoa_spatial_state_init_dev(h, state.data_ptr(), S, st)
        gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(s), pred_prob=0.0) for s in range(64)]
        sides = []
        for f in range(4):
            a = np.array([gen.next() for gen in gens], hoasp_cases.SIDE_DTYPE)
            a = np.tile(a, S // 64)
            sides.append(torch.from_numpy(a.view(np.uint8).reshape(S, -1).copy()).cuda())
        x = torch.randn(S, T, 1024, device="cuda") * 2000
        coef = torch.empty(S, 25, 1024, device="cuda")
        ms = timeit(lambda i: lib.hoa_spatial_dev(h, sides[i % 4].data_ptr(), x.data_ptr(), coef.data_ptr(),
                                                  state.data_ptr(), S, st))
        report(f"hoa_spatial order 4, T={T}, 1024 streams (plan + render kernels)", ms, S * (T + 25) * 4096,
               S * 25 * 1024, "coefficient_samples")
        r = lib.hoa_renderer_create(gr[kr + "_M"], gr[kr + "_nfc"])
        hout = torch.empty(S, 24, 1024, device="cuda")
        hst = torch.zeros(S * 25 * 14, device="cuda")

        def chain(i):
            lib.hoa_spatial_dev(h, sides[i % 4].data_ptr(), x.data_ptr(), coef.data_ptr(), state.data_ptr(), S, st)
            lib.hoa_render_dev(r, coef.data_ptr(), hout.data_ptr(), hst.data_ptr(), S, st)
        ms = timeit(chain)
        report(f"config 3 chain: hoa_spatial + NFC + render to 22.2, T={T}, 1024 streams", ms, S * (T + 24) * 4096,
               S * 24 * 1024, "output_channel_samples",
               {"stream_frames_per_s": S / (ms * 1e-3), "audio_s_per_wall_s": S * 1024 / 48000.0 / (ms * 1e-3)})
        lib.hoa_renderer_destroy(r)
        lib.hoa_spatial_destroy(h)


def fc_stage():
    """config 5 bed: 22.2 -> 7.1.4 format conversion, 512 streams."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "fc_golden.npz"))
    D = g["c13_c19_D"]
    S = 512
    h = lib.format_conv_create(D)
    x = torch.randn(S, 24, 1024, device="cuda") * 3000
    y = torch.empty(S, 12, 1024, device="cuda")
    state = torch.zeros(S * lib.format_conv_state_floats(h), device="cuda")
    ms = timeit(lambda i: lib.format_conv_dev(h, x.data_ptr(), y.data_ptr(), state.data_ptr(), S, st))
    per_unit = 24 * 4096 + 12 * 4096 + 2 * 24 * 256 * 4 + 2 * 12 * 256 * 4 + 2 * 2 * 12 * 58 * 4
    report("format_conv 22.2 -> 7.1.4, 512 streams", ms, S * per_unit, S * 24 * 1024, "input_channel_samples",
           {"stream_frames_per_s": S / (ms * 1e-3)})
    lib.format_conv_destroy(h)


def mct_stage():
    """MCT boxes ahead of the seam: 256 streams x 24 channels, 12 boxes per stream-frame."""
    import mct_cases
    from libmpegh_b200.lib import MCT_BOX_DTYPE
    S = 256
    rng = np.random.default_rng(2)
    recs, first = [], [0]
    for s_ in range(S):
        boxes = []
        while len(boxes) < 12:
            boxes += mct_cases.random_group(rng)
        recs += mct_cases.resolve_group(lib, boxes[:12], s_ * 24)
        first.append(len(recs))
    spec = torch.randn(S * 24, 1024, device="cuda") * 500
    d_rec = torch.from_numpy(np.array(recs, MCT_BOX_DTYPE).view(np.uint8).reshape(-1)).cuda()
    d_first = torch.from_numpy(np.array(first, np.int32)).cuda()
    ms = timeit(lambda i: lib.mct_dev(spec.data_ptr(), d_rec.data_ptr(), d_first.data_ptr(), S, st))
    touched = sum(int(r["len"][:int(r["n_seg"])].sum()) for r in recs)
    report("mct boxes, 256 streams x 24 ch, 12 boxes per stream-frame", ms, len(recs) * (4 * 4096 + 400),
           touched, "bins_rotated_or_predicted", {"boxes": len(recs)})


def drc_stage():
    """DRC in the STFT domain (domain switcher t2f -> sub-band gains -> f2t), 6144 channel-frames, one 3-band set."""
    from libmpegh_b200.lib import DRC_SB_CHANNEL_DTYPE
    U = 6144
    rng = np.random.default_rng(4)
    rec = np.zeros(U, DRC_SB_CHANNEL_DTYPE)
    rec["first_seq"] = 3 * (np.arange(U) // 24)
    rec["band_count"] = 3
    rec["weight_row"] = 0
    x = torch.randn(U, 1024, device="cuda") * 3000
    state = torch.zeros(U * 512, device="cuda")
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
    d_g = torch.from_numpy((rng.random((3 * (U // 24), 4)) * 0.2 + 0.9).astype(np.float32)).cuda()
    w = rng.random((3, 256)).astype(np.float32)
    d_w = torch.from_numpy((w / w.sum(0)).astype(np.float32)).cuda()
    ms = timeit(lambda i: lib.drc_stft_dev(x.data_ptr(), state.data_ptr(), d_rec.data_ptr(), 1, d_g.data_ptr(),
                                           d_w.data_ptr(), U, st))
    report("drc in the STFT domain (t2f + 3-band gains + f2t), 6144 channel-frames", ms, U * (8192 + 4096), U * 1024,
           "channel_samples")


def drc_td_stage():
    """Time-domain multi-band DRC: 6144 channel-frames, 3-band Linkwitz-Riley banks, gain curves, band sum."""
    import drc_cases
    from libmpegh_b200.lib import DRC_TD_BANK_DTYPE, DRC_TD_CHANNEL_DTYPE
    U = 6144
    rng = np.random.default_rng(6)
    banks = np.zeros(1, DRC_TD_BANK_DTYPE)
    nbands = int(os.environ.get("DRC_TD_BANDS", "3"))
    banks[0]["n_bands"], banks[0]["n_cascade"] = nbands, int(os.environ.get("DRC_TD_CASCADE", "0"))
    banks[0]["ap"] = np.stack([drc_cases.random_biquad(rng) for _ in range(9)])
    banks[0]["f"] = np.stack([drc_cases.random_biquad(rng) for _ in range(8)])
    chan = np.zeros(U, DRC_TD_CHANNEL_DTYPE)
    chan["first_seq"] = 3 * (np.arange(U) // 24)
    x = torch.randn(U, 1024, device="cuda") * 3000
    state = torch.zeros(U * lib.drc_td_state_floats(), device="cuda")
    d_chan = torch.from_numpy(chan.view(np.uint8).reshape(-1)).cuda()
    d_banks = torch.from_numpy(banks.view(np.uint8).reshape(-1)).cuda()
    d_g = torch.rand(3 * (U // 24), 1024, device="cuda") * 0.2 + 0.9
    ms = timeit(lambda i: lib.drc_td_dev(x.data_ptr(), state.data_ptr(), d_chan.data_ptr(), d_banks.data_ptr(),
                                         d_g.data_ptr(), U, st))
    report(f"drc time-domain {nbands}-band (filter bank + gains + sum), 6144 channel-frames", ms, U * 8192, U * 1024,
           "channel_samples")


def tns_rs_stage():
    import tns_cases
    from libmpegh_b200.lib import TNS_FILTER_DTYPE
    U = 6144
    rng = np.random.default_rng(1)
    coef = torch.randn(U, 1024, device="cuda") * 500
    chains = []
    bins = 0
    for u in range(U):
        if rng.random() < 0.5:
            continue
        islong = int(rng.random() < 0.9)
        n_filt, filt = tns_cases.random_tns(rng, islong)
        ch = tns_cases.resolve_filters(lib, u, islong, n_filt, filt, 49 if islong else 14, lib.tns_lpc_from_coef)
        chains += ch
        bins += sum(int(r["size"]) for c in ch for r in c)
    recs = np.array([r for c in chains for r in c], TNS_FILTER_DTYPE)
    first = np.cumsum([0] + [len(c) for c in chains]).astype(np.int32)
    d_f = torch.from_numpy(recs.view(np.uint8).reshape(len(recs), -1).copy()).cuda()
    d_first = torch.from_numpy(first).cuda()
    ms = timeit(lambda i: lib.tns_dev(coef.data_ptr(), d_f.data_ptr(), d_first.data_ptr(), len(chains), st))
    report(f"tns 6144 channel-frames, {len(chains)} chains, {bins} filtered bins", ms, bins * 8 + recs.nbytes, bins,
           "filtered_bins")
    g = np.load(os.path.join(ROOT, "tests", "golden", "tns_rs_golden.npz"))
    for up, down in ((3, 2), (2, 1)):
        h = lib.resampler_create(up, down, g["rs_filter_3"] if up == 3 else g["rs_filter_2"], g["rs_filter_2"])
        x = torch.randn(U, 1024, device="cuda") * 5000
        y = torch.empty(U, lib.resampler_out_len(h), device="cuda")
        state = torch.zeros(U * lib.resampler_state_floats(h), device="cuda")
        ms = timeit(lambda i: lib.resample_dev(h, x.data_ptr(), y.data_ptr(), state.data_ptr(), U, st))
        report(f"resampler x{up}/{down}, 6144 channel-frames", ms, U * (4096 + 4 * lib.resampler_out_len(h)), U * 1024,
               "input_samples")
        lib.resampler_destroy(h)


def obj_stage():
    # ---- config 5: 32 objects -> 7.1.4 (512 streams per GPU) + limiter + 24-bit PCM --------------------------
    S, O = 512, 32
    lay = lib.obj_layout_create(loader.obj_layout_golden(19))
    n_spk = 12
    rng = np.random.default_rng(4567)
    import obj_cases  # noqa: E402
    params = obj_cases.random_params(rng, 2, S * O, spread_prob=0.0).reshape(2, S, O, 7)
    sides = [torch.from_numpy(lib.obj_side_from_polar(params[k]).view(np.float32).reshape(S, O, 16).copy()).cuda() for k in range(2)]
    x = torch.randn(S, O, 1024, device="cuda") * 3000
    out = torch.empty(S, n_spk, 1024, device="cuda")
    state = torch.empty(S * lib.obj_state_floats(lay, O), device="cuda")
    scratch = torch.empty(S * lib.obj_scratch_floats(lay, O), device="cuda")
    lib.obj_state_init_dev(lay, O, state.data_ptr(), S)
    ms = timeit(lambda i: lib.obj_render_dev(lay, O, sides[i % 2].data_ptr(), x.data_ptr(), out.data_ptr(), state.data_ptr(),
                                             scratch.data_ptr(), S, st))
    report("obj_render 32 obj -> 7.1.4, 512 streams", ms, S * (O * 4096 + n_spk * 4096), S * O * 1024, "object_samples",
           {"fp32_unfused_Gops": S * O * 11 * 1024 * 3 / (ms * 1e-3) / 1e9})

    p = lib.limiter_params(n_spk)
    lst = torch.empty(S * lib.limiter_state_floats(p), device="cuda")
    lib.limiter_state_init_dev(p, lst.data_ptr(), S)
    lim_out = torch.empty_like(out)
    out.normal_(0, 9000.0)
    ms = timeit(lambda i: lib.limiter_dev(p, out.data_ptr(), lim_out.data_ptr(), lst.data_ptr(), S, None, st))
    report("limiter 12 ch, 512 streams (engaged)", ms, S * (2 * n_spk * 4096 + 2 * 240 * (n_spk + 1) * 4), S * n_spk * 1024,
           "channel_samples")
    out.normal_(0, 300.0)
    lib.limiter_state_init_dev(p, lst.data_ptr(), S)
    ms = timeit(lambda i: lib.limiter_dev(p, out.data_ptr(), lim_out.data_ptr(), lst.data_ptr(), S, None, st))
    report("limiter 12 ch, 512 streams (quiet: smoother skipped)", ms, S * (2 * n_spk * 4096 + 2 * 240 * (n_spk + 1) * 4),
           S * n_spk * 1024, "channel_samples")
    pcm = torch.empty(S, 1024 * n_spk * 3, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda i: lib.pcm_pack_dev(lim_out.data_ptr(), n_spk, 24, pcm.data_ptr(), S, None, None, None, st))
    report("pcm_pack 24-bit 12 ch, 512 streams", ms, S * (n_spk * 4096 + n_spk * 1024 * 3), S * n_spk * 1024, "channel_samples")



def hoa_render_stage():
    # ---- config 3: HOA order 4 (25 coeffs) + NFC -> 22.2, 1024 streams --------------------------------------
    import hoa_cases  # noqa: E402
    g = np.load(os.path.join(ROOT, "tests", "golden", "hoa_ren_golden.npz"))
    k = hoa_cases.key(hoa_cases.CONFIGS[0])
    for use_nfc, tensor in ((1, 0), (0, 0), (1, 1), (0, 1), (0, 2)):
        S = 1024
        h = lib.hoa_renderer_create(g[k + "_M"], g[k + "_nfc"] if use_nfc else None)
        lib.hoa_renderer_set_tensor_mode(h, tensor)
        hin = torch.randn(S, 25, 1024, device="cuda") * 2000
        hout = torch.empty(S, 24, 1024, device="cuda")
        hst = torch.zeros(S * 25 * 14, device="cuda")
        ms = timeit(lambda i: lib.hoa_render_dev(h, hin.data_ptr(), hout.data_ptr(), hst.data_ptr(), S, st))
        report(f"hoa_render order 4 -> 22.2, 1024 streams, nfc={use_nfc}" + (["", ", tensor mode 1 (mma.sync 3xTF32)", ", tensor mode 2 (tcgen05 + TMEM, exact split)"][tensor]), ms, S * (25 + 24) * 4096, S * 24 * 1024,
               "output_channel_samples", {"fp32_unfused_Gops": S * 24 * 25 * 1024 * 2 / (ms * 1e-3) / 1e9})
        lib.hoa_renderer_destroy(h)


def spectral_stage():
    """Joint-stereo tools and IGF at the config-2 shape: 256 streams x 24 ch = 3072 pairs / 6144 channel-frames."""
    import igf_cases
    import stereo_cases
    bands = lib.bands_create(stereo_cases.SFB_LONG, stereo_cases.SFB_SHORT)
    P = 3072
    rng = np.random.default_rng(11)
    coef = torch.randn(2 * P, 1024, device="cuda") * 500
    sds = np.zeros(P, stereo_cases.STEREO_SIDE)
    side = np.zeros((P, 3), np.int64)
    for p in range(P):
        seq = int(rng.choice([0, 1, 2, 3], p=[0.9, 0.03, 0.04, 0.03]))
        sd, _ = stereo_cases.random_frame(rng, seq, int(rng.integers(2)), int(rng.integers(2)), tool=3)
        sd["complex_coef"], sd["use_prev_frame"], sd["prev_dmx"] = 1, 1, 1
        sd["left"], sd["right"], sd["slot"] = 2 * p, 2 * p + 1, p
        sds[p] = sd
    d_pairs = torch.from_numpy(sds.view(np.uint8).reshape(P, -1).copy()).cuda()
    state = torch.randn(P, lib.stereo_state_floats(), device="cuda")
    ms = timeit(lambda i: (lib.stereo_dev(bands, coef.data_ptr(), d_pairs.data_ptr(), state.data_ptr(), P, st),
                           lib.stereo_save_dev(bands, coef.data_ptr(), d_pairs.data_ptr(), state.data_ptr(), P, st)))
    # per pair: spectra 2 x 4096 read + written, saved spectra 2 x 4096 read (previous-frame downmix) + written, side
    report("stereo tools (complex prediction with previous-frame downmix) + save, 3072 pairs", ms,
           P * (4 * 4096 + 4 * 4096 + 664), 2 * P * 1024, "channel_samples")
    sds["tool"] = 1
    d_pairs = torch.from_numpy(sds.view(np.uint8).reshape(P, -1).copy()).cuda()
    ms = timeit(lambda i: lib.stereo_dev(bands, coef.data_ptr(), d_pairs.data_ptr(), state.data_ptr(), P, st))
    report("stereo tools (M/S by mask), 3072 pairs", ms, P * (4 * 4096 + 664), 2 * P * 1024, "channel_samples")

    g = np.load(os.path.join(ROOT, "tests", "golden", "igf_golden.npz"))
    U = 6144
    cfg, grid = igf_cases.CONFIGS[1], g["c1_grid"]
    isd = np.zeros(U, igf_cases.IGF_SIDE)
    base = [igf_cases.random_frame(rng, cfg, grid, win_type=int(k % 16 == 15), all_zero_prob=0.0) for k in range(64)]
    icoef = np.stack([base[u % 64][1] for u in range(U)])
    itiles = np.stack([base[u % 64][2] for u in range(U)])
    for u in range(U):
        isd[u] = base[u % 64][0]
        isd[u]["unit"] = u
    d_isd = torch.from_numpy(isd.view(np.uint8).reshape(U, -1).copy()).cuda()
    d_icoef0 = torch.from_numpy(icoef).cuda()
    d_icoef = d_icoef0.clone()
    d_tiles = torch.from_numpy(itiles).cuda()
    d_tnf = torch.zeros(U, 2048, device="cuda")
    d_seed = torch.zeros(U, dtype=torch.int32, device="cuda")

    def igf_step(i):
        d_icoef.copy_(d_icoef0)  # the gap must be empty again; the copy is inside the timed region (4 KB r + w per unit)
        lib.igf_dev(bands, d_icoef.data_ptr(), d_tiles.data_ptr(), d_isd.data_ptr(), d_tnf.data_ptr(), d_seed.data_ptr(),
                    U, st)
    ms = timeit(igf_step)
    report("igf mono (whitening, noise, TNF on long blocks) incl. a 25 MB spectrum reset copy, 6144 channel-frames", ms,
           U * (2 * 4096 + 4 * 4096 + 548 + 8192), U * 1024, "channel_samples")
    lib.bands_destroy(bands)

    # LTPF at the config-2 shape, every channel filtering with a new pitch each frame (ZIR transition, the costly mode)
    from libmpegh_b200.lib import LTPF_SIDE_DTYPE
    audio = torch.randn(U, 1024, device="cuda") * 3000
    lst = torch.zeros(U, lib.ltpf_state_floats(48000), device="cuda")
    lsd = []
    for k in range(2):
        sd = np.zeros(U, LTPF_SIDE_DTYPE)
        sd["unit"], sd["data_present"] = np.arange(U), 1
        sd["pitch_lag_idx"], sd["gain_idx"] = rng.integers(0, 512, U), rng.integers(0, 4, U)
        lsd.append(torch.from_numpy(sd.view(np.uint8).reshape(U, -1).copy()).cuda())
    ms = timeit(lambda i: lib.ltpf_dev(48000, audio.data_ptr(), lsd[i % 2].data_ptr(), lst.data_ptr(), U, st))
    report("ltpf, new pitch every frame (ZIR transition), 6144 channel-frames", ms,
           U * (2 * 4096 + 2 * 4 * lib.ltpf_state_floats(48000)), U * 1024, "channel_samples")
    same = lsd[0]
    ms = timeit(lambda i: lib.ltpf_dev(48000, audio.data_ptr(), same.data_ptr(), lst.data_ptr(), U, st))
    report("ltpf, steady pitch, 6144 channel-frames", ms, U * (2 * 4096 + 2 * 4 * lib.ltpf_state_floats(48000)),
           U * 1024, "channel_samples")


def fd_path_stage():
    """Subsystem (1) of BASELINE.json north_star as one chain at the config-2 shape (256 streams x 24 ch): IGF on every
    channel, TNS on half of them (realistic orders <= 8 over the upper half of the spectrum), joint stereo on every
    pair (complex prediction with previous-frame downmix), FD synthesis, LTPF with a steady pitch."""
    import fd_cases
    import igf_cases
    import stereo_cases
    import tns_cases
    from libmpegh_b200.lib import LTPF_SIDE_DTYPE, TNS_FILTER_DTYPE
    U, P = 6144, 3072
    rng = np.random.default_rng(21)
    bands = lib.bands_create(stereo_cases.SFB_LONG, stereo_cases.SFB_SHORT)
    g = np.load(os.path.join(ROOT, "tests", "golden", "igf_golden.npz"))
    cfg, grid = igf_cases.CONFIGS[1], g["c1_grid"]
    base = [igf_cases.random_frame(rng, cfg, grid, win_type=0, all_zero_prob=0.0) for _ in range(64)]
    isd = np.zeros(U, igf_cases.IGF_SIDE)
    for u in range(U):
        isd[u] = base[u % 64][0]
        isd[u]["unit"] = u
    d_isd = torch.from_numpy(isd.view(np.uint8).reshape(U, -1).copy()).cuda()
    coef0 = torch.from_numpy(np.stack([base[u % 64][1] for u in range(U)])).cuda()
    coef = coef0.clone()
    d_tiles = torch.from_numpy(np.stack([base[u % 64][2] for u in range(U)])).cuda()
    d_tnf = torch.zeros(U, 2048, device="cuda")
    d_seed = torch.zeros(U, dtype=torch.int32, device="cuda")
    # TNS: one or two filters of order 4..8 over the bands above 4 kHz on half of the channels
    chains = []
    for u in range(0, U, 2):
        n_filt, filt = np.zeros(8, np.int32), np.zeros((8, 3, 20), np.int32)
        n_filt[0] = int(rng.integers(1, 3))
        top = 49
        for f in range(n_filt[0]):
            order = int(rng.integers(4, 9))
            length = int(rng.integers(6, 12))
            filt[0, f, :5] = (order, top - length, top, int(rng.integers(2)), 4)
            filt[0, f, 5:5 + order] = rng.integers(-8, 8, order)
            top -= length
        chains += tns_cases.resolve_filters(lib, u, 1, n_filt, filt, 49, lib.tns_lpc_from_coef)
    recs = np.array([r for c in chains for r in c], TNS_FILTER_DTYPE)
    first = np.cumsum([0] + [len(c) for c in chains]).astype(np.int32)
    d_f = torch.from_numpy(recs.view(np.uint8).reshape(len(recs), -1).copy()).cuda()
    d_first = torch.from_numpy(first).cuda()
    sds = np.zeros(P, stereo_cases.STEREO_SIDE)
    for p in range(P):
        sd, _ = stereo_cases.random_frame(rng, 0, int(rng.integers(2)), int(rng.integers(2)), tool=3)
        sd["complex_coef"], sd["use_prev_frame"], sd["prev_dmx"] = 1, 1, 1
        sd["left"], sd["right"], sd["slot"] = 2 * p, 2 * p + 1, p
        sds[p] = sd
    d_pairs = torch.from_numpy(sds.view(np.uint8).reshape(P, -1).copy()).cuda()
    pstate = torch.zeros(P, lib.stereo_state_floats(), device="cuda")
    side = torch.from_numpy(fd_cases.pack_side(np.stack([np.zeros(U, np.int32), rng.integers(0, 2, U).astype(np.int32),
                                                         rng.integers(0, 2, U).astype(np.int32), np.zeros(U, np.int32),
                                                         np.zeros(U, np.int32)], -1)).astype(np.int32)).cuda()
    ovl = torch.zeros(U, 1024, device="cuda")
    pcm = torch.empty(U, 1024, device="cuda")
    lsd = np.zeros(U, LTPF_SIDE_DTYPE)
    lsd["unit"], lsd["data_present"] = np.arange(U), 1
    lsd["pitch_lag_idx"], lsd["gain_idx"] = rng.integers(0, 512, U), rng.integers(0, 4, U)
    d_lsd = torch.from_numpy(lsd.view(np.uint8).reshape(U, -1).copy()).cuda()
    lst = torch.zeros(U, lib.ltpf_state_floats(48000), device="cuda")

    def step(i):
        coef.copy_(coef0)  # the frame's dequantised spectra "arrive" (stands in for the upload)
        lib.igf_dev(bands, coef.data_ptr(), d_tiles.data_ptr(), d_isd.data_ptr(), d_tnf.data_ptr(), d_seed.data_ptr(), U, st)
        lib.tns_dev(coef.data_ptr(), d_f.data_ptr(), d_first.data_ptr(), len(chains), st)
        lib.stereo_dev(bands, coef.data_ptr(), d_pairs.data_ptr(), pstate.data_ptr(), P, st)
        lib.stereo_save_dev(bands, coef.data_ptr(), d_pairs.data_ptr(), pstate.data_ptr(), P, st)
        lib.fd_synth_dev(coef.data_ptr(), side.data_ptr(), ovl.data_ptr(), pcm.data_ptr(), U, st)
        lib.ltpf_dev(48000, pcm.data_ptr(), d_lsd.data_ptr(), lst.data_ptr(), U, st)
    ms = timeit(step)
    per_unit = 4096 + 4 * 4096 + 4096 + 2 * 4096 + 2 * 4 * lib.ltpf_state_floats(48000)  # coef, tiles, pcm, overlap r/w, ltpf state
    report("FD path chain: IGF + TNS + joint stereo + FD synthesis + LTPF, 256 streams x 24 ch (7 launches)", ms,
           U * per_unit, U * 1024, "channel_samples", {"audio_s_per_wall_s": 256 * 1024 / 48000 / (ms * 1e-3),
                                                       "tns_chains": len(chains)})
    lib.bands_destroy(bands)


STAGES = {"mct": mct_stage, "drc": drc_stage, "drc_td": drc_td_stage, "binaural": binaural_stage, "hoa": hoa_chain_stage, "fc": fc_stage, "tns_rs": tns_rs_stage,
          "fd_path": fd_path_stage,
          "spectral": spectral_stage,
          "obj": obj_stage, "hoa_render": hoa_render_stage}
if ONLY in STAGES:
    STAGES[ONLY]()
    lib.close()
    sys.exit(0)

obj_stage()
hoa_render_stage()
hoa_chain_stage()
fc_stage()
tns_rs_stage()
mct_stage()
drc_stage()
drc_td_stage()
spectral_stage()
fd_path_stage()
binaural_stage()
lib.close()
