"""Shared TNS / resampler test cases."""
import numpy as np

SFB_LONG = np.array([4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 48, 56, 64, 72, 80, 88, 96, 108, 120, 132, 144, 160, 176,
                     196, 216, 240, 264, 292, 320, 352, 384, 416, 448, 480, 512, 544, 576, 608, 640, 672, 704, 736,
                     768, 800, 832, 864, 896, 928, 1024], np.int16)
SFB_SHORT = np.array([4, 8, 12, 16, 20, 28, 36, 44, 56, 68, 80, 96, 112, 128], np.int16)
TNS_MAX_LONG, TNS_MAX_SHORT = 40, 14


def random_tns(rng, islong):
    """-> (n_filt [n_win], filt [n_win, 3, 20]) in the harness layout {order, start, stop, dir, coef_res, coef[15]}"""
    n_win = 1 if islong else 8
    nb = len(SFB_LONG) if islong else len(SFB_SHORT)
    n_filt = np.zeros(8, np.int32)
    filt = np.zeros((8, 3, 20), np.int32)
    for w in range(n_win):
        nf = int(rng.integers(0, 4)) if islong else int(rng.integers(0, 2))
        n_filt[w] = nf
        top = nb
        res = int(rng.integers(3, 5))
        for f in range(nf):
            length = int(rng.integers(1, max(2, top)))
            order = int(rng.integers(0, 16 if islong else 8))
            filt[w, f, 0] = order
            filt[w, f, 1] = max(top - length, 0)
            filt[w, f, 2] = top
            filt[w, f, 3] = int(rng.integers(0, 2))
            filt[w, f, 4] = res
            lim = 1 << (res - 1)
            filt[w, f, 5:5 + order] = rng.integers(-lim, lim, order)
            top = max(top - length, 0)
    return n_filt, filt


def spectrum(rng, n=1024, sigma=500.0):
    return rng.normal(0, sigma, n).astype(np.float32)


def resolve_filters(lib, unit, islong, n_filt, filt, nbands, lpc_fn):
    """Host-side resolution of one channel's TNS side info into mpeghb200_tns_filter records, grouped in chains
    (one per window), the way ia_core_coder_tns_apply (decoder/ia_core_coder_tns.c:240-297) walks them.
    lpc_fn(order, coef_res, coef) -> lpc[16]."""
    from libmpegh_b200.lib import TNS_FILTER_DTYPE
    sfb = SFB_LONG if islong else SFB_SHORT
    tmax = TNS_MAX_LONG if islong else TNS_MAX_SHORT
    n_win, bins = (1, 1024) if islong else (8, 128)
    chains = []
    for w in range(n_win):
        recs = []
        for f in range(int(n_filt[w])):
            order, sb, eb, direction, res = (int(v) for v in filt[w, f, :5])
            if order == 0:
                continue
            start = min(sb, tmax, nbands)
            stop = min(eb, tmax, nbands)
            a = int(sfb[start - 1]) if start > 0 else 0
            b = int(sfb[stop - 1]) if stop > 0 else 0
            if b - a <= 0:
                continue
            r = np.zeros((), TNS_FILTER_DTYPE)
            r["unit"], r["start"], r["size"], r["inc"], r["order"] = unit, w * bins + a, b - a, -1 if direction else 1, order
            r["lpc"] = lpc_fn(order, res, filt[w, f, 5:20])
            recs.append(r)
        if recs:
            chains.append(recs)
    return chains
