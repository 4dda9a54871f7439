"""Seeded multichannel-coding-tool cases shared by the oracle and GPU tests (decoder/impeghd_multichannel.c:899,
window loop decoder/ia_core_coder_process.c:1043-1075)."""
import numpy as np

from tns_cases import SFB_LONG, SFB_SHORT


def random_group(rng, n_ch=24, max_boxes=12):
    """One stream-frame: a list of boxes.  All channels of a group share the block type here (the reference rejects
    a pair of different types, process.c:954)."""
    islong = bool(rng.random() < 0.8)
    signal_type = int(rng.integers(0, 2))  # per tool instance: 0 prediction, 1 rotation
    sfb = SFB_LONG if islong else SFB_SHORT
    n_win, bins = (1, 1024) if islong else (8, 128)
    total_sfb = len(sfb)
    boxes = []
    for _ in range(int(rng.integers(1, max_boxes + 1))):
        a, b = rng.choice(n_ch, 2, replace=False)
        fullband = bool(rng.random() < 0.3)
        bpw = int(rng.integers(1, (total_sfb + 1) // 2 + 1))  # bands per window, two sfb each
        if not islong:
            bpw = min(bpw, 8)
        nb = max(bpw, 1) * n_win
        if signal_type:
            coef = rng.integers(0, 65, nb).astype(np.int32)
        else:
            coef = rng.integers(-30, 31, nb).astype(np.int32)
        mask = (rng.random(nb) < 0.7).astype(np.int32)
        boxes.append(dict(ch_l=int(a), ch_r=int(b), signal_type=signal_type, pred_dir=int(rng.integers(0, 2)),
                          fullband=int(fullband), n_win=n_win, bins=bins, bpw=bpw, total_sfb=total_sfb, sfb=sfb,
                          coef=coef, mask=mask))
    return boxes


def apply_group(process, spec, boxes):
    """spec [n_ch][1024] float32, in place, through a per-window process(dmx, res, coef, mask, bpw, total_sfb, sfb,
    bins, signal_type, pred_dir, fullband) -- the oracle's or the reference's."""
    for bx in boxes:
        l, r = spec[bx["ch_l"]], spec[bx["ch_r"]]
        for w in range(bx["n_win"]):
            d = np.ascontiguousarray(l[w * bx["bins"]:(w + 1) * bx["bins"]])
            q = np.ascontiguousarray(r[w * bx["bins"]:(w + 1) * bx["bins"]])
            c = np.ascontiguousarray(bx["coef"][w * bx["bpw"]:(w + 1) * bx["bpw"]])
            m = np.ascontiguousarray(bx["mask"][w * bx["bpw"]:(w + 1) * bx["bpw"]])
            process(d, q, c, m, bx["bpw"], bx["total_sfb"], bx["sfb"], bx["bins"], bx["signal_type"],
                    bx["pred_dir"], bx["fullband"])
            l[w * bx["bins"]:(w + 1) * bx["bins"]] = d
            r[w * bx["bins"]:(w + 1) * bx["bins"]] = q
    return spec


def resolve_group(lib, boxes, unit0):
    """-> MCT_BOX_DTYPE records for a group whose channel c is row unit0 + c of the spectrum array"""
    return [lib.mct_resolve(bx["signal_type"], bx["pred_dir"], bx["fullband"], bx["n_win"], bx["bins"], bx["bpw"],
                            bx["sfb"], bx["coef"], bx["mask"], unit0 + bx["ch_l"], unit0 + bx["ch_r"])
            for bx in boxes]


def apply_records_numpy(spec, recs, sin_t, cos_t):
    """What the kernel does, in numpy float32 (every product and sum rounded separately): checks mct_resolve on CPU."""
    f = np.float32
    for r in recs:
        for s in range(int(r["n_seg"])):
            a, n, c = int(r["start"][s]), int(r["len"][s]), int(r["coef"][s])
            d = spec[r["unit_l"], a:a + n].copy()
            q = spec[r["unit_r"], a:a + n].copy()
            if r["op"] == 2:
                cs, sn = f(cos_t[c]), f(sin_t[c])
                spec[r["unit_l"], a:a + n] = d * cs - q * sn
                spec[r["unit_r"], a:a + n] = d * sn + q * cs
            else:
                alpha = f(f(c) * f(0.1))
                ad = alpha * d
                spec[r["unit_l"], a:a + n] = (d + ad) + q
                spec[r["unit_r"], a:a + n] = ((d - ad) - q) if r["op"] == 0 else (((-d) + ad) + q)
    return spec
