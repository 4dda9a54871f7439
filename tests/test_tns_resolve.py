"""mpeghb200_tns_resolve (host half of ia_core_coder_tns_apply, no GPU needed) against the reference function and
the oracle's band resolution."""
import numpy as np
import pytest

import tns_cases
from oracle import loader
from oracle.loader import bits_equal


def channel_record(unit, islong, n_filt, filt, max_sfb, igf=(0, 0, 0, 0)):
    from libmpegh_b200.lib import TNS_CHANNEL_DTYPE
    ch = np.zeros((), TNS_CHANNEL_DTYPE)
    ch["unit"], ch["is_long"], ch["max_sfb"] = unit, islong, max_sfb
    ch["igf_active"], ch["igf_after_tns_synth"], ch["igf_sfb_start"], ch["igf_sfb_stop"] = igf
    ch["n_filt"] = n_filt
    for w in range(8):
        res = 0
        for f in range(3):
            ch["filt"][w, f]["order"], ch["filt"][w, f]["start_band"] = filt[w, f, 0], filt[w, f, 1]
            ch["filt"][w, f]["stop_band"], ch["filt"][w, f]["direction"] = filt[w, f, 2], filt[w, f, 3]
            ch["filt"][w, f]["coef"][:15] = filt[w, f, 5:20]
            res = int(filt[w, f, 4]) or res  # the harness layout carries coef_res per filter; one value per window
        ch["coef_res"][w] = res or 3
    return ch


def apply_chains(oracle, x, chains):
    y = x.copy()
    for chain in chains:
        for r in chain:
            seg = y[int(r["start"]):int(r["start"]) + int(r["size"])]
            oracle.oracle_tns_filter(seg.ctypes.data, int(r["size"]), int(r["inc"]), int(r["order"]),
                                     np.ascontiguousarray(r["lpc"], np.float32))
    return y


@pytest.mark.parametrize("islong", [1, 0])
def test_resolve_matches_reference(so_lib, oracle, islong):
    rng = np.random.default_rng(40 + islong)
    sfb = tns_cases.SFB_LONG if islong else tns_cases.SFB_SHORT
    tmax = tns_cases.TNS_MAX_LONG if islong else tns_cases.TNS_MAX_SHORT
    ref = loader.load_ref()
    for k in range(60):
        n_filt, filt = tns_cases.random_tns(rng, islong)
        nbands = int(rng.integers(len(sfb) // 2, len(sfb) + 1))
        x = tns_cases.spectrum(rng)
        chains = so_lib.tns_resolve(channel_record(7, islong, n_filt, filt, nbands), sfb, tmax)
        assert all(int(r["unit"]) == 7 for c in chains for r in c)
        got = apply_chains(oracle, x, chains)
        want, _ = loader.oracle_tns_apply(x, islong, n_filt, filt, sfb, tmax, nbands)
        assert bits_equal(got, want), k
        if ref is not None:
            y = x.copy()
            assert ref.ref_tns_apply(y, islong, n_filt, np.ascontiguousarray(filt.reshape(-1)), sfb, len(sfb), tmax,
                                     nbands) == 0
            assert bits_equal(got, y), k


def test_igf_band_clamps_and_errors(so_lib):
    from libmpegh_b200.lib import MpeghError
    sfb = tns_cases.SFB_LONG
    n_filt = np.zeros(8, np.int32)
    filt = np.zeros((8, 3, 20), np.int32)
    n_filt[0] = 1
    filt[0, 0, :5] = (4, 10, 45, 0, 4)
    filt[0, 0, 5:9] = (1, -2, 3, -1)
    # no IGF: stop clamped by tns_max_band (40) and max_sfb (30)
    (c,), = [so_lib.tns_resolve(channel_record(0, 1, n_filt, filt, 30), sfb, 40)]
    assert int(c[0]["start"]) == sfb[9] and int(c[0]["start"]) + int(c[0]["size"]) == sfb[29]
    # IGF after TNS synthesis: nbands drops to the IGF start band (20); stop is not clamped by tns_max_band
    (c,), = [so_lib.tns_resolve(channel_record(0, 1, n_filt, filt, 30, igf=(1, 1, 20, 49)), sfb, 40)]
    assert int(c[0]["start"]) + int(c[0]["size"]) == sfb[19]
    # IGF before TNS: nbands rises to the IGF stop band (49): stop = min(45, 49) = 45 bands
    (c,), = [so_lib.tns_resolve(channel_record(0, 1, n_filt, filt, 30, igf=(1, 0, 20, 49)), sfb, 40)]
    assert int(c[0]["start"]) + int(c[0]["size"]) == sfb[44]
    # a band index beyond the table is the reference's fatal INVALID_TNS_SFB
    filt[0, 0, 2] = 60
    with pytest.raises(MpeghError):
        so_lib.tns_resolve(channel_record(0, 1, n_filt, filt, 64, igf=(1, 0, 20, 64)), sfb, 40)
    # order 0 and empty ranges produce no chain
    filt[0, 0, :3] = (0, 10, 20)
    assert so_lib.tns_resolve(channel_record(0, 1, n_filt, filt, 30), sfb, 40) == []
