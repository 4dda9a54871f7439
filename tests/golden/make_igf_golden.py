"""IGF (mono path) vectors from the unmodified reference: grids derived by ia_core_coder_igf_init for the test
configurations, whitening gain ROM, and ia_core_coder_igf_mono + ia_core_coder_igf_tnf outputs
(oracle/ref_harness_cpe.c: ref_igf_*), into tests/golden/igf_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import igf_cases  # noqa: E402
import stereo_cases  # noqa: E402
from oracle import loader  # noqa: E402

ref = loader.load_ref()


def rom(name, n):
    return np.ctypeslib.as_array((ctypes.c_float * n).in_dll(ref, name)).copy()


def reference_grid(cfg):
    h = ref.ref_igf_create(stereo_cases.SAMPLING_RATE_IDX, 48000, *cfg)
    grid = np.zeros(8, np.int32)
    ref.ref_igf_grid(h, grid)
    return h, grid.reshape(2, 4)


if __name__ == "__main__":
    out = {"whiten_pos": rom("ia_core_coder_pow_igf_whitening", 64),
           "whiten_neg": rom("ia_core_coder_pow_neg_igf_whitening", 64),
           "tnf_acf_win": rom("ia_core_coder_tnf_acf_win", 8)}
    N = 8
    for ci, cfg in enumerate(igf_cases.CONFIGS):
        h, grid = reference_grid(cfg)
        out[f"c{ci}_grid"] = grid
        rng = np.random.default_rng(8800 + ci)
        sds, cin, cout, tl, seeds = [], [], [], [], []
        for k in range(N):
            sd, coef, tiles = igf_cases.random_frame(rng, cfg, grid, win_type=int(k % 4 == 3))
            y = coef.copy()
            seeds.append(ref.ref_igf_process(h, sd.tobytes(), y, tiles))
            sds.append(sd), cin.append(coef), cout.append(y), tl.append(tiles)
        ref.ref_igf_destroy(h)
        out[f"c{ci}_sd"] = np.array(sds, igf_cases.IGF_SIDE).view(np.uint8).reshape(N, -1)
        out[f"c{ci}_in"], out[f"c{ci}_out"], out[f"c{ci}_tiles"] = np.stack(cin), np.stack(cout), np.stack(tl)
        out[f"c{ci}_seed"] = np.array(seeds, np.uint32)
    # joint-stereo path: pairs over consecutive frames (TNF buffers carried in the handle)
    NP = 6
    for ci, cfg in enumerate(igf_cases.CONFIGS):
        h, grid = reference_grid(cfg)
        rng = np.random.default_rng(9900 + ci)
        rec = {k: [] for k in ("sdl", "sdr", "mask", "in", "tiles", "out", "seeds", "ms")}
        for k in range(NP):
            sd_l, sd_r, mask, coef, tiles = igf_cases.random_pair_frame(rng, cfg, grid, win_type=int(k % 3 == 2))
            ms_after = int(rng.integers(2))
            y = coef.copy()
            seeds = np.zeros(2, np.uint32)
            ref.ref_igf_stereo_process(h, sd_l.tobytes(), sd_r.tobytes(), igf_cases.mask_reference_layout(mask, sd_l),
                                       ms_after, y[0], y[1], np.ascontiguousarray(tiles[0]).reshape(-1),
                                       np.ascontiguousarray(tiles[1]).reshape(-1), seeds)
            for key, v in zip(rec, (sd_l, sd_r, mask, coef, tiles, y, seeds, ms_after)):
                rec[key].append(v)
        ref.ref_igf_destroy(h)
        out[f"s{ci}_sdl"] = np.array(rec["sdl"], igf_cases.IGF_SIDE).view(np.uint8).reshape(NP, -1)
        out[f"s{ci}_sdr"] = np.array(rec["sdr"], igf_cases.IGF_SIDE).view(np.uint8).reshape(NP, -1)
        for key in ("mask", "in", "tiles", "out", "seeds"):
            out[f"s{ci}_{key}"] = np.stack(rec[key])
        out[f"s{ci}_ms"] = np.array(rec["ms"], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "igf_golden.npz"), **out)
    print("wrote igf_golden.npz", {k: v.tolist() for k, v in out.items() if k.endswith("grid")})
