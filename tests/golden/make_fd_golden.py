#!/usr/bin/env python3
"""Generate tests/golden/fd_golden.npz with the UNMODIFIED reference (oracle/_ref/libmpeghref.so).

TEST INFRASTRUCTURE.  Two sets of ia_core_coder_fd_frm_dec (decoder/ia_core_coder_imdct.c:832) vectors:
  combos_*  one frame for each of the 80 (sequence, shape, shape_prev, prev_sym, curr_sym) combinations
            with a random non-zero overlap state: pins every windowing branch and transform kernel type
  chain_*   12 consecutive frames x 3 channels with legal random window-sequence transitions and occasional
            kernel switching, state carried: pins the overlap hand-over between sequences
Inputs are float32 drawn from numpy.random.default_rng(seed); outputs are stored as raw float32.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import loader  # noqa: E402
import fd_cases  # noqa: E402


def main():
    ref = loader.load_ref()
    if ref is None:
        sys.exit("needs oracle/_ref (make -C oracle ref)")
    h = ref.ref_fd_create()
    rng = np.random.default_rng(20240601)

    combos = fd_cases.all_combinations()
    n = len(combos)
    c_coef = fd_cases.random_coef(rng, (n, 1024))
    c_ov_in = fd_cases.random_coef(rng, (n, 1024), 300.0)
    c_ov_out = c_ov_in.copy()
    c_out = np.zeros((n, 1024), np.float32)
    for i, s in enumerate(combos):
        err = ref.ref_fd_frm_dec(h, c_coef[i], c_ov_out[i], c_out[i], *[int(v) for v in s])
        assert err == 0

    F, C = 12, 3
    side = fd_cases.random_side(rng, F, C, kernel_switch_prob=0.15)
    coef = fd_cases.random_coef(rng, (F, C, 1024))
    ov = np.zeros((C, 1024), np.float32)
    out = np.zeros((F, C, 1024), np.float32)
    err = ref.ref_fd_frm_dec_batch(h, coef, ov, out, side, F, C)
    assert err == 0
    ref.ref_fd_destroy(h)

    np.savez_compressed(os.path.join(HERE, "fd_golden.npz"), combos_side=combos, combos_coef=c_coef,
                        combos_ov_in=c_ov_in, combos_ov_out=c_ov_out, combos_out=c_out, chain_side=side,
                        chain_coef=coef, chain_out=out, chain_ov_final=ov)
    print("wrote fd_golden.npz", n, "combos,", F, "x", C, "chain")


if __name__ == "__main__":
    main()
