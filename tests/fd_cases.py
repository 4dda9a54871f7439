"""Shared FD-synthesis test-case generation (seeded, used by golden generation, CPU and GPU tests)."""
import numpy as np

SEQ_NAMES = ["ONLY_LONG", "LONG_START", "EIGHT_SHORT", "LONG_STOP", "STOP_START"]

# legal successor window sequences (ISO/IEC 23003-3 table; LPD transitions excluded)
NEXT = {0: [0, 1], 1: [2, 3], 2: [2, 3, 4], 3: [0, 1], 4: [2, 3]}


def random_side(rng, n_frames, n_ch, kernel_switch_prob=0.0, legal=True):
    """side[frame][ch] = (sequence, shape, shape_prev, prev_sym, curr_sym) with state-consistent history."""
    side = np.zeros((n_frames, n_ch, 5), dtype=np.int32)
    seq = np.zeros(n_ch, dtype=np.int64)
    shape_prev = np.zeros(n_ch, dtype=np.int64)
    sym_prev = np.zeros(n_ch, dtype=np.int64)
    for f in range(n_frames):
        for c in range(n_ch):
            if f > 0 or True:
                choices = NEXT[int(seq[c])] if (legal and f > 0) else [0, 1, 2, 3, 4]
                seq[c] = choices[int(rng.integers(len(choices)))]
            shape = int(rng.integers(2))
            sym = int(rng.random() < kernel_switch_prob)
            side[f, c] = (seq[c], shape, shape_prev[c], sym_prev[c], sym)
            shape_prev[c] = shape
            sym_prev[c] = sym
    return side


def pack_side(side):
    from libmpegh_b200.lib import fd_side
    return fd_side(side[..., 0], side[..., 1], side[..., 2], side[..., 3], side[..., 4])


def random_coef(rng, shape, sigma=500.0):
    return rng.normal(0.0, sigma, size=shape).astype(np.float32)


def all_combinations():
    """Every (sequence, shape, shape_prev, prev_sym, curr_sym): 5 * 2 * 2 * 4 = 80 single-frame cases."""
    out = []
    for seq in range(5):
        for shape in range(2):
            for sp in range(2):
                for tkt in range(4):
                    out.append((seq, shape, sp, tkt >> 1, tkt & 1))
    return np.array(out, dtype=np.int32)
