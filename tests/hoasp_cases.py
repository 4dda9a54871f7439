"""Shared HOA spatial-decoder test-case generation: plausible sequences of per-frame side info
(what impeghd_hoa_spatial_decode_frame_side_info leaves in ia_hoa_dec_frame_param_str, stale slots included)."""
import numpy as np

MAX_CH, MAX_COEF, MAX_PRED = 16, 49, 4
SIDE_DTYPE = np.dtype([
    ("n_vec", np.int32), ("n_dir", np.int32), ("n_enable", np.int32), ("n_disable", np.int32),
    ("exponents", np.int32, MAX_CH), ("is_exception", np.int32, MAX_CH), ("prev_gain_reset", np.float32, MAX_CH),
    ("amb_hoa_assign", np.int32, MAX_CH), ("vec_index", np.int32, MAX_CH),
    ("dir_ch", np.int32, MAX_CH + 1), ("dir_id", np.int32, MAX_CH + 1), ("pad0", np.int32, 2),
    ("enable", np.int32, MAX_COEF), ("disable", np.int32, MAX_COEF), ("non_en_dis", np.int32, MAX_COEF),
    ("pred_type", np.int32, MAX_COEF), ("pred_idx", np.int32, (MAX_PRED, MAX_COEF)),
    ("pred_quant", np.int32, (MAX_PRED, MAX_COEF)), ("vec", np.float32, (MAX_CH, MAX_COEF))])
assert SIDE_DTYPE.itemsize == 4 * (4 + 5 * 16 + 34 + 2 + 4 * 49 + 8 * 49 + 16 * 49)

# (order, n_transport, n_addnl, min_amb_order, coded_v_vec_length, spat_interp_time, spat_interp_method,
#  pred_dir_sigs, num_bits_per_sf, use_phase_shift_decorr); BASELINE config 3 first
CONFIGS = [
    (4, 16, 12, 1, 0, 256, 0, 2, 8, 0),
    (4, 8, 4, 1, 1, 256, 1, 2, 8, 0),
    (3, 9, 8, 0, 2, 512, 0, 3, 8, 0),
    (6, 16, 7, 2, 1, 1024, 0, 4, 8, 0),
    (1, 4, 0, 1, 0, 256, 0, 1, 8, 0),
    (2, 6, 5, 0, 0, 0, 0, 2, 4, 0),
]


def key(cfg):
    return "o%d_t%d_a%d_m%d_v%d_i%d_%d_p%d_b%d" % cfg[:9]


def cfg_array(cfg):
    return np.array(list(cfg) + [0, 0], np.int32)


class SideGen:
    """Random walk over channel types with the enable / steady / disable life cycle of additional ambient
    coefficients; keeps the persistent (stale) slots the reference keeps."""

    def __init__(self, cfg, rng, pred_prob=0.3, wild_exp=False):
        self.cfg, self.rng = cfg, rng
        self.order, self.T, self.n_addnl, self.min_amb = cfg[0], cfg[1], cfg[2], cfg[3]
        self.nc = (self.order + 1) ** 2
        self.low = (self.min_amb + 1) ** 2
        self.vec_start = self.low if cfg[4] > 0 else 0
        self.pred_prob, self.wild_exp = pred_prob, wild_exp
        self.kind = ["empty"] * self.n_addnl       # empty, vec, dir, amb
        self.amb = [None] * self.n_addnl           # (coeff, state) state 1 enable, 0 steady, 2 disable
        self.side = np.zeros((), SIDE_DTYPE)
        self.side["enable"] = -1
        self.side["disable"] = -1
        self.side["non_en_dis"] = -1
        self.first = True

    def next(self):
        rng, sd = self.rng, self.side
        used = {a[0] for a in self.amb if a}
        for ch in range(self.n_addnl):
            if self.kind[ch] == "amb":
                c, stt = self.amb[ch]
                if stt == 2:
                    self.kind[ch], self.amb[ch] = "empty", None
                elif stt == 1:
                    self.amb[ch] = (c, 0)
                elif rng.random() < 0.3:
                    self.amb[ch] = (c, 2)
                continue
            if rng.random() < 0.35:
                k = rng.choice(["empty", "vec", "dir", "amb"], p=[0.15, 0.4, 0.25, 0.2])
                if k == "amb":
                    free = [c for c in range(self.low + 1, self.nc + 1) if c not in used]
                    if not free:
                        k = "vec"
                    else:
                        c = int(rng.choice(free))
                        used.add(c)
                        self.amb[ch] = (c, 1)
                self.kind[ch] = k
        sd["amb_hoa_assign"] = 0
        off = max(self.T - self.low, 0)
        for i in range(self.low):
            if off + i < MAX_CH:
                sd["amb_hoa_assign"][off + i] = i + 1
        en, dis, non = [], [], []
        nv = nd = 0
        sd["vec"] = 0
        for ch in range(self.n_addnl):
            k = self.kind[ch]
            if k == "vec":
                v = rng.normal(0, 1, self.nc - self.vec_start)
                v = v / np.linalg.norm(v) * (self.order + 1)
                sd["vec"][nv, :self.nc - self.vec_start] = v.astype(np.float32)
                sd["vec_index"][nv] = ch + 1
                nv += 1
            elif k == "dir":
                sd["dir_ch"][nd] = ch + 1
                sd["dir_id"][nd] = 0 if rng.random() < 0.1 else int(rng.integers(1, 901))
                nd += 1
            elif k == "amb":
                c, stt = self.amb[ch]
                sd["amb_hoa_assign"][ch] = c
                (en if stt == 1 else dis if stt == 2 else non).append(c)
        sd["n_vec"], sd["n_dir"] = nv, nd
        for name, lst in (("enable", en), ("disable", dis), ("non_en_dis", non)):
            sd[name] = -1
            sd[name][:len(lst)] = sorted(lst)
        sd["n_enable"], sd["n_disable"] = len(en), len(dis)
        # gain correction
        e = rng.choice([-2, -1, 0, 1, 2], size=MAX_CH, p=[0.1, 0.25, 0.3, 0.25, 0.1]).astype(np.int32)
        if self.wild_exp:
            wild = rng.random(MAX_CH) < 0.08
            e[wild] = rng.choice([-8, -7, 7, 9], size=int(wild.sum()))
        sd["exponents"] = e
        sd["is_exception"] = (rng.random(MAX_CH) < 0.1).astype(np.int32)
        sd["prev_gain_reset"] = 0
        if not self.first and rng.random() < 0.15:   # hoa_independency_flag
            sd["prev_gain_reset"][:self.T] = np.exp2(rng.integers(-3, 4, self.T)).astype(np.float32)
        # spatial prediction
        if self.n_addnl and rng.random() < self.pred_prob:
            sd["pred_type"] = 0
            for g in rng.choice(self.nc, size=int(rng.integers(1, 4)), replace=False):
                sd["pred_type"][g] = 1
                n = int(rng.integers(1, self.cfg[7] + 1))
                sd["pred_idx"][:, g] = 0
                sd["pred_idx"][:n, g] = rng.integers(1, self.T + 1, n)
                sd["pred_quant"][:n, g] = rng.integers(-128, 128, n)
        else:
            sd["pred_type"] = 0
        self.first = False
        return sd.copy()


def signals(rng, n_frames, T, sigma=2000.0):
    return rng.normal(0.0, sigma, size=(n_frames, T, 1024)).astype(np.float32)


def golden_tables(g, cfg):
    """Tables of one configuration from tests/golden/hoasp_golden.npz, re-expanded to the strides the API takes."""
    k = key(cfg)
    nc = (cfg[0] + 1) ** 2
    fine = np.zeros((900, 49), np.float32)
    fine[:, :nc] = g[k + "_fine"]
    return {"info": g[k + "_info"], "order_matrix": g[k + "_order_matrix"], "fine": fine.reshape(-1),
            "coarse": g[k + "_coarse"], "wins": g[k + "_wins"], "dyn_pow": g["dyn_pow"], "dyn_win": g["dyn_win"],
            "pow2": g["pow2"]}
