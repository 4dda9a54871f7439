"""Whole-configuration chains on the CPU: the oracle's stage functions composed in the reference's order
(decoder/ia_core_coder_decode_main.c:2736-2893, ia_core_coder_dec_ext_ele_proc :1949-2316), one stream at a time.
TEST INFRASTRUCTURE: the checker the mpeghb200_chain_* GPU sessions are compared with, byte for byte, on the final PCM.
Every stage used here is pinned to the unmodified reference by the tests/test_oracle_*.py files."""
import numpy as np

from oracle import loader


class OracleChain:
    """Mirror of mpeghb200_chain_desc.  Per frame: core = list of per-region arrays [S][n][1024] in the order bed,
    objects, HOA transport (regions with n = 0 omitted from nothing: pass None)."""

    def __init__(self, S, n_bed=0, n_obj=0, obj_cicp=None, hoa=None, fc_D=None, binaural=None, limiter=1, pcm_bits=24,
                 start_delay=0, spectra=False, loudness_gain=None, ch_map=None):
        self.S, self.n_bed, self.n_obj, self.spectra = S, n_bed, n_obj, spectra
        self.o = loader.load_oracle()
        self.pcm_bits, self.ch_map = pcm_bits, ch_map
        self.delay = [start_delay] * S
        self.gain = loudness_gain
        self.n_mix = n_bed
        if n_obj:
            self.L = loader.obj_layout_from_arrays(loader.obj_layout_golden(obj_cicp))
            self.obj_st = [loader.oracle_obj_state_new(24) for _ in range(S)]
            self.n_mix = self.L.n_spk
        self.hoa = hoa
        if hoa:
            cfg, tables, self.M, self.nfc = hoa
            self.T = cfg[1]
            self.hoasp = [loader.OracleHoasp(cfg, tables) for _ in range(S)]
            self.nfc_st = [np.zeros((self.M.shape[1], 14), np.float32) for _ in range(S)]
            self.n_mix = self.M.shape[0]
        else:
            self.T = 0
        self.n_core = n_bed + n_obj + self.T
        if spectra:
            self.ov = [np.zeros((S * n, 1024), np.float32) for n in (n_bed, n_obj, self.T)]
        self.fc = [loader.OracleFc(fc_D) for _ in range(S)] if fc_D is not None else None
        self.n_out = self.n_mix if fc_D is None else fc_D.shape[0]
        self.bin = None
        if binaural is not None:
            filt = binaural if isinstance(binaural, list) else [binaural] * S
            self.bin = [loader.OracleBinaural(f["taps_direct"], f["taps_diffuse"], f["direct_fc"], f["diffuse_fc"],
                                              f["weight"]) for f in filt]
            self.B = self.bin[0].B
            self.acc = [[] for _ in range(S)]
            self.n_pcm = 2
        else:
            self.n_pcm = self.n_out
            self.lim = [loader.oracle_limiter_new(self.n_out, 48000, 1 if limiter == 1 else 0) for _ in range(S)] \
                if limiter else None

    def _fd(self, region, x, side5):
        """x [S, n, 1024] spectra, side5 [S, n, 5] -> time signals, overlap carried"""
        S, n = x.shape[:2]
        out = np.zeros((1, S * n, 1024), np.float32)
        assert self.o.oracle_fd_frm_dec_batch(np.ascontiguousarray(x.reshape(1, S * n, 1024)), self.ov[region], out,
                                              np.ascontiguousarray(side5.reshape(1, S * n, 5)), 1, S * n) == 0
        return out.reshape(S, n, 1024)

    def frame(self, bed=None, obj=None, hoa=None, fd_side5=None, obj_params=None, hoa_sides=None):
        """-> list over streams of the PCM bytes this frame produces (empty array = nothing yet)"""
        if self.spectra:
            sides = fd_side5 or {}
            bed = self._fd(0, bed, sides["bed"]) if bed is not None else None
            obj = self._fd(1, obj, sides["obj"]) if obj is not None else None
            hoa = self._fd(2, hoa, sides["hoa"]) if hoa is not None else None
        out = []
        for s in range(self.S):
            mix = None if bed is None else bed[s].astype(np.float32)
            if self.n_obj:
                y = loader.oracle_obj_render(self.L, self.obj_st[s], loader.oracle_obj_side(obj_params[s]), obj[s])
                mix = y if mix is None else (mix + y).astype(np.float32)
            if self.hoa:
                c = self.hoasp[s].process(hoa_sides[s:s + 1], hoa[s])
                y = loader.oracle_hoa_render(self.M, c, self.nfc, self.nfc_st[s] if self.nfc is not None else None)
                mix = y if mix is None else (y + mix).astype(np.float32)
            if self.fc:
                mix = self.fc[s].process(mix)
            if self.bin:
                self.acc[s].append(mix)
                if len(self.acc[s]) * 1024 < self.B:
                    out.append(np.zeros(0, np.uint8))
                    continue
                x = np.concatenate(self.acc[s], axis=1)
                self.acc[s] = []
                y = self.bin[s].block((x * np.float32(1.0 / 32768.0)).astype(np.float32))
                mix = (y * np.float32(32768.0)).astype(np.float32)
            elif self.lim:
                if self.gain is not None:
                    mix = (mix * np.float32(self.gain)).astype(np.float32)
                mix = loader.oracle_limiter_run(self.lim[s], mix)
            pcm, self.delay[s] = loader.oracle_samples_sat(mix, self.pcm_bits, self.ch_map, self.delay[s])
            out.append(pcm)
        return out


def core_buffer(bed=None, obj=None, hoa=None):
    """region-major host layout of mpeghb200_chain_frame.core"""
    parts = [np.ascontiguousarray(a, np.float32).reshape(-1) for a in (bed, obj, hoa) if a is not None]
    return np.concatenate(parts)


def side_buffer(fd_side5, regions=("bed", "obj", "hoa")):
    import fd_cases
    parts = [fd_cases.pack_side(fd_side5[r]).astype(np.uint32).reshape(-1) for r in regions if r in fd_side5]
    return np.concatenate(parts)
