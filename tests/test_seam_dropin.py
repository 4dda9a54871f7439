"""Library-level drop-in (BASELINE.json configs[0]): the reference testbench decodes the smoke_test_suite streams
through the reference's own ia_mpegh_dec_* API while libmpeghb200_seam.so (libmpegh_b200/seam/seam_interpose.c)
takes over FD synthesis, TNS filtering, the LTPF pass, the format converter (-cicp: retargeting) and the peak
limiter on the GPU.  The WAV files must be byte-identical to
what the unmodified reference writes (tests/golden/smoke_wav.json, made by tests/golden/make_smoke_golden.py;
the cicp1 entry equals the WAV the reference repository ships).
"""
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import seam_cases

ROOT = seam_cases.ROOT
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
TB = os.path.join(REF_DIR, "ia_mpeghd_testbench")
TB_DYN = os.path.join(REF_DIR, "ia_mpeghd_testbench_dyn")
SEAM = os.path.join(ROOT, "libmpegh_b200", "libmpeghb200_seam.so")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "smoke_wav.json")))


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip(f"{os.path.relpath(p, ROOT)} not built (needs /root/reference at build time)")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def _decode(binary, stream, extra, wav, preload=None):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
        env["MPEGHB200_SEAM_STATS"] = "1"
    return subprocess.run([binary, f"-ifile:{seam_cases.stream_path(stream)}", f"-ofile:{wav}", *extra], env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=600)


def test_golden_covers_cases():
    assert set(n for n, _, _ in seam_cases.CASES) == set(k for k in GOLD if not k.startswith("_"))
    assert GOLD["cicp1_mhas"]["sha256"] == GOLD["_shipped_ref_wav"]["sha256"]
    # MHAS and MP4 of the same stream must decode identically (reference README.md:276)
    assert GOLD["cicp1_mhas"]["sha256"] == GOLD["cicp1_mp4"]["sha256"]
    assert GOLD["cicp19_mhas"]["sha256"] == GOLD["cicp19_mp4"]["sha256"]


@pytest.mark.parametrize("name,stream,extra", seam_cases.CASES[:3], ids=[c[0] for c in seam_cases.CASES[:3]])
def test_reference_testbench_matches_golden(tmp_path, name, stream, extra):
    """CPU: the shared-library build of the reference testbench (the binary the seam is preloaded into)
    reproduces the goldens when nothing is interposed."""
    _need(TB_DYN, seam_cases.stream_path(stream))
    wav = str(tmp_path / "o.wav")
    r = _decode(TB_DYN, stream, extra, wav)
    assert r.returncode == 0, r.stderr[-400:]
    assert _sha(wav) == GOLD[name]["sha256"]


def test_seam_has_no_cpu_fallback(tmp_path):
    """Without a GPU the interposed decode must stop loudly, never fall back to the reference functions."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _need(TB_DYN, SEAM, seam_cases.stream_path("sine_1khz_000_cicp1.mhas"))
    r = _decode(TB_DYN, "sine_1khz_000_cicp1.mhas", [], str(tmp_path / "o.wav"), preload=SEAM)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,stream,extra", seam_cases.CASES, ids=[c[0] for c in seam_cases.CASES])
def test_seam_dropin_bit_exact(tmp_path, name, stream, extra):
    _need(TB_DYN, SEAM, seam_cases.stream_path(stream))
    wav = str(tmp_path / "o.wav")
    r = _decode(TB_DYN, stream, extra, wav, preload=SEAM)
    assert r.returncode == 0, r.stderr[-600:]
    m = re.search(r"fd_frm_dec gpu=(\d+) host\(LPD transition\)=(\d+) tns_filters gpu=(\d+) "
                  r"limiter_frames gpu=(\d+) ltpf_frames gpu=(\d+) format_conv_frames gpu=(\d+) "
                  r"mct_boxes gpu=(\d+) kernel_launches=(\d+)", r.stderr)
    assert m, r.stderr[-600:]
    fd_gpu, fd_host, _tns, lim, ltpf, fc, mct, launches = map(int, m.groups())
    print(f"{name}: seam counters {m.group(0)}")
    assert fd_gpu > 0 and fd_host == 0 and lim > 0 and ltpf == fd_gpu and launches >= 2 * fd_gpu + lim + fc + mct
    assert fc == 0  # none of the shipped streams reaches the format converter (see test_seam_format_converter)
    assert os.path.getsize(wav) == GOLD[name]["bytes"]
    assert _sha(wav) == GOLD[name]["sha256"], f"{name}: PCM differs from the reference decoder"
    # and against the unmodified static testbench run on this very box
    if os.path.exists(TB):
        wav2 = str(tmp_path / "r.wav")
        assert _decode(TB, stream, extra, wav2).returncode == 0
        assert _sha(wav2) == _sha(wav)


FC_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import loader
ref = loader.load_ref()
cicp_in, cicp_out, out_path = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
import ctypes
err = ctypes.c_int(0)
h = ref.ref_fc_create(cicp_in, cicp_out, 48000, ctypes.byref(err))
assert err.value == 0, err.value
info = np.zeros(2, np.int32)
ref.ref_fc_info(h, info)
n_in, n_out = int(info[0]), int(info[1])
rng = np.random.default_rng(100 * cicp_in + cicp_out)
outs = []
for f in range(4):
    x = rng.normal(0, 3000, (n_in, 1024)).astype(np.float32)
    y = np.zeros((n_out, 1024), np.float32)
    assert ref.ref_fc_process(h, x.reshape(-1), y.reshape(-1)) == 0
    outs.append(y)
np.save(out_path, np.stack(outs))
"""


@pytest.mark.gpu
@pytest.mark.parametrize("cicp_in,cicp_out", [(13, 6), (6, 2), (14, 19)])
def test_seam_format_converter(tmp_path, cicp_in, cicp_out):
    """The interposed impeghd_format_converter_process against the reference's own, both driven through the
    reference's API struct by oracle/ref_harness_fc.c (none of the shipped streams enables the converter): the same
    script runs once plainly (reference code) and once with the seam preloaded (GPU), outputs must be identical."""
    libref = os.path.join(REF_DIR, "libmpeghref.so")
    _need(libref, SEAM)
    script = tmp_path / "fc.py"
    script.write_text(FC_SCRIPT)
    outs = []
    for preload in (None, SEAM):
        env = dict(os.environ)
        if preload:
            env["LD_PRELOAD"], env["MPEGHB200_SEAM_STATS"] = preload, "1"
        out = str(tmp_path / ("gpu.npy" if preload else "ref.npy"))
        r = subprocess.run([sys.executable, str(script), ROOT, str(cicp_in), str(cicp_out), out], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        if preload:
            assert "format_conv_frames gpu=4" in r.stderr, r.stderr[-400:]
        outs.append(np.load(out))
    assert outs[0].shape == outs[1].shape and outs[0].shape[0] == 4
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


MCT_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import mct_cases
from oracle import loader
ref = loader.load_ref()
rng = np.random.default_rng(77)
outs = []
for _ in range(6):
    boxes = mct_cases.random_group(rng)
    x = (rng.standard_normal((24, 1024)) * 500).astype(np.float32)
    outs.append(mct_cases.apply_group(ref.ref_mct_process, x, boxes))
np.save(sys.argv[2], np.stack(outs))
"""


@pytest.mark.gpu
def test_seam_mct(tmp_path):
    """The interposed impeghd_mc_process against the reference's own (none of the shipped streams uses the
    multichannel coding tool): the harness of oracle/ref_harness_mct.c drives it once plainly and once with the
    seam preloaded."""
    libref = os.path.join(REF_DIR, "libmpeghref.so")
    _need(libref, SEAM)
    script = tmp_path / "mct.py"
    script.write_text(MCT_SCRIPT)
    outs = []
    for preload in (None, SEAM):
        env = dict(os.environ)
        if preload:
            env["LD_PRELOAD"], env["MPEGHB200_SEAM_STATS"] = preload, "1"
        out = str(tmp_path / ("gpu.npy" if preload else "ref.npy"))
        r = subprocess.run([sys.executable, str(script), ROOT, out], env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        if preload:
            m = re.search(r"mct_boxes gpu=(\d+)", r.stderr)
            assert m and int(m.group(1)) > 6, r.stderr[-400:]
        outs.append(np.load(out))
    assert outs[0].shape == outs[1].shape
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


OBJ_SUB_SCRIPT = r"""
import sys, ctypes, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import obj_cases
from oracle import loader
ref = loader.load_ref()
frame_len, out_path = int(sys.argv[2]), sys.argv[3]
err = ctypes.c_int(0)
h = ref.ref_obj_create(19, ctypes.byref(err))
assert err.value == 0
n_obj, n_sub = 7, 1024 // frame_len
rng = np.random.default_rng(frame_len)
outs = []
for f in range(5):
    params = obj_cases.random_params(rng, n_sub, n_obj)
    present = (rng.random(n_sub) < 0.7).astype(np.int32)
    x = obj_cases.object_audio(rng, 1, n_obj)[0]
    y = np.zeros((12, 1024), np.float32)
    used = np.zeros(n_sub, np.int32)
    assert ref.ref_obj_render_subframes(h, params.reshape(-1), present, n_obj, frame_len, x.reshape(-1), y.reshape(-1), used) == 0
    outs.append(y)
np.save(out_path, np.stack(outs))
"""


@pytest.mark.gpu
@pytest.mark.parametrize("frame_len", [256, 512])
def test_seam_oam_subframes(tmp_path, frame_len):
    """OAM sub-frames through the interposed impeghd_obj_renderer_dec (none of the shipped streams uses them): the
    sub-frame loop of oracle/ref_harness_obj.c drives the reference's function once plainly and once with the seam
    preloaded; 4 / 2 GPU renders per core frame, identical output."""
    libref = os.path.join(REF_DIR, "libmpeghref.so")
    _need(libref, SEAM)
    script = tmp_path / "objsub.py"
    script.write_text(OBJ_SUB_SCRIPT)
    outs = []
    for preload in (None, SEAM):
        env = dict(os.environ)
        if preload:
            env["LD_PRELOAD"], env["MPEGHB200_SEAM_STATS"] = preload, "1"
        out = str(tmp_path / ("gpu.npy" if preload else "ref.npy"))
        r = subprocess.run([sys.executable, str(script), ROOT, str(frame_len), out], env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        if preload:
            m = re.search(r"obj_render gpu=(\d+) host\(sub-frames\)=(\d+)", r.stderr)
            assert m and int(m.group(1)) == 5 * (1024 // frame_len) and int(m.group(2)) == 0, r.stderr[-400:]
        outs.append(np.load(out))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def _run_pair(tmp_path, script_text, args, counter_regex, expect):
    """Runs a harness script once plainly (reference code) and once with the seam preloaded (GPU); returns both arrays."""
    libref = os.path.join(REF_DIR, "libmpeghref.so")
    _need(libref, SEAM)
    script = tmp_path / "h.py"
    script.write_text(script_text)
    outs = []
    for preload in (None, SEAM):
        env = dict(os.environ)
        if preload:
            env["LD_PRELOAD"], env["MPEGHB200_SEAM_STATS"] = preload, "1"
        out = str(tmp_path / ("gpu.npy" if preload else "ref.npy"))
        r = subprocess.run([sys.executable, str(script), ROOT, *[str(a) for a in args], out], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        if preload:
            m = re.search(counter_regex, r.stderr)
            assert m and [int(x) for x in m.groups()] == expect, r.stderr[-400:]
        outs.append(np.load(out))
    return outs


HOA_REN_SCRIPT = r"""
import sys, ctypes, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import hoa_cases
from oracle import loader
ref = loader.load_ref()
cicp, order, lfe, radius, out_path = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), sys.argv[6]
err = ctypes.c_int(0)
h = ref.ref_hoa_ren_create(cicp, order, lfe, radius, 48000, ctypes.byref(err))
assert err.value == 0
M, nfc, n_out = loader.ref_hoa_ren_tables(ref, h, order)
rng = np.random.default_rng(cicp + order)
x = hoa_cases.hoa_signal(rng, 4, M.shape[1])
outs = []
for f in range(4):
    y = np.zeros((n_out, 1024), np.float32)
    assert ref.ref_hoa_ren_process(h, x[f].reshape(-1), M.shape[1], 1024, 1, y.reshape(-1)) == 0
    outs.append(y)
np.save(out_path, np.stack(outs))
"""


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(13, 4, 1, 1.5), (6, 3, 1, 0.0), (19, 4, 0, 0.8)])
def test_seam_hoa_renderer(tmp_path, cfg):
    """The interposed impeghd_hoa_ren_block_process (no shipped stream carries HOA): oracle/ref_harness_hoa.c designs the
    matrix / NFC filters with the reference's own init and drives the function, once plainly and once with the seam
    preloaded; the NFC memories travel through the reference struct, so four consecutive frames must agree."""
    outs = _run_pair(tmp_path, HOA_REN_SCRIPT, cfg, r"hoa_render gpu=(\d+) host=(\d+)", [4, 0])
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


RS_SCRIPT = r"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import loader
# the interposer reads the reference library's filter ROM: make its symbols visible (an application that links the
# reference library has them in the global scope anyway)
ctypes.CDLL(os.path.join(sys.argv[1], "oracle", "_ref", "libmpeghref.so"), mode=ctypes.RTLD_GLOBAL)
ref = loader.load_ref()
up, down, out_path = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
h = [ref.ref_rs_create(up, down) for _ in range(3)]
rng = np.random.default_rng(10 * up + down)
outs = []
for f in range(4):
    for c in range(3):
        x = rng.normal(0, 5000, 1024).astype(np.float32)
        y = np.zeros(1024 * up // down, np.float32)
        ref.ref_rs_process(h[c], x, y)
        outs.append(y)
np.save(out_path, np.stack(outs))
"""


@pytest.mark.gpu
@pytest.mark.parametrize("up,down", [(2, 1), (3, 1), (3, 2)])
def test_seam_resampler(tmp_path, up, down):
    """The interposed impeghd_resample (no shipped stream is below 48 kHz): three channels x four frames through
    oracle/ref_harness_rs.c, delay lines carried in ia_resampler_struct."""
    outs = _run_pair(tmp_path, RS_SCRIPT, (up, down), r"resample gpu=(\d+)", [12])
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


HOASP_SCRIPT = r"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import hoasp_cases
from oracle import loader
# the interposer reads the reference library's HOA ROM tables (see RS_SCRIPT)
ctypes.CDLL(os.path.join(sys.argv[1], "oracle", "_ref", "libmpeghref.so"), mode=ctypes.RTLD_GLOBAL)
ref = loader.load_ref()
cfg, out_path = hoasp_cases.CONFIGS[int(sys.argv[2])], sys.argv[3]
S, F = 3, 10
hs = []
for s in range(S):
    err = ctypes.c_int()
    hs.append(ref.ref_hoasp_create(hoasp_cases.cfg_array(cfg), ctypes.byref(err)))
    assert err.value == 0
n_coef = (cfg[0] + 1) ** 2
rng = np.random.default_rng(sum(cfg) + 3)
gens = [hoasp_cases.SideGen(cfg, np.random.default_rng(17 * s + 1), wild_exp=True) for s in range(S)]
outs = np.zeros((F, S, n_coef, 1024), np.float32)
for f in range(F):
    x = hoasp_cases.signals(rng, S, cfg[1])
    for s in range(S):                     # three decoders interleaved: the state is found by struct, not by call order
        sd = gens[s].next()
        assert ref.ref_hoasp_process(hs[s], sd.ctypes.data, x[s].reshape(-1), outs[f, s].reshape(-1)) == 0
np.save(out_path, outs)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_idx", range(6))
def test_seam_hoa_spatial(tmp_path, cfg_idx):
    """The four interposed stage functions of impeghd_hoa_spatial_process (inverse gain control, ambience synthesis,
    channel reassignment, predominant-sound synthesis; no shipped stream carries HOA): oracle/ref_harness_hoasp.c fills
    the frame parameters the reference's side-info parser would and calls them in the reference's order, three decoders
    interleaved over ten frames with independency-frame gain resets, plainly and with the seam preloaded."""
    outs = _run_pair(tmp_path, HOASP_SCRIPT, (cfg_idx,), r"hoa_spatial gpu=(\d+) host=(\d+)", [30, 0])
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


BIN_SCRIPT = r"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import bin_cases
from oracle import loader
ref = loader.load_ref()
n_in, length, nb, feed, out_path = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
rng = np.random.default_rng(n_in * 7 + length + nb)
f = bin_cases.with_cutoffs(rng, bin_cases.brir_set(rng, n_in, length, nb))
hs = []
for s in range(2):   # two renderers with the same BRIR set (one GPU renderer, two state slots), interleaved
    err = ctypes.c_int()
    td, tdf = f["taps_direct"], f["taps_diffuse"]
    hs.append(ref.ref_bin_create(n_in, length, nb, td.reshape(-1), np.ascontiguousarray(tdf).reshape(-1),
                                 f["direct_fc"].reshape(-1).copy(), np.ascontiguousarray(f["diffuse_fc"]).reshape(-1),
                                 f["weight"], 0, ctypes.byref(err)))
    assert err.value == 0
info = np.zeros(8, np.int32)
ref.ref_bin_info(hs[0], info)
B = int(info[2])
T = 4
x = [bin_cases.signal(rng, T, n_in, B) for _ in range(2)]
streams = [np.concatenate(list(xx), axis=1) for xx in x]   # [n_in, T*B]
outs = [[], []]
buf = np.zeros((2, 32768), np.float32)
pos = 0
while pos < T * B:
    n = min(feed, T * B - pos)
    for s in range(2):
        got = ref.ref_bin_process(hs[s], np.ascontiguousarray(streams[s][:, pos:pos + n]).reshape(-1), n, buf.reshape(-1), 32768)
        if got:
            outs[s].append(buf[:, :got].copy())
    pos += n
res = np.stack([np.concatenate(o, axis=1) for o in outs])
assert res.shape[2] >= (T - 1) * B
np.save(out_path, res)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,blocks", [((24, 4096, 1, 1024), 8), ((5, 1024, 1, 1024), 8), ((21, 4096, 0, 1024), 8),
                                        ((11, 3000, 1, 1000), 8), ((3, 700, 1, 2500), 8),
                                        ((4, 300, 1, 1024), 8), ((2, 5000, 1, 1024), 8)])
def test_seam_binaural(tmp_path, cfg, blocks):
    """The interposed impeghd_binaural_struct_process_ld_ola (no shipped stream selects binaural rendering): the
    reference's own init designs the filter spectra, the interposer takes them from ia_binaural_ren_pers_mem_str as
    they are.  Fed 1024 samples per call like the decoder does, and with call lengths that leave partial blocks
    (1000) or complete two blocks in one call (2500 at B = 1024); two renderers interleaved share one GPU renderer."""
    outs = _run_pair(tmp_path, BIN_SCRIPT, cfg, r"binaural_blocks gpu=(\d+) host\(calls\)=(\d+)", [blocks, 0])
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


CPE_SCRIPT = r"""
import sys, os, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import stereo_cases, tns_cases
from oracle import loader
which, out_path = sys.argv[2], sys.argv[3]
ref = loader.load_ref() if which == "stock" else loader.load_ref_ns()
outs = []
for seed in range(6):
    rng = np.random.default_rng(9000 + seed)
    frames = stereo_cases.random_pair_sequence(rng, 8)
    coef = rng.normal(0, 800, (8, 2, 1024)).astype(np.float32)
    h = ref.ref_cpe_create(stereo_cases.SAMPLING_RATE_IDX, tns_cases.TNS_MAX_LONG, tns_cases.TNS_MAX_SHORT)
    pcm, spec = np.zeros((8, 2, 1024), np.float32), np.zeros((8, 2, 1024), np.float32)
    for f, fr in enumerate(frames):
        err = ref.ref_cpe_process(h, fr["sd"].tobytes(), fr["grp"], fr["tns_on_lr"],
                                  np.ascontiguousarray(fr["n_filt"].reshape(-1)),
                                  np.ascontiguousarray(fr["filt"].reshape(-1)), fr["max_sfb"], coef[f, 0], coef[f, 1],
                                  pcm[f, 0], pcm[f, 1], spec[f, 0], spec[f, 1])
        assert err == 0, err
    ref.ref_cpe_destroy(h)
    outs.append(np.stack([pcm, spec]))
np.save(out_path, np.stack(outs))
"""


def _cpe_tool_counts():
    import stereo_cases
    ms = cplx = 0
    for seed in range(6):
        rng = np.random.default_rng(9000 + seed)
        for fr in stereo_cases.random_pair_sequence(rng, 8):
            tool = int(fr["sd"]["tool"])
            ms += tool in (1, 2)
            cplx += tool == 3
    return ms, cplx


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["stock", "ns"])
def test_seam_stereo_tools(tmp_path, which):
    """ia_core_coder_data_process of the unmodified reference on a channel pair (oracle/ref_harness_cpe.c fills
    ia_usac_data_struct the way the parser does): M/S and complex prediction, long and short blocks with window groups,
    TNS before or after the stereo tools, spectrum save, FD synthesis.  With the seam preloaded impeghd_ms_stereo
    becomes a GPU call; ia_core_coder_cplx_pred_upmixing too in the reference build that gives it external linkage
    ("ns"), while against the stock library complex prediction stays in the reference.  PCM and saved spectra agree
    bit for bit either way."""
    if which == "ns":
        _need(os.path.join(REF_DIR, "libmpeghref_ns.so"))
    ms, cplx = _cpe_tool_counts()
    assert ms > 5 and cplx > 5
    outs = _run_pair(tmp_path, CPE_SCRIPT, (which,), r"stereo_tools gpu=(\d+) host=(\d+)",
                     [ms + (cplx if which == "ns" else 0), 0])
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


IGF_SCRIPT = r"""
import sys, os, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import igf_cases
import make_igf_golden as mg
from oracle import loader
ref = loader.load_ref()
mode, out_path = sys.argv[2], sys.argv[3]
outs, seeds = [], []
for ci, cfg in enumerate(igf_cases.CONFIGS):
    h, grid = mg.reference_grid(cfg)
    rng = np.random.default_rng((12000 if mode == "mono" else 14000) + ci)
    for k in range(24):
        wt = int(rng.random() < 0.3)
        if mode == "mono":
            sd, coef, tiles = igf_cases.random_frame(rng, cfg, grid, win_type=wt)
            y = coef.copy()
            seeds.append([ref.ref_igf_process(h, sd.tobytes(), y, tiles), 0])
            outs.append(np.stack([y, y]))
        else:
            sd_l, sd_r, mask, coef, tiles = igf_cases.random_pair_frame(rng, cfg, grid, win_type=wt)
            ms_after = int(rng.integers(2))
            y = coef.copy()
            sv = np.zeros(2, np.uint32)
            ref.ref_igf_stereo_process(h, sd_l.tobytes(), sd_r.tobytes(), igf_cases.mask_reference_layout(mask, sd_l),
                                       ms_after, y[0], y[1], np.ascontiguousarray(tiles[0]).reshape(-1),
                                       np.ascontiguousarray(tiles[1]).reshape(-1), sv)
            seeds.append([int(sv[0]), int(sv[1])])
            outs.append(y)
    ref.ref_igf_destroy(h)
res = np.concatenate([np.stack(outs).reshape(len(outs), -1),
                      np.array(seeds, np.uint32).view(np.float32).reshape(len(outs), 2)], axis=1)
np.save(out_path, res)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["mono", "stereo"])
def test_seam_igf(tmp_path, mode):
    """The interposed IGF entry points of the unmodified reference through oracle/ref_harness_cpe.c: the mono path
    (ia_core_coder_igf_mono + ia_core_coder_igf_tnf, the side record filled from igf_config / igf_dec_data / seed_value
    of the real ia_usac_data_struct, TNF spectrum and mask carried in the struct) and the joint-stereo path
    (ia_core_coder_igf_apply_stereo, then impeghd_ms_stereo over the IGF range on the spectra and on the TNF spectra,
    then ia_core_coder_igf_tnf per channel), long and short blocks, all configurations of igf_cases.CONFIGS; spectra
    and noise seeds agree bit for bit."""
    rx = r"igf mono=(\d+) tnf=(\d+) stereo=(\d+) host=(\d+)"
    libref = os.path.join(REF_DIR, "libmpeghref.so")
    _need(libref, SEAM)
    script = tmp_path / "h.py"
    script.write_text(IGF_SCRIPT)
    outs = []
    for preload in (None, SEAM):
        env = dict(os.environ)
        if preload:
            env["LD_PRELOAD"], env["MPEGHB200_SEAM_STATS"] = preload, "1"
        out = str(tmp_path / ("gpu.npy" if preload else "ref.npy"))
        r = subprocess.run([sys.executable, str(script), ROOT, mode, out], env=env, capture_output=True, text=True,
                           timeout=900)
        assert r.returncode == 0, r.stderr[-800:]
        if preload:
            m = re.search(rx, r.stderr)
            assert m, r.stderr[-400:]
            mono, tnf, st, host = (int(x) for x in m.groups())
            import igf_cases
            n = 24 * len(igf_cases.CONFIGS)
            print(mode, dict(mono=mono, tnf=tnf, stereo=st, host=host))
            assert host == 0
            if mode == "mono":
                assert mono == n and st == 0
            else:
                assert 0 < st <= n and mono == 0  # pairs with both channels silent never reach the tool
        outs.append(np.load(out))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


DRC_SCRIPT = r"""
import sys, os, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import drc_cases
from oracle import loader
mode, out_path = sys.argv[2], sys.argv[3]
outs = []
n_channels = 0
if mode == "subband":
    ref = loader.load_ref()
    for seed in range(12):
        rng = np.random.default_rng(600 + seed)
        fr = drc_cases.random_frame(rng)
        re, im = drc_cases.run_oracle(ref.ref_drc_gains_subband, fr, (int(rng.integers(0, 2)), int(rng.integers(0, 400))))
        pad = np.zeros((2, 16, 1024), np.float32)
        pad[0, :fr["n_ch"]], pad[1, :fr["n_ch"]] = re, im
        outs.append(pad)
        n_channels += fr["n_ch"]
elif mode == "stft_chain":
    ref = loader.load_ref()
    for seed in range(4):
        rng = np.random.default_rng(800 + seed)
        n_ch = int(rng.integers(1, 13))
        y = drc_cases.stft_chain(None, None, rng, 3, n_ch, ref=ref)   # t2f -> gains of 1..2 sets -> f2t, 3 frames
        pad = np.zeros((3, 12, 1024), np.float32)
        pad[:, :n_ch] = y
        outs.append(pad)
        n_channels += 3 * n_ch
elif mode == "td_whole":
    # impd_drc_td_process as the decoder calls it, STOCK library; the interposer calls the reference's impd_drc_get_gain
    import ctypes
    ctypes.CDLL(os.path.join(sys.argv[1], "oracle", "_ref", "libmpeghref.so"), mode=ctypes.RTLD_GLOBAL)
    ref = loader.load_ref()
    for seed in range(8):
        rng = np.random.default_rng(700 + seed)
        setup = drc_cases.random_td_frame_setup(rng)
        frames = drc_cases.td_frames(rng, setup, 3)
        y = drc_cases.td_reference_whole(ref, setup, frames)
        pad = np.zeros((3, 12, 1024), np.float32)
        pad[:, :setup["n_ch"]] = y
        outs.append(pad)
        n_channels += 3 * setup["n_ch"]
else:
    # the interposers look the reference's own functions up behind themselves (see RS_SCRIPT)
    import ctypes
    ctypes.CDLL(os.path.join(sys.argv[1], "oracle", "_ref", "libmpeghref_ns.so"), mode=ctypes.RTLD_GLOBAL)
    ref = loader.load_ref_ns()
    for seed in range(8):
        rng = np.random.default_rng(700 + seed)
        setup = drc_cases.random_td_frame_setup(rng)
        frames = drc_cases.td_frames(rng, setup, 3)
        y = drc_cases.td_reference(ref, setup, frames)
        pad = np.zeros((3, 12, 1024), np.float32)
        pad[:, :setup["n_ch"]] = y
        outs.append(pad)
        n_channels += 3 * setup["n_ch"]
np.save(out_path, np.stack(outs))
print("CHANNELS", n_channels)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["subband", "stft_chain", "td_whole", "td_ns", "td_host"])
def test_seam_drc(tmp_path, mode):
    """Application of the DRC gains through the unmodified reference functions (oracle/ref_harness_drc.c,
    ref_harness_ns.c fill ia_drc_gain_dec_struct / ia_drc_instructions_struct the way the gain decoder leaves them):
    the chain of ia_core_coder_decode_main.c:2128-2141 -- impeghd_domain_switcher_process time -> STFT256, the gains of
    one or two selected sets, impeghd_domain_switcher_process back, three frames with the carried input slot and output
    overlap ("stft_chain"); impd_drc_td_process as the decoder calls it, with the STOCK library ("td_whole": 2 / 3 / 4-band
    banks, phase-alignment cascades, three frames, the reference's own impd_drc_get_gain in front); impd_drc_apply_gains_subband (STFT256 domain: single- and multi-band groups, channels without a group, both delay
    modes, gain delays) with the stock library; impd_drc_filter_banks_process + impd_drc_apply_gains_and_add (time domain:
    2 / 3 / 4-band Linkwitz-Riley banks, phase-alignment cascades, three frames with carried filter memories) with the
    reference build that exports the second, `static`, function; against the stock library that pair is forwarded
    to the reference, which "td_host" forces (MPEGHB200_SEAM_DRC_TD_HOST=1) to cover the forwarding path."""
    libref = os.path.join(REF_DIR, "libmpeghref.so")
    _need(libref, SEAM)
    if mode in ("td_ns", "td_host"):
        _need(os.path.join(REF_DIR, "libmpeghref_ns.so"))
    script = tmp_path / "h.py"
    script.write_text(DRC_SCRIPT)
    outs, n_channels = [], 0
    for preload in (None, SEAM):
        env = dict(os.environ)
        if preload:
            env["LD_PRELOAD"], env["MPEGHB200_SEAM_STATS"] = preload, "1"
            if mode == "td_host":
                env["MPEGHB200_SEAM_DRC_TD_HOST"] = "1"
        out = str(tmp_path / ("gpu.npy" if preload else "ref.npy"))
        r = subprocess.run([sys.executable, str(script), ROOT, mode, out], env=env, capture_output=True, text=True,
                           timeout=900)
        assert r.returncode == 0, r.stderr[-800:]
        n_channels = int(re.search(r"CHANNELS (\d+)", r.stdout).group(1))
        if preload:
            m = re.search(r"drc_channels subband=(\d+) time_domain=(\d+) host\(calls\)=(\d+) stft=(\d+)", r.stderr)
            if mode == "td_host" and not m:  # nothing reached the GPU, so no context and no statistics line
                outs.append(np.load(out))
                continue
            assert m, r.stderr[-400:]
            sb, td, host, stft = (int(x) for x in m.groups())
            print(mode, dict(subband=sb, time_domain=td, host=host, stft=stft, channels=n_channels))
            if mode == "subband":
                assert sb == n_channels and td == 0 and host == 0
            elif mode == "stft_chain":  # both halves of the domain switcher per channel-frame, gains of 1..2 sets
                assert stft == 2 * n_channels and n_channels <= sb <= 2 * n_channels and host == 0
            elif mode in ("td_ns", "td_whole"):
                assert td == n_channels and host == 0
            else:
                assert td == 0 and host > 0  # the pair stays in the reference
        outs.append(np.load(out))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
