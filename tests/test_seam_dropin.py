"""Library-level drop-in (BASELINE.json configs[0]): the reference testbench decodes the smoke_test_suite streams
through the reference's own ia_mpegh_dec_* API while libmpeghb200_seam.so (libmpegh_b200/seam/seam_interpose.c)
takes over FD synthesis, TNS filtering, the LTPF pass and the peak limiter on the GPU.  The WAV files must be byte-identical to
what the unmodified reference writes (tests/golden/smoke_wav.json, made by tests/golden/make_smoke_golden.py;
the cicp1 entry equals the WAV the reference repository ships).
"""
import hashlib
import json
import os
import re
import subprocess

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
                  r"limiter_frames gpu=(\d+) ltpf_frames gpu=(\d+) kernel_launches=(\d+)", r.stderr)
    assert m, r.stderr[-600:]
    fd_gpu, fd_host, _tns, lim, ltpf, launches = map(int, m.groups())
    assert fd_gpu > 0 and fd_host == 0 and lim > 0 and ltpf == fd_gpu and launches >= 2 * fd_gpu + lim
    assert os.path.getsize(wav) == GOLD[name]["bytes"]
    assert _sha(wav) == GOLD[name]["sha256"], f"{name}: PCM differs from the reference decoder"
    # and against the unmodified static testbench run on this very box
    if os.path.exists(TB):
        wav2 = str(tmp_path / "r.wav")
        assert _decode(TB, stream, extra, wav2).returncode == 0
        assert _sha(wav2) == _sha(wav)
