"""The drop-in library libmpeghgpu.so (libmpegh_b200/seam/seam_core.c): the reference's own API in front of the GPU
signal path, with the additive mpeghb200_dec_execute_batch.

  * the reference testbench LINKED AGAINST libmpeghgpu.so (not preloaded) decodes the smoke_test_suite cases byte-
    identically, with FD synthesis, LTPF, the object renderer, the limiter and the PCM packer on the GPU;
  * mpeghb200_batch_decode advances 256 handles per mpeghb200_dec_execute_batch call: every handle's PCM equals
    handle 0's, and handle 0's WAV equals the reference decoder's (tests/golden/smoke_wav.json);
  * a batch that mixes different streams (1, 6, 10 and 12 output channels) still decodes every stream exactly.
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
TB_GPU = os.path.join(REF_DIR, "ia_mpeghd_testbench_gpu")
TOOL = os.path.join(ROOT, "libmpegh_b200", "mpeghb200_batch_decode")
GPULIB = os.path.join(ROOT, "libmpegh_b200", "libmpeghgpu.so")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "smoke_wav.json")))


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip(f"{os.path.relpath(p, ROOT)} not built (needs /root/reference at build time)")


def _env():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = REF_DIR + os.pathsep + env.get("LD_LIBRARY_PATH", "")  # where the reference decoder lives
    env["MPEGHB200_SEAM_STATS"] = "1"
    return env


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def _counters(stderr):
    m = re.search(r"fd_frm_dec gpu=(\d+) host\(LPD transition\)=(\d+) tns_filters gpu=(\d+) limiter_frames gpu=(\d+) "
                  r"ltpf_frames gpu=(\d+) format_conv_frames gpu=(\d+) mct_boxes gpu=(\d+) kernel_launches=(\d+) "
                  r"obj_render gpu=(\d+) host\(sub-frames\)=(\d+) pcm_pack gpu=(\d+) batch_calls=(\d+) "
                  r"backend_rounds=(\d+) max_batch=(\d+)", stderr)
    assert m, stderr[-600:]
    keys = ("fd_frm_dec gpu", "fd host", "tns_filters gpu", "limiter_frames gpu", "ltpf_frames gpu",
            "format_conv_frames gpu", "mct_boxes gpu", "kernel_launches", "obj_render gpu", "host(sub-frames)",
            "pcm_pack gpu", "batch_calls", "backend_rounds", "max_batch")
    return dict(zip(keys, map(int, m.groups())))


def test_dropin_exports_reference_api():
    """CPU: the five reference entry points plus the batch call are exported, and the reference decoder is a
    dependency found by soname only (no path baked in)."""
    _need(GPULIB)
    syms = subprocess.run(["nm", "-D", "--defined-only", GPULIB], capture_output=True, text=True).stdout
    for s in ("ia_mpegh_dec_get_lib_id_strings", "ia_mpegh_dec_create", "ia_mpegh_dec_init", "ia_mpegh_dec_execute",
              "ia_mpegh_dec_delete", "mpeghb200_dec_execute_batch"):
        assert re.search(rf"\bT {s}\b", syms), s
    dyn = subprocess.run(["readelf", "-d", GPULIB], capture_output=True, text=True).stdout
    assert "libmpeghb200.so" in dyn and re.search(r"NEEDED.*libmpeghref", dyn)
    assert "oracle" not in dyn


def test_dropin_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _need(TOOL)
    r = subprocess.run([TOOL, f"-o:{tmp_path}/x", seam_cases.stream_path("sine_1khz_000_cicp1.mhas")], env=_env(),
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,stream,extra", seam_cases.CASES, ids=[c[0] for c in seam_cases.CASES])
def test_testbench_linked_against_dropin(tmp_path, name, stream, extra):
    _need(TB_GPU, GPULIB, seam_cases.stream_path(stream))
    wav = str(tmp_path / "o.wav")
    r = subprocess.run([TB_GPU, f"-ifile:{seam_cases.stream_path(stream)}", f"-ofile:{wav}", *extra], env=_env(),
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-600:]
    c = _counters(r.stderr)
    print(name, c)
    assert c["fd_frm_dec gpu"] > 0 and c["limiter_frames gpu"] > 0 and c["pcm_pack gpu"] > 0
    if "cicp1" not in name:
        assert c["obj_render gpu"] > 0 and c["host(sub-frames)"] == 0  # the streams exist to exercise the renderer
    assert _sha(wav) == GOLD[name]["sha256"], f"{name}: PCM differs from the reference decoder"


@pytest.mark.gpu
@pytest.mark.parametrize("stream,gold", [("sine_1khz_000_cicp1.mhas", "cicp1_mhas"), ("sine_1khz_cicp6.mhas", "cicp6_mhas"),
                                         ("sine_1khz_cicp16.mhas", "cicp16_mhas"), ("sine_1khz_cicp19.mp4", "cicp19_mp4")])
def test_batch_of_256_copies(tmp_path, stream, gold):
    """256 handles of one stream, one mpeghb200_dec_execute_batch per frame: a launch serves all 256."""
    _need(TOOL, seam_cases.stream_path(stream))
    r = subprocess.run([TOOL, "-n:256", "-write:1", f"-o:{tmp_path}/b", seam_cases.stream_path(stream)], env=_env(),
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    c = _counters(r.stderr)
    print(gold, {k: res[k] for k in ("handles", "frames", "run_s", "audio_s_per_wall_s")}, c)
    assert res["handles"] == 256 and c["max_batch"] == 256
    hashes = {s["fnv1a64"] for s in res["streams"]}
    assert len(hashes) == 1, "handles of the same stream produced different PCM"
    assert all(s["bytes"] == GOLD[gold]["bytes"] - 44 for s in res["streams"])
    assert _sha(f"{tmp_path}/b_0.wav") == GOLD[gold]["sha256"]
    # batching is real: far fewer back-end rounds (= launches of a stage) than stage calls
    assert c["backend_rounds"] * 100 < c["fd_frm_dec gpu"] + c["limiter_frames gpu"] + c["pcm_pack gpu"]


@pytest.mark.gpu
def test_batch_of_mixed_streams(tmp_path):
    """Four different streams x 8 copies in ONE batch: the back-ends split a round by channel count / layout."""
    streams = ["sine_1khz_000_cicp1.mhas", "sine_1khz_cicp6.mhas", "sine_1khz_cicp16.mhas", "sine_1khz_cicp19.mhas"]
    golds = ["cicp1_mhas", "cicp6_mhas", "cicp16_mhas", "cicp19_mhas"]
    _need(TOOL, *[seam_cases.stream_path(s) for s in streams])
    r = subprocess.run([TOOL, "-n:8", "-write:32", f"-o:{tmp_path}/m", *[seam_cases.stream_path(s) for s in streams]],
                       env=_env(), capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["handles"] == 32
    for k, g in enumerate(golds):
        for j in range(8):
            assert _sha(f"{tmp_path}/m_{8 * k + j}.wav") == GOLD[g]["sha256"], (g, j)


@pytest.mark.gpu
def test_handle_at_a_time_equals_batch(tmp_path):
    _need(TOOL)
    s = seam_cases.stream_path("sine_1khz_cicp19.mhas")
    out = []
    for b in (0, 1):
        r = subprocess.run([TOOL, "-n:3", f"-batch:{b}", "-write:0", f"-o:{tmp_path}/x", "-pcmsz:24", s], env=_env(),
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-800:]
        out.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert [x["fnv1a64"] for x in out[0]["streams"]] == [x["fnv1a64"] for x in out[1]["streams"]]


@pytest.mark.gpu
def test_batch_with_a_truncated_stream(tmp_path):
    """Handles of one batch end at different times and one of them ends on a damaged frame: a copy of the 6-channel stream
    cut in the middle of a frame decodes next to full streams.  The full streams stay byte-identical to the reference
    decoder; the truncated handle delivers a prefix of its reference PCM and its failure stays its own."""
    full6, full19 = seam_cases.stream_path("sine_1khz_cicp6.mhas"), seam_cases.stream_path("sine_1khz_cicp19.mhas")
    _need(TOOL, full6, full19)
    cut = tmp_path / "cut.mhas"
    data = open(full6, "rb").read()
    cut.write_bytes(data[:int(len(data) * 0.43) + 7])
    r = subprocess.run([TOOL, "-n:3", "-write:9", f"-o:{tmp_path}/t", full6, str(cut), full19], env=_env(),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["handles"] == 9
    for j in range(3):
        assert _sha(f"{tmp_path}/t_{j}.wav") == GOLD["cicp6_mhas"]["sha256"], j
        assert _sha(f"{tmp_path}/t_{6 + j}.wav") == GOLD["cicp19_mhas"]["sha256"], j
    # what the unmodified reference testbench makes of the same damaged file (it stops with "insufficient input bytes")
    tb = os.path.join(REF_DIR, "ia_mpeghd_testbench")
    _need(tb)
    subprocess.run([tb, f"-ifile:{cut}", f"-ofile:{tmp_path}/cut_ref.wav"], stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, timeout=300)
    ref_cut = open(f"{tmp_path}/cut_ref.wav", "rb").read()[44:]
    full = open(f"{tmp_path}/t_0.wav", "rb").read()[44:]
    assert 0.3 * len(full) < len(ref_cut) < 0.6 * len(full) and ref_cut == full[:len(ref_cut)]
    for j in range(3):
        st = res["streams"][3 + j]
        got = open(f"{tmp_path}/t_{3 + j}.wav", "rb").read()[44:]
        assert st["bytes"] == len(got)
        n = min(len(got), len(ref_cut))
        assert got[:n] == ref_cut[:n], j
        assert abs(len(got) - len(ref_cut)) <= 6 * 1024 * 2, (len(got), len(ref_cut))  # at most the damaged frame apart


@pytest.mark.gpu
def test_one_application_thread_per_handle(tmp_path):
    """The reference's threading contract (one handle per thread is safe) holds for the drop-in: twelve
    threads, each decoding its own handle to the end through the unchanged ia_mpegh_dec_execute at the same time
    (four different streams x 3), produce the reference decoder's PCM."""
    streams = ["sine_1khz_000_cicp1.mhas", "sine_1khz_cicp6.mhas", "sine_1khz_cicp16.mhas", "sine_1khz_cicp19.mp4"]
    golds = ["cicp1_mhas", "cicp6_mhas", "cicp16_mhas", "cicp19_mp4"]
    _need(TOOL, *[seam_cases.stream_path(s) for s in streams])
    r = subprocess.run([TOOL, "-n:3", "-batch:2", "-write:12", f"-o:{tmp_path}/th",
                        *[seam_cases.stream_path(s) for s in streams]], env=_env(), capture_output=True, text=True,
                       timeout=1800)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["handles"] == 12 and res["batch"] == 2
    for k, g in enumerate(golds):
        for j in range(3):
            assert _sha(f"{tmp_path}/th_{3 * k + j}.wav") == GOLD[g]["sha256"], (g, j)


@pytest.mark.gpu
def test_handles_come_and_go(tmp_path):
    """A server's handles are created and deleted all the time, and the allocator hands the next handle the addresses of
    the last one: three create / decode / delete cycles of 8 + 8 handles in one process.  Every cycle decodes the
    reference's PCM (the device state keyed by addresses inside a deleted handle is dropped with it, so the next handle
    at that address starts from fresh limiter / object-renderer state), and the state blocks are freed with the handles."""
    streams = ["sine_1khz_cicp19.mhas", "sine_1khz_cicp6.mhas"]
    _need(TOOL, *[seam_cases.stream_path(s) for s in streams])
    r = subprocess.run([TOOL, "-n:8", "-cycles:3", "-write:16", f"-o:{tmp_path}/c",
                        *[seam_cases.stream_path(s) for s in streams]], env=_env(), capture_output=True, text=True,
                       timeout=1800)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [json.loads(l) for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 3
    for res in lines:
        assert res["handles"] == 16
        assert {s["fnv1a64"] for s in res["streams"][:8]} == {lines[0]["streams"][0]["fnv1a64"]}
        assert {s["fnv1a64"] for s in res["streams"][8:]} == {lines[0]["streams"][8]["fnv1a64"]}
    assert _sha(f"{tmp_path}/c_0.wav") == GOLD["cicp19_mhas"]["sha256"]
    assert _sha(f"{tmp_path}/c_8.wav") == GOLD["cicp6_mhas"]["sha256"]
    m = re.search(r"device state blocks created=(\d+) freed with their handles=(\d+)", r.stderr)
    assert m, r.stderr[-400:]
    made, freed = int(m.group(1)), int(m.group(2))
    assert made >= 3 and freed == made, (made, freed)


@pytest.mark.gpu
def test_configuration_change_mid_stream(tmp_path):
    """Three MHAS streams back to back in one file (12-, 6-, 12-channel output): every new configuration packet makes
    the reference flush and re-initialise the decoder in place inside ia_mpegh_dec_execute
    (ia_core_coder_decode_main.c:2519-2533 -> ia_core_coder_decoder_flush_api).  The drop-in drops the device state of the
    handle at that point, so the third part -- same structs, same addresses, same channel count as the first -- starts
    from fresh limiter and object-renderer state like the reference does: byte-identical WAV."""
    parts = ["sine_1khz_cicp19.mhas", "sine_1khz_cicp6.mhas", "sine_1khz_cicp19.mhas"]
    tb = os.path.join(REF_DIR, "ia_mpeghd_testbench")
    _need(TB_GPU, tb, *[seam_cases.stream_path(s) for s in parts])
    cat = tmp_path / "cat.mhas"
    cat.write_bytes(b"".join(open(seam_cases.stream_path(s), "rb").read() for s in parts))
    subprocess.run([tb, f"-ifile:{cat}", f"-ofile:{tmp_path}/ref.wav"], stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, timeout=600)
    ref_bytes = os.path.getsize(f"{tmp_path}/ref.wav")
    assert ref_bytes > 2 * (GOLD["cicp19_mhas"]["bytes"] - 44)  # the reference really decoded all three parts
    r = subprocess.run([TB_GPU, f"-ifile:{cat}", f"-ofile:{tmp_path}/gpu.wav"], env=_env(), stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-600:]
    c = _counters(r.stderr)
    assert c["obj_render gpu"] > 1000 and c["limiter_frames gpu"] > 1000
    assert _sha(f"{tmp_path}/gpu.wav") == _sha(f"{tmp_path}/ref.wav")


def test_integration_doc_lists_what_is_exported():
    """CPU: every reference function INTEGRATION.md's table names as interposed is a defined export of libmpeghgpu.so, and
    every reference-named export of the library appears in the document (no drift between the two)."""
    _need(GPULIB)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    table = doc[doc.index("| interposed reference function"):doc.index("**Batching without a source patch.**")]
    named = set()
    for row in table.splitlines()[2:]:
        cells = row.split("|")
        if len(cells) > 2:
            named |= set(re.findall(r"`((?:ia_core_coder|impeghd|impd_drc|ia_mpegh)_[a-z0-9_]+)`", cells[1]))
    syms = subprocess.run(["nm", "-D", "--defined-only", GPULIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT ((?:ia_core_coder|impeghd|impd_drc)_[a-z0-9_]+)\b", syms))
    callers = {"impeghd_hoa_spatial_process"}  # named in the table only as the caller of the four interposed stages
    assert len(named - callers) >= 26
    missing = {n for n in named - callers if n not in exported}
    assert not missing, f"named in INTEGRATION.md but not exported: {sorted(missing)}"
    helpers = {"ia_core_coder_decoder_flush_api"}  # lifecycle hook, described in the tests section of the document
    undocumented = {e for e in exported if e not in named and e not in helpers and e not in doc}
    assert not undocumented, f"exported but not in INTEGRATION.md: {sorted(undocumented)}"
