#!/usr/bin/env python3
"""Library-level goldens: sha256 of the WAV files the UNMODIFIED reference testbench (oracle/_ref/ia_mpeghd_testbench,
built by oracle/Makefile from /root/reference in place) writes for the smoke_test_suite streams.

Run in the build container (needs oracle/_ref and /root/reference):  python tests/golden/make_smoke_golden.py
The one WAV the reference ships (smoke_test_suite/ref/sine_1khz_000_cicp1.wav) is hashed too and must equal the
testbench's output for that stream (checked here and in tests/test_seam_dropin.py).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import seam_cases  # noqa: E402


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    tb = os.path.join(ROOT, "oracle", "_ref", "ia_mpeghd_testbench")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, stream, extra in seam_cases.CASES:
            wav = os.path.join(tmp, name + ".wav")
            subprocess.run([tb, f"-ifile:{seam_cases.stream_path(stream)}", f"-ofile:{wav}", *extra], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[name] = {"stream": stream, "args": extra, "bytes": os.path.getsize(wav), "sha256": sha(wav)}
    shipped = "/root/reference/smoke_test_suite/ref/sine_1khz_000_cicp1.wav"
    out["_shipped_ref_wav"] = {"file": "smoke_test_suite/ref/sine_1khz_000_cicp1.wav", "sha256": sha(shipped)}
    assert out["cicp1_mhas"]["sha256"] == out["_shipped_ref_wav"]["sha256"]
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "smoke_wav.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
