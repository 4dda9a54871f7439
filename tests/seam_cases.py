"""smoke_test_suite decode cases shared by the golden generator and the drop-in tests.
(name, stream file under oracle/_ref/smoke/, extra testbench arguments)"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("cicp1_mhas", "sine_1khz_000_cicp1.mhas", []),
    ("cicp1_mp4", "sine_1khz_000_cicp1.mp4", []),
    ("cicp6_mhas", "sine_1khz_cicp6.mhas", []),
    ("cicp16_mhas", "sine_1khz_cicp16.mhas", []),
    ("cicp19_mhas", "sine_1khz_cicp19.mhas", []),
    ("cicp19_mp4", "sine_1khz_cicp19.mp4", []),
    ("cicp6_pcm24", "sine_1khz_cicp6.mhas", ["-pcmsz:24"]),
    ("cicp16_pcm32", "sine_1khz_cicp16.mhas", ["-pcmsz:32"]),
    # loudspeaker retargeting: the format converter (interposed too) sits between FD synthesis and the limiter
    ("cicp16_to_cicp6", "sine_1khz_cicp16.mhas", ["-cicp:6"]),
    ("cicp19_to_cicp2", "sine_1khz_cicp19.mhas", ["-cicp:2"]),
]


def stream_path(stream):
    return os.path.join(ROOT, "oracle", "_ref", "smoke", stream)
