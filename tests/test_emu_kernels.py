"""The -m gpu parity tests, run here (no GPU) on the KERNEL SOURCES through the CPU SIMT interpreter of tests/dev/simt_emu.

This is a check of the kernels' logic (index arithmetic, walks, table handling, the f32 / f64 operation order) against the
oracle before the code reaches a B200 -- not a substitute for the GPU run, and not a product path: the interpreter is built
from mp3rgain_b200/csrc by tests/dev/simt_emu/build_emu.py into tests/dev/simt_emu/, and only a pytest session started
with RG_SIMT_EMU=1 (see conftest.py) ever loads it.  The slowest tests (full-length tracks, the bigger corpora) are left to
the GPU; `RG_SIMT_EMU=1 python -m pytest tests -m gpu` runs everything but the two 3-minute-track tests in about two minutes."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SLOW_OR_SUBPROCESS = [
    "tests/test_gpu_mp3.py::test_full_length_tracks_time_blocked_equals_exact",
    "tests/test_gpu_stage_b.py::test_long_track_fast_equals_exact",
    "tests/test_gpu_mp3.py::test_mixed_corpus_through_the_pipeline",
    "tests/test_gpu_gain_fields.py::test_three_minute_track_and_analysis_consistency",
    "tests/test_gpu_mp3.py::test_segmented_frame_index_equals_serial_walk",
    "tests/test_gpu_mp3.py::test_batch_of_many_tracks_and_small_scratch",
    "tests/test_gpu_mp3.py::test_peak_only_path_equals_the_analysis_peak",
]


@pytest.mark.skipif(shutil.which("g++") is None or os.uname().machine != "x86_64", reason="needs g++ on x86-64")
def test_gpu_parity_tests_pass_on_the_interpreted_kernel_sources():
    env = dict(os.environ, RG_SIMT_EMU="1")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
           "-k", "not below_96_bytes or kw0"]
    for t in SLOW_OR_SUBPROCESS:
        cmd += ["--deselect", t]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = p.stdout[-3000:] + p.stderr[-2000:]
    assert p.returncode == 0, tail
    assert " passed" in p.stdout and "failed" not in p.stdout, tail
