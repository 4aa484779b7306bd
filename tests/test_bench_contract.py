"""CPU test of bench.py's reference arm: one JSON line on stdout (nothing else), carrying the keys the driver's contract
names.  The GPU arm shares the emit path; its keys are checked on the B200 runs kept under profiles/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "64"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio-hours/sec analysed" and d["unit"] == "audio-h/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-h/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms describe the workload with the same dict
    import shutil
    if shutil.which("g++") is not None and os.uname().machine == "x86_64":
        q = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dev", "simt_emu", "run_bench_emu.py"), "--tracks", "2", "--frames", "64",
                            "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900)
        assert q.returncode == 0, q.stderr[-3000:]
        g = json.loads(q.stdout.splitlines()[0])
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "64",
                            "--tracks", "2"], capture_output=True, text=True, timeout=600)
        assert json.loads(r.stdout.splitlines()[0])["config"] == g["config"]


def test_last_gpu_bench_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r2_bench_final_1gpu.json")
    d = json.load(open(path))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_final_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["config"] == d["config"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    for name in ("album", "mixed", "stream"):
        assert "error" not in d["configs"][name] and d["configs"][name]["value"] > 0, name
    assert d["configs"]["album"]["histogram_is_the_sum"] is True
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_gpu_arm_runs_end_to_end_on_the_interpreted_kernels():
    """bench.py's own arm at toy sizes, with the engine built from the kernel sources by the CPU SIMT interpreter
    (tests/dev/simt_emu/run_bench_emu.py): every code path of the arm runs and the record carries the contract's keys.  The
    numbers mean nothing; the parity fields must still be clean."""
    import shutil
    import pytest
    if shutil.which("g++") is None or os.uname().machine != "x86_64":
        pytest.skip("needs g++ on x86-64")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dev", "simt_emu", "run_bench_emu.py"), "--tracks", "8", "--frames", "64",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "delta_db_vs_ref", "gain_fields"):
        assert k in d, k
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert "error" not in d["gain_fields"], d["gain_fields"]
    for k in ("album", "mixed", "stream"):        # BASELINE.json configs[2..4] beside the headline
        assert "error" not in d["configs"][k] and d["configs"][k]["value"] > 0, d["configs"][k]
    assert d["configs"]["album"]["histogram_is_the_sum"] and d["configs"]["mixed"]["sample_rates"] == [44100, 48000]
    assert d["configs"]["mixed"]["mono_tracks"] > 0
    assert d["gain_fields"]["frames_modified"] == 8 * 64
    assert d["delta_db_vs_ref"]["max_abs"] == 0.0 and d["cpu_baseline"]["loudness_mismatches_vs_gpu"] == 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
