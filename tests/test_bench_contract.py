"""bench.py's contract on CPU: the reference arm (the reference's own CPU code on the host cores) prints one JSON line with
the keys the driver reads; a non-zero rank under torchrun stays silent; the roofline traffic comes from a committed capture."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_of(1)     # the reference arm prints exactly the config object of the other arm
    import re
    m = re.search(r"on (\d+) threads", d["cpu_baseline"]["sample"])
    assert m and int(m.group(1)) == d["cpu_baseline"]["cores"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; rank 0's reference run must not inherit that"""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_traffic_comes_from_a_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    traffic, source = bench.measured_traffic("integrate_pipelined_kernel")
    assert source and os.path.exists(os.path.join(ROOT, source)) and source.startswith("profiles/")
    assert 4e8 < traffic < 7e8          # read + write of the bench frame's updated voxels
    assert bench.measured_traffic("no_such_kernel") == (None, None)


def test_clock_sampler_says_why_when_no_gpu_answers():
    """NVML first (fast enough for a 16 ms timed region), nvidia-smi second; with neither the line carries the reason, not a crash"""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    c = s.stop()
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if c["sm_mhz"] is None:
        assert c.get("samples", 0) == 0
    else:
        assert c["samples"] >= 1 and c["sm_mhz"] <= c["sm_max_mhz"] + 1


def test_committed_bench_line_carries_the_contract_keys():
    """profiles/r02_bench_ours.json is the line the final tree printed on the B200"""
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_ours.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1 and d["roofline"]["traffic"]
    assert abs(d["roofline"]["achieved"] / d["roofline"]["peak"] - d["roofline"]["frac"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["clocks"]["samples"] >= 1 and d["clocks"]["reasons"] == [] and d["gpu_launches"] > 0
    assert d["parity_check"]["ok"] is True
