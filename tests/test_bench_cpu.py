"""CPU: the parts of bench.py that do not need a GPU -- the reference arm (the oracle port of the reference's numpy path
on a bounded sample) prints one JSON line with the contract's keys and the same config key set as our arm, and the clock
sampler degrades gracefully on a box without a driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

OUR_CONFIG_KEYS = {"workload", "particles_per_gpu", "nXY", "nV", "accumulate", "accumulators", "smoothing_parity",
                   "parallelism", "l2", "streams"}


def test_reference_arm_line():
    env = dict(os.environ, PNM_CPU_SAMPLE="60000")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                    # ONE JSON line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "particles/s" and j["higher_is_better"] is True
    assert j["metric"] == "particles/s binned to maps+LOSVD cube" and j["steps"] == 2 and j["warmup"] == 1
    assert set(j["config"]) == OUR_CONFIG_KEYS                # the driver compares the two arms' configs key by key
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "60000" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is the real time of one sample step (binning + smoothing), not an extrapolation
    assert abs(j["ms_per_step"] - (cb["t_bin_sample_s"] + cb["t_smooth_s"]) * 1e3) < 1e-6
    assert j["ms_per_step_whole_workload_extrapolated"] > j["ms_per_step"]


def test_our_arm_uses_the_same_config_keys():
    """Static check: run_b200's config dict literal carries exactly OUR_CONFIG_KEYS."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    start = src.index('"config": {"workload": workload_name(n_gpus, n_local)')
    block = src[start:src.index('"roofline":', start)]
    import re
    keys = set(re.findall(r'"([A-Za-z_0-9]+)":', block)) - {"config"}
    assert keys == OUR_CONFIG_KEYS, keys ^ OUR_CONFIG_KEYS


def test_clock_sampler_without_a_driver():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.sample_now()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
