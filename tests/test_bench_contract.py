"""bench.py contract: one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout          # exactly ONE line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--sites", "3000", "--cpu-sample-sites", "3000")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "DRACH sites/sec at num_iterations=1000" and d["unit"] == "sites/s" and d["higher_is_better"] is True
    # the reference's own code (oracle/_ref, installed by oracle/make_ref.sh) when it is there, else the oracle's port
    ref_there = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "m6anet"))
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if ref_there else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_both_arms_print_the_same_config():
    """The driver compares the `config` of the two arms: it is a function of the command line only."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    argv = sys.argv
    try:
        sys.argv = ["bench.py", "--gpus", "1", "--steps", "2"]
        a = bench.parse_args()
        sys.argv = ["bench.py", "--gpus", "1", "--steps", "2", "--impl", "reference"]
        b = bench.parse_args()
    finally:
        sys.argv = argv
    assert bench.config_dict(a) == bench.config_dict(b)
    assert bench.config_dict(a)["sites"] == 1_000_000 and bench.config_dict(a)["reads_per_site"] == 50
    assert {c: bench.CONFIGS[c]["sites"] for c in bench.CONFIGS} == {2: 100_000, 3: 1_000_000, 4: 500_000, 5: 250_000}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--sites", "60000", "--cpu-sample-sites", "2000")
    assert BASE_KEYS | {"roofline", "clocks", "parity", "result_digest"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["dtype"] == "fp32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 1e6 and d["gpu_launches"] == 6     # tile_bounds_kernel + mil_infer(_tc)_kernel per step
    assert d["parity"]["ok"] is True and d["parity"]["sites"] >= 50_000 and d["parity"]["mod_count_mismatch"] == 0
    assert d["parity"]["max_abs_site"] <= 1e-4 and d["parity"]["max_abs_read"] <= 3e-6
    assert len(d["result_digest"]["site_prob_mod_count_sha256"]) == 64 and len(d["result_digest"]["read_prob_sha256"]) == 64
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["algorithmic_bytes_per_launch"] == 60000 * 2028
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 60000 * 50 * 36 and e["d2h_bytes_per_step"] > 60000 * 50 * 4
    assert e["matches_resident_path"] is True
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.parametrize("ragged", [False, True])
def test_synthetic_job_does_not_depend_on_the_sharding(ragged):
    """The job bench.py scores is generated in fixed blocks of sites: the shards of a 1-, 2-, 3-, 4- or 8-rank run put together
    are the same arrays bit for bit -- which is what makes the `result_digest` of the SCALE lines comparable."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    total, n_reads = 4003, 23
    whole = bench.synth_shard(0, total, n_reads, 7, ragged=ragged, total_sites=total)
    assert whole[1][0] == 0 and whole[1][-1] == len(whole[0]) and whole[2].shape == (total, 3)
    for world in (2, 3, 4, 8):
        cuts = [total * r // world for r in range(world + 1)]
        parts = [bench.synth_shard(cuts[r], cuts[r + 1], n_reads, 7, ragged=ragged, total_sites=total) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), whole[0])
        assert np.array_equal(np.concatenate([p[2] for p in parts]), whole[2])
        assert np.array_equal(np.concatenate([np.diff(p[1]) for p in parts]), np.diff(whole[1]))
