"""bench.py contract checks that need no GPU: the reference arm (the oracle port on host cores) prints exactly
one JSON line with the keys the driver reads; the B200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "8192", "--steps", "2", "--warmup", "1",
                        "--cpu-seconds", "1.5"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "billion_interactions_per_s" and d["unit"] == "G interactions/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    assert d["ms_per_full_step_extrapolated"] >= d["ms_per_step"] > 0


def test_reference_arm_uses_all_cores_under_torchrun_and_words_the_workload_like_the_b200_arm(built):
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers (round 1: the reference arm of the scaling runs sat on one
    core and the speed-up against it came out 18x too large): the leg asks for the box's cores itself.  The driver also
    compares config.workload of the two arms: one function words both."""
    sys.path.insert(0, ROOT)
    import bench
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == bench.host_cores() and d["cpu_baseline"]["omp_num_threads_env"] == "1"
    assert d["config"]["workload"] == bench.workload_string(bench.N_DEFAULT, "f32") and d["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_quietly(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "4096", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu(built):
    from conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", "4096", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr or "no CUDA device" in r.stderr
