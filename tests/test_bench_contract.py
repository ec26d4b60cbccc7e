"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys, on rank 0 only.
(The GPU arm needs a device; its line is produced and recorded under profiles/ by the gpurun runs.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["metric"].startswith("pseudo-point logpdf+grad evals/s") and d["n_gpus"] == 1 and d["steps"] >= 1 and d["warmup"] >= 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) <= 1e-6 * 1e3
    assert "workload" in d["config"] and d["data"] == "synthetic" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "N=" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    p = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still time the CPU path on all host cores
    (round-1 SCALE ratios at N >= 2 were inflated by a single-threaded reference arm)."""
    p = run({"OMP_NUM_THREADS": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["blas_threads"] == (os.cpu_count() or 1)
