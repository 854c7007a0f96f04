"""CPU-only: the parts of bench.py's contract that run without a GPU -- the `--impl reference` arm (the oracle port timed
on the host cores on the benchmark's own graph; here shrunk so the test takes seconds) and the shape of its JSON line."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_runs_the_stated_graph_and_reports_what_it_ran(monkeypatch):
    bench = _load_bench()
    monkeypatch.setattr(bench, "N_PER_GPU", 16_000)
    monkeypatch.setattr(bench, "E_PER_GPU", 320_000)
    base, t, n_timed, n_warm = bench.cpu_reference_run(steps=4, warmup=2, budget_s=30.0)
    assert base["kind"] == "port" and base["nodes"] == 16_000 and 300_000 < base["edges"] <= 320_000
    assert base["timed_forwards"] == n_timed == 4 and n_warm == 2
    assert abs(base["value"] - base["edges"] / t) < 1e-6 * base["value"]
    assert base["eighth_size_sample"]["nodes"] == 2_000          # the thread-count trial is a second key, not the value
    assert base["cores"] >= 1 and base["cpu_model"] and "cold" in base["sample"]
    # the time budget bounds the number of timed forwards but never below 3
    _, _, n_timed, _ = bench.cpu_reference_run(steps=50, warmup=5, budget_s=1e-9)
    assert n_timed == 3


def test_reference_arm_prints_one_json_line_with_the_contract_keys(tmp_path):
    code = ("import bench, sys; bench.N_PER_GPU, bench.E_PER_GPU = 8000, 160000; "
            "sys.argv = ['bench.py', '--impl', 'reference', '--gpus', '4', '--steps', '3', '--warmup', '3']; bench.main()")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True and d["n_gpus"] == 4
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["gpu_launches"] == 0
    assert "reference_ran" in d["config"] and d["steps"] == 3
    # the other ranks of a torchrun launch exit without work and without output
    r2 = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT,
                        env=dict(os.environ, RANK="2", WORLD_SIZE="4"))
    assert r2.returncode == 0 and not [ln for ln in r2.stdout.splitlines() if ln.startswith("{")]
