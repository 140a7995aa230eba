"""bench.py's reference arm runs on CPU: check the JSON contract of the line it prints (keys the driver
and the judge read), on a small shape so the test takes seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--users", "3000", "--items", "700", "--batch", "512"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BPR interactions/sec" and d["unit"] == "interactions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "train_single_batch" in cb["sample"]
    assert d["config"]["workload"].startswith("configs[1]") and d["dtype"] == "f32" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3",
           "--warmup", "1", "--users", "3000", "--items", "700", "--batch", "512"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_of_the_other_configs_prints_a_contract_line():
    """--config 3 / 4 / 5 have their own CPU arms (bench_configs.cpu_*): same line contract, their own metric."""
    for config, metric, extra in ((5, "embedding gather HBM GB/s (D=128)", ["--rows", "200000"]),
                                  (4, "BPR interactions/sec (LightGCN, whole-graph propagate per batch)",
                                   ["--users", "3000", "--items", "700", "--edges", "20000", "--batch", "512"])):
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(config), "--steps", "3",
               "--warmup", "1"] + extra
        out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == metric and d["value"] > 0 and d["gpu_launches"] == 0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["unit"] == d["unit"]
