"""bench.py's contract on the CPU side: the reference arm (`--impl reference`: the CPU chain on the host cores) prints ONE JSON line
whose `config` object is, key for key and value for value, the one the repo arm prints for the same flags (the driver compares them),
and carries the keys the driver reads.  The repo arm itself needs a GPU (there is no CPU fallback) and says so."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_shared_config():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-500:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    args = types.SimpleNamespace(content="natural", float_mode="c", opt=[])
    assert d["config"] == bench.config_dict(dict(bench.CONFIGS["cfg1"]), args, 1)          # what the repo arm would print
    assert set(d["config"]) == {"workload", "content", "frames_per_launch", "float_mode", "options", "l2", "sharding"}


def test_repo_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
