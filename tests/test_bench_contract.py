"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the required keys, and the
byte model / configs of the measured workloads are the ones SURVEY.md section 8d states."""
import json
import os
import subprocess
import sys

from tests.helpers import ROOT


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 1e5
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms print the SAME config object for the same command line (the driver compares them key by key)
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.config_dict("as", bench.N_PER_GPU, bench.N_PER_GPU)


def test_reference_arm_non_zero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                          "--warmup", "3"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_byte_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench

    # B = w * (A + 2*S + O + 1);  AS 52/104 B, Hawkes 76/152 B, OE 60/120 B
    assert bench.algorithmic_bytes_per_env_step(2, 4, 4) == 52 and bench.algorithmic_bytes_per_env_step(2, 4, 8) == 104
    assert bench.algorithmic_bytes_per_env_step(2, 6, 4) == 76 and bench.algorithmic_bytes_per_env_step(2, 6, 8) == 152
    assert bench.algorithmic_bytes_per_env_step(1, 5, 4) == 60 and bench.algorithmic_bytes_per_env_step(1, 5, 8) == 120
    assert bench.N_PER_GPU == 1 << 20
