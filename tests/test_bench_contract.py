"""bench.py's host-side contract, without a GPU: the work model of SURVEY.md 8(d), the sub-batch / stream policy, the
reference arm's JSON line (bounded to two coder-blocks here) and the shape of the config both arms must share."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_work_model_matches_survey_8d():
    # t = 0 scores S candidates against the one empty beam, every later variable S * B' (beam_search_coder.py:97-106, 79-95);
    # W = 10 + 24 / B' lane-instructions per candidate-dim
    cand, cd, parts, instr = bench.work_model([3], [1000], s=36, nbeams=20)
    assert cand == 36 + 2 * 36 * 20 and cd == cand * 1000 and parts == 3
    assert np.isclose(instr, 1000 * (36 * 34.0 + 2 * 36 * 20 * (10 + 24 / 20)))
    cand1, _, _, instr1 = bench.work_model([5, 1], [64, 64], s=36, nbeams=1)        # n_beams = 1: B' = 1 throughout
    assert cand1 == 5 * 36 + 36 and np.isclose(instr1, 64 * 36 * 34.0 * 6)
    cand_s, _, _, _ = bench.work_model([4], [10], s=7, nbeams=20)                    # S < B: B' = min(B, S) in this model
    assert cand_s == 7 + 3 * 7 * 7


def test_stream_policy():
    # a rank's launch is split into two sub-batch streams below ~8 coder-blocks per block context (2 contexts x 148 SMs)
    assert bench.auto_streams(128) == 2 and bench.auto_streams(256) == 2          # N = 8 and N = 4 shares of configs[3]
    assert bench.auto_streams(512) == 1 and bench.auto_streams(1024) == 1         # N = 2, N = 1
    assert bench.auto_streams(1) == 2                                            # never more streams than images: clamped by the caller


def test_both_arms_name_the_same_workload():
    cfg1, cfg8 = bench.config_dict(1), bench.config_dict(8)
    assert cfg1["workload"] == cfg8["workload"] and cfg1["images_total"] == cfg8["images_total"] == 1024
    assert "configs[3]" in cfg1["workload"] and "n_beams=20" in cfg1["workload"] and "S=36" in cfg1["workload"]
    assert cfg1["parallelism"].startswith("dp1") and cfg8["parallelism"].startswith("dp8")


def test_reference_arm_line(built, monkeypatch):
    """`bench.py --impl reference` (the C port on the host cores): one JSON line with the contract's keys; ranks > 0 print nothing"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
    monkeypatch.setattr(bench, "cpu_sample_blocks", lambda cores: 2)
    lines = []
    monkeypatch.setattr("builtins.print", lambda *a, **k: lines.append(a[0]))
    bench.run_reference(type("A", (), {"steps": 1, "warmup": 0, "gpus": 1})())
    line = json.loads(lines[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "candidates/s" and line["value"] > 0
    assert line["config"] == bench.config_dict(1) and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
