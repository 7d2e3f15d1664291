"""bench.py contract checks on the tiny CI workload: the reference (CPU) arm here, the B200 arm on the GPU box."""

import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"bench.py must print exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = run(["--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "loss_fwd_bwd_tokens_per_sec" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_json_contract():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = run(["--workload", "tiny", "--steps", "3", "--warmup", "3"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["dtype"] == "bf16" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] == "tensor"
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    # ONE run times both backward modes; the headline (top level) is the recompute mode (no [T, V] block in HBM), the
    # saved-logits mode is a labelled sub-record that launches fewer kernels per micro-batch
    assert d["config"]["headline_mode"] == "recompute" and d["config"]["lm_head_backward"].startswith("recompute")
    rec, sav = d["modes"]["recompute"], d["modes"]["saved_logits"]
    assert rec["value"] == d["value"] and rec["e2e"] == d["e2e"] and "hbm_kernels" not in rec
    assert sav["lm_head_backward"].startswith("saved-logits") and "dz_from_saved_logits_kernel" in sav["hbm_kernels"]
    assert sav["value"] > 0 and sav["e2e"]["value"] > 0 and {"bound", "achieved", "peak", "frac", "traffic"} <= set(sav["roofline"])
    assert rec["frac_of_peak"]["executed_over_algorithmic"] > sav["frac_of_peak"]["executed_over_algorithmic"] == 1.0
    r = run(["--workload", "tiny", "--steps", "3", "--warmup", "3", "--lm-head-backward", "saved", "--no-cpu-baseline", "--no-torch-baseline"])
    assert r["config"]["lm_head_backward"].startswith("saved-logits") and r["value"] > 0 and set(r["modes"]) == {"saved_logits"}
