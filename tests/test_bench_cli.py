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


def test_evidence_tool_names_kernels_like_the_host_code():
    """tools/make_evidence.py maps mangled (cuobjdump) and demangled (ncu) kernel names to the launch names bench.py uses as keys of
    profiles/*traffic.json, and bench.py finds the committed capture of the headline kernels."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_evidence", os.path.join(ROOT, "tools", "make_evidence.py"))
    me = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(me)
    assert me.friendly("_ZN3prl11gemm_kernelILi2ELi1ELb0ELb0ELi0EEEvNS_10GemmParamsE") == "lmhead_fwd_kernel"
    assert me.friendly("void gemm_kernel<2, 1, 0, 0, 1>(GemmParams)") == "lmhead_bwd_dz_kernel"
    assert me.friendly("void prl::gemm_kernel<(int)2, (int)2, (bool)0, (bool)1, (int)2>(prl::GemmParams)") == "lmhead_bwd_dhidden_kernel"
    assert me.friendly("void gemm_kernel<2, 2, 1, 1, 2>(GemmParams)") == "lmhead_bwd_dweight_kernel"
    assert me.friendly("void gemm_kernel<1, 1, 0, 0, 2>(GemmParams)") == "gemm_kernel<1,1,0,0,2>"
    assert me.friendly("dz_from_saved_logits_kernel(__nv_bfloat16 *, long, long)") == "dz_from_saved_logits_kernel"
    assert me.friendly("_ZN3prl15loss_seq_kernelEPKfS1_S1_PKhPKiiffPfS6_S6_S6_Pd") == "loss_seq_kernel"
    assert me.to_bytes("1.5", "Gbyte") == 1.5e9 and me.to_ns("2", "ms") == 2e6
    sys.path.insert(0, ROOT)
    saved_stdout = os.dup(1)  # importing bench.py points fd 1 at stderr (its stdout carries exactly one JSON line): undo that here
    try:
        import bench
    finally:
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    for mode, kernel in (("recompute", "lmhead_bwd_dz_kernel"), ("recompute", "lmhead_fwd_kernel"), ("saved", "lmhead_bwd_dweight_kernel")):
        rec = bench.ncu_traffic(mode, kernel, "qwen3-8b-hendrycks")
        assert rec and rec["bytes"] > bench.algorithmic_bytes(kernel, 16384, 4096, 151936, 16384 if mode == "recompute" else 151936) * 0.9
    assert bench.ncu_traffic("recompute", "lmhead_fwd_kernel", "tiny") is None
