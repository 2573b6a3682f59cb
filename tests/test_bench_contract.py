"""bench.py's reference arm (CPU) prints ONE JSON line with the keys the driver reads; the B200 arm's keys are checked
on the GPU box by running it."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-hyp", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "hyp/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


@pytest.mark.gpu
def test_b200_arm_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "2", "--warmup", "3", "--hyp", "64",
                        "--chunk", "65", "--layer", "2", "--no-cpu-baseline"], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert (BASE_KEYS - {"cpu_baseline"}) | {"roofline", "gpu_launches", "clocks", "kernels"} <= set(d)
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and 0 < rf["frac"] < 1.5 and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9


@pytest.mark.gpu
def test_b200_arm_parity_block_and_cpu_baseline():
    """The same-run parity gates of SURVEY.md section 8d (small sample so the test stays short)."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3", "--hyp", "16",
                        "--chunk", "17", "--layer", "2", "--ref-hyp", "4", "--cpu-hyp", "4", "--cpu-samples", "2"],
                       capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    p = d["parity"]
    assert p["rgb_equal"] and p["depth_equal"] and p["argmax_equal"] and p["top3_equal"]
    assert p["token_rel_l2"] < 4e-3 and p["scores_max_bf16_ulp"] <= 1.0 and p["tco_max_abs_diff"] < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert "forward_mesh" in d["e2e"]["api"]


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_b200_arm_strong_scaling_two_gpus(exchange):
    """--scaling strong under torchrun: hypotheses of one proposal sharded over 2 ranks; the scores travel either through
    peer memory (fp_score_publish / fp_topk_after_exchange) or through fp_allgather_scores (NCCL from the C ABI); top-k
    after the exchange, identical on both ranks and equal to the single-GPU result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "bench.py"), "--gpus", "2",
                        "--scaling", "strong", "--exchange", exchange, "--steps", "2", "--warmup", "3", "--hyp", "33",
                        "--chunk", "34", "--layer", "2"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["scaling"] == "strong" and d["n_gpus"] == 2
    assert d["strong_scaling"]["identical_on_all_ranks"] and d["strong_scaling"]["equal_to_single_gpu"]
    assert d["strong_scaling"]["hypotheses_per_rank"] == 17
