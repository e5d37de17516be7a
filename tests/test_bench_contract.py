"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    """`--impl reference` times the oracle port of the reference's CPU path (the reference is Python and cannot travel to
    the GPU box) on a bounded sample and prints the same metric / unit / config as the B200 arm."""
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "vq_tokens_per_sec_fwd_bwd_K8192_D32" and d["unit"] == "tokens/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_b200_arm_line_on_gpu():
    d = _run("--steps", "4", "--warmup", "3", "--profile-every", "2")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "hbm_side"} <= set(d)
    assert d["metric"] == "vq_tokens_per_sec_fwd_bwd_K8192_D32" and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert d["gpu_launches"] >= 4 * 4 and d["value"] > 1e8
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and 0.2 < r["frac"] < 1.0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["cores"] >= 1 and "sample" in c
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]


def test_reference_arm_line_of_the_fused_pre_quant_workload_on_cpu():
    """`--config cfg3pre` (pre_quant + quantiser, indices only) keeps the same contract; its reference
    arm is the oracle port of `pre_quant` + `Codebook.forward` on a bounded sample."""
    d = _run("--impl", "reference", "--config", "cfg3pre", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "vq_tokens_per_sec_prequant_encode_K8192_D32_C512"
    assert d["unit"] == "tokens/s" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["C"] == 512 and "model" not in d["config"]
