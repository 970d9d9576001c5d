"""bench.py contract (task statement): one JSON line on stdout with the agreed keys.

CPU: the reference arm (`--impl reference`, the CPU oracle port on the host cores) end to end, and the
algorithmic-FLOP bookkeeping against SURVEY.md §8(d).  GPU (-m gpu): a short run of the B200 arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=900):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines          # exactly ONE line on stdout, everything else goes to stderr
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"] == "unet256_train_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    from oracle import build_ref
    # "reference" = the reference's own U-Net module staged under oracle/_ref, "port" = the oracle restatement
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "unet21_256x256_b16_train_single_task"


def test_algorithmic_flops_match_the_survey_numbers():
    sys.path.insert(0, ROOT)
    import bench
    total, conv_fd = bench.igemm_flops_per_step(16, 256, 256)
    assert abs(total / 16 / 1e9 - 289.28) < 0.01          # train step GFLOP per 256x256 image
    total512, _ = bench.igemm_flops_per_step(16, 512, 512)
    assert abs(total512 / total - 4.0) < 1e-9
    assert 0.55 < conv_fd / total < 0.70                   # conv3x3 forward + dgrad share of the step's tensor work


@pytest.mark.gpu
def test_b200_arm_prints_one_json_line_with_the_contract_keys(lib_built):
    d = run_bench("--steps", "4", "--warmup", "3", "--no-cpu-baseline")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["steps"] == 4 and d["n_gpus"] == 1 and d["dtype"] == "bf16" and d["scaling"] == "weak"
    assert d["value"] > 500 and d["e2e"]["value"] > 500
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * 3 * 256 * 256 * 4 + 16 * 256 * 256 * 8 and d["e2e"]["d2h_bytes_per_step"] == 16
    assert d["gpu_launches"] >= 4 * 150
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.3 < r["frac"] < 1.0 and r["traffic"] > 0
