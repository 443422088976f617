"""bench.py's Python, end to end, without a GPU (tests/bench_dry_driver.py): the result line of the headline arm carries every
key of the contract, the two other workloads produce their lines, the reference arm runs for real (it is CPU work)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, script="tests/bench_dry_driver.py"):
    p = subprocess.run([sys.executable, os.path.join(ROOT, script)] + list(argv), capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    return json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1]), p.stderr


@pytest.mark.timeout(900)
def test_headline_arm_line_has_the_contract_keys():
    line, err = _run("train")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["unit"] == "rays/s" and line["n_gpus"] == 1 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"
    assert "error" not in line["gpu_reference_port"] and line["gpu_reference_port"]["value"] > 0
    assert "error" not in line["psnr_vs_ref"] and line["psnr_vs_ref"]["rays"] == 1024 and "rgb_db" in line["psnr_vs_ref"] and "trained_psnr_db" in line["psnr_vs_ref"]
    assert "e2e done" in err and "timed steps" in err                      # the phase breadcrumbs


@pytest.mark.timeout(900)
@pytest.mark.parametrize("workload", ["render", "train_lpips"])
def test_other_workloads_produce_their_lines(workload):
    line, _ = _run(workload)
    assert line["value"] > 0 and line["unit"] == "rays/s" and workload.split("_")[0] in line["metric"] + line["config"]["workload"]


@pytest.mark.timeout(900)
def test_reference_arm_runs_on_the_host():
    line, _ = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu_rays", "8", script="bench.py")
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
