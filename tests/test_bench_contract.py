"""bench.py output contract (one JSON line with the keys the driver reads): the CPU reference arm here,
the B200 arm on the GPU box (small configs[0] workload so it takes seconds)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, timeout=600):
    p = subprocess.run([sys.executable, str(REPO / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = _run(["--impl", "reference", "--config", "A", "--steps", "1", "--warmup", "0", "--literal-seconds", "2"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # "reference" = the unmodified reference pip-installed under baseline/_ref by build(); "port" only when that install is absent
    from oracle import reference_arm as ra
    assert cb["kind"] == ("reference" if ra.available() else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the literal per-user path of the reference is timed beside the batched forward
    assert 0 < cb["literal_value"] < cb["value"] and cb["literal_users_per_sec"] > 0 and "get_recommendations" in cb["literal_sample"]
    # both arms print the same config dict for the same command line
    sys.path.insert(0, str(REPO))
    import bench
    assert d["config"] == bench.config_dict("A", 1, 4096)


@pytest.mark.gpu
def test_b200_arm_json_contract():
    d = _run(["--config", "A", "--steps", "2", "--warmup", "3", "--user-block", "256", "--cpu-seconds", "1", "--literal-seconds", "1"])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "gpu_launches", "parity_sample"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["dtype"] == "bf16" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["run"]["kernel_path"] == "tcgen05" and d["run"]["exact_rescore"] is True and d["gpu_launches"] >= 2
    sys.path.insert(0, str(REPO))
    import bench
    assert d["config"] == bench.config_dict("A", 1, 256)          # identical to what --impl reference prints
    ps = d["parity_sample"]
    assert ps["raw16_vs_exact_top50_overlap_mean"] >= 45 and ps["raw16_max_abs_score_error_vs_fp32"] < 0.1
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["kernel_launches"] == 2 and 0.0 < r["kernel_share_of_step"] <= 1.0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 256 * 50 * 8
    from oracle import reference_arm as ra
    assert d["cpu_baseline"]["kind"] == ("reference" if ra.available() else "port") and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
