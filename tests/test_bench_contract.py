"""CPU: the reference arm of bench.py (`--impl reference`: the reference's HostExecutor from oracle/_ref on the host cores)
prints ONE JSON line with the keys the bench contract names.  The GPU arm's line is checked where it runs (profiles/)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmatx_ref_host.so")):
        pytest.skip("oracle/_ref is not built on this box")
    # test hook: 2^24 of the 2^30 elements (the arm itself runs the whole workload; here only the line's shape is checked)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       env=dict(os.environ, MXB_BENCH_REF_ELEMS=str(1 << 24)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "reference"
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_map_the_product_library():
    """`--impl reference` must run none of this repo's engine: bench.py's reference arm imports nothing of matx_b200."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    arm = src[src.index("class _RefHost"):src.index("# this repo's arm")]
    assert "matx_b200" not in arm.replace("libmatx_b200.so", "").replace("of matx_b200 is imported", "") and "oracle_harness" not in arm


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])
