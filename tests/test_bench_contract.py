"""bench.py's reference arm runs on host cores only, so its JSON contract can be checked without a GPU; the B200 arm
must refuse to run without one (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import oracle_bindings as ob  # noqa: E402


@pytest.mark.skipif(not ob.have_ref(), reason="needs oracle/_ref (the reference build)")
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    # (the default workload, R-MAT-22, takes minutes on the CPU test box: the contract is checked on BASELINE configs[1])
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "rmat16"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grank_node_iterations_per_s" and d["unit"] == "node-iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert "R-MAT scale 16" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["reference_sample"] == "full job"  # R-MAT-16: all 30 iterations, nothing extrapolated
    # the reference process never maps the product library (its graph comes from the numpy generator)
    assert "libppr_b200" not in r.stderr


def test_default_workload_is_the_configuration_the_target_is_quoted_on():
    """BENCH / SCALE measure GRank on R-MAT scale 22 (BASELINE configs[3]) with the MC R-MAT-20 job under "mc"; big graphs
    bound the reference arm to one sweep of each partition, labelled as extrapolated"""
    sys.path.insert(0, str(ROOT))
    import bench
    assert bench.DEFAULT_WORKLOAD == "rmat22"
    w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert w["kind"] == "grank" and w["scale"] == 22 and (w["K"], w["L"], w["iterations"]) == (50, 100, 30)
    assert bench.WORKLOADS["rmat20mc"]["kind"] == "mc" and bench.WORKLOADS["rmat20mc"]["iterations"] == 1000
    assert bench.REFERENCE_SAMPLE_ITERATIONS == 2 and (1 << 22) > bench.REFERENCE_FULL_JOB_NODES


def test_committed_ncu_traffic_covers_the_default_workload():
    """roofline.traffic of the default line comes from the committed ncu launch list (profiles/traffic.json, made by
    tools/mk_profile.py from profiles/r2/launches_bench_rmat22.csv.gz): the entry must exist, name its source, and stay
    below the algorithmic bytes of the job (the L2 serves the popular baskets; well above would mean wasted re-reads)"""
    sys.path.insert(0, str(ROOT))
    import bench
    prof = bench.committed_profile(bench.DEFAULT_WORKLOAD)
    assert prof and prof["dram_bytes_per_step"] > 0 and (ROOT / prof["source"].split(" ")[0]).exists()
    line = json.loads((ROOT / "profiles" / "r2" / "bench_r2_n1_rmat22.json").read_text().splitlines()[-1])
    alg = line["roofline"]["algorithmic_bytes_per_step"]
    assert 0.2 * alg < prof["dram_bytes_per_step"] < 1.5 * alg


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--workload", "ring"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout) or "cuda" in (r.stderr + r.stdout)
