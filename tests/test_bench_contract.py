"""CPU: the parts of bench.py's contract that do not need a GPU -- the cpu_baseline object (oracle port, torch.cdist and
cKDTree legs on a bounded sample), the workload constants BASELINE.json names, and the argument parser defaults."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_is_baseline_config_c2():
    sys.path.insert(0, ROOT)
    import bench

    assert (bench.B, bench.N, bench.M) == (32, 2048, 16384)          # BASELINE.json configs[1]
    assert bench.FLOP_PER_PAIR == 8 and abs(bench.FP32_NOMINAL_TFLOPS - 74.4) < 0.1
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "Chamfer fwd+bwd point-pairs/sec" in base["metric"]


def test_cpu_baseline_object():
    sys.path.insert(0, ROOT)
    import bench
    from genpc_b200.synthetic import pcn_batch

    part, comp = pcn_batch(0, 1, bench.N, bench.M)
    out = bench.cpu_baseline(part, comp, budget_b=1)
    assert out["kind"] == "port" and out["unit"] == "pairs/s" and out["value"] > 0 and out["cores"] >= 1 and out["sample"]
    for leg in ("torch_cdist", "scipy_ckdtree"):
        assert out[leg].get("value", 0) > 0, out[leg]


def test_help_and_defaults():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in r.stdout
