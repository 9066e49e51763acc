"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs the unmodified reference binary on the box's
host cores and prints ONE JSON line with the keys the driver reads.  (The GPU arm's line is checked on the GPU box by the
driver itself; its keys are built next to these in bench.py.)"""
import json
import os
import subprocess
import sys

import pytest

from conftest import REF_BIN, ROOT


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ref_sigma not built")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--vars", "40000", "--clauses", "200000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sigma_pass_literals_per_s" and d["unit"] == "literals/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_gpu_arm_refuses_to_run_without_a_device():
    """No CPU fallback: without CUDA the product arm fails loudly instead of measuring something else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--vars", "2000", "--clauses", "8000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
