"""bench.py contract on CPU: the reference arm (`--impl reference`, the CPU oracle port timed on the host cores) prints the
JSON line the driver expects, rank != 0 prints nothing, and the workload description is shared by both arms."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402


def _args(**kw):
    base = dict(gpus=1, steps=1, warmup=0, micro=32, impl="reference", cpu_baseline_seconds=1.0, graph=0, light=False)
    base.update(kw)
    return argparse.Namespace(**base)


def test_reference_arm_json_contract(monkeypatch, capsys):
    monkeypatch.setattr(bench, "SIZE", 32)         # same code path, small networks: seconds on CPU
    monkeypatch.setattr(bench, "DEC_SIZE", 64)
    monkeypatch.setattr(bench, "N_MLP", 2)
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(_args())
    line = capsys.readouterr().out.strip().splitlines()[-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("batch-sharded restoration inference") and d["steps"] == 1


def test_reference_arm_other_ranks_stay_silent(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(_args(gpus=2))
    assert capsys.readouterr().out == ""


def test_own_arm_refuses_to_run_without_cuda(monkeypatch):
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--light"])
    with pytest.raises(RuntimeError):
        bench.main()
