"""CPU: the bench.py JSON contract of the reference arm (the CPU-oracle timing itself is stubbed: it takes minutes)."""
import json
import sys
import types


def test_reference_arm_line(monkeypatch, capsys):
    sys.path.insert(0, ".")
    import bench
    monkeypatch.setattr(bench, "cpu_oracle_seconds_per_image", lambda verbose=False: (200.0, 150.0, "stub sample"))
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, ref_budget_s=1e9)
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert abs(line["value"] - 1 / 200.0) < 1e-12 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["metric"] == bench.METRIC and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent(monkeypatch, capsys):
    sys.path.insert(0, ".")
    import bench
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(types.SimpleNamespace(gpus=2, steps=1, warmup=0, ref_budget_s=1.0))
    assert capsys.readouterr().out == ""
