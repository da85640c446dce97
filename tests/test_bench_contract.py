"""CPU: pieces of bench.py's JSON contract that do not need a GPU (the driver-recorded line named the wrong
workload because a nested helper rebound `stage`)."""
import ast
import os

import bench_step

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_default_workload_label_names_the_self_training_step():
    c3 = bench_step.workload_config(3, 16, 1, True, "tcgen05")
    assert "self-training step" in c3["workload"] and "tools/train_ssl_reg.py" in c3["workload"]
    assert c3["global_batch"] == 16 and c3["parallelism"] == "dp1"
    c2 = bench_step.workload_config(2, 16, 1, True, "tcgen05")
    assert "tools/train_align_reg.py" in c2["workload"] and "self-training" not in c2["workload"]


def test_run_does_not_rebind_stage():
    """no nested def / assignment inside run() may shadow the names the JSON line is built from"""
    tree = ast.parse(open(os.path.join(ROOT, "bench_step.py")).read())
    run = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run")
    nested = [n.name for n in ast.walk(run) if isinstance(n, ast.FunctionDef) and n is not run]
    assert "stage" not in nested and "world" not in nested and "imgs" not in nested
    stage_assignments = [n for n in ast.walk(run) if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "stage" for t in n.targets)]
    assert len(stage_assignments) == 1


def test_reference_arm_line_shape():
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "lrh", "--steps", "1", "--warmup", "1",
                        "--regions", "50"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, REGDA_REF_LRH_TILES="2"))
    assert r.returncode == 0, r.stderr[-1500:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_step_reference_arm_names_our_workload(monkeypatch, capsys):
    """the CPU arm reports the SAME metric / workload label as our arm (the driver compares the two lines), the bounded sample stated beside it"""
    import json
    import types
    monkeypatch.setattr(bench_step, "cpu_step_baseline",
                        lambda reps=2, batch=2: dict(value=2.5, unit="images/s", cores=4, kind="port", sample="2+2 tiles", ms_per_step=1600.0))
    bench_step.reference(types.SimpleNamespace(steps=2, warmup=1), 0, 1)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    ours = bench_step.workload_config(3, 16, 1, True, "tcgen05")
    assert line["impl"] == "reference" and line["metric"] == "train images/sec (512x512, 6-class)" and line["unit"] == "images/s"
    assert line["config"]["workload"] == ours["workload"] and line["config"]["sample"] == "2+2 tiles"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"] == {"value": 2.5, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    bench_step.reference(types.SimpleNamespace(steps=2, warmup=1), 1, 2)          # other ranks: no work, no line
    assert capsys.readouterr().out == ""
