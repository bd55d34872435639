"""The CPU-only parts of bench.py (driver contract): the `--impl reference` arm prints one JSON line with
the agreed keys, rank != 0 stays silent under a torchrun-style environment, and the cpu_baseline leg
reports which CPU code ran (the reference's own sources when oracle/_ref holds them, else the port)."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "24", "--steps", "2",
                        "--warmup", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mcell-iters/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["iterations_per_step"] == 50 and "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_ignores_torchruns_omp_setting():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm sizes its own team (one thread per
    physical core it may run on) and says how many it used"""
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    d = json.loads(_run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})[0])
    assert d["cpu_baseline"]["cores"] == bench.host_cores()
    assert "numa" in d["cpu_baseline"]


def test_cpu_baseline_leg_reports_the_code_that_ran():
    """the cpu_baseline object of the GPU arm's line comes from the CPU arm run in a fresh process"""
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    cpu = bench.cpu_baseline_subprocess(20, 6)
    from oracle import ref_ldu
    assert cpu["kind"] == ("reference" if ref_ldu.omp_available() else "port")
    assert cpu["value"] > 0 and cpu["cores"] >= 1 and "stock_dic_serial" in cpu
    if cpu["kind"] == "reference":
        assert cpu["port"]["value"] > 0
