"""The runnable examples stay runnable (CPU variants)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    e.update(env or {})
    return subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=280, env=e)


def test_xgboost_example():
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import xgboost_pandas
    res = xgboost_pandas.main(rows=1500)
    assert res["accuracy"] > 0.8 and res["reloaded_identical"] and res["rounds_after_warm_start"] > 5


def test_collectives_tour_cpu():
    r = _run(["examples/collectives_tour.py", "--np", "-2"], env={"SPARKDL_NUM_GPUS": "0", "CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert '"allreduce_sum"' in r.stdout and "tour finished on 2 ranks" in r.stdout


def test_all_examples_compile():
    import py_compile
    for f in sorted(os.listdir(os.path.join(ROOT, "examples"))):
        if f.endswith(".py"):
            py_compile.compile(os.path.join(ROOT, "examples", f), doraise=True)
