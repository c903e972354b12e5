"""T1 - launcher plumbing on CPU: cloudpickle round trip, rank-0 return value,
failure propagation, stdout multiplexing, log_to_driver, no orphans."""
import os
import time

import pytest

from sparkdl import HorovodRunner
from spark_deep_learning_b200.parallel.launcher import WorkerFailure


def test_return_value_lambda():
    # a lambda: only cloudpickle can ship it (reference test_return_value)
    assert HorovodRunner(np=-1).run(lambda: 42) == 42


def test_kwargs_and_closure():
    k = 7

    def main(a, b):
        return a * b + k
    assert HorovodRunner(np=-1).run(main, a=3, b=4) == 19


def test_rank0_value_and_env():
    def main():
        import os
        return (int(os.environ["SPARKDL_RANK"]), int(os.environ["SPARKDL_SIZE"]), os.getpid())
    r, s, pid = HorovodRunner(np=-3).run(main)
    assert (r, s) == (0, 3) and pid != os.getpid()


def test_np_positive_runs_locally(monkeypatch):
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    assert HorovodRunner(np=2, driver_log_verbosity="all").run(lambda: "ok") == "ok"


def test_exception_propagates_with_traceback():
    def main():
        import os
        if os.environ["SPARKDL_RANK"] == "1":
            raise ValueError("boom from rank one")
        import time
        time.sleep(30)
    t0 = time.time()
    with pytest.raises(WorkerFailure) as ei:
        HorovodRunner(np=-2).run(main)
    assert "boom from rank one" in str(ei.value) and ei.value.rank == 1
    assert time.time() - t0 < 20, "siblings must be torn down, not waited for"


def test_fault_injection_kill(monkeypatch):
    monkeypatch.setenv("SPARKDL_FAULT", "kill:rank=1:at=start")
    with pytest.raises(WorkerFailure) as ei:
        HorovodRunner(np=-2).run(lambda: __import__("time").sleep(20))
    assert ei.value.rank == 1 and ei.value.returncode == 137


def test_stdout_mux_and_log_to_driver(capfd):
    def main():
        import os
        from sparkdl.horovod import log_to_driver
        print("hello from %s" % os.environ["SPARKDL_RANK"])
        if os.environ["SPARKDL_RANK"] == "0":
            log_to_driver("driver-visible " + "y" * 5000)
        return 1
    HorovodRunner(np=-2).run(main)
    out = capfd.readouterr().out
    assert "[0] hello from 0" in out and "[1] hello from 1" in out
    line = [l for l in out.splitlines() if l.startswith("driver-visible")][0]
    assert len(line) == 4000


def test_log_callback_only_hides_worker_stdout(capfd, monkeypatch):
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")

    def main():
        from sparkdl.horovod import log_to_driver
        print("noisy worker line")
        log_to_driver("selected line")
    HorovodRunner(np=1).run(main)           # np>0 honours driver_log_verbosity
    out = capfd.readouterr().out
    assert "selected line" in out and "noisy worker line" not in out


def test_fault_injection_mid_training(monkeypatch):
    """Rank 1 dies at its 3rd collective: the gang is torn down and the driver raises quickly (no hang)."""
    monkeypatch.setenv("SPARKDL_FAULT", "kill:rank=1:op=3")
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")

    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        for i in range(10):
            hvd.allreduce(torch.ones(8), name="step%d" % i)
        return "finished"
    t0 = time.time()
    with pytest.raises(WorkerFailure) as ei:
        HorovodRunner(np=-2).run(main)
    assert ei.value.rank == 1 and ei.value.returncode == 137 and time.time() - t0 < 60


def test_np_zero_uses_all_slots_and_warns(monkeypatch, caplog):
    """np==0 (deprecated, /root/reference/README.md:57-61): every task slot."""
    import logging
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "2")

    def main():
        import os
        return int(os.environ["SPARKDL_SIZE"])
    with caplog.at_level(logging.WARNING, logger="HorovodRunner"):
        assert HorovodRunner(np=0).run(main) == 2
    assert any("deprecated" in r.message for r in caplog.records)


def test_driver_log_file_and_kept_job_dir(monkeypatch, tmp_path):
    """stdout/stderr of every worker is also tee'd into the driver log (ref runner_base.py:49-51)."""
    log = tmp_path / "driver.log"
    monkeypatch.setenv("SPARKDL_DRIVER_LOG", str(log))

    def main():
        import os
        import sys
        print("out-%s" % os.environ["SPARKDL_RANK"])
        sys.stderr.write("err-%s\n" % os.environ["SPARKDL_RANK"])
        from sparkdl.horovod import log_to_driver
        log_to_driver("cb-%s" % os.environ["SPARKDL_RANK"])
    HorovodRunner(np=-2).run(main)
    text = log.read_text()
    for needle in ("[0]<stdout> out-0", "[1]<stdout> out-1", "[0]<stderr> err-0", "[1]<stderr> err-1",
                   "<log_to_driver> cb-0", "<log_to_driver> cb-1"):
        assert needle in text, needle


def test_no_orphan_processes_after_failure():
    import psutil

    def main():
        import os
        import time
        if os.environ["SPARKDL_RANK"] == "0":
            raise RuntimeError("rank 0 fails")
        time.sleep(60)
    before = {p.pid for p in psutil.Process().children(recursive=True)}
    with pytest.raises(WorkerFailure):
        HorovodRunner(np=-3).run(main)
    time.sleep(0.5)
    after = [p for p in psutil.Process().children(recursive=True) if p.pid not in before and p.is_running()
             and p.status() != psutil.STATUS_ZOMBIE]
    assert not after, after


def test_external_gang_under_torchrun(tmp_path):
    """bench.py's multi-GPU launch mode: torchrun starts the ranks, every rank calls HorovodRunner.run with
    SPARKDL_EXTERNAL_GANG=1 and joins the existing gang (store hosted by rank 0) instead of spawning workers."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ext.py"
    script.write_text(
        "import os, sys, json\n"
        "sys.path.insert(0, %r)\n"
        "os.environ['SPARKDL_EXTERNAL_GANG'] = '1'\n"
        "from sparkdl import HorovodRunner\n"
        "def main(scale):\n"
        "    import torch\n"
        "    import sparkdl.hvd as hvd\n"
        "    hvd.init()\n"
        "    out = hvd.allreduce(torch.full((4,), float(hvd.rank() + 1)) * scale, op=hvd.Sum)\n"
        "    objs = hvd.allgather_object({'r': hvd.rank()})\n"
        "    res = (hvd.rank(), hvd.size(), out.tolist(), [o['r'] for o in objs])\n"
        "    hvd.shutdown()\n"
        "    return res\n"
        "res = HorovodRunner(np=-2).run(main, scale=2.0)\n"
        "open(os.path.join(%r, 'out.%%s.json' %% os.environ['RANK']), 'w').write(json.dumps(res))\n" % (root, str(tmp_path)))
    port = 20000 + os.getpid() % 20000
    env = dict(os.environ, SPARKDL_NUM_GPUS="0", CUDA_VISIBLE_DEVICES="", SPARKDL_FORCE_CPU="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    import json
    for rank in (0, 1):
        got = json.loads((tmp_path / ("out.%d.json" % rank)).read_text())
        assert got == [rank, 2, [6.0] * 4, [0, 1]]
