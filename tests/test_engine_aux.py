"""Auxiliary subsystems of the engine: Chrome-trace timeline, stall watchdog,
Horovod import shim, config env spellings."""
import json
import os

import pytest

from sparkdl import HorovodRunner


def test_timeline_chrome_trace(tmp_path, monkeypatch):
    path = str(tmp_path / "timeline.json")
    monkeypatch.setenv("HOROVOD_TIMELINE", path)          # Horovod spelling is honoured

    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        for i in range(3):
            hvd.allreduce(torch.ones(10), name="t%d" % i)
        hvd.allgather(torch.ones(2, 2), name="g")
        hvd.shutdown()
        return True
    assert HorovodRunner(np=-2).run(main)
    ev = json.load(open(path))
    names = {e["name"] for e in ev}
    assert {"NEGOTIATE_ALLREDUCE", "SHM_ALLREDUCE", "NEGOTIATE_ALLGATHER"} <= names
    assert {e["ph"] for e in ev} == {"B", "E"} and all("ts" in e and "pid" in e for e in ev)
    assert os.path.exists(path + ".1")                    # rank 1 writes its own file


def test_stall_watchdog_warns_and_shuts_down(monkeypatch, capfd):
    monkeypatch.setenv("SPARKDL_STALL_CHECK_TIME_SECONDS", "1")
    monkeypatch.setenv("SPARKDL_STALL_SHUTDOWN_TIME_SECONDS", "3")

    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        msg = "no error"
        if hvd.rank() == 0:
            try:
                hvd.allreduce(torch.ones(4), name="only_rank0_submits")
            except hvd.HorovodInternalError as e:
                msg = str(e)
        else:
            import time
            time.sleep(6)
        return msg
    msg = HorovodRunner(np=-2).run(main)
    assert "stalled" in msg and "only_rank0_submits" in msg
    err = capfd.readouterr().err
    assert "has been waiting" in err and "ranks [1]" in err


def test_horovod_shim_import(monkeypatch):
    monkeypatch.setenv("SPARKDL_HOROVOD_SHIM", "1")

    def main():
        import horovod.torch as hvd          # unmodified Horovod-style script
        import torch
        hvd.init()
        out = hvd.allreduce(torch.full((3,), float(hvd.rank() + 1)), op=hvd.Sum)
        return out.tolist(), hvd.size(), hvd.nccl_built()
    vals, size, nccl = HorovodRunner(np=-2).run(main)
    assert vals == [3.0, 3.0, 3.0] and size == 2 and nccl == 0


def test_config_env_spellings(monkeypatch):
    from spark_deep_learning_b200.utils.config import get_config
    monkeypatch.setenv("HOROVOD_FUSION_THRESHOLD", "1048576")
    monkeypatch.setenv("SPARKDL_CYCLE_TIME", "2.5")
    cfg = get_config(refresh=True)
    assert cfg.fusion_threshold == 1048576 and cfg.cycle_time_ms == 2.5
    monkeypatch.setenv("SPARKDL_FUSION_THRESHOLD", "2048")   # SPARKDL_ spelling wins
    assert get_config(refresh=True).fusion_threshold == 2048


def test_nvtx_annotation_is_free_when_off_and_balanced_when_on(monkeypatch):
    """SPARKDL_NVTX: off -> the function object is returned untouched; on -> push/pop pairs even when the body raises."""
    import torch
    from spark_deep_learning_b200.utils import nvtx

    def body(fail):
        if fail:
            raise ValueError("x")
        return 7
    assert nvtx.annotate("r", enabled=False)(body) is body
    calls = []
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda name: calls.append(name))
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda: calls.append(None))
    wrapped = nvtx.annotate("sparkdl.test", enabled=True)(body)
    assert wrapped(False) == 7
    with pytest.raises(ValueError):
        wrapped(True)
    assert calls == ["sparkdl.test", None, "sparkdl.test", None]


def test_timeline_can_be_started_and_stopped_at_run_time(tmp_path):
    """hvd.start_timeline / stop_timeline (Horovod's run-time switch): events only while recording, valid JSON, and a
    second start while active is refused."""
    path = str(tmp_path / "rt.json")

    def main(path):
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        hvd.allreduce(torch.ones(4), name="before")
        hvd.start_timeline(path)
        refused = False
        try:
            hvd.start_timeline(path + ".again")
        except ValueError:
            refused = True
        hvd.allreduce(torch.ones(4), name="during")
        hvd.stop_timeline()
        hvd.allreduce(torch.ones(4), name="after")
        hvd.start_timeline(path + ".second")            # restartable
        hvd.allreduce(torch.ones(4), name="second")
        hvd.stop_timeline()
        return refused
    os.environ["SPARKDL_NUM_GPUS"] = "0"
    try:
        assert HorovodRunner(np=-2).run(main, path=path) is True
    finally:
        del os.environ["SPARKDL_NUM_GPUS"]
    for p, inside, outside in ((path, "during", ("before", "after")), (path + ".second", "second", ("during",))):
        ev = json.load(open(p))
        names = {e["cat"] for e in ev}
        assert any(inside in n for n in names), names
        assert not any(o in n for n in names for o in outside), names
    assert os.path.exists(path + ".1")                  # rank 1 wrote its own file
