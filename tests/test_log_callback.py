"""LogCallback contract: per-epoch / per-batch messages with elapsed time, routed
through log_to_driver to the driver's stdout."""
from sparkdl import HorovodRunner


def test_log_callback_inprocess(capsys):
    from sparkdl.horovod.tensorflow.keras import LogCallback
    cb = LogCallback()
    cb.on_epoch_begin(0)
    cb.on_batch_end(3, {"loss": 1.0})
    cb.on_epoch_end(0, {"loss": 0.5, "acc": 0.75})
    out = capsys.readouterr().out
    assert "Epoch 0 begin" in out and "batch 3" not in out
    assert "Epoch 0 end" in out and "seconds elapsed" in out and "acc: 0.7500 - loss: 0.5000" in out
    cb = LogCallback(per_batch_log=True)
    cb.on_epoch_begin(1)
    cb.on_batch_end(3, {"loss": 1.0})
    assert "Epoch 1 batch 3: loss: 1.0000" in capsys.readouterr().out


def test_log_callback_reaches_driver_from_worker(capfd, monkeypatch):
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")

    def main():
        from sparkdl.horovod.torch import LogCallback
        cb = LogCallback(per_batch_log=True)
        cb.on_epoch_begin(0)
        cb.on_batch_end(0, {"loss": 2.0})
        cb.on_epoch_end(0, {"loss": 1.0})
        print("worker stdout that must stay hidden")
    HorovodRunner(np=1, driver_log_verbosity="log_callback_only").run(main)
    out = capfd.readouterr().out
    assert "Epoch 0 batch 0: loss: 2.0000" in out and "Epoch 0 end" in out and "must stay hidden" not in out
