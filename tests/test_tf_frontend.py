"""``horovod.tensorflow.keras``-style frontend (compat shim) on CPU workers.  TensorFlow is not installable here, so the
Keras protocol is played by small stand-ins: variables with ``numpy()`` / ``assign()``, an optimizer class with
``apply_gradients`` / ``get_config`` / ``from_config``, a model with ``get_weights`` / ``set_weights``."""
from sparkdl import HorovodRunner


def _main():
    import numpy as np
    from spark_deep_learning_b200 import compat
    compat.install_horovod_shim()
    import horovod.tensorflow.keras as hvd          # what an unmodified Keras Horovod main imports

    class Var(object):
        def __init__(self, v):
            self.v = np.array(v, dtype=np.float32)

        def numpy(self):
            return self.v

        def assign(self, v):
            self.v = np.array(v, dtype=np.float32)

    class SGD(object):
        def __init__(self, learning_rate=0.1):
            self.learning_rate = learning_rate
            self.applied = []

        def get_config(self):
            return {"learning_rate": self.learning_rate}

        @classmethod
        def from_config(cls, cfg):
            return cls(**cfg)

        def apply_gradients(self, grads_and_vars):
            for g, v in grads_and_vars:
                v.assign(v.numpy() - self.learning_rate * np.asarray(g))
            self.applied.append(len(list(grads_and_vars)) if not isinstance(grads_and_vars, list) else len(grads_and_vars))

    class Model(object):
        def __init__(self, r):
            self.w = [np.full((3,), float(r + 1), dtype=np.float32), np.full((2, 2), float(r + 10), dtype=np.float32)]
            self.optimizer = None

        def get_weights(self):
            return [w.copy() for w in self.w]

        def set_weights(self, ws):
            self.w = [np.array(w) for w in ws]

    hvd.init()
    r, n = hvd.rank(), hvd.size()
    out = {}
    out["allreduce"] = hvd.allreduce(np.full((4,), float(r), dtype=np.float32)).tolist()
    out["allreduce_sum_scalar"] = hvd.allreduce(float(r + 1), op=hvd.Sum)
    out["allgather"] = hvd.allgather(np.full((r + 1, 2), float(r), dtype=np.float32)).shape[0]
    out["broadcast"] = hvd.broadcast(np.full((3,), float(r)), root_rank=n - 1).tolist()
    vs = [Var([r, r]), Var([[r]])]
    hvd.broadcast_variables(vs, root_rank=0)
    out["bvars"] = [v.numpy().tolist() for v in vs]
    # DistributedOptimizer: gradients are averaged before the wrapped class applies them
    opt = hvd.DistributedOptimizer(SGD(learning_rate=0.5))
    out["opt_class"] = (type(opt).__name__, isinstance(opt, SGD), opt.learning_rate)
    v = Var([1.0, 1.0])
    opt.apply_gradients([(np.array([float(r), 2.0 * r], dtype=np.float32), v)])
    out["after_step"] = v.numpy().tolist()            # 1 - 0.5 * mean(r), 1 - 0.5 * mean(2r)
    opt2 = hvd.DistributedOptimizer(SGD(learning_rate=1.0), backward_passes_per_step=2)
    v2 = Var([0.0])
    opt2.apply_gradients([(np.array([1.0 + r], dtype=np.float32), v2)])
    out["after_pass1"] = v2.numpy().tolist()
    opt2.apply_gradients([(np.array([3.0 + r], dtype=np.float32), v2)])
    out["after_pass2"] = v2.numpy().tolist()          # -(mean over ranks of (1+r + 3+r)/2)
    # callbacks
    m = Model(r)
    cb = hvd.callbacks.BroadcastGlobalVariablesCallback(0)
    cb.model = m
    cb.on_batch_end(0)
    out["bgv"] = [w.tolist() for w in m.get_weights()]
    logs = {"loss": float(r), "acc": 1.0, "note": "x"}
    hvd.callbacks.MetricAverageCallback().on_epoch_end(0, logs)
    out["metric"] = logs
    hvd.shutdown()
    return out


def test_keras_style_frontend_on_two_cpu_workers(monkeypatch):
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    out = HorovodRunner(np=-2).run(_main)
    assert out["allreduce"] == [0.5] * 4 and out["allreduce_sum_scalar"] == 3.0
    assert out["allgather"] == 3 and out["broadcast"] == [1.0, 1.0, 1.0]
    assert out["bvars"] == [[0.0, 0.0], [[0.0]]]
    assert out["opt_class"] == ("SGD", True, 0.5)
    assert out["after_step"] == [0.75, 0.5]
    assert out["after_pass1"] == [0.0] and out["after_pass2"] == [-2.5]
    assert out["bgv"] == [[1.0, 1.0, 1.0], [[10.0, 10.0], [10.0, 10.0]]]
    assert out["metric"] == {"loss": 0.5, "acc": 1.0, "note": "x"}


def _tape_main():
    import numpy as np
    from spark_deep_learning_b200 import compat
    compat.install_horovod_shim()
    import horovod.tensorflow as hvd                 # what an unmodified TF2 custom-loop main imports
    hvd.init()
    r, n = hvd.rank(), hvd.size()

    class Tape(object):                               # stands in for tf.GradientTape
        def __init__(self):
            self.entered = 0

        def __enter__(self):
            self.entered += 1
            return self

        def __exit__(self, *exc):
            return False

        def watch(self, v):
            self.watched = v

        def gradient(self, target, sources):
            return [np.full((2, 3), float(r + 1), dtype=np.float32), None, np.float32(10.0 * (r + 1))][:len(sources)]

    with hvd.DistributedGradientTape(Tape()) as tape:
        tape.watch("w")                               # forwarded to the wrapped tape
    g = tape.gradient(None, ["a", "b", "c"])
    single = hvd.DistributedGradientTape(Tape(), op=hvd.Sum).gradient(None, ["a"])
    objs = hvd.allgather_object({"rank": r})
    mean = sum(range(1, n + 1)) / n
    return (float(g[0][0, 0]), g[1], float(g[2]), float(single[0][0, 0]), [o["rank"] for o in objs], mean, tape._tape.watched)


def test_distributed_gradient_tape_averages_the_gradients():
    g0, gnone, g2, ssum, ranks, mean, watched = HorovodRunner(np=-2).run(_tape_main)
    assert g0 == mean and gnone is None and g2 == 10.0 * mean and ssum == 3.0 and ranks == [0, 1] and watched == "w"
